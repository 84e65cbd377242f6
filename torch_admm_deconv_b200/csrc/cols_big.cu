// cols_big.cu -- column pass of the ADMM iteration for LARGE mixed-radix heights (H = 2160 = 15*12*12, the
// 2160x3840 single-frame configuration) on sm_100a.
//
//   COLS_ITER:  S0 = F_col^-1[ A + Bm * F_col(S1) ]          (deconv.py:104-106 with freq_c and rho folded into A, Bm)
//   COLS_INIT:  A  = Mul * F_col(S1),  S0 = F_col^-1[A]      (deconv.py:57,99,104: freq_c * rfftn(H_t(xin)))
//
// One CTA owns a tile of C = 4 packed columns (32-byte global runs) of one plane: 4 x 180 threads, one radix-12/15
// butterfly of one column per thread per pass, the tile (69 KB) in shared memory once.  The first forward pass loads
// straight from global memory, the last forward pass (radix 12) leaves exactly the inputs of the first inverse pass
// (radix 12) in the same thread's registers, so the spectral update happens in registers between the two, and the last
// inverse pass stores straight to global memory: four shared-memory exchanges per tile instead of six.
// The inverse passes use the padded map (one spare slot per 12 entries) because their first radix is even.
#include "common.cuh"
#include "fft_big.cuh"

namespace admm {

template <int H> struct ColBig;
template <> struct ColBig<2160> { static constexpr int R0 = 15, R1 = 12, R2 = 12; };

constexpr int kColBigTile = 4;

template <int H> struct ColBigCfg {
    using CB = ColBig<H>;
    static constexpr int C = kColBigTile;
    static constexpr int T0 = H / CB::R0, T1 = H / CB::R1, T2 = H / CB::R2;
    static constexpr int TN = (T0 > T1 ? T0 : T1) > T2 ? (T0 > T1 ? T0 : T1) : T2;
    static constexpr int NT = C * TN;
    static constexpr size_t smem = (size_t)(H + H / CB::R2) * C * sizeof(float2);
};

template <int H, int MODE>
__global__ void __launch_bounds__(ColBigCfg<H>::NT, 1)
k_cols_big(ColArgs a, int Wc, int ntiles) {
    using CB = ColBig<H>;
    using CF = ColBigCfg<H>;
    constexpr int R0 = CB::R0, R1 = CB::R1, R2 = CB::R2, C = CF::C;
    using F1 = BigPass<H, R0, 1, -1, C>;
    using F2 = BigPass<H, R1, R0, -1, C>;
    using F3 = BigPass<H, R2, R0 * R1, -1, C>;
    using I1 = BigPass<H, R2, 1, +1, C, R2>;
    using I2 = BigPass<H, R1, R2, +1, C, R2>;
    using I3 = BigPass<H, R0, R2 * R1, +1, C, R2>;
    constexpr int RMAX = (R1 > R2 ? R1 : R2) > R0 ? (R1 > R2 ? R1 : R2) : R0;
    extern __shared__ float2 smem[];

    const int c = threadIdx.x % C;
    const int j = threadIdx.x / C;
    const int tile = blockIdx.x % ntiles;
    const int p = blockIdx.x / ntiles;
    const int c0 = tile * C;
    const size_t plane = (size_t)p * H * Wc;
    const float2* __restrict__ in = a.spec_in + plane + c0 + c;
    float2* __restrict__ out = a.spec_out + plane + c0 + c;
    const float2* __restrict__ tw = a.tw;
    float2* buf = smem + c;

    float2 v[RMAX];
    // ---- forward pass 1: global -> registers -> shared
    if (j < F1::T) {
#pragma unroll
        for (int r = 0; r < R0; ++r) v[r] = __ldg(in + (size_t)(j + r * F1::T) * Wc);
        dft_big<R0, -1>(v);
        F1::store(buf, j, v);
    }
    __syncthreads();
    // ---- forward pass 2 (in place)
    if (j < F2::T) { F2::load(buf, j, v); F2::butterfly(v, j, tw); }
    __syncthreads();
    if (j < F2::T) F2::store(buf, j, v);
    __syncthreads();
    // ---- forward pass 3 + spectral update + inverse pass 1, all in registers: entry r is frequency u = j + r T2
    constexpr int T2 = F3::T;
    const bool act = j < T2;
    float2 Av[R2];
    if (act) {
        // issue the table reads before the shared-memory pass so that their latency overlaps it
#pragma unroll
        for (int r = 0; r < R2; ++r) {
            const size_t o = (size_t)(j + r * T2) * Wc + c0 + c;
            if (MODE == COLS_ITER) {
                Av[r] = __ldg(a.A + plane + o);
            } else {
                Av[r] = __ldg(a.Mul + o);
            }
        }
        F3::load(buf, j, v);
        F3::butterfly(v, j, tw);
    }
    const bool col0 = (c0 == 0);                   // packed column 0 (DC + Nyquist) needs the mirrored frequency
    if (col0) {
        __syncthreads();                           // pass-3 loads done
        if (act && c == 0) {
#pragma unroll
            for (int r = 0; r < R2; ++r) smem[j + r * T2] = v[r];
        }
        __syncthreads();                           // the mirrored entries are read from shared memory where they are used
    }
    if (act) {
#pragma unroll
        for (int r = 0; r < R2; ++r) {
            const int u = j + r * T2;
            float2 o;
            if (MODE == COLS_ITER) {
                const float bm = __ldg(a.Bm + (size_t)u * Wc + c0 + c);
                o = make_float2(fmaf(bm, v[r].x, Av[r].x), fmaf(bm, v[r].y, Av[r].y));
                if (col0 && c == 0) {
                    const float bq = __ldg(a.Bq + u);
                    const float2 zm = smem[u == 0 ? 0 : H - u];
                    o.x = fmaf(bq, zm.x, o.x);
                    o.y = fmaf(-bq, zm.y, o.y);
                }
            } else {
                o = cmul(Av[r], v[r]);
                if (col0 && c == 0) o = cadd(o, cmul(__ldg(a.Mq + u), cconj(smem[u == 0 ? 0 : H - u])));
                a.A[plane + (size_t)u * Wc + c0 + c] = o;
            }
            v[r] = o;
        }
        dft_big<R2, +1>(v);
    }
    __syncthreads();                               // pass-3 loads (and the column-0 exchange) done
    if (act) I1::store(buf, j, v);
    __syncthreads();
    // ---- inverse pass 2 (in place)
    if (j < I2::T) { I2::load(buf, j, v); I2::butterfly(v, j, tw); }
    __syncthreads();
    if (j < I2::T) I2::store(buf, j, v);
    __syncthreads();
    // ---- inverse pass 3: shared -> registers -> global
    if (j < I3::T) {
        I3::load(buf, j, v);
        I3::butterfly(v, j, tw);
#pragma unroll
        for (int r = 0; r < R0; ++r) out[(size_t)(j + r * I3::T) * Wc] = v[r];
    }
}

bool cols_big_supported(const Geometry& g) {
    if (options().force_generic || !(options().use_big & 2)) return false;
    return g.H == 2160 && (g.Wc % kColBigTile == 0);
}

template <int H>
static int launch_cols_big_h(ColMode mode, const Geometry& g, const ColArgs& a, cudaStream_t st) {
    using CF = ColBigCfg<H>;
    const int ntiles = g.Wc / CF::C;
    dim3 grid((unsigned)((size_t)ntiles * g.P));
    int dev = 0;
    ADMM_CUDA_CHECK(cudaGetDevice(&dev));
    static bool attr_set[64] = {};
    if (dev >= 64 || !attr_set[dev]) {
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_cols_big<H, COLS_ITER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::smem));
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_cols_big<H, COLS_INIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::smem));
        if (dev < 64) attr_set[dev] = true;
    }
    ProfScope ps(mode == COLS_ITER ? PROF_COLS : PROF_OTHER, st);
    if (mode == COLS_ITER) k_cols_big<H, COLS_ITER><<<grid, CF::NT, CF::smem, st>>>(a, g.Wc, ntiles);
    else k_cols_big<H, COLS_INIT><<<grid, CF::NT, CF::smem, st>>>(a, g.Wc, ntiles);
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_cols_big(ColMode mode, const Geometry& g, const ColArgs& a, cudaStream_t st) {
    if (mode != COLS_ITER && mode != COLS_INIT) return fail(4, "large-column kernel: unsupported mode");
    switch (g.H) {
        case 2160: return launch_cols_big_h<2160>(mode, g, a, st);
        default: return fail(4, "no large-column kernel for this height");
    }
}

}  // namespace admm
