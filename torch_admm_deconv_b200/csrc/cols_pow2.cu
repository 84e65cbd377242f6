// cols_pow2.cu -- specialised column-pass kernel of the ADMM iteration for H in {128, 256, 512} (sm_100a).
//
//   packed row spectrum of v  --FFT along H-->  V  -->  X = A + Bm * V  --inverse FFT along H-->  row spectrum of x
//   (deconv.py:104-106: freq_c * rfftn(...) then irfftn; A and Bm fold H_t(xin), rho and 1/(HW))
//
// One CTA owns a tile of T packed columns of one plane.  Thread (t, tc) keeps 16 points of column tc in
// registers (positions t + q*H/16); the tile lives in shared memory "column fastest" so every access of a
// warp is a run of consecutive float2.  The last forward pass leaves the spectrum in exactly the register
// layout the first inverse pass consumes, so the spectral update happens in registers and the tile makes
// four shared-memory round trips in total.  Global accesses are T*8-byte row segments, 16 loads in flight
// per thread.
#include "common.cuh"
#include "fft_pow2.cuh"

namespace admm {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// forward radices (F0, F1, F2); the inverse runs them in reverse order (F2, F1, F0)
template <int H> struct ColRadix;
template <> struct ColRadix<512> { static constexpr int F0 = 8, F1 = 8, F2 = 8; };
template <> struct ColRadix<256> { static constexpr int F0 = 4, F1 = 8, F2 = 8; };
template <> struct ColRadix<128> { static constexpr int F0 = 4, F1 = 4, F2 = 8; };

template <int H> struct ColCfg {
    using CR = ColRadix<H>;
    static constexpr int TPS = H / kPT;                // threads per column
    static constexpr int T = 256 / TPS;                // columns per tile (256 threads)
    // tables: fwd pass 1 (F1, Ns=F0), fwd pass 2 (F2, Ns=F0*F1), inv pass 1 (F1, Ns=F2), inv pass 2 (F0, Ns=F2*F1);
    // identical tables are shared (all four collapse to two when F0 == F2)
    static constexpr bool kShare = (CR::F0 == CR::F2);
    static constexpr int TAB_F1 = 0;
    static constexpr int TAB_F2 = TAB_F1 + tab_size(CR::F1, CR::F0);
    static constexpr int TAB_F_END = TAB_F2 + tab_size(CR::F2, CR::F0 * CR::F1);
    static constexpr int TAB_I1 = kShare ? TAB_F1 : TAB_F_END;
    static constexpr int TAB_I2 = kShare ? TAB_F2 : TAB_I1 + tab_size(CR::F1, CR::F2);
    static constexpr int TAB_END = kShare ? TAB_F_END : TAB_I2 + tab_size(CR::F0, CR::F2 * CR::F1);
    // buf (H*T) + prefetched A tile (H*T) + tables + mirror copy of packed column 0 (H)
    static constexpr size_t bytes = (size_t)(2 * H * T + TAB_END + H) * sizeof(float2);
};

template <int H>
__global__ void __launch_bounds__(256, 3)
k_cols_iter_pow2(ColArgs a, int Wc, int ntiles) {
    using C = ColCfg<H>;
    using CR = ColRadix<H>;
    constexpr int TPS = C::TPS, T = C::T;
    extern __shared__ float2 smem[];
    float2* buf = smem;                       // H*T
    float2* abuf = buf + H * T;               // H*T: A tile, prefetched with cp.async
    float2* tabs = abuf + H * T;
    float2* zcol = tabs + C::TAB_END;         // H: copy of packed column 0 for the mirrored term
    const int tid = threadIdx.x;
    const int tc = tid % T;
    const int t = tid / T;
    const int tile = blockIdx.x % ntiles;
    const int p = blockIdx.x / ntiles;
    const int c = tile * T + tc;
    const size_t plane = (size_t)p * H * Wc;
    ColMap<T> map; map.tc = tc;

    float2 d[kPT];
    // forward pass 0 straight from global memory: slot q <-> row u = t + q*TPS
    const float2* in = a.spec_in + plane + c;
#pragma unroll
    for (int q = 0; q < kPT; ++q) d[q] = __ldg(in + (size_t)(t + q * TPS) * Wc);
    // A tile -> shared memory, asynchronously (consumed by the spectral update after the forward FFT)
    {
        const float2* Ag = a.A + plane + tile * T;
        constexpr int CH = T / 2;                          // 16-byte chunks per row segment
#pragma unroll
        for (int k = tid; k < H * CH; k += 256) {
            const int u = k / CH, part = k - u * CH;
            cp_async16(abuf + u * T + 2 * part, Ag + (size_t)u * Wc + 2 * part);
        }
        cp_async_commit();
    }
    build_tab<H, CR::F1, CR::F0>(tabs + C::TAB_F1, a.tw);
    build_tab<H, CR::F2, CR::F0 * CR::F1>(tabs + C::TAB_F2, a.tw);
    if (!C::kShare) {
        build_tab<H, CR::F1, CR::F2>(tabs + C::TAB_I1, a.tw);
        build_tab<H, CR::F0, CR::F2 * CR::F1>(tabs + C::TAB_I2, a.tw);
    }
    pass_compute<H, CR::F0, 1, -1>(d, t, nullptr);
    pass_store<H, CR::F0, 1>(d, t, buf, map);
    __syncthreads();
    pass_load<H>(d, t, buf, map);
    pass_compute<H, CR::F1, CR::F0, -1>(d, t, tabs + C::TAB_F1);
    __syncthreads();
    pass_store<H, CR::F1, CR::F0>(d, t, buf, map);
    __syncthreads();
    pass_load<H>(d, t, buf, map);
    pass_compute<H, CR::F2, CR::F0 * CR::F1, -1>(d, t, tabs + C::TAB_F2);
    // d[m + r*NB] = V[u], u = (t + m*TPS) + r*(H/F2): exactly the input layout of the first inverse pass

    // spectral update  X = A + Bm V   (+ Bq conj(V[-u]) on packed column 0, which carries DC and Nyquist)
    {
        constexpr int NB = kPT / CR::F2;
        const float* __restrict__ Bp = a.Bm + c;
        float bmv[kPT];
#pragma unroll
        for (int m = 0; m < NB; ++m)
#pragma unroll
            for (int r = 0; r < CR::F2; ++r) bmv[m + r * NB] = __ldg(Bp + (size_t)((t + m * TPS) + r * (H / CR::F2)) * Wc);
        cp_async_wait_all();
        if (tile == 0) {                                   // CTA-uniform
            if (tc == 0) {
#pragma unroll
                for (int m = 0; m < NB; ++m)
#pragma unroll
                    for (int r = 0; r < CR::F2; ++r) zcol[(t + m * TPS) + r * (H / CR::F2)] = d[m + r * NB];
            }
        }
        __syncthreads();                                   // A tile (and zcol) visible to every thread
#pragma unroll
        for (int m = 0; m < NB; ++m) {
#pragma unroll
            for (int r = 0; r < CR::F2; ++r) {
                const int u = (t + m * TPS) + r * (H / CR::F2);
                const float2 Av = abuf[map.at(u)];
                const float bm = bmv[m + r * NB];
                float2 Z = d[m + r * NB];
                float2 o = make_float2(fmaf(bm, Z.x, Av.x), fmaf(bm, Z.y, Av.y));
                if (tile == 0 && tc == 0) {
                    const float2 Zm = zcol[(H - u) & (H - 1)];
                    const float bq = a.Bq[u];
                    o.x = fmaf(bq, Zm.x, o.x);
                    o.y = fmaf(-bq, Zm.y, o.y);
                }
                d[m + r * NB] = o;
            }
        }
    }
    // inverse pass 0 (radix F2, no twiddles) from registers
    pass_compute<H, CR::F2, 1, +1>(d, t, nullptr);
    // every thread passed the barrier above after its last read of buf (forward pass 2 loads)
    pass_store<H, CR::F2, 1>(d, t, buf, map);
    __syncthreads();
    pass_load<H>(d, t, buf, map);
    pass_compute<H, CR::F1, CR::F2, +1>(d, t, tabs + C::TAB_I1);
    __syncthreads();
    pass_store<H, CR::F1, CR::F2>(d, t, buf, map);
    __syncthreads();
    pass_load<H>(d, t, buf, map);
    pass_compute<H, CR::F0, CR::F2 * CR::F1, +1>(d, t, tabs + C::TAB_I2);
    // natural order: slot (m, r) -> row u = (t + m*TPS) + r*(H/F0)
    {
        constexpr int NB = kPT / CR::F0;
        float2* out = a.spec_out + plane + c;
#pragma unroll
        for (int m = 0; m < NB; ++m)
#pragma unroll
            for (int r = 0; r < CR::F0; ++r)
                out[(size_t)((t + m * TPS) + r * (H / CR::F0)) * Wc] = d[m + r * NB];
    }
}

template <int H>
static int launch_cols_pow2_t(const Geometry& g, const ColArgs& a, cudaStream_t st) {
    using C = ColCfg<H>;
    const int ntiles = g.Wc / C::T;
    static bool attr_set = false;
    if (!attr_set) {
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_cols_iter_pow2<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::bytes));
        attr_set = true;
    }
    ProfScope ps(PROF_COLS, st);
    k_cols_iter_pow2<H><<<(unsigned)((size_t)ntiles * g.P), 256, C::bytes, st>>>(a, g.Wc, ntiles);
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

bool cols_pow2_supported(const Geometry& g) {
    if (options().force_generic) return false;
    int T = 0;
    switch (g.H) {
        case 128: T = ColCfg<128>::T; break;
        case 256: T = ColCfg<256>::T; break;
        case 512: T = ColCfg<512>::T; break;
        default: return false;
    }
    return (g.W % 2 == 0) && (g.Wc % T == 0);
}

int launch_cols_pow2(const Geometry& g, const ColArgs& a, cudaStream_t st) {
    switch (g.H) {
        case 128: return launch_cols_pow2_t<128>(g, a, st);
        case 256: return launch_cols_pow2_t<256>(g, a, st);
        case 512: return launch_cols_pow2_t<512>(g, a, st);
    }
    return fail(4, "cols_pow2: unsupported height");
}

}  // namespace admm
