// rows_big.cu -- row pass of the ADMM iteration for LARGE mixed-radix widths (W = 3840 = 15*16*16, the
// 2160x3840 single-frame configuration; also 1920, 1024, 2048, 4096) on sm_100a.
//
//   packed row spectrum of x_k  --C2R-->  x_k  --prox / dual / divergence-->  v_{k+1}  --R2C-->  packed spectrum
//   (deconv.py:106 irfftn rows, :108-115 Dx/Dy/soft_thresh/dual update, :104 Dx_t/Dy_t + rfftn rows)
//
// One CTA (NT = W/15 threads) owns a band of Rb (even) image rows of one plane and marches down it one row PAIR at a
// time; the whole CTA works on one complex FFT of length W (two real rows, z = row_a + i row_b), 15 or 16 points per
// thread in registers.  Only two copies of a row pair live in shared memory (2 x W x 8 bytes = 60 KB for W = 3840, three
// CTAs per SM):
//   * inverse FFT of x pair m+1: the first pass (radix 15, every thread) takes its inputs straight from global memory
//     and does the Hermitian merge of the two packed half spectra on the fly; passes 2, 3 run in place in buffer F;
//   * spatial step: the thread that owns butterfly j of the first FORWARD pass computes v for exactly the columns
//     j + r W/15 that butterfly consumes, so v never goes through shared memory; x comes from buffers P (rows
//     ra-1, ra) and F (rows ra+1, ra+2), the pre-clamp state q from global (coalesced 128-byte runs per warp);
//   * forward passes 2, 3 run in place in P (x pair m is dead by then), the split into two packed half spectra reads P;
//   * P and F swap roles.
// q_y of the first row of the next pair is recomputed there (one more 4-byte read per pixel pair, an L2 hit) instead of
// being carried, so no third buffer is needed.
#include "common.cuh"
#include "fft_big.cuh"

namespace admm {

template <int W> struct RowBig;
template <> struct RowBig<3840> { static constexpr int R0 = 15, R1 = 16, R2 = 16; };

__device__ __forceinline__ float clampf3(float q, float tau) { return fminf(fmaxf(q, -tau), tau); }
// w = z - u with z = soft_thresh(q), u = q - z  ==>  w = q - 2 clamp(q)          (deconv.py:15-16, 104, 114-115)
__device__ __forceinline__ float wfun3(float q, float tau) { return fmaf(-2.0f, clampf3(q, tau), q); }

template <int W>
__global__ void __launch_bounds__(W / RowBig<W>::R0, 3)
k_rows_big(RowArgs a, int H, int nbands) {
    using RB = RowBig<W>;
    constexpr int R0 = RB::R0, R1 = RB::R1, R2 = RB::R2;
    constexpr int NT = W / R0;
    constexpr int Wc = W / 2;
    using I1 = BigPass<W, R0, 1, +1>;
    using I2 = BigPass<W, R1, R0, +1>;
    using I3 = BigPass<W, R2, R0 * R1, +1>;
    using F1 = BigPass<W, R0, 1, -1>;
    using F2 = BigPass<W, R1, R0, -1>;
    using F3 = BigPass<W, R2, R0 * R1, -1>;
    constexpr int RMAX = (R1 > R2 ? R1 : R2) > R0 ? (R1 > R2 ? R1 : R2) : R0;
    extern __shared__ float2 smem[];
    float2* P = smem;            // x pair m   (.x = row ra-1, .y = row ra)
    float2* F = smem + W;        // x pair m+1 (.x = row rb,   .y = row rb+1)

    const int j = threadIdx.x;
    const int band = blockIdx.x % nbands;
    const int p = blockIdx.x / nbands;
    const int hh = H >> 1;
    const int r0 = 2 * ((band * hh) / nbands);
    const int r1 = 2 * (((band + 1) * hh) / nbands);
    const int npv = (r1 - r0) / 2;
    const size_t plane_real = (size_t)p * H * W;
    const float2* __restrict__ spec = a.spec_in + (size_t)p * H * Wc;
    float2* __restrict__ sout = a.spec_out + (size_t)p * H * Wc;
    const float* __restrict__ qxi = a.qx_in ? a.qx_in + plane_real : nullptr;
    const float* __restrict__ qyi = a.qy_in ? a.qy_in + plane_real : nullptr;
    float* __restrict__ qxo = a.qx_out + plane_real;
    float* __restrict__ qyo = a.qy_out + plane_real;
    const float tau = a.lmbd[0] / a.rho[0];                     // deconv.py:44
    const float2* __restrict__ tw = a.tw;

    float2 v[RMAX];

    // inverse FFT of the row pair (rowa, rowb) into dst; the caller guarantees nobody still reads dst
    auto inverse_pair = [&](int rowa, int rowb, float2* __restrict__ dst) {
        const float2* __restrict__ Sa = spec + (size_t)rowa * Wc;
        const float2* __restrict__ Sb = spec + (size_t)rowb * Wc;
#pragma unroll
        for (int r = 0; r < R0; ++r) {
            const int n = j + r * NT;
            const bool hi = n > Wc;
            const int c = hi ? W - n : n;
            if (c == 0 || c == Wc) {
                const float2 A0 = __ldg(Sa), B0 = __ldg(Sb);
                v[r] = (c == 0) ? make_float2(A0.x, B0.x) : make_float2(A0.y, B0.y);
            } else {
                const float2 A = __ldg(Sa + c), B = __ldg(Sb + c);
                // Z[n] = Xa[n] + i Xb[n];  upper half from the Hermitian symmetry of the two real rows
                v[r] = hi ? make_float2(A.x + B.y, B.x - A.y) : make_float2(A.x - B.y, A.y + B.x);
            }
        }
        dft_big<R0, +1>(v);
        __syncthreads();                       // every reader of dst (split of the previous step) is done
        I1::store(dst, j, v);
        __syncthreads();
        if (j < I2::T) { I2::load(dst, j, v); I2::butterfly(v, j, tw); }
        __syncthreads();
        if (j < I2::T) I2::store(dst, j, v);
        __syncthreads();
        if (j < I3::T) { I3::load(dst, j, v); I3::butterfly(v, j, tw); }
        __syncthreads();
        if (j < I3::T) I3::store(dst, j, v);
        __syncthreads();
    };

    { int rm = r0 - 1; if (rm < 0) rm += H; inverse_pair(rm, r0, P); }

    for (int m = 0; m < npv; ++m) {
        const int ra = r0 + 2 * m, rb = ra + 1;
        int rc = rb + 1;
        if (rc >= H) rc -= H;
        inverse_pair(rb, rc, F);

        // ---- spatial step for the columns of this thread's first forward butterfly
        const size_t oa = (size_t)ra * W, ob = (size_t)rb * W, oc = (size_t)rc * W;
#pragma unroll
        for (int r = 0; r < R0; ++r) {
            const int c = j + r * NT;
            const int cl = (c == 0) ? W - 1 : c - 1;
            const int cr = (c == W - 1) ? 0 : c + 1;
            const float2 Pc = P[c], Pl = P[cl], Pr = P[cr];
            const float2 Fc = F[c], Fl = F[cl], Fr = F[cr];
            float uxa = 0.f, uxar = 0.f, uxb = 0.f, uxbr = 0.f, uya = 0.f, uyb = 0.f, uyc = 0.f;
            if (qxi) {                                          // previous dual u = clamp(q_prev); zero on the first iteration
                uxa  = clampf3(__ldg(qxi + oa + c), tau);
                uxar = clampf3(__ldg(qxi + oa + cr), tau);
                uxb  = clampf3(__ldg(qxi + ob + c), tau);
                uxbr = clampf3(__ldg(qxi + ob + cr), tau);
                uya  = clampf3(__ldg(qyi + oa + c), tau);
                uyb  = clampf3(__ldg(qyi + ob + c), tau);
                uyc  = clampf3(__ldg(qyi + oc + c), tau);
            }
            const float qx_a  = Pc.y - Pl.y + uxa;              // deconv.py:108,111,114
            const float qx_ar = Pr.y - Pc.y + uxar;
            const float qx_b  = Fc.x - Fl.x + uxb;
            const float qx_br = Fr.x - Fc.x + uxbr;
            const float qy_a  = Pc.y - Pc.x + uya;              // deconv.py:109,112,115
            const float qy_b  = Fc.x - Pc.y + uyb;
            const float qy_c  = Fc.y - Fc.x + uyc;
            const float wyb = wfun3(qy_b, tau);
            qxo[oa + c] = qx_a; qxo[ob + c] = qx_b;
            qyo[oa + c] = qy_a; qyo[ob + c] = qy_b;
            // v = Dx^T w_x + Dy^T w_y                           (deconv.py:104)
            v[r] = make_float2(wfun3(qx_a, tau) - wfun3(qx_ar, tau) + wfun3(qy_a, tau) - wyb,
                               wfun3(qx_b, tau) - wfun3(qx_br, tau) + wyb - wfun3(qy_c, tau));
        }
        dft_big<R0, -1>(v);
        __syncthreads();                       // all reads of P (x pair m) are done
        F1::store(P, j, v);
        __syncthreads();
        if (j < F2::T) { F2::load(P, j, v); F2::butterfly(v, j, tw); }
        __syncthreads();
        if (j < F2::T) F2::store(P, j, v);
        __syncthreads();
        if (j < F3::T) { F3::load(P, j, v); F3::butterfly(v, j, tw); }
        __syncthreads();
        if (j < F3::T) F3::store(P, j, v);
        __syncthreads();
        // ---- split Z = Va + i Vb into the two packed half spectra
        float2* __restrict__ Oa = sout + (size_t)ra * Wc;
        float2* __restrict__ Ob = sout + (size_t)rb * Wc;
        for (int c = j; c < Wc; c += NT) {
            const float2 Z = P[c];
            if (c == 0) {
                const float2 Zn = P[Wc];
                Oa[0] = make_float2(Z.x, Zn.x);
                Ob[0] = make_float2(Z.y, Zn.y);
            } else {
                const float2 Zm = P[W - c];
                Oa[c] = make_float2(0.5f * (Z.x + Zm.x), 0.5f * (Z.y - Zm.y));
                Ob[c] = make_float2(0.5f * (Z.y + Zm.y), 0.5f * (Zm.x - Z.x));
            }
        }
        float2* t = P; P = F; F = t;           // the next inverse_pair() barriers before it overwrites the old P
    }
}

bool rows_big_supported(const Geometry& g) {
    if (options().force_generic || !(options().use_big & 1)) return false;
    return g.W == 3840 && (g.H % 2 == 0) && g.H >= 4;
}

template <int W>
static int launch_rows_big_w(const Geometry& g, const RowArgs& a, cudaStream_t st) {
    constexpr int NT = W / RowBig<W>::R0;
    const size_t smem = (size_t)2 * W * sizeof(float2);
    static bool attr_set[64] = {};
    int dev = 0;
    ADMM_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_rows_big<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    } else if (dev >= 64) {
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_rows_big<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    // one wave: as many bands per plane as fill the resident-CTA slots (3 per SM), even band heights
    const int occ = (int)std::min<size_t>(3, (227 * 1024) / (smem + 1024));
    int R = options().rows_per_band;
    int nbands;
    const int hh = g.H / 2;
    if (R > 0) {
        nbands = std::max(1, std::min(hh, (g.H + R - 1) / R));
    } else {
        const int slots = 148 * occ;
        nbands = std::max(1, slots / g.P);
        // at least 8 rows per band (halo = one extra inverse FFT per band), more waves instead when P is large
        nbands = std::min(nbands, std::max(1, hh / 4));
    }
    nbands = std::max(1, std::min(nbands, hh));
    dim3 grid((unsigned)((size_t)nbands * g.P));
    ProfScope ps(PROF_ROWS, st);
    k_rows_big<W><<<grid, NT, smem, st>>>(a, g.H, nbands);
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_rows_big(const Geometry& g, const RowArgs& a, cudaStream_t st) {
    switch (g.W) {
        case 3840: return launch_rows_big_w<3840>(g, a, st);
        default: return fail(4, "no large-row kernel for this width");
    }
}

}  // namespace admm
