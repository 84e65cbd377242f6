// cols_common.cuh -- shared pieces of the power-of-two column kernels (cols_pow2.cu, coop_small.cu): radix
// schedules, shared-memory budget, and the two-columns-per-thread Stockham passes on 16-byte words.
#pragma once
#include "common.cuh"
#include "fft_pow2.cuh"

namespace admm {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

constexpr int kCP = 8;    // points per thread per column
// threads per CTA of the register-staged column kernel: 512 for H = 512 (tiles of 16 columns = 128-byte global runs)
template <int H> constexpr int col_threads() { return H >= 512 ? 512 : 256; }

// forward radices (F0, F1, F2); the inverse runs them in reverse order (F2, F1, F0)
template <int H> struct ColRadix;
template <> struct ColRadix<512> { static constexpr int F0 = 8, F1 = 8, F2 = 8; };
template <> struct ColRadix<256> { static constexpr int F0 = 4, F1 = 8, F2 = 8; };
template <> struct ColRadix<128> { static constexpr int F0 = 4, F1 = 4, F2 = 8; };

template <int H, int NT = col_threads<H>()> struct ColCfg {
    static constexpr int kThreads = NT;
    using CR = ColRadix<H>;
    static constexpr int TPS = H / kCP;                // threads per column pair
    static constexpr int NPAIRS = NT / TPS;            // column pairs per tile
    static constexpr int T = 2 * NPAIRS;               // columns per tile
    // tables: fwd pass 1 (F1, Ns=F0), fwd pass 2 (F2, Ns=F0*F1), inv pass 1 (F1, Ns=F2), inv pass 2 (F0, Ns=F2*F1);
    // identical tables are shared (all four collapse to two when F0 == F2)
    static constexpr bool kShare = (CR::F0 == CR::F2);
    static constexpr int TAB_F1 = 0;
    static constexpr int TAB_F2 = TAB_F1 + tab_size(CR::F1, CR::F0);
    static constexpr int TAB_F_END = TAB_F2 + tab_size(CR::F2, CR::F0 * CR::F1);
    static constexpr int TAB_I1 = kShare ? TAB_F1 : TAB_F_END;
    static constexpr int TAB_I2 = kShare ? TAB_F2 : TAB_I1 + tab_size(CR::F1, CR::F2);
    static constexpr int TAB_END = kShare ? TAB_F_END : TAB_I2 + tab_size(CR::F0, CR::F2 * CR::F1);
    // buf (H*T) + tables + mirror copy of packed column 0 (H)
    static constexpr size_t bytes = (size_t)(H * T + TAB_END + H) * sizeof(float2);
};

// The column passes read only the first power of every twiddle (tab[k] = e^{-2 pi i k/(Ns R)}, k < Ns; the other powers are
// formed by multiplication): build just those Ns entries (8 + 64 instead of 504 loads per CTA for H = 512)
template <int N, int R, int NS>
__device__ __forceinline__ void build_tab1(float2* tab, const float2* __restrict__ tw_global) {
    if (NS > 1) {
        constexpr int tws = N / (NS * R);
        for (int k = threadIdx.x; k < NS; k += blockDim.x) tab[k] = tw_global[k * tws];
    }
}

// ---- passes on two columns at once: d[q] = (col0.re, col0.im, col1.re, col1.im) of slot q <-> position t + q*TPS
template <int H, int R, int NS, int DIR>
__device__ __forceinline__ void cpass_compute(float4 (&d)[kCP], int t, const float2* __restrict__ tab) {
    constexpr int TPS = H / kCP, NB = kCP / R;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * TPS;
        float2 a[R], b[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            a[r] = make_float2(d[m + r * NB].x, d[m + r * NB].y);
            b[r] = make_float2(d[m + r * NB].z, d[m + r * NB].w);
        }
        if (NS > 1) {
            const int k = j & (NS - 1);
            // one table read per butterfly, the other powers by multiplication (relieves the shared-memory pipe)
            float2 w1 = tab[k];
            if (DIR > 0) w1.y = -w1.y;
            float2 wp[R];
            wp[1] = w1;
#pragma unroll
            for (int r = 2; r < R; ++r) wp[r] = (r & 1) ? cmul(wp[r - 1], w1) : cmul(wp[r / 2], wp[r / 2]);
#pragma unroll
            for (int r = 1; r < R; ++r) { a[r] = cmul(a[r], wp[r]); b[r] = cmul(b[r], wp[r]); }
        }
        dftR<R, DIR>(a);
        dftR<R, DIR>(b);
#pragma unroll
        for (int r = 0; r < R; ++r) d[m + r * NB] = make_float4(a[r].x, a[r].y, b[r].x, b[r].y);
    }
}

// shared tile as float4 words: word index = position * NPAIRS + pair
template <int H, int R, int NS, int NPAIRS>
__device__ __forceinline__ void cpass_store(const float4 (&d)[kCP], int t, int pr, float4* __restrict__ buf) {
    constexpr int TPS = H / kCP, NB = kCP / R;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * TPS;
        const int k = j & (NS - 1);
        const int b = ((j - k) * R + k) * NPAIRS + pr;
#pragma unroll
        for (int r = 0; r < R; ++r) buf[b + r * NS * NPAIRS] = d[m + r * NB];
    }
}

template <int H, int NPAIRS>
__device__ __forceinline__ void cpass_load(float4 (&d)[kCP], int t, int pr, const float4* __restrict__ buf) {
    constexpr int TPS = H / kCP;
    const int b = t * NPAIRS + pr;
#pragma unroll
    for (int q = 0; q < kCP; ++q) d[q] = buf[b + q * TPS * NPAIRS];
}


}  // namespace admm
