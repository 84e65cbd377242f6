// fft_pow2.cuh -- register-resident Stockham passes for power-of-two lengths (sm_100a).
//
// Every thread owns PT = 16 complex points of one length-N sequence, logical slot q <-> position
// t + q*TPS (TPS = N/16 threads per sequence).  A pass with radix R groups the 16 slots into 16/R
// butterflies (butterfly m takes slots q = m + r*(16/R), i.e. positions j + r*N/R with j = t + m*TPS),
// multiplies by the Stockham twiddles, runs the radix-R DFT in registers and scatters the results to the
// autosort positions (j/Ns)*Ns*R + j%Ns + r*Ns.  Twiddles come from small per-pass shared-memory tables
// tab[(r-1)*Ns + k] = e^{-2 pi i k r/(Ns R)} built once per CTA from the fp64-generated global table, so
// consecutive lanes read consecutive table entries.
//
// Two shared-memory index maps are used:
//   RowMap: one sequence (a pair of image rows) per region, one pad slot every 16 complex (contiguous runs
//           and the stride-8 scatter of the first pass are bank-conflict free);
//   ColMap: T sequences interleaved "column fastest" (position i of column tc at i*T + tc), conflict free
//           for any pattern because consecutive lanes are consecutive columns.
#pragma once
#include "fft_engine.cuh"

namespace admm {

constexpr int kPT = 16;   // complex points per thread

template <int X> struct Log2 { static constexpr int value = 1 + Log2<X / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

// Index maps expose at(i) plus the split at(b + off) == base(b) + delta(off) that holds for the
// (base, compile-time offset) combinations the passes use, so every shared-memory access is
// "register + immediate".
struct RowMapObj {           // pad one complex slot every 16: contiguous runs of 16 stay conflict free
    __device__ __forceinline__ int at(int i) const { return i + (i >> 4); }
    __device__ __forceinline__ int base(int i) const { return i + (i >> 4); }
    __device__ __forceinline__ static constexpr int delta(int off) { return off + (off >> 4); }
};
template <int W> constexpr int row_region() { return W + W / 16; }      // complex slots per row pair

template <int T> struct ColMap {
    int tc;
    __device__ __forceinline__ int at(int i) const { return i * T + tc; }
    __device__ __forceinline__ int base(int i) const { return i * T + tc; }
    __device__ __forceinline__ static constexpr int delta(int off) { return off * T; }
};

// number of table entries of a pass (radix R, Ns previous product)
constexpr int tab_size(int R, int Ns) { return Ns > 1 ? (R - 1) * Ns : 0; }

// Build tab[(r-1)*Ns + k] = tw[k r N/(Ns R)] (forward sign) cooperatively.
template <int N, int R, int NS>
__device__ __forceinline__ void build_tab(float2* tab, const float2* __restrict__ tw_global) {
    if (NS > 1) {
        constexpr int tws = N / (NS * R);
        for (int e = threadIdx.x; e < (R - 1) * NS; e += blockDim.x) {
            const int r = e / NS + 1, k = e - (r - 1) * NS;
            tab[e] = tw_global[k * r * tws];
        }
    }
}

// Twiddle + radix-R DFTs on the 16 register slots (in place: slot m + r*NB <- output r of butterfly m).
template <int N, int R, int NS, int DIR>
__device__ __forceinline__ void pass_compute(float2 (&d)[kPT], int t, const float2* __restrict__ tab) {
    constexpr int TPS = N / kPT, NB = kPT / R;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * TPS;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = d[m + r * NB];
        if (NS > 1) {
            const int k = j & (NS - 1);
#pragma unroll
            for (int r = 1; r < R; ++r) {
                float2 w = tab[(r - 1) * NS + k];
                if (DIR > 0) w.y = -w.y;
                v[r] = cmul(v[r], w);
            }
        }
        dftR<R, DIR>(v);
#pragma unroll
        for (int r = 0; r < R; ++r) d[m + r * NB] = v[r];
    }
}

// Autosort scatter of the pass results: slot (m, r) -> position (j/Ns)*Ns*R + j%Ns + r*Ns.
template <int N, int R, int NS, class Map>
__device__ __forceinline__ void pass_store(const float2 (&d)[kPT], int t, float2* __restrict__ reg, const Map& map) {
    constexpr int TPS = N / kPT, NB = kPT / R;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * TPS;
        const int k = j & (NS - 1);
        const int b = map.base((j - k) * R + k);
#pragma unroll
        for (int r = 0; r < R; ++r) reg[b + Map::delta(r * NS)] = d[m + r * NB];
    }
}

// Gather the 16 slots: slot q <- position t + q*TPS.
template <int N, class Map>
__device__ __forceinline__ void pass_load(float2 (&d)[kPT], int t, const float2* __restrict__ reg, const Map& map) {
    constexpr int TPS = N / kPT;
    const int b = map.base(t);
#pragma unroll
    for (int q = 0; q < kPT; ++q) d[q] = reg[b + Map::delta(q * TPS)];
}

// ---- row-pass schedule shared by rows_pow2.cu and cluster_pow2.cu
template <int W> struct RowRadix;
template <> struct RowRadix<512> { static constexpr int IB = 8, IC = 8, FA = 8, FB = 8; };
template <> struct RowRadix<256> { static constexpr int IB = 8, IC = 4, FA = 4, FB = 8; };
template <> struct RowRadix<128> { static constexpr int IB = 4, IC = 4, FA = 4, FB = 4; };

template <int W> struct RowSmem {
    using RR = RowRadix<W>;
    static constexpr int kThreads = 256;
    static constexpr int TPS = W / kPT;                       // threads per row pair
    static constexpr int NPAIR = kThreads / TPS;              // row pairs per CTA (x side, halo included)
    static constexpr int REGION = row_region<W>();            // padded complex slots per pair
    static constexpr int RMAX = 2 * NPAIR - 2;                // band rows per CTA
    // twiddle tables (forward sign; the inverse passes conjugate): inverse pass B (IB, Ns=8), inverse pass C
    // (IC, Ns=8*IB), forward pass B (FB, Ns=FA), forward pass C (8, Ns=W/8); identical tables are shared
    static constexpr bool kShareB = (RR::FB == RR::IB) && (RR::FA == 8);
    static constexpr bool kShareC = (RR::IC == 8) && (8 * RR::IB == W / 8);
    static constexpr int TAB_IB = 0;
    static constexpr int TAB_IC = TAB_IB + tab_size(RR::IB, 8);
    static constexpr int TAB_I_END = TAB_IC + tab_size(RR::IC, 8 * RR::IB);
    static constexpr int TAB_FB = kShareB ? TAB_IB : TAB_I_END;
    static constexpr int TAB_FB_END = kShareB ? TAB_I_END : TAB_FB + tab_size(RR::FB, RR::FA);
    static constexpr int TAB_FC = kShareC ? TAB_IC : TAB_FB_END;
    static constexpr int TAB_END = kShareC ? TAB_FB_END : TAB_FC + tab_size(8, W / 8);
    static constexpr int SIDE = (kThreads / 32) * NPAIR * 2;      // per warp and pair slot: the two columns next to the warp's strip
    static constexpr size_t bytes = (size_t)(NPAIR * REGION + TAB_END + SIDE) * sizeof(float2);
};

}  // namespace admm
