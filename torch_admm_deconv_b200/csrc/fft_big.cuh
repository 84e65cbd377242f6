// fft_big.cuh -- compile-time mixed-radix Stockham passes for the LARGE transform lengths (3840 = 15*16*16,
// 2160 = 15*12*12, ...) of the single-frame configuration (BASELINE.json configs[2]: 2160x3840, "large-FFT path").
//
// One transform of length N is spread over N/R threads per pass, every thread holding one radix-R butterfly (12..16
// complex points) in registers; passes exchange data through ONE shared-memory copy of the sequence (in place: all
// loads of a pass are in registers before a block barrier, the stores follow it).  The first radix is odd (15), so the
// stride-R stores of the first pass fall on distinct banks without padding; later passes store runs of >= 15
// consecutive entries.  Twiddles come from the per-call global table e^{-2 pi i n/N} (L1 resident).
//
// Pass p with radix R and NS = product of the earlier radices (same indexing as oracle/packed_layout_model.py
// `stockham_fft`): butterfly j in [0, N/R): k = j % NS, inputs src[j + r N/R] * w_N^{k r N/(NS R)}, outputs
// dst[(j-k) R + k + r NS].
#pragma once
#include "fft_engine.cuh"

namespace admm {

template <int DIR> __device__ __forceinline__ void dft12(float2* v) {
    // N1 = 4 (n1), N2 = 3 (n2): n = 3 n1 + n2, k = k1 + 4 k2
    const float c1 = 0.86602540378443864676f, s1 = 0.5f;     // w12^1 = cos30 -+ i sin30
    float2 y[3][4];
#pragma unroll
    for (int n2 = 0; n2 < 3; ++n2) {
        float2 t[4] = {v[n2], v[n2 + 3], v[n2 + 6], v[n2 + 9]};
        dft4<DIR>(t);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) y[n2][k1] = t[k1];
    }
    // twiddles w12^{n2 k1}: n2=1: 1,2,3 ; n2=2: 2,4,6
    y[1][1] = mul_tw<DIR>(y[1][1], c1, s1);
    y[1][2] = mul_tw<DIR>(y[1][2], s1, c1);
    y[1][3] = mul_dir_i<DIR>(y[1][3]);
    y[2][1] = mul_tw<DIR>(y[2][1], s1, c1);
    y[2][2] = mul_tw<DIR>(y[2][2], -s1, c1);
    y[2][3] = make_float2(-y[2][3].x, -y[2][3].y);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        float2 t[3] = {y[0][k1], y[1][k1], y[2][k1]};
        dft3<DIR>(t);
#pragma unroll
        for (int k2 = 0; k2 < 3; ++k2) v[k1 + 4 * k2] = t[k2];
    }
}

// radix 6 and 10 (2 x 3, 2 x 5): n = N2 n1 + n2 with N1 = 2, k = k1 + 2 k2
template <int DIR> __device__ __forceinline__ void dft6(float2* v) {
    float2 y[3][2];
#pragma unroll
    for (int n2 = 0; n2 < 3; ++n2) { y[n2][0] = cadd(v[n2], v[n2 + 3]); y[n2][1] = csub(v[n2], v[n2 + 3]); }
    y[1][1] = mul_tw<DIR>(y[1][1], 0.5f, 0.86602540378443864676f);
    y[2][1] = mul_tw<DIR>(y[2][1], -0.5f, 0.86602540378443864676f);
#pragma unroll
    for (int k1 = 0; k1 < 2; ++k1) {
        float2 t[3] = {y[0][k1], y[1][k1], y[2][k1]};
        dft3<DIR>(t);
#pragma unroll
        for (int k2 = 0; k2 < 3; ++k2) v[k1 + 2 * k2] = t[k2];
    }
}
template <int DIR> __device__ __forceinline__ void dft10(float2* v) {
    float2 y[5][2];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) { y[n2][0] = cadd(v[n2], v[n2 + 5]); y[n2][1] = csub(v[n2], v[n2 + 5]); }
    y[1][1] = mul_tw<DIR>(y[1][1], 0.80901699437494742410f, 0.58778525229247312917f);
    y[2][1] = mul_tw<DIR>(y[2][1], 0.30901699437494742410f, 0.95105651629515357212f);
    y[3][1] = mul_tw<DIR>(y[3][1], -0.30901699437494742410f, 0.95105651629515357212f);
    y[4][1] = mul_tw<DIR>(y[4][1], -0.80901699437494742410f, 0.58778525229247312917f);
#pragma unroll
    for (int k1 = 0; k1 < 2; ++k1) {
        float2 t[5] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1], y[4][k1]};
        dft5<DIR>(t);
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) v[k1 + 2 * k2] = t[k2];
    }
}

template <int R, int DIR> __device__ __forceinline__ void dft_big(float2* v) {
    if (R == 12) dft12<DIR>(v);
    else if (R == 10) dft10<DIR>(v);
    else if (R == 6) dft6<DIR>(v);
    else dftR<R, DIR>(v);
}

// STRIDE = distance (in float2) between consecutive sequence entries in shared memory: 1 for the row kernels (one
// sequence per CTA), the tile width for the column kernels (entry n of column c at n * STRIDE + c).
// PAD = q > 0 places entry n at slot n + n/q: one spare slot per q entries, so that the stride-R stores of a first pass
// with an EVEN radix R = q spread over the banks (needs q | N/R of the passes that load, and q | NS or NS == 1, R == q
// for the passes that store -- all divisions then fold into constants).
template <int N, int R, int NS, int DIR, int STRIDE = 1, int PAD = 0>
struct BigPass {
    static constexpr int T = N / R;                 // butterflies = active threads per sequence
    static constexpr int TWS = N / (NS * R);
    static_assert(T * R == N && TWS * NS * R == N, "radix schedule does not multiply up to N");

    static __device__ __forceinline__ void load(const float2* __restrict__ src, int j, float2* v) {
        if (PAD == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = src[(j + r * T) * STRIDE];
        } else {
            static_assert(PAD == 0 || T % (PAD ? PAD : 1) == 0, "padded load needs PAD | N/R");
            const int b = j + j / (PAD ? PAD : 1);
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = src[(b + r * (T + T / (PAD ? PAD : 1))) * STRIDE];
        }
    }
    static __device__ __forceinline__ void twiddle(float2* v, int j, const float2* __restrict__ tw) {
        if (NS == 1) return;
        const int k = j % NS;
#pragma unroll
        for (int r = 1; r < R; ++r) {
            float2 w = __ldg(tw + k * r * TWS);
            if (DIR > 0) w.y = -w.y;
            v[r] = cmul(v[r], w);
        }
    }
    static __device__ __forceinline__ void butterfly(float2* v, int j, const float2* __restrict__ tw) {
        twiddle(v, j, tw);
        dft_big<R, DIR>(v);
    }
    // twiddles from a shared-memory table laid out per thread: entry r-1 at tab[(r-1) * TSTRIDE] (forward sign)
    template <int TSTRIDE>
    static __device__ __forceinline__ void butterfly_tab(float2* v, const float2* __restrict__ tab) {
#pragma unroll
        for (int r = 1; r < R; ++r) {
            float2 w = tab[(r - 1) * TSTRIDE];
            if (DIR > 0) w.y = -w.y;
            v[r] = cmul(v[r], w);
        }
        dft_big<R, DIR>(v);
    }
    // twiddles held in registers by the caller (w[r-1] = forward twiddle of input r)
    static __device__ __forceinline__ void butterfly_reg(float2* v, const float2* w) {
#pragma unroll
        for (int r = 1; r < R; ++r) {
            float2 t = w[r - 1];
            if (DIR > 0) t.y = -t.y;
            v[r] = cmul(v[r], t);
        }
        dft_big<R, DIR>(v);
    }
    static __device__ __forceinline__ void store(float2* __restrict__ dst, int j, const float2* v) {
        const int k = j % NS;
        const int j0 = (j - k) * R + k;
        if (PAD == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) dst[(j0 + r * NS) * STRIDE] = v[r];
        } else if (NS == 1) {
            static_assert(PAD == 0 || NS != 1 || R == PAD, "padded first pass needs R == PAD");
#pragma unroll
            for (int r = 0; r < R; ++r) dst[(j0 + j + r) * STRIDE] = v[r];
        } else {
            static_assert(PAD == 0 || NS == 1 || NS % (PAD ? PAD : 1) == 0, "padded store needs PAD | NS");
            const int b = j0 + j0 / (PAD ? PAD : 1);
#pragma unroll
            for (int r = 0; r < R; ++r) dst[(b + r * (NS + NS / (PAD ? PAD : 1))) * STRIDE] = v[r];
        }
    }
};

}  // namespace admm
