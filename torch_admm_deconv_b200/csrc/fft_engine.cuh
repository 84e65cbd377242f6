// fft_engine.cuh -- generic (any length) batched shared-memory FFT for sm_100a.
//
// Layout: NB independent length-N complex sequences live in shared memory "batch fastest":
//     buf[n * BS + b],  n in [0,N), b in [0,NB), BS >= NB.
// Consecutive threads take consecutive b for the same butterfly index, so every shared-memory access of
// a warp is a run of consecutive float2 (conflict-free for any N, any radix) and the twiddle is
// warp-uniform.  The transform is an autosort Stockham: natural order in, natural order out, ping-pong
// between two buffers, one __syncthreads per pass.  Radices 2,3,4,5,8,9,15,16 are unrolled in registers; any
// other prime factor runs a generic O(p^2) pass, so every length is supported (D7 of SURVEY.md: the
// reference transforms exactly (H, W), deconv.py:49,104-106, no padding to a power of two allowed).
//
// This is the fallback / arbitrary-size engine.  The specialised power-of-two kernels are in
// fft_pow2.cuh.
#pragma once
#include <cuda_runtime.h>

namespace admm {

constexpr int kMaxPasses = 24;

struct FftPlan {
    int n;
    int npass;
    int radix[kMaxPasses];
};

__host__ inline bool make_plan(int n, FftPlan& p) {
    p.n = n; p.npass = 0;
    if (n < 1) return false;
    int m = n;
    // big composite radices first (fewest shared-memory passes), then what is left of the powers of 2, 3, 5
    const int pref[9] = {16, 15, 9, 8, 4, 2, 3, 5, 7};
    for (int i = 0; i < 9; ++i) {
        while (m > 1 && m % pref[i] == 0) {
            if (p.npass >= kMaxPasses) return false;
            p.radix[p.npass++] = pref[i]; m /= pref[i];
        }
    }
    for (int f = 11; m > 1; f += 2) {
        while (m % f == 0) {
            if (p.npass >= kMaxPasses) return false;
            p.radix[p.npass++] = f; m /= f;
        }
        if (f > 46341) return false;
    }
    return true;
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i*DIR... : DIR=-1 (forward) -> multiply by -i ; DIR=+1 (inverse) -> multiply by +i
template <int DIR> __device__ __forceinline__ float2 mul_dir_i(float2 a) {
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
// forward twiddle table holds e^{-2 pi i n/N}; the inverse transform conjugates it
template <int DIR> __device__ __forceinline__ float2 tw_load(const float2* tw, int idx) {
    float2 w = tw[idx];
    if (DIR > 0) w.y = -w.y;
    return w;
}

// ---------------------------------------------------------------- in-register DFTs, natural order out
template <int DIR> __device__ __forceinline__ void dft2(float2& a, float2& b) {
    float2 t = a; a = cadd(t, b); b = csub(t, b);
}
template <int DIR> __device__ __forceinline__ void dft3(float2* v) {
    const float c = -0.5f, s = 0.86602540378443864676f;
    float2 t1 = cadd(v[1], v[2]);
    float2 t2 = make_float2(fmaf(c, t1.x, v[0].x), fmaf(c, t1.y, v[0].y));
    float2 d = csub(v[1], v[2]);
    float2 t3 = mul_dir_i<DIR>(make_float2(s * d.x, s * d.y));   // -i s d (fwd) / +i s d (inv)
    v[0] = cadd(v[0], t1);
    v[1] = cadd(t2, t3);
    v[2] = csub(t2, t3);
}
template <int DIR> __device__ __forceinline__ void dft4(float2* v) {
    float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]), t3 = mul_dir_i<DIR>(csub(v[1], v[3]));
    v[0] = cadd(t0, t2); v[2] = csub(t0, t2);
    v[1] = cadd(t1, t3); v[3] = csub(t1, t3);
}
template <int DIR> __device__ __forceinline__ void dft5(float2* v) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    float2 r1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    float2 r2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    float2 i1 = mul_dir_i<DIR>(make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
    float2 i2 = mul_dir_i<DIR>(make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
    v[0] = cadd(v[0], cadd(a1, a2));
    v[1] = cadd(r1, i1); v[4] = csub(r1, i1);
    v[2] = cadd(r2, i2); v[3] = csub(r2, i2);
}
template <int DIR> __device__ __forceinline__ void dft8(float2* v) {
    const float h = 0.70710678118654752440f;
    // radix-2 stage
    float2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
    float2 a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
    float2 a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]);
    float2 a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
    // twiddles w8^k on the odd half: w8 = (1 -+ i)/sqrt2, w8^2 = -+i, w8^3 = (-1 -+ i)/sqrt2
    float2 t5 = mul_dir_i<DIR>(a5);                       // -+i a5
    a5 = make_float2(h * (a5.x + t5.x), h * (a5.y + t5.y));   // (1 -+ i)/sqrt2 * a5
    a6 = mul_dir_i<DIR>(a6);
    float2 t7 = mul_dir_i<DIR>(a7);
    a7 = make_float2(h * (t7.x - a7.x), h * (t7.y - a7.y));   // (-1 -+ i)/sqrt2 * a7
    // two radix-4 on (a0,a1,a2,a3) and (a4,a5,a6,a7)
    float2 e[4] = {a0, a1, a2, a3};
    float2 o[4] = {a4, a5, a6, a7};
    dft4<DIR>(e); dft4<DIR>(o);
    v[0] = e[0]; v[2] = e[1]; v[4] = e[2]; v[6] = e[3];
    v[1] = o[0]; v[3] = o[1]; v[5] = o[2]; v[7] = o[3];
}
// multiply by e^{-+ 2 pi i m / n} given (cos, sin) of the positive angle: forward uses (c, -s), inverse (c, +s)
template <int DIR> __device__ __forceinline__ float2 mul_tw(float2 a, float c, float s) {
    return DIR < 0 ? make_float2(fmaf(a.x, c, a.y * s), fmaf(a.y, c, -a.x * s))
                   : make_float2(fmaf(a.x, c, -a.y * s), fmaf(a.y, c, a.x * s));
}
// Composite radices (Cooley-Tukey N1 x N2 in registers, n = N2 n1 + n2, k = k1 + N1 k2): they halve the number of
// shared-memory passes of the mixed-radix sizes (3840 = 16*16*15, 2160 = 16*15*9).
template <int DIR> __device__ __forceinline__ void dft16(float2* v) {
    const float c1 = 0.92387953251128673848f, s1 = 0.38268343236508978178f, h = 0.70710678118654752440f;
    float2 y[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        float2 t[4] = {v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]};
        dft4<DIR>(t);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) y[n2][k1] = t[k1];
    }
    // twiddles w16^{n2 k1}
    y[1][1] = mul_tw<DIR>(y[1][1], c1, s1);  y[1][2] = mul_tw<DIR>(y[1][2], h, h);    y[1][3] = mul_tw<DIR>(y[1][3], s1, c1);
    y[2][1] = mul_tw<DIR>(y[2][1], h, h);    y[2][2] = mul_dir_i<DIR>(y[2][2]);       y[2][3] = mul_tw<DIR>(y[2][3], -h, h);
    y[3][1] = mul_tw<DIR>(y[3][1], s1, c1);  y[3][2] = mul_tw<DIR>(y[3][2], -h, h);   y[3][3] = mul_tw<DIR>(y[3][3], -c1, -s1);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        float2 t[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
        dft4<DIR>(t);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = t[k2];
    }
}
template <int DIR> __device__ __forceinline__ void dft9(float2* v) {
    float2 y[3][3];
#pragma unroll
    for (int n2 = 0; n2 < 3; ++n2) {
        float2 t[3] = {v[n2], v[n2 + 3], v[n2 + 6]};
        dft3<DIR>(t);
#pragma unroll
        for (int k1 = 0; k1 < 3; ++k1) y[n2][k1] = t[k1];
    }
    y[1][1] = mul_tw<DIR>(y[1][1], 0.76604444311897801345f, 0.64278760968653925190f);
    y[1][2] = mul_tw<DIR>(y[1][2], 0.17364817766693041445f, 0.98480775301220802032f);
    y[2][1] = mul_tw<DIR>(y[2][1], 0.17364817766693041445f, 0.98480775301220802032f);
    y[2][2] = mul_tw<DIR>(y[2][2], -0.93969262078590831688f, 0.34202014332566887944f);
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1) {
        float2 t[3] = {y[0][k1], y[1][k1], y[2][k1]};
        dft3<DIR>(t);
#pragma unroll
        for (int k2 = 0; k2 < 3; ++k2) v[k1 + 3 * k2] = t[k2];
    }
}
template <int DIR> __device__ __forceinline__ void dft15(float2* v) {
    // N1 = 3 (n1), N2 = 5 (n2): n = 5 n1 + n2, k = k1 + 3 k2
    float2 y[5][3];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        float2 t[3] = {v[n2], v[n2 + 5], v[n2 + 10]};
        dft3<DIR>(t);
#pragma unroll
        for (int k1 = 0; k1 < 3; ++k1) y[n2][k1] = t[k1];
    }
    // twiddles w15^{n2 k1}: exponents 1,2 | 2,4 | 3,6 | 4,8
    y[1][1] = mul_tw<DIR>(y[1][1], 0.91354545764260086660f, 0.40673664307580015276f);
    y[1][2] = mul_tw<DIR>(y[1][2], 0.66913060635885823757f, 0.74314482547739413310f);
    y[2][1] = mul_tw<DIR>(y[2][1], 0.66913060635885823757f, 0.74314482547739413310f);
    y[2][2] = mul_tw<DIR>(y[2][2], -0.10452846326765333207f, 0.99452189536827340088f);
    y[3][1] = mul_tw<DIR>(y[3][1], 0.30901699437494745126f, 0.95105651629515353118f);
    y[3][2] = mul_tw<DIR>(y[3][2], -0.80901699437494734024f, 0.58778525229247324813f);
    y[4][1] = mul_tw<DIR>(y[4][1], -0.10452846326765333207f, 0.99452189536827340088f);
    y[4][2] = mul_tw<DIR>(y[4][2], -0.97814760073380568883f, -0.20791169081775906502f);
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1) {
        float2 t[5] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1], y[4][k1]};
        dft5<DIR>(t);
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) v[k1 + 3 * k2] = t[k2];
    }
}
template <int R, int DIR> __device__ __forceinline__ void dftR(float2* v) {
    if (R == 2) dft2<DIR>(v[0], v[1]);
    else if (R == 3) dft3<DIR>(v);
    else if (R == 4) dft4<DIR>(v);
    else if (R == 5) dft5<DIR>(v);
    else if (R == 8) dft8<DIR>(v);
    else if (R == 9) dft9<DIR>(v);
    else if (R == 15) dft15<DIR>(v);
    else if (R == 16) dft16<DIR>(v);
}

// ---------------------------------------------------------------- one Stockham pass
template <int R, int DIR>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ src, float2* __restrict__ dst,
                                              int N, int Ns, int NB, int BS, const float2* __restrict__ tw) {
    const int T = N / R;
    const int tws = N / (Ns * R);
    const int work = T * NB;
    for (int w = threadIdx.x; w < work; w += blockDim.x) {
        const int j = w / NB;
        const int b = w - j * NB;
        const int k = j % Ns;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = src[(j + r * T) * BS + b];
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw_load<DIR>(tw, k * r * tws));
        }
        dftR<R, DIR>(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) dst[(j0 + r * Ns) * BS + b] = v[r];
    }
}

// generic radix (any prime p): out[m] = sum_r src[j + rT] * w_N^{ r (k + Ns m) N/(Ns p) }
template <int DIR>
__device__ __forceinline__ void stockham_pass_generic(const float2* __restrict__ src, float2* __restrict__ dst,
                                                      int N, int R, int Ns, int NB, int BS,
                                                      const float2* __restrict__ tw) {
    const int T = N / R;
    const int tws = N / (Ns * R);
    const int work = T * NB * R;          // one output per work item
    for (int w = threadIdx.x; w < work; w += blockDim.x) {
        const int b = w % NB;
        const int jm = w / NB;
        const int j = jm % T;
        const int m = jm / T;
        const int k = j % Ns;
        const long long step = (long long)(k + Ns * m) * tws;    // < N
        float2 acc = make_float2(0.f, 0.f);
        int idx = 0;
        const int st = (int)(step % N);
        for (int r = 0; r < R; ++r) {
            float2 x = src[(j + r * T) * BS + b];
            float2 wv = tw_load<DIR>(tw, idx);
            acc.x = fmaf(x.x, wv.x, fmaf(-x.y, wv.y, acc.x));
            acc.y = fmaf(x.x, wv.y, fmaf(x.y, wv.x, acc.y));
            idx += st; if (idx >= N) idx -= N;
        }
        const int j0 = (j - k) * R + k;
        dst[(j0 + m * Ns) * BS + b] = acc;
    }
}

// Transforms NB sequences (see layout above).  The caller must have synchronised after filling `src`.
// Returns the buffer that holds the result (either src or dst); a __syncthreads() has been executed
// after the last pass, so the result is visible to every thread of the block.
template <int DIR>
__device__ float2* fft_batched(float2* src, float2* dst, const FftPlan& plan, int NB, int BS,
                               const float2* __restrict__ tw) {
    const int N = plan.n;
    int Ns = 1;
    for (int p = 0; p < plan.npass; ++p) {
        const int R = plan.radix[p];
        switch (R) {
            case 2: stockham_pass<2, DIR>(src, dst, N, Ns, NB, BS, tw); break;
            case 3: stockham_pass<3, DIR>(src, dst, N, Ns, NB, BS, tw); break;
            case 4: stockham_pass<4, DIR>(src, dst, N, Ns, NB, BS, tw); break;
            case 5: stockham_pass<5, DIR>(src, dst, N, Ns, NB, BS, tw); break;
            case 8: stockham_pass<8, DIR>(src, dst, N, Ns, NB, BS, tw); break;
            case 9: stockham_pass<9, DIR>(src, dst, N, Ns, NB, BS, tw); break;
            case 15: stockham_pass<15, DIR>(src, dst, N, Ns, NB, BS, tw); break;
            case 16: stockham_pass<16, DIR>(src, dst, N, Ns, NB, BS, tw); break;
            default: stockham_pass_generic<DIR>(src, dst, N, R, Ns, NB, BS, tw); break;
        }
        __syncthreads();
        float2* t = src; src = dst; dst = t;
        Ns *= R;
    }
    return src;
}

}  // namespace admm
