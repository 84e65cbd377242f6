// cluster_pow2.cu -- cluster-resident ADMM-TV solver for small planes (H, W in {128, 256}) on sm_100a.
//
// The two-kernel iteration (rows_pow2.cu + cols_pow2.cu) exchanges the packed spectrum through HBM / L2 twice per
// iteration and needs two dependent launches; for one or a few small planes (BASELINE configs[0]: a single 256 x 256
// image, 50 iterations) that is pure latency: ~11 us per iteration against 0.4 us of memory time.  Here ONE launch runs
// the whole solve (deconv.py:103-115, all maxit iterations) for a plane inside a thread-block cluster of NC = 16 CTAs:
//
//   * every CTA owns RB = H/NC image rows for the row phase (C2R rows -> prox / dual / divergence -> R2C rows, the code
//     of rows_pow2.cu) and CB = (W/2)/NC packed columns for the column phase (FFT along H -> X = A + Bm V -> inverse FFT,
//     the code of cols_pow2.cu);
//   * the duals u_x, u_y of its rows (ping-pong), and A, Bm of its columns stay in its shared memory for all iterations;
//   * the two transposes per iteration go through distributed shared memory: each CTA stores its results straight into
//     the receiving CTA's input buffer (st.shared::cluster via cluster.map_shared_rank) and a cluster barrier
//     (barrier.cluster arrive.release / wait.acquire) separates the phases.  The row phase needs one halo row of x above
//     and below its band: the column owners send those rows to both neighbours; the halo row of u_y is read from the
//     neighbour's shared memory.
//
// HBM is touched for y (once), the tables (once) and x (once).  The arithmetic and its order are those of the two-kernel
// path (results agree to the last fp32 digits; tested against that path and the fp64 oracle).  Inference only (no saved
// state), iso = 0.
#include <cooperative_groups.h>

#include "cols_common.cuh"

namespace cg = cooperative_groups;

// timing experiment only (-DCLUSTER_SYNC_EXPERIMENT): per-phase cluster barriers become CTA barriers -- wrong results, but
// an upper bound of what replacing them by mbarrier / st.async dataflow synchronisation could gain
#ifdef CLUSTER_SYNC_EXPERIMENT
#define CLUSTER_PHASE_SYNC() __syncthreads()
#else
#define CLUSTER_PHASE_SYNC() cluster.sync()
#endif

namespace admm {

template <int H, int W, int NC> struct ClusterCfg {
    using RR = RowRadix<W>;
    using CR = ColRadix<H>;
    using RS = RowSmem<W>;
    static constexpr int kThreads = 512;
    static constexpr int PT = 8;                              // complex points per thread in the row FFTs
    static constexpr int Wc = W / 2;
    static constexpr int RB = H / NC;                         // image rows per CTA
    static constexpr int CB = Wc / NC;                        // packed columns per CTA
    static constexpr int NPV = RB / 2;                        // v pairs: rows (r0 + 2m, r0 + 2m + 1)
    static constexpr int NPX = RB / 2 + 1;                    // x pairs: rows (r0 - 1 + 2m, r0 + 2m), halo included
    static constexpr int TPS_R = W / PT;                      // threads per row pair (= W/8: one radix-8 edge butterfly each)
    static constexpr int REGION = row_region<W>();
    static constexpr int TPS_C = H / kCP;                     // threads per column (8 points each)
    static_assert(RB >= 2 && RB % 2 == 0 && CB >= 2 && CB % 2 == 0, "band / column block too small for this cluster size");
    static_assert(NPX * TPS_R <= kThreads && CB * TPS_C <= kThreads, "a phase does not fit the CTA");
    // column twiddle tables as in ColCfg
    static constexpr bool kShareC = (CR::F0 == CR::F2);
    static constexpr int CT_F1 = 0;
    static constexpr int CT_F2 = CT_F1 + tab_size(CR::F1, CR::F0);
    static constexpr int CT_F_END = CT_F2 + tab_size(CR::F2, CR::F0 * CR::F1);
    static constexpr int CT_I1 = kShareC ? CT_F1 : CT_F_END;
    static constexpr int CT_I2 = kShareC ? CT_F2 : CT_I1 + tab_size(CR::F1, CR::F2);
    static constexpr int CT_END = kShareC ? CT_F_END : CT_I2 + tab_size(CR::F0, CR::F2 * CR::F1);
    static constexpr int SIDE = (kThreads / 32) * NPX * 2;
    // shared memory, in float2 slots
    static constexpr int O_RIN = 0;                                   // [(RB + 2) rows][Wc]   incoming spectrum of x (row 0 = r0 - 1)
    static constexpr int O_REGX = O_RIN + (RB + 2) * Wc;              // [NPX][REGION]
    static constexpr int O_CIN = O_REGX + NPX * REGION;               // [H][CB]               incoming spectrum of v (my columns)
    static constexpr int O_CBUF = O_CIN + H * CB;                     // [H][CB]               FFT tile of the column passes
    static constexpr int O_A = O_CBUF + H * CB;                       // [H][CB]
    static constexpr int O_BM = O_A + H * CB;                         // [H][CB] floats = H * CB / 2 slots
    static constexpr int O_BQ = O_BM + H * CB / 2;                    // [H] floats
    static constexpr int O_ZCOL = O_BQ + H / 2;                       // [H]
    static constexpr int O_U = O_ZCOL + H;                            // [2 buffers][2 fields][RB][W] floats
    static constexpr int O_RTAB = O_U + 2 * RB * W;                   // row twiddle tables
    static constexpr int O_SIDE = O_RTAB + RS::TAB_END;
    static constexpr int O_CTAB = O_SIDE + SIDE;
    static constexpr int O_END = O_CTAB + CT_END;
    static constexpr size_t bytes = (size_t)O_END * sizeof(float2);
};

__device__ __forceinline__ float2 mk_c(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }       // a + i b
__device__ __forceinline__ float2 mkc_c(float2 a, float2 b) { return make_float2(a.x + b.y, b.x - a.y); }      // conj(a) + i conj(b)

// Stockham passes of fft_pow2.cuh for PT (instead of 16) points per thread: slot q <-> position t + q * (N / PT)
template <int N, int PT, class Map>
__device__ __forceinline__ void xpass_load(float2 (&d)[PT], int t, const float2* __restrict__ reg, const Map& map) {
    constexpr int TPS = N / PT;
    const int b = map.base(t);
#pragma unroll
    for (int q = 0; q < PT; ++q) d[q] = reg[b + Map::delta(q * TPS)];
}
template <int N, int PT, int R, int NS, int DIR>
__device__ __forceinline__ void xpass_compute(float2 (&d)[PT], int t, const float2* __restrict__ tab) {
    constexpr int TPS = N / PT, NB = PT / R;
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
template <int N, int PT, int R, int NS, class Map>
__device__ __forceinline__ void xpass_store(const float2 (&d)[PT], int t, float2* __restrict__ reg, const Map& map) {
    constexpr int TPS = N / PT, NB = PT / R;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * TPS;
        const int k = j & (NS - 1);
        const int b = map.base((j - k) * R + k);
#pragma unroll
        for (int r = 0; r < R; ++r) reg[b + Map::delta(r * NS)] = d[m + r * NB];
    }
}
// column passes of cols_common.cuh for ONE column per thread: tile word = position * NCOLS + column
template <int H, int R, int NS, int DIR>
__device__ __forceinline__ void c1pass_compute(float2 (&d)[kCP], int t, const float2* __restrict__ tab) {
    constexpr int TPS = H / kCP, NB = kCP / R;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * TPS;
        float2 a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = d[m + r * NB];
        if (NS > 1) {
            const int k = j & (NS - 1);
            float2 w1 = tab[k];
            if (DIR > 0) w1.y = -w1.y;
            float2 wp[R];
            wp[1] = w1;
#pragma unroll
            for (int r = 2; r < R; ++r) wp[r] = (r & 1) ? cmul(wp[r - 1], w1) : cmul(wp[r / 2], wp[r / 2]);
#pragma unroll
            for (int r = 1; r < R; ++r) a[r] = cmul(a[r], wp[r]);
        }
        dftR<R, DIR>(a);
#pragma unroll
        for (int r = 0; r < R; ++r) d[m + r * NB] = a[r];
    }
}
template <int H, int R, int NS, int NCOLS>
__device__ __forceinline__ void c1pass_store(const float2 (&d)[kCP], int t, int col, float2* __restrict__ buf) {
    constexpr int TPS = H / kCP, NB = kCP / R;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * TPS;
        const int k = j & (NS - 1);
        const int b = ((j - k) * R + k) * NCOLS + col;
#pragma unroll
        for (int r = 0; r < R; ++r) buf[b + r * NS * NCOLS] = d[m + r * NB];
    }
}
template <int H, int NCOLS>
__device__ __forceinline__ void c1pass_load(float2 (&d)[kCP], int t, int col, const float2* __restrict__ buf) {
    constexpr int TPS = H / kCP;
    const int b = t * NCOLS + col;
#pragma unroll
    for (int q = 0; q < kCP; ++q) d[q] = buf[b + q * TPS * NCOLS];
}

enum ClusterRowMode { CL_R2C = 0, CL_FULL = 1, CL_C2R = 2 };
enum ClusterColMode { CL_INIT = 0, CL_ITER = 1 };

// ------------------------------------------------------------------------------------------ row phase
// CL_R2C : y rows -> spectrum of y, pushed to the column owners                        (deconv.py:104, first rfftn)
// CL_FULL: rin (spectrum of x) -> x -> u, v -> spectrum of v, pushed to the column owners   (deconv.py:106-115, 104)
// CL_C2R : rin -> x -> out (+ bias, activation)                                         (deconv.py:117, admmdeconv.py:64)
template <int H, int W, int NC, int MODE>
__device__ __forceinline__ void cluster_row_phase(const ClusterArgs& a, float2* smem, cg::cluster_group& cluster, int rank,
                                                  int p, int it) {
    using C = ClusterCfg<H, W, NC>;
    using RR = RowRadix<W>;
    using S = RowSmem<W>;
    constexpr int TPS = C::TPS_R, REGION = C::REGION, Wc = C::Wc, RB = C::RB, NPX = C::NPX, NPV = C::NPV;
    constexpr int T8 = W / 8;
    float2* rin = smem + C::O_RIN;
    float2* regX = smem + C::O_REGX;
    float2* regV = (MODE == CL_FULL) ? regX + REGION : regX;
    float2* tabs = smem + C::O_RTAB;
    float2* side = smem + C::O_SIDE;
    float* ubase = reinterpret_cast<float*>(smem + C::O_U);
    const RowMapObj map;
    const int tid = threadIdx.x;
    const int pair = tid / TPS;
    const int t = tid % TPS;
    const int r0 = rank * RB;
    const bool t0 = (t == 0);
    constexpr int PT = C::PT;
    static_assert(TPS == T8, "one radix-8 edge butterfly per thread");
    float2 d[PT];
    float2* myX = regX + pair * REGION;
    const unsigned pmask = (TPS >= 32) ? 0xffffffffu : (((1u << (TPS & 31)) - 1u) << (((tid & 31) / TPS) * TPS));

    // ---------------------------------------------------------------- C2R: merge + inverse FFT of the x pairs
    if (MODE != CL_R2C && pair < NPX) {
        const float2* Sa = rin + (2 * pair) * Wc;          // row r0 - 1 + 2 pair
        const float2* Sb = Sa + Wc;                        // row r0 + 2 pair
        // first inverse pass: radix 8, no twiddles, butterfly j = t: inputs Z[n], n = t + r T8, of z = x_a + i x_b, merged
        // from the two packed half spectra: n < W/2: P_a[n] + i P_b[n];  n > W/2: conj(P_a[W-n]) + i conj(P_b[W-n]);
        // packed column 0 = (DC, Nyquist), both real
        float2 v[8];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = mk_c(Sa[t + r * T8], Sb[t + r * T8]);
#pragma unroll
        for (int r = 4; r < 8; ++r) {
            // W - n = (T8 - t) + (7 - r) T8 for t > 0;  (8 - r) T8 for t == 0
            const int c = t0 ? ((8 - r) & 3) * T8 : (T8 - t) + (7 - r) * T8;
            v[r] = mkc_c(Sa[c], Sb[c]);
        }
        if (t0) {
            const float2 a0 = Sa[0], b0 = Sb[0];
            v[0] = make_float2(a0.x, b0.x);                // Z[0]
            v[4] = make_float2(a0.y, b0.y);                // Z[W/2]
        }
        dft8<+1>(v);
        {
            const int b1 = map.base(8 * t);
#pragma unroll
            for (int r = 0; r < 8; ++r) myX[b1 + r] = v[r];
        }
        __syncwarp(pmask);
        xpass_load<W, PT>(d, t, myX, map);
        xpass_compute<W, PT, RR::IB, 8, +1>(d, t, tabs + S::TAB_IB);
        __syncwarp(pmask);
        xpass_store<W, PT, RR::IB, 8>(d, t, myX, map);
        __syncwarp(pmask);
        xpass_load<W, PT>(d, t, myX, map);
        xpass_compute<W, PT, RR::IC, 8 * RR::IB, +1>(d, t, tabs + S::TAB_IC);
        __syncwarp(pmask);
        xpass_store<W, PT, RR::IC, 8 * RR::IB>(d, t, myX, map);
    }

    if (MODE == CL_C2R) {
        // x pairs -> real rows (+ bias, activation): pair pp = rows (r0 - 1 + 2 pp, r0 + 2 pp); halo rows are not written
        __syncthreads();
        constexpr int CP = W / 2;
        float* __restrict__ out = a.out + out_plane_offset(a, p, H, W);
        const float bias = a.bias ? __ldg(a.bias) : 0.f;
        const int act = a.act;
        for (int e = tid; e < NPX * CP; e += C::kThreads) {
            const int pp = e / CP, c = 2 * (e - pp * CP);
            const int pc = map.at(c);
            const float2 X0 = regX[pp * REGION + pc], X1 = regX[pp * REGION + pc + 1];
            const int ra = r0 - 1 + 2 * pp;
            if (pp > 0)
                *reinterpret_cast<float2*>(out + (size_t)ra * W + c) = make_float2(act_apply(X0.x + bias, act), act_apply(X1.x + bias, act));
            if (pp < NPX - 1)
                *reinterpret_cast<float2*>(out + (size_t)(ra + 1) * W + c) = make_float2(act_apply(X0.y + bias, act), act_apply(X1.y + bias, act));
        }
        return;
    }

    if (MODE == CL_R2C) {
        // y rows r0 .. r0 + RB - 1 -> complex pairs (row a + i row b)
        constexpr int CP = W / 2;
        const size_t plane_real = (size_t)p * H * W;
        const float* __restrict__ in = a.y ? a.y + plane_real : nullptr;
        const unsigned char* __restrict__ in8 = a.y8 ? a.y8 + plane_real : nullptr;
        for (int e = tid; e < NPV * CP; e += C::kThreads) {
            const int pp = e / CP, c = 2 * (e - pp * CP);
            const int pc = map.at(c);
            const size_t o = (size_t)(r0 + 2 * pp) * W + c;
            float2 ra_, rb_;
            if (in8) {
                const uchar2 ua = __ldg(reinterpret_cast<const uchar2*>(in8 + o)), ub = __ldg(reinterpret_cast<const uchar2*>(in8 + o + W));
                ra_ = make_float2((float)ua.x / 255.0f, (float)ua.y / 255.0f);
                rb_ = make_float2((float)ub.x / 255.0f, (float)ub.y / 255.0f);
            } else {
                ra_ = __ldg(reinterpret_cast<const float2*>(in + o)); rb_ = __ldg(reinterpret_cast<const float2*>(in + o + W));
            }
            regX[pp * REGION + pc] = make_float2(ra_.x, rb_.x);
            regX[pp * REGION + pc + 1] = make_float2(ra_.y, rb_.y);
        }
    }

    // ---------------------------------------------------------------- prox / dual update / divergence (rows_pow2.cu march)
    if (MODE == CL_FULL) {
        constexpr int CP = W / 2;
        constexpr int NG = C::kThreads / CP;
        const int g = tid / CP;
        const int c = 2 * (tid % CP);
        const int m_lo = (g * NPV) / NG, m_hi = ((g + 1) * NPV) / NG;
        const float tau = __ldg(a.lmbd) / __ldg(a.rho);                 // deconv.py:44
        const bool have_q = (it > 1);                                   // u = 0 before the first iteration
        // duals: buffer (it - 1) & 1 holds u of the previous iteration, buffer it & 1 receives the new ones;
        // layout [buffer][field x / y][RB rows][W]
        const int prev = (it - 1) & 1, cur = it & 1;
        const float* __restrict__ uxi = ubase + (size_t)(prev * 2 + 0) * RB * W + c;
        const float* __restrict__ uyi = ubase + (size_t)(prev * 2 + 1) * RB * W + c;
        float* __restrict__ uxo = ubase + (size_t)(cur * 2 + 0) * RB * W + c;
        float* __restrict__ uyo = ubase + (size_t)(cur * 2 + 1) * RB * W + c;
        // u_y of the row below the band (row r0 + RB): first row of the next CTA's previous buffer, read through DSMEM
        const float* __restrict__ uy_next = cluster.map_shared_rank(ubase + (size_t)(prev * 2 + 1) * RB * W, (rank + 1) % NC) + c;
        const int c2 = (c + 2 == W) ? (2 - W) : 2;
        const int cl = (c == 0) ? W - 1 : c - 1;
        const int pl = map.at(cl), pc = map.at(c), pr2 = map.at((c + 2) & (W - 1));
        const int pc1 = pc + 1;
        struct QRegs { float2 qxa, qxb, qyb, qyc; float qxa2, qxb2; };
        auto ld2 = [](const float* q_) { return *reinterpret_cast<const float2*>(q_); };
        auto load_q = [&](int m, QRegs& q) {
            // local rows a = 2m, b = a + 1, row below = b + 1 (the next CTA's first row when it leaves the band)
            const float* xa = uxi + (size_t)(2 * m) * W;
            const float* ya = uyi + (size_t)(2 * m) * W;
            q.qxa = ld2(xa); q.qxa2 = xa[c2];
            q.qxb = ld2(xa + W); q.qxb2 = xa[W + c2];
            q.qyb = ld2(ya + W);
            q.qyc = (2 * m + 2 < RB) ? ld2(ya + 2 * W) : ld2(uy_next);
        };
        QRegs q0, q1;
        float2 qya = make_float2(0.f, 0.f);
        q0.qxa = q0.qxb = q0.qyb = q0.qyc = make_float2(0.f, 0.f); q0.qxa2 = q0.qxb2 = 0.f;
        q1 = q0;
        if (have_q && m_lo < m_hi) {
            qya = ld2(uyi + (size_t)(2 * m_lo) * W);
            load_q(m_lo, q0);
            if (m_lo + 1 < m_hi) load_q(m_lo + 1, q1);
        }
        __syncthreads();                                         // x of every pair is in shared memory

        const int lane = tid & 31, wid = tid >> 5;
        float2* sideW = side + wid * (NPX * 2);
        {
            const int cfirst = 2 * ((tid & ~31) % CP);
            const int cL = (cfirst == 0) ? W - 1 : cfirst - 1;
            const int cR = (cfirst + 64) & (W - 1);
            const int nslots = m_hi - m_lo + 1;
            for (int e = lane; e < 2 * nslots; e += 32) {
                const int ps = e >> 1;
                sideW[e] = regX[(m_lo + ps) * REGION + map.at((e & 1) ? cR : cL)];
            }
        }
        __syncwarp();
        const float2* baseL = (lane == 0) ? sideW : regX + m_lo * REGION + pl;
        const float2* baseR = (lane == 31) ? sideW + 1 : regX + m_lo * REGION + pr2;
        const int strideL = (lane == 0) ? 2 : REGION;
        const int strideR = (lane == 31) ? 2 : REGION;
        const int steps = (NPV + NG - 1) / NG;
        auto march = [&](auto tsign) {
            constexpr bool NEG = decltype(tsign)::neg;
            auto sof = [tau](float q_) { return dual_of<NEG>(q_, tau); };            // the state arrays hold u = u(q)
            auto wfun2 = [](float q_, float tau_) { return fmaf(-2.0f, dual_of<NEG>(q_, tau_), q_); };
            const float2* X = regX + m_lo * REGION;
            float2 Pl = baseL[0], P0 = X[pc], P1 = X[pc1], P2 = baseR[0];
            __syncthreads();                                       // every group holds its first pair; side buffers complete
            float qy0 = P0.y - P0.x + qya.x;
            float qy1 = P1.y - P1.x + qya.y;
            float wy0 = wfun2(qy0, tau), wy1 = wfun2(qy1, tau);
            for (int s_ = 0; s_ < steps; ++s_) {
                const int m = m_lo + s_;
                if (m >= m_hi) break;
                const QRegs q = q0;
                q0 = q1;
                if (have_q && m + 2 < m_hi) load_q(m + 2, q1);
                X += REGION;
                const float2 Nl = baseL[(s_ + 1) * strideL], N0 = X[pc], N1 = X[pc1], N2 = baseR[(s_ + 1) * strideR];
                __syncwarp();
                const float qxa0 = P0.y - Pl.y + q.qxa.x;
                const float qxa1 = P1.y - P0.y + q.qxa.y;
                const float qxa2 = P2.y - P1.y + q.qxa2;
                const float wxa0 = wfun2(qxa0, tau), wxa1 = wfun2(qxa1, tau), wxa2 = wfun2(qxa2, tau);
                const float qyb0 = N0.x - P0.y + q.qyb.x;
                const float qyb1 = N1.x - P1.y + q.qyb.y;
                const float wyb0 = wfun2(qyb0, tau), wyb1 = wfun2(qyb1, tau);
                const float qxb0 = N0.x - Nl.x + q.qxb.x;
                const float qxb1 = N1.x - N0.x + q.qxb.y;
                const float qxb2 = N2.x - N1.x + q.qxb2;
                const float wxb0 = wfun2(qxb0, tau), wxb1 = wfun2(qxb1, tau), wxb2 = wfun2(qxb2, tau);
                const float qyc0 = N0.y - N0.x + q.qyc.x;
                const float qyc1 = N1.y - N1.x + q.qyc.y;
                const float wyc0 = wfun2(qyc0, tau), wyc1 = wfun2(qyc1, tau);
                const float va0 = wxa0 - wxa1 + wy0 - wyb0;
                const float va1 = wxa1 - wxa2 + wy1 - wyb1;
                const float vb0 = wxb0 - wxb1 + wyb0 - wyc0;
                const float vb1 = wxb1 - wxb2 + wyb1 - wyc1;
                const size_t oa = (size_t)(2 * m) * W;
                *reinterpret_cast<float2*>(uxo + oa) = make_float2(sof(qxa0), sof(qxa1));
                *reinterpret_cast<float2*>(uyo + oa) = make_float2(sof(qy0), sof(qy1));
                *reinterpret_cast<float2*>(uxo + oa + W) = make_float2(sof(qxb0), sof(qxb1));
                *reinterpret_cast<float2*>(uyo + oa + W) = make_float2(sof(qyb0), sof(qyb1));
                float2* V = regV + m * REGION;
                V[pc] = make_float2(va0, vb0);
                V[pc1] = make_float2(va1, vb1);
                Pl = Nl; P0 = N0; P1 = N1; P2 = N2;
                qy0 = qyc0; qy1 = qyc1; wy0 = wyc0; wy1 = wyc1;
            }
        };
        if (tau < 0.f) march(TauNeg{}); else march(TauPos{});
    }
    __syncthreads();

    // ---------------------------------------------------------------- R2C: forward FFT + split, pushed to the column owners
    if (pair < NPV) {
        float2* myV = regV + pair * REGION;
        xpass_load<W, PT>(d, t, myV, map);
        xpass_compute<W, PT, RR::FA, 1, -1>(d, t, nullptr);
        __syncwarp(pmask);
        xpass_store<W, PT, RR::FA, 1>(d, t, myV, map);
        __syncwarp(pmask);
        xpass_load<W, PT>(d, t, myV, map);
        xpass_compute<W, PT, RR::FB, RR::FA, -1>(d, t, tabs + S::TAB_FB);
        __syncwarp(pmask);
        xpass_store<W, PT, RR::FB, RR::FA>(d, t, myV, map);
        __syncwarp(pmask);
        // last pass: radix 8, Ns = T8, butterfly j = t: inputs t + r T8, twiddle k = t;  v[r] = Z[t + r T8]
        float2 v[8];
        {
            const int b1 = map.base(t);
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = myV[b1 + RowMapObj::delta(r * T8)];
        }
        const float2* tabC = tabs + S::TAB_FC;
#pragma unroll
        for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], tabC[(r - 1) * T8 + t]);
        dft8<-1>(v);
        // split: X_a[c] = (Z[c] + conj(Z[W-c])) / 2, X_b[c] = (Z[c] - conj(Z[W-c])) / 2i for c = t + r T8 < W/2; the partner
        // Z[W-c] = Z[(T8 - t) + (7 - r) T8] sits in slot 7 - r of lane T8 - t of this pair (t == 0: own slot 8 - r)
        const int ra = r0 + 2 * pair;
        float2* cin_local = smem + C::O_CIN;
        auto push = [&](int col, float2 Xa, float2 Xb) {
            float2* dst = cluster.map_shared_rank(cin_local, col / C::CB) + (size_t)ra * C::CB + (col % C::CB);
            dst[0] = Xa;
            dst[C::CB] = Xb;
        };
        const int plane_lane = (tid & 31) - t + ((T8 - t) & (T8 - 1));      // lane of butterfly T8 - t (t == 0: itself)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float2 Z1 = v[r];
            float2 M1;
            M1.x = __shfl_sync(pmask, v[7 - r].x, plane_lane);
            M1.y = __shfl_sync(pmask, v[7 - r].y, plane_lane);
            if (t0) M1 = v[(8 - r) & 7];
            float2 Xa = make_float2(0.5f * (Z1.x + M1.x), 0.5f * (Z1.y - M1.y));
            float2 Xb = make_float2(0.5f * (Z1.y + M1.y), 0.5f * (M1.x - Z1.x));
            if (r == 0 && t0) {
                Xa = make_float2(v[0].x, v[4].x);
                Xb = make_float2(v[0].y, v[4].y);
            }
            push(t + r * T8, Xa, Xb);
        }
    }
}

// ------------------------------------------------------------------------------------------ column phase
// CL_INIT: cin (row spectrum of y) -> FFT along H -> A = Mul Z (kept in shared memory) -> inverse FFT -> row owners
// CL_ITER: cin (row spectrum of v) -> FFT along H -> X = A + Bm Z -> inverse FFT -> row owners      (deconv.py:104-106)
template <int H, int W, int NC, int MODE>
__device__ __forceinline__ void cluster_col_phase(const ClusterArgs& a, float2* smem, cg::cluster_group& cluster, int rank) {
    using C = ClusterCfg<H, W, NC>;
    using CR = ColRadix<H>;
    constexpr int TPS = C::TPS_C, Wc = C::Wc, RB = C::RB, CB = C::CB;
    constexpr int NB2 = kCP / CR::F2;
    const float2* cin = smem + C::O_CIN;                                         // word = row * CB + column
    float2* buf = smem + C::O_CBUF;
    float2* As = smem + C::O_A;
    const float* Bms = reinterpret_cast<const float*>(smem + C::O_BM);
    const float* Bqs = reinterpret_cast<const float*>(smem + C::O_BQ);
    float2* zcol = smem + C::O_ZCOL;
    const float2* tabs = smem + C::O_CTAB;
    const int tid = threadIdx.x;
    const bool active = tid < CB * TPS;
    const int col = tid % CB;
    const int t = tid / CB;
    const int c = rank * CB + col;                   // packed column of this thread
    const bool col0 = (c == 0);                      // packed column 0 carries DC and Nyquist
    float2 d[kCP];
    if (active) {
#pragma unroll
        for (int q = 0; q < kCP; ++q) d[q] = cin[(t + q * TPS) * CB + col];
        c1pass_compute<H, CR::F0, 1, -1>(d, t, nullptr);
        c1pass_store<H, CR::F0, 1, CB>(d, t, col, buf);
    }
    __syncthreads();
    if (active) {
        c1pass_load<H, CB>(d, t, col, buf);
        c1pass_compute<H, CR::F1, CR::F0, -1>(d, t, tabs + C::CT_F1);
    }
    __syncthreads();
    if (active) c1pass_store<H, CR::F1, CR::F0, CB>(d, t, col, buf);
    __syncthreads();
    if (active) {
        c1pass_load<H, CB>(d, t, col, buf);
        c1pass_compute<H, CR::F2, CR::F0 * CR::F1, -1>(d, t, tabs + C::CT_F2);
        if (col0) {
#pragma unroll
            for (int m = 0; m < NB2; ++m)
#pragma unroll
                for (int r = 0; r < CR::F2; ++r) zcol[(t + m * TPS) + r * (H / CR::F2)] = d[m + r * NB2];
        }
    }
    __syncthreads();                                   // all reads of buf done, zcol visible
    if (active) {
#pragma unroll
        for (int m = 0; m < NB2; ++m) {
#pragma unroll
            for (int r = 0; r < CR::F2; ++r) {
                const int u = (t + m * TPS) + r * (H / CR::F2);
                const float2 Z = d[m + r * NB2];
                float2 o;
                if (MODE == CL_ITER) {
                    const float2 Av = As[u * CB + col];
                    const float bm = Bms[u * CB + col];
                    o = make_float2(fmaf(bm, Z.x, Av.x), fmaf(bm, Z.y, Av.y));
                    if (col0) {
                        const float2 Zm = zcol[(H - u) & (H - 1)];
                        const float bq = Bqs[u];
                        o.x = fmaf(bq, Zm.x, o.x);
                        o.y = fmaf(-bq, Zm.y, o.y);
                    }
                } else {
                    o = cmul(__ldg(a.Mul + c + (size_t)u * Wc), Z);
                    if (col0) {
                        const float2 Zm = zcol[(H - u) & (H - 1)];
                        const float2 e = cmul(__ldg(a.Mq + u), cconj(Zm));
                        o.x += e.x; o.y += e.y;
                    }
                    As[u * CB + col] = o;
                }
                d[m + r * NB2] = o;
            }
        }
        c1pass_compute<H, CR::F2, 1, +1>(d, t, nullptr);
        c1pass_store<H, CR::F2, 1, CB>(d, t, col, buf);
    }
    __syncthreads();
    if (active) {
        c1pass_load<H, CB>(d, t, col, buf);
        c1pass_compute<H, CR::F1, CR::F2, +1>(d, t, tabs + C::CT_I1);
    }
    __syncthreads();
    if (active) c1pass_store<H, CR::F1, CR::F2, CB>(d, t, col, buf);
    __syncthreads();
    if (active) {
        c1pass_load<H, CB>(d, t, col, buf);
        c1pass_compute<H, CR::F0, CR::F2 * CR::F1, +1>(d, t, tabs + C::CT_I2);
        // natural order: slot (m, r) -> row u = (t + m TPS) + r (H / F0); row u belongs to CTA u / RB (local row u % RB + 1
        // of its rin); the first / last row of a band is also the halo row of the neighbour above / below
        constexpr int NB = kCP / CR::F0;
        float2* rin_local = smem + C::O_RIN;
#pragma unroll
        for (int m = 0; m < NB; ++m)
#pragma unroll
            for (int r = 0; r < CR::F0; ++r) {
                const int u = (t + m * TPS) + r * (H / CR::F0);
                const int k = u / RB, l = u - k * RB;
                const float2 val = d[m + r * NB];
                cluster.map_shared_rank(rin_local, k)[(size_t)(l + 1) * Wc + c] = val;
                if (l == 0) cluster.map_shared_rank(rin_local, (k + NC - 1) % NC)[(size_t)(RB + 1) * Wc + c] = val;
                if (l == RB - 1) cluster.map_shared_rank(rin_local, (k + 1) % NC)[c] = val;
            }
    }
}

template <int H, int W, int NC>
__global__ void __launch_bounds__(512, 1)
k_cluster_solve(ClusterArgs a) {
    using C = ClusterCfg<H, W, NC>;
    using RR = RowRadix<W>;
    using S = RowSmem<W>;
    using CR = ColRadix<H>;
    extern __shared__ float2 smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / NC;
    const int ncl = gridDim.x / NC;
    const int tid = threadIdx.x;

    // per-CTA constants: twiddle tables, Bm / Bq of my columns
    float2* rtabs = smem + C::O_RTAB;
    build_tab<W, RR::IB, 8>(rtabs + S::TAB_IB, a.twW);
    build_tab<W, RR::IC, 8 * RR::IB>(rtabs + S::TAB_IC, a.twW);
    if (!S::kShareB) build_tab<W, RR::FB, RR::FA>(rtabs + S::TAB_FB, a.twW);
    if (!S::kShareC) build_tab<W, 8, W / 8>(rtabs + S::TAB_FC, a.twW);
    float2* ctabs = smem + C::O_CTAB;
    build_tab<H, CR::F1, CR::F0>(ctabs + C::CT_F1, a.twH);
    build_tab<H, CR::F2, CR::F0 * CR::F1>(ctabs + C::CT_F2, a.twH);
    if (!C::kShareC) {
        build_tab<H, CR::F1, CR::F2>(ctabs + C::CT_I1, a.twH);
        build_tab<H, CR::F0, CR::F2 * CR::F1>(ctabs + C::CT_I2, a.twH);
    }
    {
        float* Bms = reinterpret_cast<float*>(smem + C::O_BM);
        float* Bqs = reinterpret_cast<float*>(smem + C::O_BQ);
        const int c0 = rank * C::CB;
        for (int e = tid; e < H * C::CB; e += C::kThreads) {
            const int u = e / C::CB, cc = e - u * C::CB;
            Bms[e] = __ldg(a.Bm + (size_t)u * C::Wc + c0 + cc);
        }
        for (int u = tid; u < H; u += C::kThreads) Bqs[u] = __ldg(a.Bq + u);
    }
    // every CTA of the cluster is running (and its shared memory is allocated) before anyone stores into it
    cluster.sync();

    for (int p = cid; p < a.P; p += ncl) {
        // x_1 = F^-1[A],  A = Mul F(y)                                       (deconv.py:104-106 with z = u = 0)
        cluster_row_phase<H, W, NC, CL_R2C>(a, smem, cluster, rank, p, 0);
        CLUSTER_PHASE_SYNC();
        cluster_col_phase<H, W, NC, CL_INIT>(a, smem, cluster, rank);
        CLUSTER_PHASE_SYNC();
        for (int it = 1; it < a.maxit; ++it) {
            cluster_row_phase<H, W, NC, CL_FULL>(a, smem, cluster, rank, p, it);
            CLUSTER_PHASE_SYNC();
            cluster_col_phase<H, W, NC, CL_ITER>(a, smem, cluster, rank);
            CLUSTER_PHASE_SYNC();
        }
        cluster_row_phase<H, W, NC, CL_C2R>(a, smem, cluster, rank, p, 0);
        // the next plane's first remote stores go to cin, which nobody reads before the next cluster barrier; its
        // column phase (which overwrites rin) starts behind that barrier, i.e. after every CTA has finished this C2R
        __syncthreads();
    }
    // no CTA may exit while a sibling can still address its shared memory
    cluster.sync();
}

template <int H, int W, int NC>
static int launch_cluster_t(const Geometry& g, const ClusterArgs& a, cudaStream_t st) {
    using C = ClusterCfg<H, W, NC>;
    static std::atomic<int> max_clusters_dev[64];
    int dev = 0;
    ADMM_CUDA_CHECK(cudaGetDevice(&dev));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(C::kThreads); cfg.dynamicSmemBytes = C::bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int maxc = max_clusters_dev[dev & 63].load();
    if (maxc == 0) {
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_cluster_solve<H, W, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::bytes));
        if (NC > 8)
            ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_cluster_solve<H, W, NC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cfg.gridDim = dim3(NC);
        int n = 0;
        ADMM_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&n, k_cluster_solve<H, W, NC>, &cfg));
        if (n < 1) { max_clusters_dev[dev & 63].store(-1); return fail(4, "cluster solver: no cluster of this size fits on the device"); }
        max_clusters_dev[dev & 63].store(n);
        maxc = n;
    }
    if (maxc < 0) return fail(4, "cluster solver: no cluster of this size fits on the device");
    const int ncl = std::min(maxc, g.P);
    cfg.gridDim = dim3((unsigned)(ncl * NC));
    ProfScope ps(PROF_OTHER, st);
    ADMM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_cluster_solve<H, W, NC>, a));
    return 0;
}

static constexpr int kClusterSize = 16;

// one plane needs H/16 >= 2 even rows and (W/2)/16 >= 2 even columns per CTA and fits 227 KB: H, W in {128, 256}
bool cluster_solver_supported(const Geometry& g, int iso, bool training) {
    if (!options().use_cluster || options().force_generic || iso || training) return false;
    if (!((g.H == 128 || g.H == 256) && (g.W == 128 || g.W == 256))) return false;
    return true;
}

// heuristic: the two-kernel path wins once its launches fill the machine.  Measured (B200, 50 iterations, ms per solve,
// two-kernel / cluster): 1 plane 256^2 0.548 / 0.353, 3 planes 0.652 / 0.353, 9 planes (100 it) 1.754 / 1.308,
// 24 planes 1.073 / 1.310, 48 planes 1.480 / 2.266; 1 plane 128^2 0.487 / 0.232.
bool cluster_solver_preferred(const Geometry& g) {
    const int want = options().use_cluster;          // 2 = always (tests)
    if (want >= 2) return true;
    return g.P <= 12;
}

int launch_cluster_solve(const Geometry& g, const ClusterArgs& a, cudaStream_t st) {
    if (g.H == 256 && g.W == 256) return launch_cluster_t<256, 256, kClusterSize>(g, a, st);
    if (g.H == 128 && g.W == 128) return launch_cluster_t<128, 128, kClusterSize>(g, a, st);
    if (g.H == 128 && g.W == 256) return launch_cluster_t<128, 256, kClusterSize>(g, a, st);
    if (g.H == 256 && g.W == 128) return launch_cluster_t<256, 128, kClusterSize>(g, a, st);
    return fail(4, "cluster solver: unsupported size");
}

}  // namespace admm
