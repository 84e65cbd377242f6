// cols_big.cu -- column pass of the ADMM iteration for the LARGE heights on sm_100a: H = 2160 = 15*12*12 (the 2160x3840
// single-frame configuration, BASELINE configs[2]), 1080 = 15*9*8, 1440 = 15*12*8, 720 = 15*8*6, 1024 = 16*8*8,
// 2048 = 16*16*8, 768 = 16*8*6, 1536 = 16*12*8.
//
//   COLS_ITER:  S0 = F_col^-1[ A + Bm * F_col(S1) ]          (deconv.py:104-106 with freq_c and rho folded into A, Bm)
//   COLS_INIT:  A  = Mul * F_col(S1),  S0 = F_col^-1[A]      (deconv.py:57,99,104: freq_c * rfftn(H_t(xin)))
//
// Persistent CTAs (one or two per SM) loop over work items of C = 4 packed columns of one plane: 4 x N/R2 threads, one
// radix-R butterfly of one column per thread per pass.  The first forward pass loads straight from global memory, the
// last forward pass (radix R2) leaves exactly the inputs of the first inverse pass (radix R2) in the same thread's
// registers, so the spectral update happens in registers between the two, and the last inverse pass stores straight to
// global memory.  The two middle exchanges ping-pong between two tile buffers in shared memory (4 block barriers per
// item); twiddles come from shared-memory tables built once per CTA.  Passes whose first radix is even use a padded
// map (one spare slot per R entries).  A and a copy of Bm are kept item-major ([item][u][4]); the spectra are row-major
// or, between the two large-frame kernels, tile-major (common.cuh, kSpecTile).
#include <cuda.h>
#include <cstring>
#include <cstdio>
#include "common.cuh"
#include "fft_big.cuh"
#include "cols_common.cuh"

namespace admm {

// ---- TMA plumbing of the one-CTA-per-SM iteration kernels (input tile fetched by cp.async.bulk.tensor one item ahead)
__device__ __forceinline__ unsigned cb_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cb_mbar_init(unsigned long long* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cb_smem(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void cb_mbar_expect(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cb_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cb_mbar_wait(unsigned long long* bar, unsigned parity) {      // bounded: traps instead of hanging
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(cb_smem(bar)), "r"(parity) : "memory");
        if (!done && spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void cb_tma_load(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(cb_smem(dst)), "l"(map), "r"(x), "r"(y), "r"(cb_smem(bar)) : "memory");
}

template <int H> struct ColBig;
// forward radices (R0, R1, R2), the inverse runs (R2, R1, R0); R0 odd; N/R2 is the largest thread count (the middle pass
// pair F3 / I1 lives in the registers of the same thread) and R2 | N/R1, R2 | N/R0 (padded inverse map)
// OCC: CTAs per SM (480..540 threads fit twice at 60 registers: two barrier domains per SM, +13 % measured; 720 and
// 1024 threads fit once).  FPAD: padded map of the FORWARD passes (0 = none: R0 odd; R0 when R0 is even, e.g. the power-of-two height 1024)
// C: packed columns per work item.  4 gives 64-byte runs in the tile-major spectrum.  Measured for the 2160-high tile
// (2 x 75 KB of ping-pong buffers at C = 4: one CTA per SM): C = 2 (360 threads, 2 x 37 KB, TWO CTAs = two barrier
// domains per SM, but 32-byte runs) is slower: column pass 0.330 vs 0.396 of the HBM roofline, cfg3 27.7 vs 30.1 G
// pixel-it/s (-DCOLS_BIG_C2160=2 rebuilds that variant).
// COLS_BIG_EXP (timing only, wrong results): bit 0 = no output stores, bit 1 = no input loads
#ifndef COLS_BIG_EXP
#define COLS_BIG_EXP 0
#endif
// 1 = the 2160- and 1536-high iteration kernels (one CTA per SM) get their (tile-major) input through TMA, one item ahead.
// Measured, column pass us per launch with / without: 2160: 108.2 / 115.1, 1536: 58.5 / 63.3, 1440: 90.4 / 89.0, 2048 (1024 threads,
// 64 registers): 120.1 / 106.1 -- so only the first two
#ifndef COLS_BIG_TMA
#define COLS_BIG_TMA 1
#endif
// 1 = the L2 prefetch of the next item's slice of A is issued half an item ahead of its use instead of a whole item
// (2160 high: 105.9 vs 107.7 us per launch)
#ifndef COLS_BIG_APF_LATE
#define COLS_BIG_APF_LATE 1
#endif
#ifndef COLS_BIG_C2160
#define COLS_BIG_C2160 4
#endif
template <> struct ColBig<2160> { static constexpr int R0 = 15, R1 = 12, R2 = 12, FPAD = 0, C = COLS_BIG_C2160, OCC = (COLS_BIG_C2160 == 2 ? 2 : 1); };
template <> struct ColBig<1080> { static constexpr int R0 = 15, R1 = 9,  R2 = 8,  FPAD = 0, C = 4, OCC = 2; };
template <> struct ColBig<1024> { static constexpr int R0 = 16, R1 = 8,  R2 = 8,  FPAD = 16, C = 4, OCC = 2; };
template <> struct ColBig<2048> { static constexpr int R0 = 16, R1 = 16, R2 = 8,  FPAD = 16, C = 4, OCC = 1; };
template <> struct ColBig<768>  { static constexpr int R0 = 16, R1 = 8,  R2 = 6,  FPAD = 16, C = 4, OCC = 2; };
template <> struct ColBig<1536> { static constexpr int R0 = 16, R1 = 12, R2 = 8,  FPAD = 16, C = 4, OCC = 1; };
template <> struct ColBig<1440> { static constexpr int R0 = 15, R1 = 12, R2 = 8,  FPAD = 0, C = 4, OCC = 1; };
template <> struct ColBig<720>  { static constexpr int R0 = 15, R1 = 8,  R2 = 6,  FPAD = 0, C = 4, OCC = 2; };

template <int H> struct ColBigCfg {
    using CB = ColBig<H>;
    static constexpr int C = CB::C;
    static_assert(kSpecTile % C == 0, "a work item is a whole fraction of a spectrum tile");
    static constexpr int T0 = H / CB::R0, T1 = H / CB::R1, T2 = H / CB::R2;
    static constexpr int TN = (T0 > T1 ? T0 : T1) > T2 ? (T0 > T1 ? T0 : T1) : T2;
    static constexpr int NT = C * TN;
    static_assert(TN == T2, "the register-resident middle pass needs the largest thread count");
    static constexpr int BUF = (H + H / (CB::FPAD && CB::FPAD < CB::R2 ? CB::FPAD : CB::R2)) * C;   // padded tile (larger of the two maps)
    // twiddle tables (forward sign): forward pass 2 (NS = R0) and inverse pass 2 (NS = R2) compact [(r-1) * NS + k];
    // forward pass 3 (NS = R0 R1 = T2) and inverse pass 3 (NS = R2 R1 = T0) per butterfly [(r-1) * T + j]
    static constexpr int TAB_F2 = 0;
    static constexpr int TAB_I2 = TAB_F2 + (CB::R1 - 1) * CB::R0;
    static constexpr int TAB_F3 = TAB_I2 + (CB::R1 - 1) * CB::R2;
    static constexpr int TAB_I3 = TAB_F3 + (CB::R2 - 1) * T2;
    static constexpr int TAB_END = TAB_I3 + (CB::R0 - 1) * T0;
    static constexpr size_t smem = (size_t)(2 * BUF + TAB_END) * sizeof(float2) + 16; // two tile buffers (ping-pong) + one mbarrier
    // TMA input (tile-major spectrum: one row pair of an item = C columns x 2 rows = 64 bytes of a 128-byte line)
    static constexpr int kPairs = H / 2;
    static constexpr int kBoxRows = (kPairs % 216 == 0) ? 216 : ((kPairs % 256 == 0) ? 256 : ((kPairs % 180 == 0) ? 180 : 0));
    static constexpr bool kTmaIn = (H == 2160 || H == 1536) && (C == 4) && kBoxRows > 0 && COLS_BIG_TMA;   // measured per height, see below
    static constexpr unsigned kItemBytes = (unsigned)(H * C * sizeof(float2));
};

// IN_T / OUT_T: spec_in / spec_out in the tile-major layout shared with the large row kernel (common.cuh, kSpecTile)
template <int H, int MODE, bool IN_T, bool OUT_T>
__global__ void __launch_bounds__(ColBigCfg<H>::NT, ColBig<H>::OCC)
k_cols_big(const __grid_constant__ CUtensorMap tmap_in, ColArgs a, int Wc, int ntiles, int nitems) {
    using CB = ColBig<H>;
    using CF = ColBigCfg<H>;
    constexpr int R0 = CB::R0, R1 = CB::R1, R2 = CB::R2, C = CF::C;
    using F1 = BigPass<H, R0, 1, -1, C, CB::FPAD>;
    using F2 = BigPass<H, R1, R0, -1, C, CB::FPAD>;
    using F3 = BigPass<H, R2, R0 * R1, -1, C, CB::FPAD>;
    using I1 = BigPass<H, R2, 1, +1, C, R2>;
    using I2 = BigPass<H, R1, R2, +1, C, R2>;
    using I3 = BigPass<H, R0, R2 * R1, +1, C, R2>;
    constexpr int RMAX = (R1 > R2 ? R1 : R2) > R0 ? (R1 > R2 ? R1 : R2) : R0;
    extern __shared__ __align__(128) float2 smem[];
    // TMA_IN: the input tile of an item lands in Y (dense, [row pair][C columns][row in pair]) while the previous item runs
    // its last pass: Y is free once inverse pass 3 has loaded it, and the first forward pass reads it before pass 2 writes it
    constexpr bool TMA_IN = CF::kTmaIn && IN_T && MODE == COLS_ITER;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem + 2 * CF::BUF + CF::TAB_END);
    float2* Yraw = smem + CF::BUF;
    auto fetch_tile = [&](int it) {                      // one thread: arm the barrier, one box per kBoxRows row pairs
        const int tl = it % ntiles, pp = it / ntiles;
        const int y0 = (pp * (Wc / kSpecTile) + tl / (kSpecTile / C)) * CF::kPairs;
        const int x0 = (tl % (kSpecTile / C)) * C * 4;   // floats: C columns x 2 rows x (re, im)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cb_mbar_expect(mbar, CF::kItemBytes);
#pragma unroll
        for (int b = 0; b < CF::kPairs / (CF::kBoxRows ? CF::kBoxRows : 1); ++b)
            cb_tma_load(Yraw + b * CF::kBoxRows * C * 2, &tmap_in, x0, y0 + b * CF::kBoxRows, mbar);
    };
    unsigned in_phase = 0;

    const int c = threadIdx.x % C;
    const int j = threadIdx.x / C;
    const float2* __restrict__ tw = a.tw;
    float2* X = smem + c;                        // ping
    float2* Y = smem + CF::BUF + c;              // pong
    float2* tabs = smem + 2 * CF::BUF;
    // tables once per (persistent) CTA
    for (int i = threadIdx.x; i < CF::TAB_END; i += CF::NT) {
        int idx;
        if (i < CF::TAB_I2) { const int e = i - CF::TAB_F2; const int r = e / R0 + 1, k = e % R0; idx = k * r * (H / (R0 * R1)); }
        else if (i < CF::TAB_F3) { const int e = i - CF::TAB_I2; const int r = e / R2 + 1, k = e % R2; idx = k * r * (H / (R2 * R1)); }
        else if (i < CF::TAB_I3) { const int e = i - CF::TAB_F3; const int r = e / CF::T2 + 1, k = e % CF::T2; idx = k * r; }
        else { const int e = i - CF::TAB_I3; const int r = e / CF::T0 + 1, k = e % CF::T0; idx = k * r; }
        tabs[i] = __ldg(tw + idx);
    }
    if (TMA_IN) {
        if (threadIdx.x == 0) cb_mbar_init(mbar);
        __syncthreads();
        if (threadIdx.x == 0 && (int)blockIdx.x < nitems) fetch_tile(blockIdx.x);
    }
    const float2* tF2 = tabs + CF::TAB_F2 + j % R0;
    const float2* tI2 = tabs + CF::TAB_I2 + j % R2;
    const float2* tF3 = tabs + CF::TAB_F3 + j;
    const float2* tI3 = tabs + CF::TAB_I3 + j;
    constexpr int T2 = F3::T;
    const bool act = j < T2;

    float2 v[RMAX];
    // -DCOLS_STATS: per-phase clock cycles of a few CTAs (development instrumentation, see tools/README.md)
#ifdef COLS_STATS
    long long st[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tp = clock64(); int nst = 0;
#define CB_ST(i) { const long long tn_ = clock64(); st[i] += tn_ - tp; tp = tn_; }
#else
#define CB_ST(i)
#endif
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int tile = item % ntiles;
        const int p = item / ntiles;
        const int c0 = tile * C;
        const size_t plane = (size_t)p * H * Wc;
        const float2* __restrict__ in = a.spec_in + plane + c0 + c;
        float2* __restrict__ out = a.spec_out + plane + c0 + c;
        // tile-major spectra: an item is one half (C of kSpecTile columns) of a contiguous block
        const size_t tb = ((size_t)p * (Wc / kSpecTile) + tile / (kSpecTile / C)) * H * kSpecTile;
        const int cc = (tile % (kSpecTile / C)) * C + c;                    // column inside the spectrum tile
        const float2* __restrict__ tin = a.spec_in + tb;
        float2* __restrict__ tout = a.spec_out + tb;
        // A and Bm live tile-major ([item][u][C] / [tile][u][C]): a warp reads 256 / 128 contiguous bytes, not 8 x 32
        float2* __restrict__ At = a.A + (size_t)item * H * C + c;
        const float* __restrict__ Bt = a.Bmt + (size_t)tile * H * C + c;
        if (item + (int)gridDim.x < nitems) {
            // pull the next item's tile and its slice of A into L2 while this item computes
            const int ni = item + gridDim.x;
            const size_t nb = (size_t)(ni / ntiles) * H * Wc + (size_t)(ni % ntiles) * C;
            if (TMA_IN) {
                // the tile itself is fetched by TMA behind the previous item's last pass
            } else if (IN_T) {
                const char* sn = (const char*)(a.spec_in + ((size_t)(ni / ntiles) * (Wc / kSpecTile) + (ni % ntiles) / (kSpecTile / C)) * H * kSpecTile);
                for (int o = threadIdx.x * 128; o < H * kSpecTile * (int)sizeof(float2); o += CF::NT * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(sn + o));
            } else {
                for (int u = threadIdx.x; u < H; u += CF::NT)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.spec_in + nb + (size_t)u * Wc));
            }
            if (MODE == COLS_ITER && !COLS_BIG_APF_LATE) {
                const char* an = (const char*)(a.A + (size_t)ni * H * C);
                for (int o = threadIdx.x * 128; o < H * C * (int)sizeof(float2); o += CF::NT * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(an + o));
            }
        }
        // ---- forward pass 1: global -> registers -> X.  (X was last read by the inverse pass 2 of the previous item,
        //      and every thread has passed the barrier that follows those loads.)
        if (TMA_IN) { cb_mbar_wait(mbar, in_phase); in_phase ^= 1; }
        if (j < F1::T) {
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                const int u = j + r * F1::T;
                if (TMA_IN) v[r] = Yraw[((u >> 1) * C + c) * 2 + (u & 1)];
                else if (!(COLS_BIG_EXP & 2) || MODE != COLS_ITER)
                    v[r] = IN_T ? __ldg(tin + ((u >> 1) * kSpecTile + cc) * 2 + (u & 1)) : __ldg(in + (size_t)u * Wc);
            }
            dft_big<R0, -1>(v);
            F1::store(X, j, v);
        }
        __syncthreads();
        CB_ST(0);
        // the reads of the spectral update (A or Mul: entry r is frequency u = j + r T2) are issued here, TWO shared-memory
        // passes ahead of their use: with one CTA per SM nothing else hides their latency (per-phase clocks: 3 K of an
        // item's 22 K cycles were spent waiting for them when they were issued one pass ahead)
        float2 Av[R2];
        if (act) {
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                const int u = j + r * T2;
                Av[r] = (MODE == COLS_ITER) ? __ldg(At + u * C) : __ldg(a.Mul + (size_t)u * Wc + c0 + c);
            }
        }
        // ---- forward pass 2: X -> Y
        if (j < F2::T) { F2::load(X, j, v); F2::template butterfly_tab<R0>(v, tF2); F2::store(Y, j, v); }
        __syncthreads();
        CB_ST(1);
        // ---- forward pass 3 + spectral update + inverse pass 1 in registers
        if (act) {
            F3::load(Y, j, v);
            F3::template butterfly_tab<T2>(v, tF3);
        }
        const bool col0 = (c0 == 0);               // packed column 0 (DC + Nyquist) needs the mirrored frequency
        float2* mir = smem;                        // column-0 spectrum, natural order, in X (free at this point)
        if (col0) {
            if (act && c == 0) {
#pragma unroll
                for (int r = 0; r < R2; ++r) mir[j + r * T2] = v[r];
            }
            __syncthreads();
        }
        if (act) {
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                const int u = j + r * T2;
                float2 o;
                if (MODE == COLS_ITER) {
                    const float bm = __ldg(Bt + u * C);
                    o = make_float2(fmaf(bm, v[r].x, Av[r].x), fmaf(bm, v[r].y, Av[r].y));
                    if (col0 && c == 0) {
                        const float bq = __ldg(a.Bq + u);
                        const float2 zm = mir[u == 0 ? 0 : H - u];
                        o.x = fmaf(bq, zm.x, o.x);
                        o.y = fmaf(-bq, zm.y, o.y);
                    }
                } else {
                    o = cmul(Av[r], v[r]);
                    if (col0 && c == 0) o = cadd(o, cmul(__ldg(a.Mq + u), cconj(mir[u == 0 ? 0 : H - u])));
                    At[u * C] = o;
                }
                v[r] = o;
            }
            dft_big<R2, +1>(v);
        }
        if (col0) __syncthreads();                 // the mirrored entries in X have been read
        if (act) I1::store(X, j, v);
        __syncthreads();
        CB_ST(2);
#if COLS_BIG_APF_LATE
        // the next item's slice of A -> L2 from here (half an item ahead of its use: a prefetch a whole item ahead has partly
        // left the L2 again by the time it is needed)
        if (MODE == COLS_ITER && item + (int)gridDim.x < nitems && threadIdx.x == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.A + (size_t)(item + gridDim.x) * H * C),
                         "r"((unsigned)(H * C * sizeof(float2))) : "memory");
#endif
        // ---- inverse pass 2: X -> Y  (Y was last read by forward pass 3, before the barrier above)
        if (j < I2::T) { I2::load(X, j, v); I2::template butterfly_tab<R2>(v, tI2); I2::store(Y, j, v); }
        __syncthreads();
        CB_ST(3);
        // ---- inverse pass 3: Y -> registers -> global
        if (TMA_IN) {
            if (j < I3::T) I3::load(Y, j, v);
            __syncthreads();                               // Y is free: the next item's tile may land in it
            if (threadIdx.x == 0 && item + (int)gridDim.x < nitems) fetch_tile(item + gridDim.x);
        }
        if (j < I3::T) {
            if (!TMA_IN) I3::load(Y, j, v);
            I3::template butterfly_tab<CF::T0>(v, tI3);
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                const int u = j + r * I3::T;
                if ((COLS_BIG_EXP & 1) && MODE == COLS_ITER && v[r].x != 12345.678f) continue;
                if (OUT_T) {                       // x spectrum: rows (2k-1, 2k) share a slot pair, row H-1 pairs with row 0
                    const int k = (u + 1 == H) ? 0 : ((u + 1) >> 1);
                    tout[(k * kSpecTile + cc) * 2 + ((u + 1) & 1)] = v[r];
                } else {
                    out[(size_t)u * Wc] = v[r];
                }
            }
        }
        // no barrier here: the next item writes X (free) first, and Y only after two more barriers
        CB_ST(4);
#ifdef COLS_STATS
        ++nst;
#endif
    }
#ifdef COLS_STATS
    if (MODE == COLS_ITER && threadIdx.x == 0 && (blockIdx.x % 37) == 3 && nst)
        printf("cta %d items %d: prefetch+load+F1 %lld | F2 %lld | F3,upd,I1 %lld | I2 %lld | I3+store %lld\n", blockIdx.x, nst, st[0] / nst,
               st[1] / nst, st[2] / nst, st[3] / nst, st[4] / nst);
#endif
}

// Bm (H x Wc, row-major, shared with the generic kernels) -> [tile][u][C]
__global__ void k_bm_tiled(const float* __restrict__ Bm, float* __restrict__ Bmt, int H, int Wc, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * Wc) return;
    const int u = i / Wc, col = i - u * Wc;
    Bmt[((size_t)(col / C) * H + u) * C + col % C] = Bm[i];
}

static int cols_big_tile(int H) { return H == 2160 ? ColBig<2160>::C : 4; }

int launch_bm_tiled(const Geometry& g, const float* Bm, float* Bmt, cudaStream_t st) {
    const int n = g.H * g.Wc;
    ProfScope ps(PROF_OTHER, st);
    k_bm_tiled<<<(n + 255) / 256, 256, 0, st>>>(Bm, Bmt, g.H, g.Wc, cols_big_tile(g.H));
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

bool cols_big_supported(const Geometry& g) {
    if (options().force_generic || !(options().use_big & 2)) return false;
    return (g.H == 2160 || g.H == 1080 || g.H == 1024 || g.H == 1440 || g.H == 720 || g.H == 2048 || g.H == 768 || g.H == 1536) && (g.Wc % kSpecTile == 0);
}

typedef CUresult (*PFN_cbEncode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_cbEncode cb_encode() {
    static const PFN_cbEncode fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return (PFN_cbEncode)p;
        return (PFN_cbEncode) nullptr;
    }();
    return fn;
}

template <int H>
static int launch_cols_big_h(ColMode mode, const Geometry& g, const ColArgs& a, cudaStream_t st) {
    using CF = ColBigCfg<H>;
    // tile-major input spectrum as a 2-D fp32 tensor: a row = one row pair of one 8-column tile (128 bytes)
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof(tmap));
    if (CF::kTmaIn && mode == COLS_ITER && a.in_tiled) {
        PFN_cbEncode enc = cb_encode();
        if (!enc) return fail(4, "cuTensorMapEncodeTiled is not available");
        const cuuint64_t dims[2] = {(cuuint64_t)kSpecTile * 4, (cuuint64_t)g.P * (g.Wc / kSpecTile) * (H / 2)};
        const cuuint64_t strides[1] = {(cuuint64_t)kSpecTile * 4 * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)(CF::C * 4), (cuuint32_t)CF::kBoxRows};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.spec_in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(3, "cuTensorMapEncodeTiled failed");
    }
    const int ntiles = g.Wc / CF::C;
    const int nitems = ntiles * g.P;
    dim3 grid((unsigned)std::min(nitems, 148 * ColBig<H>::OCC));
    int dev = 0;
    ADMM_CUDA_CHECK(cudaGetDevice(&dev));
    ProfScope ps(mode == COLS_ITER ? PROF_COLS : PROF_OTHER, st);
#define ADMM_LAUNCH_COLS_BIG(M, IT, OT)                                                                              \
    do {                                                                                                            \
        static std::atomic<bool> attr_set[64];             /* per instantiation and device; the call costs microseconds */  \
        if (dev >= 64 || !attr_set[dev]) {                                                                          \
            ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_cols_big<H, M, IT, OT>,                                          \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::smem));      \
            if (dev < 64) attr_set[dev] = true;                                                                     \
        }                                                                                                           \
        k_cols_big<H, M, IT, OT><<<grid, CF::NT, CF::smem, st>>>(tmap, a, g.Wc, ntiles, nitems);                          \
    } while (0)
    if (mode == COLS_ITER) {
        if (a.in_tiled && a.out_tiled) ADMM_LAUNCH_COLS_BIG(COLS_ITER, true, true);
        else if (a.in_tiled) ADMM_LAUNCH_COLS_BIG(COLS_ITER, true, false);
        else if (a.out_tiled) return fail(4, "large-column kernel: tiled output needs tiled input in COLS_ITER");
        else ADMM_LAUNCH_COLS_BIG(COLS_ITER, false, false);
    } else {
        if (a.in_tiled) return fail(4, "large-column kernel: COLS_INIT reads the row-major spectrum");
        if (a.out_tiled) ADMM_LAUNCH_COLS_BIG(COLS_INIT, false, true);
        else ADMM_LAUNCH_COLS_BIG(COLS_INIT, false, false);
    }
#undef ADMM_LAUNCH_COLS_BIG
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_cols_big(ColMode mode, const Geometry& g, const ColArgs& a, cudaStream_t st) {
    if (mode != COLS_ITER && mode != COLS_INIT) return fail(4, "large-column kernel: unsupported mode");
    switch (g.H) {
        case 2160: return launch_cols_big_h<2160>(mode, g, a, st);
        case 1080: return launch_cols_big_h<1080>(mode, g, a, st);
        case 1024: return launch_cols_big_h<1024>(mode, g, a, st);
        case 1440: return launch_cols_big_h<1440>(mode, g, a, st);
        case 720: return launch_cols_big_h<720>(mode, g, a, st);
        case 2048: return launch_cols_big_h<2048>(mode, g, a, st);
        case 768: return launch_cols_big_h<768>(mode, g, a, st);
        case 1536: return launch_cols_big_h<1536>(mode, g, a, st);
        default: return fail(4, "no large-column kernel for this height");
    }
}

}  // namespace admm
