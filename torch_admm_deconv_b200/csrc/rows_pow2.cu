// rows_pow2.cu -- specialised row-pass kernel of the ADMM iteration for W in {128, 256, 512} (sm_100a).
//
//   packed row spectrum of x_k  --C2R-->  x_k  --prox / dual / divergence-->  v_{k+1}  --R2C-->  packed spectrum
//   (deconv.py:106 irfftn rows, :108-115 Dx/Dy/soft_thresh/dual update, :104 Dx_t/Dy_t + rfftn rows)
//
// One CTA (256 threads) owns a band of Rb image rows of one plane plus one halo row above and below.
// Two image rows share one complex FFT (z = row_a + i row_b); TPS = W/16 consecutive threads of one warp
// own one row pair and keep 16 complex points each in registers, so the FFT passes need only __syncwarp.
// The first inverse pass and the last forward pass give every thread the two butterflies j and W/8 - j:
// the Hermitian merge (two packed half spectra -> full spectrum of z) and split (back) then happen
// entirely in registers and the global accesses are 256-byte coalesced runs, 8 bytes per lane.
#include "common.cuh"
#include "fft_pow2.cuh"

namespace admm {

template <int W> struct RowRadix;
template <> struct RowRadix<512> { static constexpr int IB = 8, IC = 8, FA = 8, FB = 8; };
template <> struct RowRadix<256> { static constexpr int IB = 8, IC = 4, FA = 4, FB = 8; };
template <> struct RowRadix<128> { static constexpr int IB = 4, IC = 4, FA = 4, FB = 4; };

template <int W> struct RowSmem {
    using RR = RowRadix<W>;
    static constexpr int kThreads = 256;
    static constexpr int TPS = W / kPT;                       // threads per row pair
    static constexpr int NPAIR = kThreads / TPS;              // row pairs per CTA (x side, halo included)
    static constexpr int REGION = W + W / 8;                  // padded complex slots per pair
    static constexpr int RMAX = 2 * NPAIR - 2;                // band rows per CTA
    // twiddle tables: inverse pass B (IB, Ns=8), inverse pass C (IC, Ns=8*IB), forward pass B (FB, Ns=FA),
    // forward pass C (8, Ns=W/8)
    static constexpr int TAB_IB = 0;
    static constexpr int TAB_IC = TAB_IB + tab_size(RR::IB, 8);
    static constexpr int TAB_FB = TAB_IC + tab_size(RR::IC, 8 * RR::IB);
    static constexpr int TAB_FC = TAB_FB + tab_size(RR::FB, RR::FA);
    static constexpr int TAB_END = TAB_FC + tab_size(8, W / 8);
    static constexpr size_t bytes = (size_t)((2 * NPAIR - 1) * REGION + TAB_END) * sizeof(float2);
};

__device__ __forceinline__ float clampf2(float q, float tau) { return fminf(fmaxf(q, -tau), tau); }
__device__ __forceinline__ float wfun2(float q, float tau) { return fmaf(-2.0f, clampf2(q, tau), q); }
__device__ __forceinline__ float2 mk(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }       // a + i b
__device__ __forceinline__ float2 mkc(float2 a, float2 b) { return make_float2(a.x + b.y, b.x - a.y); }      // conj(a) + i conj(b)

template <int W>
__global__ void __launch_bounds__(256, 3)
k_rows_full_pow2(RowArgs a, int H, int nbands) {
    using S = RowSmem<W>;
    using RR = RowRadix<W>;
    constexpr int TPS = S::TPS, NPAIR = S::NPAIR, REGION = S::REGION;
    constexpr int T8 = W / 8;                 // butterflies of the radix-8 edge passes
    constexpr int Wc = W / 2;
    extern __shared__ float2 smem[];
    float2* regX = smem;                                    // NPAIR regions
    float2* regV = regX + NPAIR * REGION;                   // NPAIR-1 regions
    float2* tabs = regV + (NPAIR - 1) * REGION;
    const RowMapObj map;

    const int tid = threadIdx.x;
    const int pair = tid / TPS;               // row pair handled by this thread in the FFT phases
    const int t = tid % TPS;
    const int band = blockIdx.x % nbands;
    const int p = blockIdx.x / nbands;
    // balanced even band sizes: rows [r0, r1)
    const int hh = H >> 1;
    const int r0 = 2 * (int)(((long long)band * hh) / nbands);
    const int r1 = 2 * (int)(((long long)(band + 1) * hh) / nbands);
    const int Rb = r1 - r0;                   // even, <= RMAX
    const int npx = Rb / 2 + 1;               // x pairs (rows r0-1 .. r0+Rb)
    const int npv = Rb / 2;                   // v pairs (rows r0 .. r0+Rb-1)
    const size_t plane_real = (size_t)p * H * W;
    const size_t plane_spec = (size_t)p * H * Wc;

    // edge-pass butterflies of this thread: j1 = t, j2 = T8 - t   (t == 0: j1 = 0, j2 = T8/2)
    const bool t0 = (t == 0);
    const int j1 = t;
    const int j2 = t0 ? (T8 / 2) : (T8 - t);
    float2 d[kPT];
    float2* myX = regX + pair * REGION;
    // the TPS lanes of one row pair synchronise among themselves only (a warp may hold several pairs, and the
    // last pair of a band can be inactive)
    const unsigned pmask = (TPS >= 32) ? 0xffffffffu
                                       : (((1u << (TPS & 31)) - 1u) << (((tid & 31) / TPS) * TPS));

    // issue the global loads of the merge first, build the twiddle tables while they are in flight
    float2 A1[4], B1[4], A2[4], B2[4];
    if (pair < npx) {
        // rows (circular): ia = r0 - 1 + 2*pair, ib = ia + 1
        int ra = r0 - 1 + 2 * pair; if (ra < 0) ra += H; if (ra >= H) ra -= H;
        int rb = ra + 1; if (rb >= H) rb -= H;
        const float2* Sa = a.spec_in + plane_spec + (size_t)ra * Wc;
        const float2* Sb = a.spec_in + plane_spec + (size_t)rb * Wc;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            A1[r] = Sa[j1 + r * T8]; B1[r] = Sb[j1 + r * T8];
            A2[r] = Sa[j2 + r * T8]; B2[r] = Sb[j2 + r * T8];
        }
    }
    build_tab<W, RR::IB, 8>(tabs + S::TAB_IB, a.tw);
    build_tab<W, RR::IC, 8 * RR::IB>(tabs + S::TAB_IC, a.tw);
    build_tab<W, RR::FB, RR::FA>(tabs + S::TAB_FB, a.tw);
    build_tab<W, 8, W / 8>(tabs + S::TAB_FC, a.tw);
    __syncthreads();

    // ------------------------------------------------------------------ C2R: merge + inverse FFT
    if (pair < npx) {
        // slot (m, r) = d[m + 2r]: m = 0 -> butterfly j1, m = 1 -> butterfly j2; point n = j + r*T8
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            d[0 + 2 * r] = mk(A1[r], B1[r]);
            d[1 + 2 * r] = mk(A2[r], B2[r]);
        }
#pragma unroll
        for (int r = 4; r < 8; ++r) {
            // n = j + r*T8 > W/2: Z[n] = conj(P_a[W-n]) + i conj(P_b[W-n]);  W - n = (T8 - j) + (7 - r) T8
            const float2 g1a = t0 ? A1[(8 - r) & 3] : A2[7 - r];      // t0: W - r*T8 = (8 - r) T8
            const float2 g1b = t0 ? B1[(8 - r) & 3] : B2[7 - r];
            const float2 g2a = t0 ? A2[7 - r] : A1[7 - r];            // t0: j2 = T8/2 is self-paired
            const float2 g2b = t0 ? B2[7 - r] : B1[7 - r];
            d[0 + 2 * r] = mkc(g1a, g1b);
            d[1 + 2 * r] = mkc(g2a, g2b);
        }
        if (t0) {
            // packed column 0 = (DC, Nyquist) of each row, both real
            d[0] = make_float2(A1[0].x, B1[0].x);                      // Z[0]
            d[0 + 2 * 4] = make_float2(A1[0].y, B1[0].y);              // Z[W/2] = Z[4*T8]
        }
        // first inverse pass: radix 8, no twiddles
        {
            float2 v0[8], v1[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) { v0[r] = d[2 * r]; v1[r] = d[1 + 2 * r]; }
            dft8<+1>(v0); dft8<+1>(v1);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                myX[map.at(8 * j1 + r)] = v0[r];
                myX[map.at(8 * j2 + r)] = v1[r];
            }
        }
        __syncwarp(pmask);
        pass_load<W>(d, t, myX, map);
        pass_compute<W, RR::IB, 8, +1>(d, t, tabs + S::TAB_IB);
        __syncwarp(pmask);
        pass_store<W, RR::IB, 8>(d, t, myX, map);
        __syncwarp(pmask);
        pass_load<W>(d, t, myX, map);
        pass_compute<W, RR::IC, 8 * RR::IB, +1>(d, t, tabs + S::TAB_IC);
        __syncwarp(pmask);
        pass_store<W, RR::IC, 8 * RR::IB>(d, t, myX, map);      // natural order: myX[at(c)] = (x_a[c], x_b[c])
    }
    __syncthreads();

    // ------------------------------------------------------------------ prox / dual update / divergence
    {
        const float tau = a.lmbd[0] / a.rho[0];                         // deconv.py:44
        const float* qxi = a.qx_in ? a.qx_in + plane_real : nullptr;
        const float* qyi = a.qy_in ? a.qy_in + plane_real : nullptr;
        float* qxo = a.qx_out + plane_real;
        float* qyo = a.qy_out + plane_real;
        for (int c = tid; c < W; c += 256) {
            const int cl = (c == 0) ? W - 1 : c - 1;
            const int cr = (c == W - 1) ? 0 : c + 1;
            const int pc = map.at(c), pl = map.at(cl), pr = map.at(cr);
            // x row i lives in pair i>>1, component i&1 (i = 0 is the halo row r0-1)
            float2 Pc = regX[pc];
            float2 Pl = regX[pl];
            float2 Pr = regX[pr];
            float xprev = Pc.x;                       // x[i = 0]
            // w_y of band row 0 (i = 1)
            int row = r0;
            float uy = qyi ? clampf2(qyi[(size_t)row * W + c], tau) : 0.f;
            float qy_cur = Pc.y - xprev + uy;
            float wy_cur = wfun2(qy_cur, tau);
            for (int m = 0; m < npv; ++m) {
                // band rows b = 2m (i = 2m+1, = component y of pair m) and b + 1 (i = 2m+2, component x of pair m+1)
                const float2 Nc = regX[(m + 1) * REGION + pc];
                const float2 Nl = regX[(m + 1) * REGION + pl];
                const float2 Nr = regX[(m + 1) * REGION + pr];
                const int ra = r0 + 2 * m, rb = ra + 1;
                int rc = rb + 1; if (rc >= H) rc -= H;
                float uxa = 0.f, uxar = 0.f, uxb = 0.f, uxbr = 0.f, uyb = 0.f, uyc = 0.f;
                if (qxi) {
                    uxa  = clampf2(qxi[(size_t)ra * W + c], tau);
                    uxar = clampf2(qxi[(size_t)ra * W + cr], tau);
                    uxb  = clampf2(qxi[(size_t)rb * W + c], tau);
                    uxbr = clampf2(qxi[(size_t)rb * W + cr], tau);
                    uyb  = clampf2(qyi[(size_t)rb * W + c], tau);
                    uyc  = clampf2(qyi[(size_t)rc * W + c], tau);
                }
                // row a: x = Pc.y, left Pl.y, right Pr.y            (deconv.py:108, 111, 114)
                const float qx_a  = Pc.y - Pl.y + uxa;
                const float qx_ar = Pr.y - Pc.y + uxar;
                // row b: x = Nc.x
                const float qy_b  = Nc.x - Pc.y + uyb;                 // deconv.py:109, 112, 115
                const float wy_b  = wfun2(qy_b, tau);
                const float qx_b  = Nc.x - Nl.x + uxb;
                const float qx_br = Nr.x - Nc.x + uxbr;
                // row below b: x = Nc.y
                const float qy_c  = Nc.y - Nc.x + uyc;
                const float wy_c  = wfun2(qy_c, tau);
                const float va = wfun2(qx_a, tau) - wfun2(qx_ar, tau) + wy_cur - wy_b;    // deconv.py:104
                const float vb = wfun2(qx_b, tau) - wfun2(qx_br, tau) + wy_b - wy_c;
                qxo[(size_t)ra * W + c] = qx_a;
                qyo[(size_t)ra * W + c] = qy_cur;
                qxo[(size_t)rb * W + c] = qx_b;
                qyo[(size_t)rb * W + c] = qy_b;
                regV[m * REGION + pc] = make_float2(va, vb);
                Pc = Nc; Pl = Nl; Pr = Nr;
                qy_cur = qy_c; wy_cur = wy_c;
            }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ R2C: forward FFT + split
    if (pair < npv) {
        float2* myV = regV + pair * REGION;
        pass_load<W>(d, t, myV, map);
        pass_compute<W, RR::FA, 1, -1>(d, t, nullptr);
        __syncwarp(pmask);
        pass_store<W, RR::FA, 1>(d, t, myV, map);
        __syncwarp(pmask);
        pass_load<W>(d, t, myV, map);
        pass_compute<W, RR::FB, RR::FA, -1>(d, t, tabs + S::TAB_FB);
        __syncwarp(pmask);
        pass_store<W, RR::FB, RR::FA>(d, t, myV, map);
        __syncwarp(pmask);
        // last pass: radix 8, Ns = T8, butterflies j1 and j2; inputs j + r*T8, twiddle k = j
        float2 v0[8], v1[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            v0[r] = myV[map.at(j1 + r * T8)];
            v1[r] = myV[map.at(j2 + r * T8)];
        }
        const float2* tabC = tabs + S::TAB_FC;
#pragma unroll
        for (int r = 1; r < 8; ++r) {
            v0[r] = cmul(v0[r], tabC[(r - 1) * T8 + j1]);
            v1[r] = cmul(v1[r], tabC[(r - 1) * T8 + j2]);
        }
        dft8<-1>(v0); dft8<-1>(v1);
        // v0[r] = Z[j1 + r T8], v1[r] = Z[j2 + r T8];  partner of n is W - n
        const int ra = r0 + 2 * pair;
        float2* Oa = a.spec_out + plane_spec + (size_t)ra * Wc;
        float2* Ob = Oa + Wc;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            // column c = j1 + r T8 (< W/2)
            const float2 Z1 = v0[r];
            const float2 M1 = t0 ? v0[(8 - r) & 7] : v1[7 - r];
            float2 Xa = make_float2(0.5f * (Z1.x + M1.x), 0.5f * (Z1.y - M1.y));
            float2 Xb = make_float2(0.5f * (Z1.y + M1.y), 0.5f * (M1.x - Z1.x));
            if (r == 0 && t0) {                                        // packed (DC, Nyquist)
                Xa = make_float2(v0[0].x, v0[4].x);
                Xb = make_float2(v0[0].y, v0[4].y);
            }
            Oa[j1 + r * T8] = Xa; Ob[j1 + r * T8] = Xb;
            // column c = j2 + r T8 (< W/2)
            const float2 Z2 = v1[r];
            const float2 M2 = t0 ? v1[7 - r] : v0[7 - r];
            Oa[j2 + r * T8] = make_float2(0.5f * (Z2.x + M2.x), 0.5f * (Z2.y - M2.y));
            Ob[j2 + r * T8] = make_float2(0.5f * (Z2.y + M2.y), 0.5f * (M2.x - Z2.x));
        }
    }
}

template <int W>
static int launch_rows_pow2_t(const Geometry& g, const RowArgs& a, cudaStream_t st) {
    using S = RowSmem<W>;
    const int nbands = (g.H + S::RMAX - 1) / S::RMAX;
    static bool attr_set = false;
    if (!attr_set) {
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_rows_full_pow2<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
        attr_set = true;
    }
    ProfScope ps(PROF_ROWS, st);
    k_rows_full_pow2<W><<<(unsigned)((size_t)nbands * g.P), 256, S::bytes, st>>>(a, g.H, nbands);
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

bool rows_pow2_supported(const Geometry& g) {
    if (options().force_generic) return false;
    if (g.H & 1) return false;
    return g.W == 128 || g.W == 256 || g.W == 512;
}

int launch_rows_pow2(const Geometry& g, const RowArgs& a, cudaStream_t st) {
    switch (g.W) {
        case 128: return launch_rows_pow2_t<128>(g, a, st);
        case 256: return launch_rows_pow2_t<256>(g, a, st);
        case 512: return launch_rows_pow2_t<512>(g, a, st);
    }
    return fail(4, "rows_pow2: unsupported width");
}

}  // namespace admm
