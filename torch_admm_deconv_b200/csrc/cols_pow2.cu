// cols_pow2.cu -- specialised column-pass kernel of the ADMM iteration for H in {128, 256, 512} (sm_100a).
//
//   packed row spectrum of v  --FFT along H-->  V  -->  X = A + Bm * V  --inverse FFT along H-->  row spectrum of x
//   (deconv.py:104-106: freq_c * rfftn(...) then irfftn; A and Bm fold H_t(xin), rho and 1/(HW))
//
// One CTA (256 threads; 512 for H = 512) owns a tile of T = 16 packed columns of one plane.  A thread owns TWO adjacent
// columns and 8 points of each (one radix-8 butterfly, or two radix-4): every shared-memory and global access is a
// 16-byte float4 = (column c, column c+1), which halves the load/store instruction count of the exchange passes, and the
// Stockham twiddle is shared by both columns.  The tile lives in shared memory "column fastest" so consecutive
// lanes touch consecutive 16-byte words for every access pattern.  The last forward pass leaves the spectrum in
// exactly the register layout the first inverse pass consumes, so Bm V happens in registers; the constant term A
// enters one butterfly later: the array `A` holds P0[A] (A after inverse pass 0), copied into the tile buffer with
// cp.async behind the last forward pass and added to the outputs of inverse pass 0 (cols_pow2_body.cuh).  On square
// planes the packed width is a template parameter (strides as immediates, no spills at 64 registers).
#include "cols_pow2_body.cuh"

namespace admm {

template <int H, int MODE, int WCT>
__global__ void __launch_bounds__(col_threads<H>(), 1024 / col_threads<H>())
k_cols_pow2(ColArgs a, int Wc, int ntiles, int pdl) {
    extern __shared__ float4 smem4[];
    constexpr int NT = ColCfg<H>::kThreads;
    cols_pow2_body<H, MODE, NT, false, WCT>(a, Wc, ntiles, pdl, blockIdx.x, smem4);
}

// WCT = H / 2 on square planes (every BASELINE configuration of these sizes), else 0
template <int H, int MODE, int WCT>
static int launch_cols_pow2_w(const Geometry& g, const ColArgs& a, cudaStream_t st) {
    using C = ColCfg<H>;
    const int ntiles = g.Wc / C::T;
    // the opt-in shared-memory limit is a per-device function attribute: remember which devices have it
    static std::atomic<bool> attr_set_dev[64];
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    std::atomic<bool>& attr_set = attr_set_dev[dev_id & 63];
    if (!attr_set) {
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_cols_pow2<H, MODE, WCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::bytes));
        attr_set = true;
    }
    ProfScope ps(MODE == COLS_ITER ? PROF_COLS : PROF_OTHER, st);
    const size_t nctas = (size_t)ntiles * g.P;
    if (MODE == COLS_ITER && options().use_pdl && nctas <= 148 * 8) {
        ADMM_CUDA_CHECK(launch_pdl(k_cols_pow2<H, MODE, WCT>, dim3((unsigned)nctas), dim3(C::kThreads), C::bytes, st, a, g.Wc, ntiles, 1));
    } else {
        // pdl < 0 carries the next-tile prefetch distance (resident CTAs) for large grids
        const int pf = (MODE == COLS_ITER && options().cols_prefetch) ? -(148 * (1024 / C::kThreads)) : 0;
        k_cols_pow2<H, MODE, WCT><<<(unsigned)nctas, C::kThreads, C::bytes, st>>>(a, g.Wc, ntiles, pf);
    }
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int H, int MODE>
static int launch_cols_pow2_m(const Geometry& g, const ColArgs& a, cudaStream_t st) {
    if (g.Wc == H / 2) return launch_cols_pow2_w<H, MODE, H / 2>(g, a, st);
    return launch_cols_pow2_w<H, MODE, 0>(g, a, st);
}

template <int H>
static int launch_cols_pow2_t(ColMode mode, const Geometry& g, const ColArgs& a, cudaStream_t st) {
    switch (mode) {
        case COLS_ITER: return launch_cols_pow2_m<H, COLS_ITER>(g, a, st);
        case COLS_INIT: return launch_cols_pow2_m<H, COLS_INIT>(g, a, st);
        case COLS_FFT_FWD: return launch_cols_pow2_m<H, COLS_FFT_FWD>(g, a, st);
        case COLS_BM_INV: return launch_cols_pow2_m<H, COLS_BM_INV>(g, a, st);
        case COLS_CMUL_INV: return launch_cols_pow2_m<H, COLS_CMUL_INV>(g, a, st);
        case COLS_INIT_SPEC: return launch_cols_pow2_m<H, COLS_INIT_SPEC>(g, a, st);
        default: break;
    }
    return fail(4, "cols_pow2: unsupported mode");
}

bool cols_pow2_mode_supported(ColMode mode) {
    return mode == COLS_ITER || mode == COLS_INIT || mode == COLS_FFT_FWD || mode == COLS_BM_INV || mode == COLS_CMUL_INV ||
           mode == COLS_INIT_SPEC;
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

int launch_cols_pow2(ColMode mode, const Geometry& g, const ColArgs& a, cudaStream_t st) {
    switch (g.H) {
        case 128: return launch_cols_pow2_t<128>(mode, g, a, st);
        case 256: return launch_cols_pow2_t<256>(mode, g, a, st);
        case 512: return launch_cols_pow2_t<512>(mode, g, a, st);
    }
    return fail(4, "cols_pow2: unsupported height");
}

}  // namespace admm

// ------------------------------------------------------------------------------------------ fused backward column pass
namespace admm {

// One kernel per backward iteration on the column side:
//     G = F_col(row spectrum of xbar);  Gs += G;  [V = F_col(row spectrum of v_k);  GVp += conj(G) V / (HW)]
//     row spectrum of vbar = F_col^-1[Bm G]
// GVp is kept per plane (plain read-modify-write by the owning thread, deterministic); packed column 0 carries the
// DC and the Nyquist column, whose products are formed on the unpacked values (mirror entries through shared
// memory) and stored in GVp[.., 0] (DC) and GVn (Nyquist).
struct ColAdjArgs {
    const float2* spec_x;     // row spectrum of xbar
    const float2* spec_v;     // row spectrum of v_k or NULL
    float2* Gs;               // per-plane sum of G (packed)
    float2* GVp;              // per-plane sum of conj(G) V / (HW) (column 0: DC product)
    float2* GVn;              // per-plane Nyquist products, planes x H
    float2* spec_out;         // row spectrum of vbar, or NULL on the last sweep step
    const float* Bm; const float* Bq;
    const float2* tw;
};

// element-wise fp32 add of four consecutive floats in global memory, no return value
__device__ __forceinline__ void red_add4(float2* addr, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#ifndef COLS_ADJ_OCC
#define COLS_ADJ_OCC 3
#endif
// WCT: the packed width as a compile-time constant (square planes), 0 = run-time value (see cols_pow2_body)
template <int H, int WCT>
__global__ void __launch_bounds__(256, COLS_ADJ_OCC)
k_cols_adj(ColAdjArgs a, int Wc_dyn, int ntiles_dyn, float inv_hw, int pdl) {
    using C = ColCfg<H, 256>;
    const int Wc = WCT ? WCT : Wc_dyn;
    const int ntiles = WCT ? WCT / C::T : ntiles_dyn;
    using CR = ColRadix<H>;
    constexpr int TPS = C::TPS, T = C::T, NPAIRS = C::NPAIRS;
    constexpr int NB2 = kCP / CR::F2;
    extern __shared__ float4 smem4[];
    float4* buf = smem4;
    float2* tabs = reinterpret_cast<float2*>(buf + H * NPAIRS);
    float2* zcolG = tabs + C::TAB_END;
    float2* zcolV = zcolG + H;
    const int tid = threadIdx.x;
    const int pr = tid % NPAIRS;
    const int t = tid / NPAIRS;
    const int tile = blockIdx.x % ntiles;
    const int p = blockIdx.x / ntiles;
    const int c = tile * T + 2 * pr;
    const size_t plane = (size_t)p * H * Wc;
    const bool has_v = (a.spec_v != nullptr);
    const bool col0 = (tile == 0 && pr == 0);

    if (pdl) {
        // programmatic dependent launch: the tables are built while the previous kernel drains; nothing it wrote (and
        // nothing it may still read) is touched before pdl_wait()
        pdl_launch_dependents();
        build_tab1<H, CR::F1, CR::F0>(tabs + C::TAB_F1, a.tw);
        build_tab1<H, CR::F2, CR::F0 * CR::F1>(tabs + C::TAB_F2, a.tw);
        if (!C::kShare) {
            build_tab1<H, CR::F1, CR::F2>(tabs + C::TAB_I1, a.tw);
            build_tab1<H, CR::F0, CR::F2 * CR::F1>(tabs + C::TAB_I2, a.tw);
        }
        pdl_wait();
    }
    float4 dG[kCP], dV[kCP];
    {
        const float2* in = a.spec_x + plane + c;
#pragma unroll
        for (int q = 0; q < kCP; ++q) dG[q] = __ldg(reinterpret_cast<const float4*>(in + (size_t)(t + q * TPS) * Wc));
        if (has_v) {
            const float2* iv = a.spec_v + plane + c;
#pragma unroll
            for (int q = 0; q < kCP; ++q) dV[q] = __ldg(reinterpret_cast<const float4*>(iv + (size_t)(t + q * TPS) * Wc));
        }
    }
    if (!pdl) {
        build_tab1<H, CR::F1, CR::F0>(tabs + C::TAB_F1, a.tw);
        build_tab1<H, CR::F2, CR::F0 * CR::F1>(tabs + C::TAB_F2, a.tw);
        if (!C::kShare) {
            build_tab1<H, CR::F1, CR::F2>(tabs + C::TAB_I1, a.tw);
            build_tab1<H, CR::F0, CR::F2 * CR::F1>(tabs + C::TAB_I2, a.tw);
        }
    }
    auto forward = [&](float4 (&d)[kCP]) {
        cpass_compute<H, CR::F0, 1, -1>(d, t, nullptr);
        cpass_store<H, CR::F0, 1, NPAIRS>(d, t, pr, buf);
        __syncthreads();
        cpass_load<H, NPAIRS>(d, t, pr, buf);
        cpass_compute<H, CR::F1, CR::F0, -1>(d, t, tabs + C::TAB_F1);
        __syncthreads();
        cpass_store<H, CR::F1, CR::F0, NPAIRS>(d, t, pr, buf);
        __syncthreads();
        cpass_load<H, NPAIRS>(d, t, pr, buf);
        cpass_compute<H, CR::F2, CR::F0 * CR::F1, -1>(d, t, tabs + C::TAB_F2);
    };
    forward(dG);
    if (has_v) {
        __syncthreads();                                   // every thread finished reading buf
        forward(dV);
    }
    // slot (m, r) <-> u = (t + m*TPS) + r*(H/F2) for both register sets
    if (col0) {
#pragma unroll
        for (int m = 0; m < NB2; ++m)
#pragma unroll
            for (int r = 0; r < CR::F2; ++r) {
                const int u = (t + m * TPS) + r * (H / CR::F2);
                zcolG[u] = make_float2(dG[m + r * NB2].x, dG[m + r * NB2].y);
                if (has_v) zcolV[u] = make_float2(dV[m + r * NB2].x, dV[m + r * NB2].y);
            }
    }
    __syncthreads();                                       // buf free for the inverse; mirror columns visible
    {
        float2* Gsp = a.Gs + plane + c;
        float2* GVpp = a.GVp + plane + c;
        const float* __restrict__ Bp = a.Bm + c;
#pragma unroll
        for (int m = 0; m < NB2; ++m) {
#pragma unroll
            for (int r = 0; r < CR::F2; ++r) {
                const int u = (t + m * TPS) + r * (H / CR::F2);
                const size_t off = (size_t)u * Wc;
                const float4 G = dG[m + r * NB2];
                // Gs += G.  The running sums are updated with vector reductions (RED.ADD.F32x4, no return value): the thread is
                // the only writer of its elements in this launch and launches are stream-ordered, so the result is the same
                // sequence of fp32 additions as a load / add / store -- without two exposed loads in the middle of the chain
                red_add4(Gsp + off, G);
                if (has_v) {
                    const float4 V = dV[m + r * NB2];
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    // second column of the pair: always an ordinary packed column
                    acc.z += inv_hw * (G.z * V.z + G.w * V.w);
                    acc.w += inv_hw * (G.z * V.w - G.w * V.z);
                    if (!col0) {
                        acc.x += inv_hw * (G.x * V.x + G.y * V.y);
                        acc.y += inv_hw * (G.x * V.y - G.y * V.x);
                    } else {
                        // unpack DC and Nyquist columns of G and V:  D = (z + conj(zm))/2,  N = (z - conj(zm))/(2i)
                        const int um = (H - u) & (H - 1);
                        const float2 gm = zcolG[um], vm = zcolV[um];
                        const float2 gD = make_float2(0.5f * (G.x + gm.x), 0.5f * (G.y - gm.y));
                        const float2 gN = make_float2(0.5f * (G.y + gm.y), -0.5f * (G.x - gm.x));
                        const float2 vD = make_float2(0.5f * (V.x + vm.x), 0.5f * (V.y - vm.y));
                        const float2 vN = make_float2(0.5f * (V.y + vm.y), -0.5f * (V.x - vm.x));
                        acc.x += inv_hw * (gD.x * vD.x + gD.y * vD.y);
                        acc.y += inv_hw * (gD.x * vD.y - gD.y * vD.x);
                        float2 n = a.GVn[(size_t)p * H + u];
                        n.x += inv_hw * (gN.x * vN.x + gN.y * vN.y);
                        n.y += inv_hw * (gN.x * vN.y - gN.y * vN.x);
                        a.GVn[(size_t)p * H + u] = n;
                    }
                    red_add4(GVpp + off, acc);
                }
                // vbar spectrum = Bm G  (+ Bq conj(G[-u]) on packed column 0)
                const float2 bm = __ldg(reinterpret_cast<const float2*>(Bp + off));
                float4 o = make_float4(bm.x * G.x, bm.x * G.y, bm.y * G.z, bm.y * G.w);
                if (col0) {
                    const float2 Zm = zcolG[(H - u) & (H - 1)];
                    const float bq = a.Bq[u];
                    o.x = fmaf(bq, Zm.x, o.x);
                    o.y = fmaf(-bq, Zm.y, o.y);
                }
                dG[m + r * NB2] = o;
            }
        }
    }
    if (a.spec_out == nullptr) return;
    cpass_compute<H, CR::F2, 1, +1>(dG, t, nullptr);
    cpass_store<H, CR::F2, 1, NPAIRS>(dG, t, pr, buf);
    __syncthreads();
    cpass_load<H, NPAIRS>(dG, t, pr, buf);
    cpass_compute<H, CR::F1, CR::F2, +1>(dG, t, tabs + C::TAB_I1);
    __syncthreads();
    cpass_store<H, CR::F1, CR::F2, NPAIRS>(dG, t, pr, buf);
    __syncthreads();
    cpass_load<H, NPAIRS>(dG, t, pr, buf);
    cpass_compute<H, CR::F0, CR::F2 * CR::F1, +1>(dG, t, tabs + C::TAB_I2);
    {
        constexpr int NB = kCP / CR::F0;
        float2* out = a.spec_out + plane + c;
#pragma unroll
        for (int m = 0; m < NB; ++m)
#pragma unroll
            for (int r = 0; r < CR::F0; ++r)
                *reinterpret_cast<float4*>(out + (size_t)((t + m * TPS) + r * (H / CR::F0)) * Wc) = dG[m + r * NB];
    }
}

template <int H, int WCT>
static int launch_cols_adj_w(const Geometry& g, const ColAdjArgs& a, cudaStream_t st) {
    using C = ColCfg<H, 256>;
    const size_t bytes = (size_t)(H * C::T + C::TAB_END + 2 * H) * sizeof(float2);
    static std::atomic<bool> attr_set_dev[64];
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    if (!attr_set_dev[dev_id & 63]) {
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_cols_adj<H, WCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        attr_set_dev[dev_id & 63] = true;
    }
    const int ntiles = g.Wc / C::T;
    ProfScope ps(PROF_OTHER, st);
    const size_t nctas = (size_t)ntiles * g.P;
    const float inv_hw = 1.0f / ((float)g.H * (float)g.W);
    if (options().use_pdl && nctas <= 148 * 8) {
        ADMM_CUDA_CHECK(launch_pdl(k_cols_adj<H, WCT>, dim3((unsigned)nctas), dim3(256), bytes, st, a, g.Wc, ntiles, inv_hw, 1));
    } else {
        k_cols_adj<H, WCT><<<(unsigned)nctas, 256, bytes, st>>>(a, g.Wc, ntiles, inv_hw, 0);
    }
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int H>
static int launch_cols_adj_t(const Geometry& g, const ColAdjArgs& a, cudaStream_t st) {
    if (g.Wc == H / 2) return launch_cols_adj_w<H, H / 2>(g, a, st);
    return launch_cols_adj_w<H, 0>(g, a, st);
}

bool cols_adj_supported(const Geometry& g) {
    if (options().force_generic) return false;
    if (g.W % 2) return false;
    switch (g.H) {
        case 128: return g.Wc % ColCfg<128, 256>::T == 0;
        case 256: return g.Wc % ColCfg<256, 256>::T == 0;
        case 512: return g.Wc % ColCfg<512, 256>::T == 0;
    }
    return false;
}

int launch_cols_adj(const Geometry& g, const float2* spec_x, const float2* spec_v, float2* Gs, float2* GVp, float2* GVn,
                    float2* spec_out, const float* Bm, const float* Bq, const float2* tw, cudaStream_t st) {
    ColAdjArgs a{spec_x, spec_v, Gs, GVp, GVn, spec_out, Bm, Bq, tw};
    switch (g.H) {
        case 128: return launch_cols_adj_t<128>(g, a, st);
        case 256: return launch_cols_adj_t<256>(g, a, st);
        case 512: return launch_cols_adj_t<512>(g, a, st);
    }
    return fail(4, "cols_adj: unsupported height");
}

}  // namespace admm
