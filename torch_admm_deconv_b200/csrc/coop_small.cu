// coop_small.cu -- persistent cooperative kernel for SMALL, latency-bound batches on the power-of-two kernels (sm_100a).
//
// The reference's own workloads are small: scripts/train.py:19-24 trains ADMMDeconv(kern_size=(), max_iters=100, iso=True) on
// 3 x 3 x 256 x 256 batches and evaluates on 8 x 3 x 256 x 256 (configs/train_cfg.json:6,11,13): 9 .. 24 planes.  There the
// two (iso: four) kernels of an iteration are each a fraction of one wave and an iteration costs its kernel BOUNDARIES:
// ~5 us per dependent launch (3.1 ms per 100-iteration solve, 2.6 ms even as a CUDA graph) against ~2 us of work per phase.
//
// Here ONE cooperative launch runs iterations 1 .. maxit-1 of the whole batch: every phase executes the body of the
// corresponding stand-alone kernel (rows_pow2_body.cuh / cols_pow2_body.cuh, instantiated with COOP = true: coherent
// loads, virtual block index) over a grid-stride loop of virtual blocks, and a grid barrier replaces the kernel boundary.
// Phases per iteration (deconv.py:103-115):
//   iso = 0:  rows  (C2R -> prox / dual / divergence -> R2C)            | cols (FFT -> A + Bm V -> iFFT)
//   iso = 1:  rows C2R | per-pixel block threshold over the planes | rows R2C of D^T((2s-1) q) | cols
// The precompute (R2C of y, COLS_INIT) and the last C2R stay ordinary launches around it.  Inference only.
//
// MEASURED AND SWITCHED OFF (option use_coop, default 0; tools/coop_time.py, 100 iterations, ms per solve, separate launches /
// cooperative): 8x3x256^2 iso 3.11 / 3.42, 3x3x256^2 iso 2.38 / 2.21, 8x3x256^2 aniso 2.17 / 2.47, 16x3x256^2 iso 4.75 / 6.79,
// 1x3x512^2 iso 2.91 / 2.76, 8x3x128^2 iso 1.90 / 2.59 (cooperative_groups' grid.sync(): another 20-30 % slower than the
// counter barrier below).  A phase here is a single wave at 2 CTAs per SM with its load latency fully exposed, which costs
// more than the kernel boundary it saves; the separate launches (programmatic dependent launch) stay the default.  The
// results are bit-identical to the separate launches (tests/test_gpu_parity.py), the file documents the experiment.
#include "cols_pow2_body.cuh"
#include "rows_pow2_body.cuh"

namespace admm {

struct CoopArgs {
    // rows
    const float2* twW; const float* lmbd; const float* rho;
    float* q[2][2];                 // ping-pong state (u in the iso = 0 path, q in the iso = 1 path)
    // cols
    const float2* twH; float2* A; const float* Bm; const float* Bq;
    float2* S0; float2* S1;
    // iso
    float* xreal; float* nmap[2]; float* sbmap;
    int P, H, maxit, iso;
    int nb_full, nb_plain, ntiles;  // virtual grids: bands per plane (march / plain row modes), column tiles per plane
    unsigned* barrier;              // grid barrier counter (zeroed before the launch)
};

__device__ __forceinline__ float coop_iso_scale(float n, float tau) { return fmaxf(1.f - tau / (n + 1e-15f), 0.f); }

// Grid-wide barrier of the co-resident CTAs of a cooperative launch: one release-add per CTA on a counter in global memory
// and an acquire-poll until all have arrived (cooperative_groups' grid.sync() measured ~2.5x slower here).  `target`
// advances by gridDim.x per barrier; the counter is zeroed by the launcher.
__device__ __forceinline__ void coop_barrier(unsigned* counter, unsigned& target) {
    target += gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned seen;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

template <int H, int W>
__global__ void __launch_bounds__(256, 2)
k_coop_solve(CoopArgs a) {
    extern __shared__ float2 smem[];
    unsigned bar_target = 0;
    constexpr int Wc = W / 2;
    const int P = a.P;
    RowArgs ra; ColArgs ca;
    // zero-initialise without memset (device): every field is assigned below or cleared here
    ra.real_in = nullptr; ra.real_out = nullptr; ra.spec_in = nullptr; ra.spec_out = nullptr;
    ra.qx_in = ra.qy_in = nullptr; ra.qx_out = ra.qy_out = nullptr;
    ra.lmbd = a.lmbd; ra.rho = a.rho; ra.bias = nullptr; ra.real_in_u8 = nullptr; ra.act = 0; ra.out_C = 1; ra.out_bstride = 0;
    ra.out_p0 = 0; ra.cmap = nullptr; ra.r2c_div = 0; ra.ubx_in = ra.uby_in = nullptr; ra.ubx_out = ra.uby_out = nullptr;
    ra.taubar = nullptr; ra.qvx = ra.qvy = nullptr; ra.spec_out2 = nullptr; ra.tw = a.twW; ra.tiled = 0;
    ca.spec_in = a.S1; ca.spec_out = a.S0; ca.A = a.A; ca.Bm = a.Bm; ca.Bq = a.Bq; ca.Bmt = nullptr; ca.Mul = nullptr;
    ca.Mq = nullptr; ca.tw = a.twH; ca.in_tiled = 0; ca.out_tiled = 0;

    for (int it = 1; it < a.maxit; ++it) {
        float* qx_new = a.q[it & 1][0]; float* qy_new = a.q[it & 1][1];
        const float* qx_prev = it > 1 ? a.q[(it - 1) & 1][0] : nullptr;
        const float* qy_prev = it > 1 ? a.q[(it - 1) & 1][1] : nullptr;
        if (!a.iso) {
            ra.spec_in = a.S0; ra.spec_out = a.S1;
            ra.qx_in = qx_prev; ra.qy_in = qy_prev; ra.qx_out = qx_new; ra.qy_out = qy_new;
            for (unsigned vb = blockIdx.x; vb < (unsigned)(a.nb_full * P); vb += gridDim.x) {
                rows_pow2_body<W, ROWS_FULL_U, true>(ra, H, a.nb_full, 0, vb, smem);
                __syncthreads();
            }
            coop_barrier(a.barrier, bar_target);
        } else {
            // x_k as a real field
            ra.spec_in = a.S0; ra.real_out = a.xreal; ra.r2c_div = 0;
            for (unsigned vb = blockIdx.x; vb < (unsigned)(a.nb_plain * P); vb += gridDim.x) {
                rows_pow2_body<W, ROWS_C2R, true>(ra, H, a.nb_plain, 0, vb, smem);
                __syncthreads();
            }
            coop_barrier(a.barrier, bar_target);
            // block threshold: one thread per pixel walks the planes (iso.cu:k_iso_prox; deconv.py:19-24, 108-115)
            {
                const float tau = a.lmbd[0] / a.rho[0];
                const size_t HW = (size_t)H * W;
                const float* n_prev = it > 1 ? a.nmap[(it - 1) & 1] : nullptr;
                float* n_new = a.nmap[it & 1];
                for (int idx = blockIdx.x * 256 + threadIdx.x; idx < H * W; idx += gridDim.x * 256) {
                    const int r = idx / W, c = idx - r * W;
                    const int cl = c == 0 ? W - 1 : c - 1, ru = r == 0 ? H - 1 : r - 1;
                    float ux_scale = 0.f, uy_scale = 0.f;
                    if (n_prev) {
                        ux_scale = 1.f - coop_iso_scale(n_prev[idx], tau);
                        uy_scale = 1.f - coop_iso_scale(n_prev[HW + idx], tau);
                    }
                    float sx = 0.f, sy = 0.f;
                    for (int p = 0; p < P; ++p) {
                        const float* X = a.xreal + p * HW;
                        const float xc = X[idx];
                        float qx = xc - X[(size_t)r * W + cl];
                        float qy = xc - X[(size_t)ru * W + c];
                        if (n_prev) {
                            qx += ux_scale * qx_prev[p * HW + idx];
                            qy += uy_scale * qy_prev[p * HW + idx];
                        }
                        qx_new[p * HW + idx] = qx; qy_new[p * HW + idx] = qy;
                        sx = fmaf(qx, qx, sx); sy = fmaf(qy, qy, sy);
                    }
                    const float nx = sqrtf(sx + 1e-15f), ny = sqrtf(sy + 1e-15f);
                    n_new[idx] = nx; n_new[HW + idx] = ny;
                    a.sbmap[idx] = 2.f * coop_iso_scale(nx, tau) - 1.f;
                    a.sbmap[HW + idx] = 2.f * coop_iso_scale(ny, tau) - 1.f;
                }
            }
            coop_barrier(a.barrier, bar_target);
            // v = D^T((2s-1) q) formed while loading, R2C rows
            ra.r2c_div = 1; ra.cmap = a.sbmap; ra.qx_in = qx_new; ra.qy_in = qy_new; ra.spec_out = a.S1;
            for (unsigned vb = blockIdx.x; vb < (unsigned)(a.nb_plain * P); vb += gridDim.x) {
                rows_pow2_body<W, ROWS_R2C, true>(ra, H, a.nb_plain, 0, vb, smem);
                __syncthreads();
            }
            coop_barrier(a.barrier, bar_target);
        }
        for (unsigned vb = blockIdx.x; vb < (unsigned)(a.ntiles * P); vb += gridDim.x) {
            cols_pow2_body<H, COLS_ITER, 256, true>(ca, Wc, a.ntiles, 0, vb, reinterpret_cast<float4*>(smem));
            __syncthreads();
        }
        coop_barrier(a.barrier, bar_target);
    }
}

template <int H, int W>
static int launch_coop_t(const Geometry& g, CoopArgs& a, cudaStream_t st) {
    using CC = ColCfg<H, 256>;
    using RS = RowSmem<W>;
    const size_t smem = std::max((size_t)RS::bytes, (size_t)CC::bytes);
    static std::atomic<int> slots_dev[64];
    int dev = 0;
    ADMM_CUDA_CHECK(cudaGetDevice(&dev));
    int slots = slots_dev[dev & 63].load();
    if (slots == 0) {
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_coop_solve<H, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        ADMM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_coop_solve<H, W>, 256, smem));
        ADMM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        slots = std::max(1, std::min(per_sm, 2) * sms);
        slots_dev[dev & 63].store(slots);
    }
    // virtual grids: about one virtual block per resident CTA in every phase
    const int hh = g.H / 2;
    const int rmax_full = RS::RMAX, rmax_plain = 2 * RS::NPAIR;
    auto bands = [&](int rmax) {
        int nb = (g.H + rmax - 1) / rmax;
        nb = std::max(nb, std::min(hh, (slots + g.P - 1) / g.P));
        return std::min(nb, hh);
    };
    a.nb_full = bands(rmax_full);
    a.nb_plain = bands(rmax_plain);
    a.ntiles = g.Wc / CC::T;
    const int need = std::max(std::max(a.nb_full, a.nb_plain) * g.P, a.ntiles * g.P);
    const int grid = std::min(slots, need);
    ADMM_CUDA_CHECK(cudaMemsetAsync(a.barrier, 0, sizeof(unsigned), st));
    void* params[] = {(void*)&a};
    ProfScope ps(PROF_OTHER, st);
    ADMM_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_coop_solve<H, W>, dim3(grid), dim3(256), params, smem, st));
    return 0;
}

// sizes: both axes on the power-of-two kernels, column tiles of the 256-thread configuration divide the packed width
bool coop_solver_supported(const Geometry& g) {
    if (!options().use_coop || options().force_generic) return false;
    if (!rows_pow2_supported(g)) return false;
    int T = 0;
    switch (g.H) {
        case 128: T = ColCfg<128, 256>::T; break;
        case 256: T = ColCfg<256, 256>::T; break;
        case 512: T = ColCfg<512, 256>::T; break;
        default: return false;
    }
    return g.W % 2 == 0 && g.Wc % T == 0;
}

// the stand-alone kernels win once a phase is several waves: the cooperative kernel keeps 2 CTAs per SM
bool coop_solver_preferred(const Geometry& g) {
    const int want = options().use_coop;             // 2 = always (tests)
    if (want >= 2) return true;
    return (size_t)g.P * g.H * g.W <= ((size_t)options().coop_max_melems << 20);
}

#define ADMM_COOP_CASE(HH, WW) if (g.H == HH && g.W == WW) return launch_coop_t<HH, WW>(g, a, st)
int launch_coop_iterations(const Geometry& g, const Workspace& ws, const float* lmbd, const float* rho, int maxit, cudaStream_t st) {
    CoopArgs a;
    a.twW = ws.twW; a.lmbd = lmbd; a.rho = rho;
    for (int i = 0; i < 2; ++i) for (int f = 0; f < 2; ++f) a.q[i][f] = ws.q[i][f];
    a.twH = ws.twH; a.A = ws.A; a.Bm = ws.Bm; a.Bq = ws.Bq; a.S0 = ws.S0; a.S1 = ws.S1;
    a.xreal = ws.xreal; a.nmap[0] = ws.nmap[0]; a.nmap[1] = ws.nmap[1]; a.sbmap = ws.sbmap;
    a.P = g.P; a.H = g.H; a.maxit = maxit; a.iso = g.iso;
    a.nb_full = a.nb_plain = a.ntiles = 0;
    a.barrier = reinterpret_cast<unsigned*>(ws.red);
    ADMM_COOP_CASE(128, 128); ADMM_COOP_CASE(128, 256); ADMM_COOP_CASE(128, 512);
    ADMM_COOP_CASE(256, 128); ADMM_COOP_CASE(256, 256); ADMM_COOP_CASE(256, 512);
    ADMM_COOP_CASE(512, 128); ADMM_COOP_CASE(512, 256); ADMM_COOP_CASE(512, 512);
    return fail(4, "cooperative solver: unsupported size");
}
#undef ADMM_COOP_CASE

}  // namespace admm
