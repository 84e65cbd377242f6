// cols_tma.cu -- persistent, TMA-fed column pass of the ADMM iteration (sm_100a), H in {128, 256, 512}.
//
// Same arithmetic as k_cols_pow2<H, COLS_ITER> (cols_pow2.cu), different data movement: the grid is sized to
// the machine (one wave of resident CTAs), every CTA walks a strided list of (plane, column tile) work items, and
// the packed-column tile of the NEXT item is fetched by the TMA engine (cp.async.bulk.tensor.2d -> shared memory,
// completion on an mbarrier) while the current item is being transformed.  The tile lands directly in the
// "column fastest" layout the Stockham passes use, so there is no register staging and no per-thread address
// arithmetic on the load path; twiddle tables are built once per CTA instead of once per tile.
#include <cuda.h>

#include "cols_common.cuh"

namespace admm {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a TMA that never completes traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int H> struct TmaCfg {
    using C = ColCfg<H, 256>;
    static constexpr int kBoxRows = (H > 256) ? 256 : H;              // TMA box dimension limit is 256
    static constexpr int kBoxes = H / kBoxRows;
    static constexpr unsigned kTileBytes = (unsigned)(H * C::T * sizeof(float2));
    // two tile buffers + tables + zcol + two mbarriers
    static constexpr size_t bytes = (size_t)(2 * H * C::T + C::TAB_END + H) * sizeof(float2) + 64;
};

template <int H>
__global__ void __launch_bounds__(256, 3)
k_cols_iter_tma(const __grid_constant__ CUtensorMap tmap, ColArgs a, int Wc, int ntiles, int nitems) {
    using C = ColCfg<H, 256>;
    using CR = ColRadix<H>;
    using TC = TmaCfg<H>;
    constexpr int TPS = C::TPS, T = C::T, NPAIRS = C::NPAIRS;
    constexpr int NB2 = kCP / CR::F2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* buf0 = reinterpret_cast<float4*>(smem_raw);                // H * NPAIRS words each
    float4* buf1 = buf0 + H * NPAIRS;
    float2* tabs = reinterpret_cast<float2*>(buf1 + H * NPAIRS);
    float2* zcol = tabs + C::TAB_END;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(zcol + H);
    const int tid = threadIdx.x;
    const int pr = tid % NPAIRS;
    const int t = tid / NPAIRS;

    auto issue = [&](int item, int slot) {
        // one elected thread: arm the barrier with the tile size, then one TMA box per 256 rows
        const int tile = item % ntiles, p = item / ntiles;
        float4* dst = slot ? buf1 : buf0;
        mbar_expect_tx(&bars[slot], TC::kTileBytes);
#pragma unroll
        for (int b = 0; b < TC::kBoxes; ++b)
            tma_load_2d(dst + b * TC::kBoxRows * NPAIRS, &tmap, tile * T * 2, p * H + b * TC::kBoxRows, &bars[slot]);
    };

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    build_tab<H, CR::F1, CR::F0>(tabs + C::TAB_F1, a.tw);
    build_tab<H, CR::F2, CR::F0 * CR::F1>(tabs + C::TAB_F2, a.tw);
    if (!C::kShare) {
        build_tab<H, CR::F1, CR::F2>(tabs + C::TAB_I1, a.tw);
        build_tab<H, CR::F0, CR::F2 * CR::F1>(tabs + C::TAB_I2, a.tw);
    }
    __syncthreads();
    int item = blockIdx.x;
    if (tid == 0 && item < nitems) issue(item, 0);
    if (item < nitems) {
        const float2* Ag = a.A + (size_t)(item / ntiles) * H * Wc + (item % ntiles) * T;
        for (int u = tid; u < H; u += 256) asm volatile("prefetch.global.L2 [%0];" ::"l"(Ag + (size_t)u * Wc));
    }

    unsigned phase0 = 0, phase1 = 0;
    int slot = 0;
    for (; item < nitems; item += gridDim.x, slot ^= 1) {
        const int tile = item % ntiles, p = item / ntiles;
        const int c = tile * T + 2 * pr;
        const size_t plane = (size_t)p * H * Wc;
        float4* buf = slot ? buf1 : buf0;
        // prefetch the next item's tile into the other buffer (free: its last reader passed the barrier that ends
        // the previous iteration) and its A tile into L2
        const int next = item + gridDim.x;
        if (tid == 0 && next < nitems) {
            fence_proxy_async();
            issue(next, slot ^ 1);
        }
        if (next < nitems) {                               // the next item's A tile -> L2
            const float2* Ag = a.A + (size_t)(next / ntiles) * H * Wc + (next % ntiles) * T;
            for (int u = tid; u < H; u += 256) asm volatile("prefetch.global.L2 [%0];" ::"l"(Ag + (size_t)u * Wc));
        }
        // wait for this item's tile
        if (slot == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1; } else { mbar_wait(&bars[1], phase1); phase1 ^= 1; }

        float4 d[kCP];
        cpass_load<H, NPAIRS>(d, t, pr, buf);
        cpass_compute<H, CR::F0, 1, -1>(d, t, nullptr);
        __syncthreads();
        cpass_store<H, CR::F0, 1, NPAIRS>(d, t, pr, buf);
        __syncthreads();
        cpass_load<H, NPAIRS>(d, t, pr, buf);
        cpass_compute<H, CR::F1, CR::F0, -1>(d, t, tabs + C::TAB_F1);
        __syncthreads();
        cpass_store<H, CR::F1, CR::F0, NPAIRS>(d, t, pr, buf);
        __syncthreads();
        cpass_load<H, NPAIRS>(d, t, pr, buf);
        cpass_compute<H, CR::F2, CR::F0 * CR::F1, -1>(d, t, tabs + C::TAB_F2);
        // spectral update X = A + Bm Z (deconv.py:104-106), packed column 0 with the mirrored entry
        {
            float2 bmv[kCP];
            const float* __restrict__ Bp = a.Bm + c;
#pragma unroll
            for (int m = 0; m < NB2; ++m)
#pragma unroll
                for (int r = 0; r < CR::F2; ++r)
                    bmv[m + r * NB2] = __ldg(reinterpret_cast<const float2*>(Bp + (size_t)((t + m * TPS) + r * (H / CR::F2)) * Wc));
            if (tile == 0 && pr == 0) {
#pragma unroll
                for (int m = 0; m < NB2; ++m)
#pragma unroll
                    for (int r = 0; r < CR::F2; ++r)
                        zcol[(t + m * TPS) + r * (H / CR::F2)] = make_float2(d[m + r * NB2].x, d[m + r * NB2].y);
            }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < NB2; ++m) {
#pragma unroll
                for (int r = 0; r < CR::F2; ++r) {
                    const int u = (t + m * TPS) + r * (H / CR::F2);
                    const float4 Z = d[m + r * NB2];
                    const float4 Av = __ldg(reinterpret_cast<const float4*>(a.A + plane + c + (size_t)u * Wc));
                    const float2 bm = bmv[m + r * NB2];
                    float4 o = make_float4(fmaf(bm.x, Z.x, Av.x), fmaf(bm.x, Z.y, Av.y), fmaf(bm.y, Z.z, Av.z), fmaf(bm.y, Z.w, Av.w));
                    if (tile == 0 && pr == 0) {
                        const float2 Zm = zcol[(H - u) & (H - 1)];
                        const float bq = a.Bq[u];
                        o.x = fmaf(bq, Zm.x, o.x);
                        o.y = fmaf(-bq, Zm.y, o.y);
                    }
                    d[m + r * NB2] = o;
                }
            }
        }
        cpass_compute<H, CR::F2, 1, +1>(d, t, nullptr);
        cpass_store<H, CR::F2, 1, NPAIRS>(d, t, pr, buf);
        __syncthreads();
        cpass_load<H, NPAIRS>(d, t, pr, buf);
        cpass_compute<H, CR::F1, CR::F2, +1>(d, t, tabs + C::TAB_I1);
        __syncthreads();
        cpass_store<H, CR::F1, CR::F2, NPAIRS>(d, t, pr, buf);
        __syncthreads();
        cpass_load<H, NPAIRS>(d, t, pr, buf);
        __syncthreads();                                   // buf is free: the next iteration may refill it by TMA
        cpass_compute<H, CR::F0, CR::F2 * CR::F1, +1>(d, t, tabs + C::TAB_I2);
        {
            constexpr int NB = kCP / CR::F0;
            float2* out = a.spec_out + plane + c;
#pragma unroll
            for (int m = 0; m < NB; ++m)
#pragma unroll
                for (int r = 0; r < CR::F0; ++r)
                    *reinterpret_cast<float4*>(out + (size_t)((t + m * TPS) + r * (H / CR::F0)) * Wc) = d[m + r * NB];
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    // function-local static: initialised exactly once, thread-safe (C++11)
    static const PFN_encodeTiled fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            return (PFN_encodeTiled)p;
        return (PFN_encodeTiled) nullptr;
    }();
    return fn;
}

template <int H>
static int launch_cols_tma_t(const Geometry& g, const ColArgs& a, cudaStream_t st) {
    using C = ColCfg<H, 256>;
    using TC = TmaCfg<H>;
    PFN_encodeTiled enc = get_encode();
    if (!enc) return fail(4, "cuTensorMapEncodeTiled is not available");
    // packed spectrum as a 2-D fp32 tensor: inner = 2*Wc floats of one image row, outer = all rows of all planes
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {(cuuint64_t)2 * g.Wc, (cuuint64_t)g.P * g.H};
    const cuuint64_t strides[1] = {(cuuint64_t)2 * g.Wc * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)(2 * C::T), (cuuint32_t)TC::kBoxRows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.spec_in, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(3, "cuTensorMapEncodeTiled failed");
    const int ntiles = g.Wc / C::T;
    const int nitems = ntiles * g.P;
    static std::atomic<int> ctas_per_sm_dev[64];
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    int ctas_per_sm = ctas_per_sm_dev[dev & 63].load();
    if (!ctas_per_sm) {
        ADMM_CUDA_CHECK(cudaFuncSetAttribute(k_cols_iter_tma<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::bytes));
        ADMM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_cols_iter_tma<H>, 256, TC::bytes));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        ctas_per_sm_dev[dev & 63].store(ctas_per_sm);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = std::min(nitems, sms * ctas_per_sm);
    ProfScope ps(PROF_COLS, st);
    k_cols_iter_tma<H><<<grid, 256, TC::bytes, st>>>(tmap, a, g.Wc, ntiles, nitems);
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

bool cols_tma_supported(const Geometry& g) {
    if (!options().use_tma) return false;
    if (options().force_generic) return false;
    if (g.H != 128 && g.H != 256 && g.H != 512) return false;
    if (g.W % 2 || g.Wc % ColCfg<128, 256>::T) return false;                 // widest tile (H = 128): 32 columns
    if (g.Wc % 2) return false;                                              // TMA global stride: multiple of 16 bytes
    return get_encode() != nullptr;
}

int launch_cols_tma(const Geometry& g, const ColArgs& a, cudaStream_t st) {
    if (((uintptr_t)a.spec_in & 15) != 0) return fail(2, "TMA source must be 16-byte aligned");
    switch (g.H) {
        case 128: return launch_cols_tma_t<128>(g, a, st);
        case 256: return launch_cols_tma_t<256>(g, a, st);
        case 512: return launch_cols_tma_t<512>(g, a, st);
    }
    return fail(4, "cols_tma: unsupported height");
}

}  // namespace admm
