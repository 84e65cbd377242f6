// tma_probe.cu -- how fast does one SM pull 512 x 128-byte-row tiles (2 KB row pitch) through TMA, and with what latency?
// One CTA per SM, one producer thread, S ring slots of 64 KB; every tile is a different (plane, column tile) of a
// [planes x 512 x 256] complex64 array (the packed row spectrum of cfg2).  Reports GB/s and issue->complete latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int S, int ROWS_PER_BOX>
__global__ void __launch_bounds__(128, 1) k_probe(const __grid_constant__ CUtensorMap tmap, int ntiles, int nitems, int inner_floats,
                                                   unsigned long long* lat_sum, unsigned* work, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + S * 65536);
    const unsigned tile_bytes = 512u * inner_floats * 4u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    long long t_issue[S];
    int live[S];
    unsigned long long lsum = 0, n = 0;
    auto issue = [&](int s) {
        int item = (int)atomicAdd(work, 1u);
        if (item >= nitems) { live[s] = 0; return; }
        const int p = item / ntiles, tile = item % ntiles;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(&bars[s])), "r"(tile_bytes) : "memory");
        for (int r = 0; r < 512; r += ROWS_PER_BOX)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(sa(smem + s * 65536 + r * inner_floats * 4)), "l"(&tmap), "r"(tile * inner_floats), "r"(p * 512 + r), "r"(sa(&bars[s])) : "memory");
        t_issue[s] = clock64();
        live[s] = 1;
    };
    for (int s = 0; s < S; ++s) issue(s);
    unsigned phase[S];
    for (int s = 0; s < S; ++s) phase[s] = 0;
    for (int k = 0;; ++k) {
        const int s = k % S;
        if (!live[s]) break;
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(sa(&bars[s])), "r"(phase[s]) : "memory");
        phase[s] ^= 1;
        lsum += (unsigned long long)(clock64() - t_issue[s]); ++n;
        issue(s);
    }
    atomicAdd(lat_sum, lsum);
    atomicAdd(lat_sum + 1, n);
    if (lsum == 1) sink[0] = reinterpret_cast<float*>(smem)[5];
}

// plain-load reference: every thread of 2 CTAs x 512 threads per SM loads 8 x 16 bytes of a tile (the register-staged pattern)
__global__ void __launch_bounds__(512, 2) k_ldg(const float4* __restrict__ in, int ntiles, int Wc4, float* sink) {
    const int tile = blockIdx.x % ntiles, p = blockIdx.x / ntiles;
    const int pr = threadIdx.x % 8, t = threadIdx.x / 8;
    const float4* src = in + (size_t)p * 512 * Wc4 + tile * 8 + pr;
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldg(src + (size_t)(t + q * 64) * Wc4);
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += v[q].x + v[q].y + v[q].z + v[q].w;
    if (acc == 12345.678f) sink[0] = acc;
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int S, int RPB>
void run(PFN_enc enc, float* d, int P, int Wc, int inner_floats, unsigned long long* lat, unsigned* work, float* sink) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)2 * Wc, (cuuint64_t)P * 512};
    cuuint64_t strides[1] = {(cuuint64_t)2 * Wc * 4};
    cuuint32_t box[2] = {(cuuint32_t)inner_floats, (cuuint32_t)RPB};
    cuuint32_t es[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return; }
    const int ntiles = 2 * Wc / inner_floats, nitems = ntiles * P;
    const size_t smem = (size_t)S * 65536 + 64;
    cudaFuncSetAttribute(k_probe<S, RPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f; unsigned long long h[2] = {0, 0};
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemset(work, 0, 8); cudaMemset(lat, 0, 16);
        cudaEventRecord(e0);
        k_probe<S, RPB><<<148, 128, smem>>>(tm, ntiles, nitems, inner_floats, lat, work, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        cudaMemcpy(h, lat, 16, cudaMemcpyDeviceToHost);
    }
    cudaError_t err = cudaGetLastError();
    const double bytes = (double)nitems * 512 * inner_floats * 4;
    printf("TMA slots=%d rows/box=%d inner=%dB: %.3f ms  %.0f GB/s  mean issue->complete %.0f cycles  (%s)\n", S, RPB, inner_floats * 4, best,
           bytes / best / 1e6, (double)h[0] / (double)(h[1] ? h[1] : 1), cudaGetErrorString(err));
}

int main() {
    const int P = 192, Wc = 256;
    float* d; cudaMalloc(&d, (size_t)P * 512 * Wc * 8); cudaMemset(d, 0, (size_t)P * 512 * Wc * 8);
    float* flush; cudaMalloc(&flush, 512u << 20);
    unsigned long long* lat; cudaMalloc(&lat, 16); unsigned* work; cudaMalloc(&work, 8); float* sink; cudaMalloc(&sink, 4);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN_enc enc = (PFN_enc)fp;
    cudaMemset(flush, 1, 512u << 20);
    run<1, 256>(enc, d, P, Wc, 32, lat, work, sink);
    run<2, 256>(enc, d, P, Wc, 32, lat, work, sink);
    run<3, 256>(enc, d, P, Wc, 32, lat, work, sink);
    run<3, 64>(enc, d, P, Wc, 32, lat, work, sink);
    run<3, 256>(enc, d, P, Wc, 16, lat, work, sink);      // 64-byte rows (half the bytes per tile slot)
    run<3, 128>(enc, d, P, Wc, 64, lat, work, sink);      // 256-byte rows, 256 rows -> still 64 KB
    // register-staged reference
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemset(flush, rep, 512u << 20);
        cudaEventRecord(e0);
        k_ldg<<<P * 16, 512>>>((const float4*)d, 16, Wc / 2, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("LDG 2 CTAs x 512 thr x 8 x 16 B: %.3f ms  %.0f GB/s\n", best, (double)P * 512 * Wc * 8 / best / 1e6);
    return 0;
}
