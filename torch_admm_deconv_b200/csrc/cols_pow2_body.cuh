// cols_pow2_body.cuh -- body of the specialised column-pass kernel (H in {128, 256, 512}); see cols_pow2.cu for the design notes.
// Included by cols_pow2.cu (the __global__ kernel k_cols_pow2 and its launchers) and by coop_small.cu.
#pragma once
#include <cstdio>
#include "cols_common.cuh"

namespace admm {

// MODE: COLS_ITER      spec -> FFT -> A + Bm Z -> iFFT -> spec          (one ADMM iteration; A is added after inverse pass 0)
//       COLS_INIT      spec -> FFT -> Mul Z -> iFFT -> spec   (x_1 = F^-1[A]; P0[Mul Z] is stored to `A` if given)
//       COLS_FFT_FWD   spec -> FFT -> full spectrum
//       COLS_BM_INV    full spectrum -> Bm Z -> iFFT -> spec                (backward: vbar = F^-1[Bm G])
//       COLS_CMUL_INV  full spectrum -> conj(Mul) Z -> iFFT -> spec         (backward: ybar)
//       COLS_INIT_SPEC full spectrum -> Mul Z -> iFFT -> spec  (COLS_INIT from a shared F(y); stores `A` the same way)
// -DCOLS_STATS prints per-phase clock cycles of a few CTAs (development instrumentation, tools/README.md).
// NT: threads per CTA (the kernel: col_threads<H>(); the cooperative small-batch kernel: 256).  COOP: the body is one phase of a
// persistent cooperative kernel (coop_small.cu): the input spectrum was written by an earlier phase of the same launch and
// is read with plain (coherent) loads; `bid` replaces the block index; shared memory is handed in.
// WCT: the packed width Wc as a compile-time constant (0 = run-time value): every global row stride becomes an immediate
template <int H, int MODE, int NT, bool COOP, int WCT = 0>
__device__ __forceinline__ void cols_pow2_body(const ColArgs& a, int Wc_dyn, int ntiles_dyn, int pdl, unsigned bid, float4* smem4) {
    using C = ColCfg<H, NT>;
    const int Wc = WCT ? WCT : Wc_dyn;
    const int ntiles = WCT ? WCT / C::T : ntiles_dyn;
    auto ld_in = [](const float4* p_) { return COOP ? *p_ : __ldg(p_); };
    using CR = ColRadix<H>;
    constexpr int TPS = C::TPS, T = C::T, NPAIRS = C::NPAIRS;
    constexpr bool kFwd = (MODE == COLS_ITER || MODE == COLS_INIT || MODE == COLS_FFT_FWD);
    constexpr bool kInv = (MODE != COLS_FFT_FWD);
    constexpr int NB2 = kCP / CR::F2;
    float4* buf = smem4;                                              // H * NPAIRS words
    float2* tabs = reinterpret_cast<float2*>(buf + H * NPAIRS);
    float2* zcol = tabs + C::TAB_END;                                 // H: packed column 0 for the mirrored term
    const int tid = threadIdx.x;
    const int pr = tid % NPAIRS;                                      // column pair inside the tile
    const int t = tid / NPAIRS;                                       // 0 .. TPS-1
    const int tile = bid % ntiles;
    const int p = bid / ntiles;
    const int c = tile * T + 2 * pr;                                  // first of this thread's two columns
    const size_t plane = (size_t)p * H * Wc;

    const bool early_tabs = (MODE == COLS_ITER) && pdl > 0;
    if (early_tabs) {
        // launched with programmatic stream serialisation (small, latency-bound problems): the tables are built
        // while the previous kernel drains; nothing the previous kernel wrote is touched before pdl_wait()
        pdl_launch_dependents();
        build_tab1<H, CR::F1, CR::F0>(tabs + C::TAB_F1, a.tw);
        build_tab1<H, CR::F2, CR::F0 * CR::F1>(tabs + C::TAB_F2, a.tw);
        if (!C::kShare) {
            build_tab1<H, CR::F1, CR::F2>(tabs + C::TAB_I1, a.tw);
            build_tab1<H, CR::F0, CR::F2 * CR::F1>(tabs + C::TAB_I2, a.tw);
        }
        pdl_wait();
    }
#ifdef COLS_STATS
    long long st[8]; st[0] = clock64();
#define COLS_ST(i) st[i] = clock64()
#else
#define COLS_ST(i)
#endif
    float4 d[kCP];
    {
        const float2* in = a.spec_in + plane + c;
        if (kFwd) {
            // forward pass 0 straight from global memory: slot q <-> row u = t + q*TPS, columns c, c+1 (16 bytes)
#pragma unroll
            for (int q = 0; q < kCP; ++q) d[q] = ld_in(reinterpret_cast<const float4*>(in + (size_t)(t + q * TPS) * Wc));
        } else {
            // the input already is a full spectrum: load it in the register layout of the last forward pass
#pragma unroll
            for (int m = 0; m < NB2; ++m)
#pragma unroll
                for (int r = 0; r < CR::F2; ++r)
                    d[m + r * NB2] = ld_in(reinterpret_cast<const float4*>(in + (size_t)((t + m * TPS) + r * (H / CR::F2)) * Wc));
        }
    }
    if (MODE == COLS_ITER) {
        // pull this tile of A into L2 now; it is consumed by the spectral update after the forward FFT
        const float2* Ag = a.A + plane + tile * T;
        for (int u = tid; u < H; u += C::kThreads) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Ag + (size_t)u * Wc));
        }
        // ... and (option cols_prefetch, default off: measured no gain) the input tile of the CTA that will take this SM slot
        // next (block index + the number of resident CTAs).  A tile prefetched a whole CTA lifetime ahead does not survive in
        // L2 at this traffic level: the same prefetch of the A tile costs +20 % (it is fetched from DRAM twice)
        if (pdl < 0) {
            const unsigned nb = bid + (unsigned)(-pdl);
            if (!COOP && nb < gridDim.x) {
                const float2* Sn = a.spec_in + (size_t)(nb / ntiles) * H * Wc + (nb % ntiles) * T;
                for (int u = tid; u < H; u += C::kThreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(Sn + (size_t)u * Wc));
            }
        }
    }
    if (!early_tabs) {
        build_tab1<H, CR::F1, CR::F0>(tabs + C::TAB_F1, a.tw);
        build_tab1<H, CR::F2, CR::F0 * CR::F1>(tabs + C::TAB_F2, a.tw);
        if (!C::kShare) {
            build_tab1<H, CR::F1, CR::F2>(tabs + C::TAB_I1, a.tw);
            build_tab1<H, CR::F0, CR::F2 * CR::F1>(tabs + C::TAB_I2, a.tw);
        }
    }
    if (kFwd) {
        cpass_compute<H, CR::F0, 1, -1>(d, t, nullptr);
        cpass_store<H, CR::F0, 1, NPAIRS>(d, t, pr, buf);
        __syncthreads();
        COLS_ST(1);
        cpass_load<H, NPAIRS>(d, t, pr, buf);
        cpass_compute<H, CR::F1, CR::F0, -1>(d, t, tabs + C::TAB_F1);
        __syncthreads();
        cpass_store<H, CR::F1, CR::F0, NPAIRS>(d, t, pr, buf);
        __syncthreads();
        cpass_load<H, NPAIRS>(d, t, pr, buf);
    }
    // Position (row of the tile buffer) of output r of this thread's inverse-pass-0 butterfly: inverse pass 0 is radix 8
    // without twiddles (Ns = 1) for every H, one butterfly per thread: outputs land at rows 8 t + r.
    static_assert(CR::F2 == kCP, "inverse pass 0: one radix-8 butterfly per thread");
    if (MODE == COLS_ITER) {
        // The constant term A enters AFTER inverse pass 0 (linearity: P0[A + Bm Z] = P0[A] + P0[Bm Z]).  The array `A` holds
        // A' = P0[A] in the positions inverse pass 0 writes (stored that way by COLS_INIT), so its tile can be copied
        // straight into the tile buffer with cp.async as soon as the last forward exchange has been read, and every thread
        // later adds its butterfly's outputs to the eight words it copied itself (no barrier for the copy).  The copy runs
        // behind the last forward pass, the spectral update and inverse pass 0 -- the latency of this 1/3 of the kernel's
        // traffic used to be exposed in the middle of the chain (measured: 31-44 us of a 157 us launch at cfg2).
        __syncthreads();                                   // every thread has taken its values of the last forward exchange
        COLS_ST(2);
        const float2* Ap = a.A + plane + c + (size_t)(kCP * t) * Wc;
#pragma unroll
        for (int r = 0; r < kCP; ++r) cp_async16(buf + (kCP * t + r) * NPAIRS + pr, Ap + (size_t)r * Wc);
        cp_async_commit();
    }
    if (kFwd) cpass_compute<H, CR::F2, CR::F0 * CR::F1, -1>(d, t, tabs + C::TAB_F2);
    // d[m + r*NB2] = Z[u], u = (t + m*TPS) + r*(H/F2): exactly the input layout of the first inverse pass
    if (MODE == COLS_FFT_FWD) {
        float2* out = a.spec_out + plane + c;
#pragma unroll
        for (int m = 0; m < NB2; ++m)
#pragma unroll
            for (int r = 0; r < CR::F2; ++r)
                *reinterpret_cast<float4*>(out + (size_t)((t + m * TPS) + r * (H / CR::F2)) * Wc) = d[m + r * NB2];
        return;
    }

    // spectral update on packed columns; packed column 0 carries DC and Nyquist and needs the mirrored entry Z[-u]
    {
        float2 bmv[kCP];
        if (MODE == COLS_ITER || MODE == COLS_BM_INV) {
            const float* __restrict__ Bp = a.Bm + c;
#pragma unroll
            for (int m = 0; m < NB2; ++m)
#pragma unroll
                for (int r = 0; r < CR::F2; ++r)
                    bmv[m + r * NB2] = __ldg(reinterpret_cast<const float2*>(Bp + (size_t)((t + m * TPS) + r * (H / CR::F2)) * Wc));
        }
        if (tile == 0) {                                   // CTA-uniform
            if (pr == 0) {
#pragma unroll
                for (int m = 0; m < NB2; ++m)
#pragma unroll
                    for (int r = 0; r < CR::F2; ++r)
                        zcol[(t + m * TPS) + r * (H / CR::F2)] = make_float2(d[m + r * NB2].x, d[m + r * NB2].y);
            }
        }
        // zcol visible; for the modes without the cp.async copy also: all reads of buf done before inverse pass 0 stores
        if (MODE != COLS_ITER || tile == 0) __syncthreads();
#pragma unroll
        for (int m = 0; m < NB2; ++m) {
#pragma unroll
            for (int r = 0; r < CR::F2; ++r) {
                const int u = (t + m * TPS) + r * (H / CR::F2);
                const float4 Z = d[m + r * NB2];
                float4 o;
                if (MODE == COLS_ITER || MODE == COLS_BM_INV) {
                    // X = A + Bm Z   (deconv.py:104-106 with freq_c, rho folded into A and Bm); here the Bm Z part
                    const float2 bm = bmv[m + r * NB2];
                    o = make_float4(bm.x * Z.x, bm.x * Z.y, bm.y * Z.z, bm.y * Z.w);
                    if (tile == 0 && pr == 0) {
                        const float2 Zm = zcol[(H - u) & (H - 1)];
                        const float bq = a.Bq[u];
                        o.x = fmaf(bq, Zm.x, o.x);
                        o.y = fmaf(-bq, Zm.y, o.y);
                    }
                } else {
                    // INIT: A = Mul Z (freq_c * rfftn(H_t(xin)), deconv.py:57,99,104); CMUL_INV: adjoint, conj(Mul) Z
                    float4 M = __ldg(reinterpret_cast<const float4*>(a.Mul + c + (size_t)u * Wc));
                    if (MODE == COLS_CMUL_INV) { M.y = -M.y; M.w = -M.w; }
                    const float2 o0 = cmul(make_float2(M.x, M.y), make_float2(Z.x, Z.y));
                    const float2 o1 = cmul(make_float2(M.z, M.w), make_float2(Z.z, Z.w));
                    o = make_float4(o0.x, o0.y, o1.x, o1.y);
                    if (tile == 0 && pr == 0) {
                        const float2 Zm = zcol[(H - u) & (H - 1)];
                        float2 mq = a.Mq[u];
                        if (MODE == COLS_CMUL_INV) mq.y = -mq.y;
                        const float2 e = cmul(mq, cconj(Zm));
                        o.x += e.x; o.y += e.y;
                    }
                }
                d[m + r * NB2] = o;
            }
        }
    }
    if (!kInv) return;
    // inverse pass 0 (radix 8, no twiddles) from registers
    cpass_compute<H, CR::F2, 1, +1>(d, t, nullptr);
    if ((MODE == COLS_INIT || MODE == COLS_INIT_SPEC) && a.A) {
        // A' = P0[A]: the constant term of every later iteration, in the positions inverse pass 0 writes (see COLS_ITER)
        float2* Ap = a.A + plane + c + (size_t)(kCP * t) * Wc;
#pragma unroll
        for (int r = 0; r < kCP; ++r) *reinterpret_cast<float4*>(Ap + (size_t)r * Wc) = d[r];
    }
    COLS_ST(3);
    if (MODE == COLS_ITER) {
        cp_async_wait_all();                               // this thread's eight words of A' are in the buffer
#pragma unroll
        for (int r = 0; r < kCP; ++r) {
            float4* w = buf + (kCP * t + r) * NPAIRS + pr;
            const float4 av = *w;
            *w = make_float4(d[r].x + av.x, d[r].y + av.y, d[r].z + av.z, d[r].w + av.w);
        }
    } else {
        cpass_store<H, CR::F2, 1, NPAIRS>(d, t, pr, buf);
    }
    __syncthreads();
    COLS_ST(4);
    cpass_load<H, NPAIRS>(d, t, pr, buf);
    cpass_compute<H, CR::F1, CR::F2, +1>(d, t, tabs + C::TAB_I1);
    __syncthreads();
    cpass_store<H, CR::F1, CR::F2, NPAIRS>(d, t, pr, buf);
    __syncthreads();
    cpass_load<H, NPAIRS>(d, t, pr, buf);
    cpass_compute<H, CR::F0, CR::F2 * CR::F1, +1>(d, t, tabs + C::TAB_I2);
    COLS_ST(5);
    // natural order: slot (m, r) -> row u = (t + m*TPS) + r*(H/F0)
    {
        constexpr int NB = kCP / CR::F0;
        float2* out = a.spec_out + plane + c;
#pragma unroll
        for (int m = 0; m < NB; ++m)
#pragma unroll
            for (int r = 0; r < CR::F0; ++r)
                *reinterpret_cast<float4*>(out + (size_t)((t + m * TPS) + r * (H / CR::F0)) * Wc) = d[m + r * NB];
    }
#ifdef COLS_STATS
    COLS_ST(6);
    if (MODE == COLS_ITER && tid == 0 && (bid % 397) == 5)
        printf("cta %u: load+F0 %lld | F1,F2ld %lld | F2,upd,I0 %lld | A wait+add %lld | I1,I2 %lld | store %lld | total %lld\n", bid, st[1] - st[0],
               st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4], st[6] - st[5], st[6] - st[0]);
#endif
}


}  // namespace admm
