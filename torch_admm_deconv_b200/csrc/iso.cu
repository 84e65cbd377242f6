// iso.cu -- spatial kernels of the iso=True (block threshold) mode, forward and backward.
//
// block_thresh (deconv.py:19-20) scales every plane of a pixel by s = max(1 - tau / (n + 1e-15), 0) with
// n = sqrt(sum over (batch, channel) of q^2 + 1e-15) (pixelnorm, deconv.py:23-24), separately for the x and the y
// gradient field.  One thread owns one pixel and walks the planes, so the reduction over planes is sequential,
// deterministic and coalesced along the row.  State carried between iterations: q (pre-prox) plus the two
// H x W norm maps; u = (1 - s) q and w = z - u = (2 s - 1) q are rebuilt from them.
#include "common.cuh"

namespace admm {

__device__ __forceinline__ float iso_scale(float n, float tau) { return fmaxf(1.f - tau / (n + 1e-15f), 0.f); }

__global__ void k_iso_prox(const float* __restrict__ x, const float* __restrict__ qxp, const float* __restrict__ qyp,
                           const float* __restrict__ n_prev, float* __restrict__ qxn, float* __restrict__ qyn,
                           float* __restrict__ n_new, float* __restrict__ c_new, const float* __restrict__ lmbd,
                           const float* __restrict__ rho, int P, int H, int W, int pdl) {
    // launched with programmatic stream serialisation for small (latency-bound) batches: the next kernel of the iteration
    // may start its prologue (twiddle tables) while this one runs; nothing the previous kernel wrote is read before the wait
    if (pdl) { pdl_launch_dependents(); pdl_wait(); }
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * W) return;
    const int r = idx / W, c = idx - r * W;
    const int cl = c == 0 ? W - 1 : c - 1, ru = r == 0 ? H - 1 : r - 1;
    const float tau = lmbd[0] / rho[0];
    float ux_scale = 0.f, uy_scale = 0.f;                     // u_prev = (1 - s_prev) q_prev
    if (n_prev) {
        ux_scale = 1.f - iso_scale(n_prev[idx], tau);
        uy_scale = 1.f - iso_scale(n_prev[(size_t)H * W + idx], tau);
    }
    float sx = 0.f, sy = 0.f;
    const size_t HW = (size_t)H * W;
    for (int p = 0; p < P; ++p) {
        const float* X = x + p * HW;
        const float xc = X[idx];
        float qx = xc - X[(size_t)r * W + cl];               // deconv.py:108
        float qy = xc - X[(size_t)ru * W + c];               // deconv.py:109
        if (n_prev) {
            qx += ux_scale * qxp[p * HW + idx];
            qy += uy_scale * qyp[p * HW + idx];
        }
        qxn[p * HW + idx] = qx; qyn[p * HW + idx] = qy;
        sx = fmaf(qx, qx, sx); sy = fmaf(qy, qy, sy);
    }
    const float nx = sqrtf(sx + 1e-15f), ny = sqrtf(sy + 1e-15f);   // deconv.py:23-24
    n_new[idx] = nx;
    n_new[HW + idx] = ny;
    if (c_new) {                                              // w = z - u = (2 s - 1) q : coefficient map for the divergence
        c_new[idx] = 2.f * iso_scale(nx, tau) - 1.f;
        c_new[HW + idx] = 2.f * iso_scale(ny, tau) - 1.f;
    }
}

// v = Dx^T w_x + Dy^T w_y,  w = z - u = (2 s - 1) q        (deconv.py:104 with z = s q, u = q - z)
// cmap (2s-1 per pixel and field) if given, else rebuilt from the norm maps.  grid = (W/256, H, planes)
__global__ void k_iso_div(const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ nmap,
                          const float* __restrict__ cmap, float* __restrict__ v, const float* __restrict__ lmbd,
                          const float* __restrict__ rho, int H, int W) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= W) return;
    const int r = blockIdx.y;
    const size_t HW = (size_t)H * W;
    const size_t pl = (size_t)blockIdx.z * HW;
    const int cr = c == W - 1 ? 0 : c + 1, rd = r == H - 1 ? 0 : r + 1;
    const size_t m00 = (size_t)r * W + c, m0r = (size_t)r * W + cr, md0 = (size_t)rd * W + c;
    float kx0, kxr, ky0, kyd;
    if (cmap) {
        kx0 = cmap[m00]; kxr = cmap[m0r]; ky0 = cmap[HW + m00]; kyd = cmap[HW + md0];
    } else {
        const float tau = lmbd[0] / rho[0];
        kx0 = 2.f * iso_scale(nmap[m00], tau) - 1.f; kxr = 2.f * iso_scale(nmap[m0r], tau) - 1.f;
        ky0 = 2.f * iso_scale(nmap[HW + m00], tau) - 1.f; kyd = 2.f * iso_scale(nmap[HW + md0], tau) - 1.f;
    }
    v[pl + m00] = (kx0 * qx[pl + m00] - kxr * qx[pl + m0r]) + (ky0 * qy[pl + m00] - kyd * qy[pl + md0]);
}

// backward of the block threshold, one kernel:
//   sb_f[pixel] = sum over planes (2 wbar_f - ubar_f) q_f                                   (reduction over planes)
//   qbar = (2s-1) wbar + (1-s) ubar + act sb tau / (n+eps)^2 q / n ;  taubar += act sb (-1/(n+eps))
// A block owns 32 consecutive pixels; its kIsoPG warps split the planes (plane p -> warp p % kIsoPG) and stream over them
// (unrolled, loads of several planes in flight).  The partial sums meet in shared memory and are added in a fixed order
// (deterministic); then every warp sweeps its planes again to apply the result (vbar / ubar / q are re-read; parking
// the first sweep's terms in shared memory or registers was measured and is not faster: the kernel is bound by the
// latency of its many concurrent plane streams, ~3 TB/s).
constexpr int kIsoPG = 4;
template <bool HAS_U>
__global__ void __launch_bounds__(32 * kIsoPG)
k_iso_bwd_fused(const float* __restrict__ vb, const float* __restrict__ ubx_in, const float* __restrict__ uby_in,
                const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ nmap,
                float* __restrict__ ubx_out, float* __restrict__ uby_out, const float* __restrict__ lmbd,
                const float* __restrict__ rho, double* __restrict__ taubar, int P, int H, int W, int pdl) {
    if (pdl) { pdl_launch_dependents(); pdl_wait(); }         // see k_iso_prox
    __shared__ float red[kIsoPG][2][32];
    const int tx = threadIdx.x & 31, g = threadIdx.x >> 5;
    const size_t HW = (size_t)H * W;
    const size_t idx = (size_t)blockIdx.x * 32 + tx;
    const bool ok = idx < HW;
    size_t ol = 0, ou = 0;
    if (ok) {
        const int r = (int)(idx / W), c = (int)(idx - (size_t)r * W);
        ol = (size_t)r * W + (c == 0 ? W - 1 : c - 1);
        ou = (size_t)(r == 0 ? H - 1 : r - 1) * W + c;
    }
    float ax = 0.f, ay = 0.f;
    if (ok) {
#pragma unroll 4
        for (int p = g; p < P; p += kIsoPG) {
            const float* V = vb + p * HW;
            const float v0 = V[idx];
            const float wbx = v0 - V[ol], wby = v0 - V[ou];
            const float ux = HAS_U ? ubx_in[p * HW + idx] : 0.f, uy = HAS_U ? uby_in[p * HW + idx] : 0.f;
            const float qxv = qx[p * HW + idx], qyv = qy[p * HW + idx];
            const float tX = 2.f * wbx - ux, tY = 2.f * wby - uy;
            ax = fmaf(tX, qxv, ax);
            ay = fmaf(tY, qyv, ay);
        }
    }
    red[g][0][tx] = ax; red[g][1][tx] = ay;
    __syncthreads();
    double tsum = 0.0;
    if (ok) {
        float sbx = 0.f, sby = 0.f;
#pragma unroll
        for (int i = 0; i < kIsoPG; ++i) { sbx += red[i][0][tx]; sby += red[i][1][tx]; }
        const float tau = lmbd[0] / rho[0];
        const float nx = nmap[idx], ny = nmap[HW + idx];
        const float sx = iso_scale(nx, tau), sy = iso_scale(ny, tau);
        const float kx = (sx > 0.f) ? sbx * tau / ((nx + 1e-15f) * (nx + 1e-15f) * nx) : 0.f;
        const float ky = (sy > 0.f) ? sby * tau / ((ny + 1e-15f) * (ny + 1e-15f) * ny) : 0.f;
        if (g == 0) {
            if (sx > 0.f) tsum -= (double)sbx / (double)(nx + 1e-15f);
            if (sy > 0.f) tsum -= (double)sby / (double)(ny + 1e-15f);
        }
#pragma unroll 4
        for (int p = g; p < P; p += kIsoPG) {
            const float* V = vb + p * HW;
            const float v0 = V[idx];
            const float wbx = v0 - V[ol], wby = v0 - V[ou];
            const float ux = HAS_U ? ubx_in[p * HW + idx] : 0.f, uy = HAS_U ? uby_in[p * HW + idx] : 0.f;
            // (2s-1) wbar + (1-s) ubar + k q  =  s (2 wbar - ubar) + (ubar - wbar) + k q
            ubx_out[p * HW + idx] = fmaf(sx, 2.f * wbx - ux, ux - wbx) + kx * qx[p * HW + idx];
            uby_out[p * HW + idx] = fmaf(sy, 2.f * wby - uy, uy - wby) + ky * qy[p * HW + idx];
        }
    }
    if (g == 0) {                                              // warp 0 holds the tau-gradient terms of the 32 pixels
        for (int o = 16; o > 0; o >>= 1) tsum += __shfl_down_sync(0xffffffffu, tsum, o);
        if (tx == 0) taubar[blockIdx.x] += tsum;        // per-block slot, single writer (deterministic)
    }
}

// coefficient map 2s-1 of the divergence (both fields) rebuilt from saved norm maps (backward recompute of v_k)
__global__ void k_iso_cmap(const float* __restrict__ nmap, float* __restrict__ cmap, const float* __restrict__ lmbd,
                           const float* __restrict__ rho, int n2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    cmap[i] = 2.f * iso_scale(nmap[i], lmbd[0] / rho[0]) - 1.f;
}

int launch_iso_cmap(const Geometry& g, const float* nmap, float* cmap, const float* lmbd, const float* rho, cudaStream_t st) {
    ProfScope ps(PROF_OTHER, st);
    const int n2 = 2 * g.H * g.W;
    k_iso_cmap<<<(n2 + 255) / 256, 256, 0, st>>>(nmap, cmap, lmbd, rho, n2);
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// xbar = Dx^T a_x + Dy^T a_y
__global__ void k_div_adjoint(const float* __restrict__ ax, const float* __restrict__ ay, float* __restrict__ xb,
                              int H, int W, size_t total) {
    const size_t HW = (size_t)H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % W);
        const size_t rowi = i / W;
        const int r = (int)(rowi % H);
        const size_t pl = (rowi / H) * HW;
        const int cr = c == W - 1 ? 0 : c + 1, rd = r == H - 1 ? 0 : r + 1;
        const size_t m00 = pl + (size_t)r * W + c;
        xb[m00] = (ax[m00] - ax[pl + (size_t)r * W + cr]) + (ay[m00] - ay[pl + (size_t)rd * W + c]);
    }
}

static int ew_grid(size_t total) { return (int)std::min<size_t>((total + 255) / 256, 148 * 16); }

int launch_iso_prox(const Geometry& g, const float* x, const float* qx_prev, const float* qy_prev, const float* n_prev,
                    float* qx_new, float* qy_new, float* n_new, float* c_new, const float* lmbd, const float* rho,
                    cudaStream_t st) {
    ProfScope ps(PROF_OTHER, st);
    const int n = g.H * g.W;
    const unsigned nctas = (unsigned)((n + 63) / 64);
    if (options().use_pdl && (size_t)g.P * g.H * g.W <= (size_t)4 << 20) {        // a few waves per kernel: launch latency counts
        ADMM_CUDA_CHECK(launch_pdl(k_iso_prox, dim3(nctas), dim3(64), 0, st, x, qx_prev, qy_prev, n_prev, qx_new, qy_new, n_new, c_new,
                                   lmbd, rho, g.P, g.H, g.W, 1));
    } else {
        k_iso_prox<<<nctas, 64, 0, st>>>(x, qx_prev, qy_prev, n_prev, qx_new, qy_new, n_new, c_new, lmbd, rho, g.P, g.H, g.W, 0);
    }
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_iso_div(const Geometry& g, const float* qx, const float* qy, const float* nmap, const float* cmap, float* v,
                   const float* lmbd, const float* rho, cudaStream_t st) {
    ProfScope ps(PROF_OTHER, st);
    const int bx = g.W >= 256 ? 256 : (g.W >= 128 ? 128 : 64);
    const dim3 grid((g.W + bx - 1) / bx, g.H, g.P);
    if (g.H > 65535 || g.P > 65535) return fail(4, "iso: H and B*C must be <= 65535");
    k_iso_div<<<grid, bx, 0, st>>>(qx, qy, nmap, cmap, v, lmbd, rho, g.H, g.W);
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_iso_bwd(const Geometry& g, const float* vb, const float* ubx_in, const float* uby_in, const float* qx,
                   const float* qy, const float* nmap, float* sbmap, float* ubx_out, float* uby_out, float* xb,
                   const float* lmbd, const float* rho, double* taubar, cudaStream_t st) {
    (void)sbmap;                                               // the per-pixel sums stay in shared memory
    ProfScope ps(PROF_OTHER, st);
    const size_t total = (size_t)g.P * g.H * g.W;
    const size_t n = (size_t)g.H * g.W;
    const unsigned nb = (unsigned)((n + 31) / 32);
    const bool pdl = options().use_pdl && total <= ((size_t)4 << 20);
    const float* nullf = nullptr;
    if (ubx_in && uby_in) {
        if (pdl) ADMM_CUDA_CHECK(launch_pdl(k_iso_bwd_fused<true>, dim3(nb), dim3(32 * kIsoPG), 0, st, vb, ubx_in, uby_in, qx, qy, nmap,
                                            ubx_out, uby_out, lmbd, rho, taubar, g.P, g.H, g.W, 1));
        else k_iso_bwd_fused<true><<<nb, 32 * kIsoPG, 0, st>>>(vb, ubx_in, uby_in, qx, qy, nmap, ubx_out, uby_out, lmbd, rho, taubar,
                                                              g.P, g.H, g.W, 0);
    } else {
        if (pdl) ADMM_CUDA_CHECK(launch_pdl(k_iso_bwd_fused<false>, dim3(nb), dim3(32 * kIsoPG), 0, st, vb, nullf, nullf, qx, qy, nmap,
                                            ubx_out, uby_out, lmbd, rho, taubar, g.P, g.H, g.W, 1));
        else k_iso_bwd_fused<false><<<nb, 32 * kIsoPG, 0, st>>>(vb, nullptr, nullptr, qx, qy, nmap, ubx_out, uby_out, lmbd, rho,
                                                               taubar, g.P, g.H, g.W, 0);
    }
    ADMM_CUDA_CHECK(cudaGetLastError());
    if (xb) {                                                  // else the caller forms D^T qbar inside its R2C row pass
        k_div_adjoint<<<ew_grid(total), 256, 0, st>>>(ubx_out, uby_out, xb, g.H, g.W, total);
        ADMM_CUDA_CHECK(cudaGetLastError());
    }
    return 0;
}

}  // namespace admm
