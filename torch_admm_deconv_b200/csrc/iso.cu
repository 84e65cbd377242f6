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
                           const float* __restrict__ rho, int P, int H, int W) {
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

// backward, pass 1: sb_f[pixel] = sum over planes (2 wbar_f - ubar_f) q_f
__global__ void k_iso_bwd_reduce(const float* __restrict__ vb, const float* __restrict__ ubx, const float* __restrict__ uby,
                                 const float* __restrict__ qx, const float* __restrict__ qy, float* __restrict__ sb,
                                 int P, int H, int W) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * W) return;
    const int r = idx / W, c = idx - r * W;
    const int cl = c == 0 ? W - 1 : c - 1, ru = r == 0 ? H - 1 : r - 1;
    const size_t HW = (size_t)H * W;
    float ax = 0.f, ay = 0.f;
    for (int p = 0; p < P; ++p) {
        const float* V = vb + p * HW;
        const float v0 = V[idx];
        const float wbx = v0 - V[(size_t)r * W + cl], wby = v0 - V[(size_t)ru * W + c];
        const float ux = ubx ? ubx[p * HW + idx] : 0.f, uy = uby ? uby[p * HW + idx] : 0.f;
        ax = fmaf(2.f * wbx - ux, qx[p * HW + idx], ax);
        ay = fmaf(2.f * wby - uy, qy[p * HW + idx], ay);
    }
    sb[idx] = ax; sb[HW + idx] = ay;
}

// backward, pass 2: qbar = (2s-1) wbar + (1-s) ubar + act sb tau / (n+eps)^2 q / n ;  taubar += act sb (-1/(n+eps))
__global__ void k_iso_bwd_apply(const float* __restrict__ vb, const float* __restrict__ ubx_in, const float* __restrict__ uby_in,
                                const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ nmap,
                                const float* __restrict__ sb, float* __restrict__ ubx_out, float* __restrict__ uby_out,
                                const float* __restrict__ lmbd, const float* __restrict__ rho, double* __restrict__ taubar,
                                int P, int H, int W) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    double tsum = 0.0;
    if (idx < H * W) {
        const int r = idx / W, c = idx - r * W;
        const int cl = c == 0 ? W - 1 : c - 1, ru = r == 0 ? H - 1 : r - 1;
        const size_t HW = (size_t)H * W;
        const float tau = lmbd[0] / rho[0];
        const float nx = nmap[idx], ny = nmap[HW + idx];
        const float sx = iso_scale(nx, tau), sy = iso_scale(ny, tau);
        const float kx = (sx > 0.f) ? sb[idx] * tau / ((nx + 1e-15f) * (nx + 1e-15f) * nx) : 0.f;
        const float ky = (sy > 0.f) ? sb[HW + idx] * tau / ((ny + 1e-15f) * (ny + 1e-15f) * ny) : 0.f;
        if (sx > 0.f) tsum -= (double)sb[idx] / (double)(nx + 1e-15f);
        if (sy > 0.f) tsum -= (double)sb[HW + idx] / (double)(ny + 1e-15f);
        for (int p = 0; p < P; ++p) {
            const float* V = vb + p * HW;
            const float v0 = V[idx];
            const float wbx = v0 - V[(size_t)r * W + cl], wby = v0 - V[(size_t)ru * W + c];
            const float ux = ubx_in ? ubx_in[p * HW + idx] : 0.f, uy = uby_in ? uby_in[p * HW + idx] : 0.f;
            ubx_out[p * HW + idx] = (2.f * sx - 1.f) * wbx + (1.f - sx) * ux + kx * qx[p * HW + idx];
            uby_out[p * HW + idx] = (2.f * sy - 1.f) * wby + (1.f - sy) * uy + ky * qy[p * HW + idx];
        }
    }
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) tsum += __shfl_down_sync(0xffffffffu, tsum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tsum;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x + 31) / 32) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0 && v != 0.0) atomicAdd(taubar, v);
    }
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
    k_iso_prox<<<(n + 63) / 64, 64, 0, st>>>(x, qx_prev, qy_prev, n_prev, qx_new, qy_new, n_new, c_new, lmbd, rho, g.P, g.H, g.W);
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
    ProfScope ps(PROF_OTHER, st);
    const int n = g.H * g.W;
    const size_t total = (size_t)g.P * g.H * g.W;
    k_iso_bwd_reduce<<<(n + 127) / 128, 128, 0, st>>>(vb, ubx_in, uby_in, qx, qy, sbmap, g.P, g.H, g.W);
    ADMM_CUDA_CHECK(cudaGetLastError());
    k_iso_bwd_apply<<<(n + 127) / 128, 128, 0, st>>>(vb, ubx_in, uby_in, qx, qy, nmap, sbmap, ubx_out, uby_out, lmbd, rho,
                                                     taubar, g.P, g.H, g.W);
    ADMM_CUDA_CHECK(cudaGetLastError());
    k_div_adjoint<<<ew_grid(total), 256, 0, st>>>(ubx_out, uby_out, xb, g.H, g.W, total);
    ADMM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace admm
