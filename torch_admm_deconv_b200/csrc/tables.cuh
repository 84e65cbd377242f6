// tables.cuh -- per-frequency quantities of the solve, shared by the table builder and the backward.
#pragma once
#include <cuda_runtime.h>

namespace admm {

struct SpecEntry { double2 sg; double2 ph; double L; double den; };

// sigma (deconv.py:49), |delta|^2 (deconv.py:51-55), den = 1/freq_c (deconv.py:57) and the H_t phase
// e^{+2 pi i s (u/H + v/W)}, s = ceil((k-1)/2) (deconv.py:88-99), all in fp64.  G is the row DFT of the PSF.
__device__ __forceinline__ SpecEntry spec_entry(int u, int v, int H, int W, int ks, const double2* __restrict__ G,
                                                const double2* __restrict__ twHd, const double2* __restrict__ twWd,
                                                double rho) {
    const int Wh = W / 2 + 1;
    SpecEntry e;
    e.sg = make_double2(1.0, 0.0);
    e.ph = make_double2(1.0, 0.0);
    if (ks > 0) {
        double2 sg = make_double2(0.0, 0.0);
        for (int a = 0; a < ks; ++a) {
            const double2 g = G[a * Wh + v];
            const double2 w = twHd[(int)(((long long)u * a) % H)];
            sg.x += g.x * w.x - g.y * w.y;
            sg.y += g.x * w.y + g.y * w.x;
        }
        e.sg = sg;
        const int s = ks / 2;                        // ceil((k-1)/2) == floor(k/2)
        const double2 a1 = twHd[(int)(((long long)s * u) % H)];
        const double2 a2 = twWd[(int)(((long long)s * v) % W)];
        e.ph = make_double2(a1.x * a2.x - a1.y * a2.y, -(a1.x * a2.y + a1.y * a2.x));   // conj(a1 a2)
    }
    e.L = (2.0 - 2.0 * twHd[u].x) + (2.0 - 2.0 * twWd[v].x);
    e.den = e.sg.x * e.sg.x + e.sg.y * e.sg.y + rho * e.L;      // no epsilon, like the reference
    return e;
}

struct TabEntry { double bm; double2 mul; };

__device__ __forceinline__ TabEntry table_entry(int u, int v, int H, int W, int ks, const double2* G,
                                                const double2* twHd, const double2* twWd, double rho) {
    const SpecEntry s = spec_entry(u, v, H, W, ks, G, twHd, twWd, rho);
    const double inv = 1.0 / (s.den * (double)H * (double)W);
    TabEntry e;
    e.bm = rho * inv;
    e.mul = make_double2((s.sg.x * s.ph.x - s.sg.y * s.ph.y) * inv, (s.sg.x * s.ph.y + s.sg.y * s.ph.x) * inv);
    return e;
}

}  // namespace admm
