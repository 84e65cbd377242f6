"""numpy model of the PACKED data flow of the CUDA kernels  --  TEST INFRASTRUCTURE ONLY.

The CUDA path never materialises a (H, W/2+1) half spectrum.  It keeps, per plane, a
"row spectrum" of H x Wc complex numbers, Wc = ceil(W/2):

    P[r, 0]   = ( Re X_r[0] , Re X_r[W/2] )      DC and Nyquist bins of row r (both real);
                                                 imaginary slot is 0 when W is odd
    P[r, v]   = X_r[v]            1 <= v < Wc    X_r = 1-D DFT of row r along W

and runs the 2-D transform as a row pass (two real rows per complex FFT) and a column pass
(complex FFT along H per packed column).  Packed column 0 holds TWO real-input columns, so its
spectral update needs the Hermitian-mirrored entry (see `col_pass`).  This module restates that
data flow with numpy FFTs so that (a) the layout algebra is pinned against the reference
independently of any CUDA code, and (b) GPU tests can compare intermediate buffers.

It also contains `stockham_fft`, a scalar model of the mixed-radix Stockham pass indexing used
by the shared-memory FFT engine (csrc/fft_engine.cuh).

Follows /root/reference/src/admmtor/eops/deconv.py:46-57 (tables) and :103-115 (loop).
"""
from __future__ import annotations

import math

import numpy as np


def wc_of(W):
    return (W + 1) // 2


# ------------------------------------------------------------------------------- row passes
def rows_r2c(v):
    """real (..., H, W) -> packed row spectrum (..., H, Wc) complex."""
    W = v.shape[-1]
    X = np.fft.fft(v, axis=-1)
    Wc = wc_of(W)
    P = X[..., :Wc].copy()
    if W % 2 == 0:
        P[..., 0] = X[..., 0].real + 1j * X[..., W // 2].real
    else:
        P[..., 0] = X[..., 0].real
    return P


def rows_c2r(P, W):
    """packed row spectrum -> real rows, UNNORMALISED inverse (sum, no 1/W)."""
    Wc = wc_of(W)
    full = np.zeros(P.shape[:-1] + (W // 2 + 1,), dtype=np.complex128)
    full[..., :Wc] = P
    full[..., 0] = P[..., 0].real
    if W % 2 == 0:
        full[..., W // 2] = P[..., 0].imag
    return np.fft.irfft(full, n=W, axis=-1) * W


# ------------------------------------------------------------------------------- tables
def build_tables(H, W, k2d, rho):
    """Per-call tables shared by all planes (all pre-scaled by 1/(H W) so the two inverse
    passes can be unnormalised):

    Bm[u, c]   real   rho/den/(HW)              c >= 1 ordinary packed columns (v = c)
    Bp[u], Bq[u]      column 0: (Bm_DC +- Bm_Nyq)/2
    Mul[u, c]  cplx   sigma*ph/den/(HW)         (A = Mul * F(y))
    Mp[u], Mq[u]      column 0: (Mul_DC +- Mul_Nyq)/2
    """
    Wc = wc_of(W)
    u = np.arange(H, dtype=np.float64)[:, None]
    vv = np.arange(W // 2 + 1, dtype=np.float64)[None, :]
    L = (2 - 2 * np.cos(2 * np.pi * u / H)) + (2 - 2 * np.cos(2 * np.pi * vv / W))
    if k2d is None:
        sigma = np.ones((H, W // 2 + 1), dtype=np.complex128)
        ph = np.ones_like(sigma)
    else:
        k = k2d.shape[0]
        pad = np.zeros((H, W)); pad[:k, :k] = k2d
        sigma = np.fft.rfft2(pad)
        s = int(math.ceil((k - 1) / 2))
        ph = np.exp(2j * np.pi * s * (u / H + vv / W))
    den = np.abs(sigma) ** 2 + rho * L
    with np.errstate(divide="ignore", invalid="ignore"):
        Bm_full = rho / den / (H * W)
        Mul_full = sigma * ph / den / (H * W)
    Bm = Bm_full[:, :Wc].copy()
    Mul = Mul_full[:, :Wc].copy()
    if W % 2 == 0:
        bn, mn = Bm_full[:, W // 2], Mul_full[:, W // 2]
    else:  # no Nyquist column: mirror DC so that Bq = Mq = 0
        bn, mn = Bm_full[:, 0], Mul_full[:, 0]
    Bp = 0.5 * (Bm_full[:, 0] + bn); Bq = 0.5 * (Bm_full[:, 0] - bn)
    Mp = 0.5 * (Mul_full[:, 0] + mn); Mq = 0.5 * (Mul_full[:, 0] - mn)
    return dict(Bm=Bm, Mul=Mul, Bp=Bp, Bq=Bq, Mp=Mp, Mq=Mq)


def _mirror(Z):
    """Z[(H-u) % H] along axis -2."""
    return np.roll(Z[..., ::-1, :], 1, axis=-2)


# ------------------------------------------------------------------------------- column passes
def col_init(S1, T):
    """packed row spectrum of y -> (A packed, S0 = packed row spectrum of x_1 * 1)."""
    Z = np.fft.fft(S1, axis=-2)
    A = T["Mul"] * Z
    A[..., 0] = T["Mp"] * Z[..., 0] + T["Mq"] * np.conj(_mirror(Z)[..., 0])
    S0 = np.fft.ifft(A, axis=-2) * S1.shape[-2]
    return A, S0


def col_pass(S1, A, T):
    """one spectral x-update: S0 = iFFT_col( A + Bm * FFT_col(S1) ), column 0 mirrored."""
    Z = np.fft.fft(S1, axis=-2)
    X = A + T["Bm"] * Z
    X[..., 0] = A[..., 0] + T["Bp"] * Z[..., 0] + T["Bq"] * np.conj(_mirror(Z)[..., 0])
    return np.fft.ifft(X, axis=-2) * S1.shape[-2]


# ------------------------------------------------------------------------------- spatial step
def spatial_step(x, u_x, u_y, tau):
    """q = Dx + u ; u' = clamp(q) ; w = q - 2u' ; v = Dx^T w_x + Dy^T w_y  (aniso)."""
    q_x = x - np.roll(x, 1, -1) + u_x
    q_y = x - np.roll(x, 1, -2) + u_y
    dual = lambda q: q - np.sign(q) * np.maximum(np.abs(q) - tau, 0)      # clip(q, -tau, tau) for tau >= 0; sign(q) tau for tau < 0
    nu_x = dual(q_x); nu_y = dual(q_y)
    w_x = q_x - 2 * nu_x; w_y = q_y - 2 * nu_y
    v = (w_x - np.roll(w_x, -1, -1)) + (w_y - np.roll(w_y, -1, -2))
    return v, nu_x, nu_y


def admm_tv_packed(y, lam, rho, k2d, maxit):
    """Whole solve through the packed data flow (aniso).  Mirrors the kernel launch sequence:
    rows_r2c(y) -> col_init -> [rows_c2r -> spatial -> rows_r2c -> col_pass] x (N-1) -> rows_c2r."""
    y = np.asarray(y, dtype=np.float64)
    B, C, H, W = y.shape
    if maxit == 0:
        return np.zeros_like(y)
    T = build_tables(H, W, k2d, rho)
    tau = lam / rho
    A, S0 = col_init(rows_r2c(y), T)
    u_x = np.zeros_like(y); u_y = np.zeros_like(y)
    for _ in range(maxit - 1):
        x = rows_c2r(S0, W)
        v, u_x, u_y = spatial_step(x, u_x, u_y, tau)
        S0 = col_pass(rows_r2c(v), A, T)
    return rows_c2r(S0, W)


# ------------------------------------------------------------------------------- pair trick
def pair_r2c(a, b):
    """Two real rows through ONE complex FFT: z = a + i b;  returns packed spectra (Pa, Pb)."""
    W = a.shape[-1]
    Z = np.fft.fft(a + 1j * b)
    Wc = wc_of(W)
    Zm = np.conj(np.roll(Z[::-1], 1))          # conj(Z[(W - v) % W])
    Xa = 0.5 * (Z + Zm); Xb = -0.5j * (Z - Zm)
    Pa = Xa[:Wc].copy(); Pb = Xb[:Wc].copy()
    if W % 2 == 0:
        Pa[0] = Z[0].real + 1j * Z[W // 2].real
        Pb[0] = Z[0].imag + 1j * Z[W // 2].imag
    else:
        Pa[0] = Z[0].real; Pb[0] = Z[0].imag
    return Pa, Pb


def pair_c2r(Pa, Pb, W):
    """Inverse of pair_r2c (unnormalised): builds Z = Xa + i Xb on all W bins, one complex iFFT."""
    Wc = wc_of(W)
    Z = np.zeros(W, dtype=np.complex128)
    Z[0] = Pa[0].real + 1j * Pb[0].real
    if W % 2 == 0:
        Z[W // 2] = Pa[0].imag + 1j * Pb[0].imag
    for v in range(1, Wc):
        Z[v] = Pa[v] + 1j * Pb[v]
        Z[W - v] = np.conj(Pa[v]) + 1j * np.conj(Pb[v])
    z = np.fft.ifft(Z) * W
    return z.real, z.imag


# ------------------------------------------------------------------------------- Stockham model
def factorize(n, radices=(16, 15, 9, 8, 4, 2, 3, 5, 7)):
    """Radix schedule used by the host planner: greedy over the preferred radices, leftover
    primes become generic-radix passes."""
    out = []
    m = n
    for r in radices:
        while m % r == 0 and m > 1:
            out.append(r); m //= r
    p = 3
    while m > 1:
        while m % p == 0:
            out.append(p); m //= p
        p += 2
    return out


def stockham_fft(x, radices, sign=-1):
    """Scalar model of the pass indexing in fft_engine.cuh: autosort Stockham, natural order in
    and out.  For pass with radix R and Ns = product of previous radices, T = N/R:
        k = j % Ns;  v[r] = src[j + r T] * w_N^{k r N/(Ns R)};  DFT_R;  dst[(j-k) R + k + r Ns] = v[r]."""
    x = np.asarray(x, dtype=np.complex128)
    N = x.shape[0]
    assert int(np.prod(radices)) == N
    tw = np.exp(sign * 2j * np.pi * np.arange(N) / N)
    src = x.copy()
    Ns = 1
    for R in radices:
        T = N // R
        tws = N // (Ns * R)
        dst = np.empty_like(src)
        dftm = np.exp(sign * 2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R)
        for j in range(T):
            k = j % Ns
            v = np.array([src[j + r * T] * tw[(k * r * tws) % N] for r in range(R)])
            o = dftm @ v
            j0 = (j - k) * R + k
            for r in range(R):
                dst[j0 + r * Ns] = o[r]
        src = dst
        Ns *= R
    return src


# ------------------------------------------------------------------------------- tile-major spectra
K_SPEC_TILE = 8


def tile_major_index(u, c, H, kind):
    """Slot (in complex entries, inside one plane) of packed-spectrum entry (row u, packed column c) in the tile-major
    layout the two large-frame kernels exchange (csrc/common.cuh, kSpecTile): [c // 8][row pair][c % 8][row in pair].
    kind 'v' (row pass -> column pass) pairs rows (2k, 2k+1); kind 'x' (column pass -> row pass) pairs rows (2k-1, 2k),
    row H-1 pairing with row 0 -- the two rows the row kernel transforms as one complex FFT."""
    if kind == "v":
        pair, slot = u // 2, u % 2
    else:
        pair, slot = ((u + 1) // 2) % (H // 2), (u + 1) % 2
    return (((c // K_SPEC_TILE) * (H // 2) + pair) * K_SPEC_TILE + c % K_SPEC_TILE) * 2 + slot


def to_tile_major(P, kind):
    """Row-major packed spectrum (H, Wc) -> flat tile-major array (Wc a multiple of 8, H even)."""
    H, Wc = P.shape
    out = np.empty(H * Wc, dtype=P.dtype)
    for u in range(H):
        for c in range(Wc):
            out[tile_major_index(u, c, H, kind)] = P[u, c]
    return out
