"""CPU tests: the oracle (oracle/admm_oracle.py, oracle/packed_layout_model.py) against the golden
fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import golden, golden_names
from oracle import admm_oracle as O
from oracle import packed_layout_model as M

FWD = golden_names(exclude=("module", "grad"))
GRAD = golden_names(prefix="grad")


@pytest.mark.parametrize("name", FWD)
def test_spectral_form_matches_reference_fp64(name):
    d = golden(name)
    out = O.admm_tv_spectral_form(d["x"].astype(np.float64), d["lam"], d["rho"], d["kern"], bool(d["iso"]), int(d["maxit"]))
    assert O.rel_err(out, d["out64"]) < 1e-11


@pytest.mark.parametrize("name", [n for n in FWD if "cfg" not in n])
def test_stencil_form_matches_reference_fp64(name):
    d = golden(name)
    out = O.admm_tv_stencil_form(d["x"].astype(np.float64), d["lam"], d["rho"], d["kern"], bool(d["iso"]), int(d["maxit"]))
    assert O.rel_err(out, d["out64"]) < 1e-11


@pytest.mark.parametrize("name", FWD)
def test_spectral_form_fp32_within_tolerance(name):
    """fp32 oracle vs the reference in fp32 AND fp64: the 1e-4 bar of BASELINE.json."""
    d = golden(name)
    out = O.admm_tv_spectral_form(d["x"], d["lam"], d["rho"], d["kern"], bool(d["iso"]), int(d["maxit"]))
    assert out.dtype == np.float32
    assert O.rel_err(out, d["out64"]) < 1e-4
    assert O.rel_err(out, d["out32"]) < 1e-4


@pytest.mark.parametrize("name", [n for n in FWD if not n.startswith("iso")])
def test_packed_layout_model_matches_reference(name):
    d = golden(name)
    k = d["kern"]
    k2 = None if k.size == 0 else k[0, 0].astype(np.float64)
    out = M.admm_tv_packed(d["x"], float(d["lam"]), float(d["rho"]), k2, int(d["maxit"]))
    assert O.rel_err(out, d["out64"]) < 1e-11


@pytest.mark.parametrize("name", GRAD)
def test_adjoint_matches_reference_autograd(name):
    d = golden(name)
    gx, gl, gr, gk = O.admm_tv_backward(d["x"], d["lam"], d["rho"], d["kern"], d["gout"], bool(d["iso"]), int(d["maxit"]))
    assert O.rel_err(gx, d["gx"]) < 1e-11
    assert abs(gl - d["glam"][0]) <= 1e-10 * max(1.0, abs(d["glam"][0]))
    assert abs(gr - d["grho"][0]) <= 1e-10 * max(1.0, abs(d["grho"][0]))
    if d["kern"].size:
        assert O.rel_err(gk, d["gkern"]) < 1e-11


def test_known_answers():
    """maxit=0 -> zeros; constant image is a fixed point; delta kernel == denoise."""
    rng = np.random.default_rng(3)
    x = rng.random((1, 2, 16, 16))
    k = rng.random((3, 3)); k /= k.sum()
    assert np.all(O.admm_tv_spectral_form(x, 0.02, 0.04, k[None, None], False, 0) == 0)
    c = np.full((1, 1, 16, 16), 0.37)
    assert O.rel_err(O.admm_tv_spectral_form(c, 0.02, 0.04, k[None, None], False, 5), c) < 1e-12
    delta = np.zeros((3, 3)); delta[1, 1] = 1.0
    a = O.admm_tv_spectral_form(x, 0.05, 0.1, delta[None, None], False, 7)
    b = O.admm_tv_spectral_form(x, 0.05, 0.1, np.zeros((0,)), False, 7)
    assert O.rel_err(a, b) < 1e-12
    # circular shift equivariance
    r = O.admm_tv_spectral_form(np.roll(x, (3, 5), (-2, -1)), 0.02, 0.04, k[None, None], False, 6)
    assert O.rel_err(r, np.roll(O.admm_tv_spectral_form(x, 0.02, 0.04, k[None, None], False, 6), (3, 5), (-2, -1))) < 1e-12


def test_non_square_kernel_raises_like_reference():
    with pytest.raises(RuntimeError):
        O.admm_tv_spectral_form(np.zeros((1, 1, 8, 8)), 0.02, 0.04, np.zeros((1, 1, 3, 5)), False, 2)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8, 9, 12, 15, 16, 31, 45, 60, 64, 90, 135, 256])
def test_stockham_index_model(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    r = M.factorize(n)
    assert np.abs(M.stockham_fft(x, r, -1) - np.fft.fft(x)).max() < 1e-11 * n
    assert np.abs(M.stockham_fft(x, r, +1) - np.fft.ifft(x) * n).max() < 1e-11 * n


@pytest.mark.parametrize("W", [4, 7, 8, 9, 16, 33])
def test_pair_trick(W):
    rng = np.random.default_rng(W)
    a, b = rng.standard_normal(W), rng.standard_normal(W)
    Pa, Pb = M.pair_r2c(a, b)
    assert np.abs(Pa - M.rows_r2c(a)).max() < 1e-12 and np.abs(Pb - M.rows_r2c(b)).max() < 1e-12
    ra, rb = M.pair_c2r(Pa, Pb, W)
    assert np.abs(ra - a * W).max() < 1e-11 and np.abs(rb - b * W).max() < 1e-11


def test_tile_major_spectrum_layout_model():
    """The tile-major layout between the two large-frame kernels: a bijection of the H x Wc entries in which (a) the two
    rows the row kernel transforms together are adjacent complex slots (one 16-byte access), (b) the 8 columns of a tile
    for one row pair are one 128-byte run, (c) a column tile is one contiguous block."""
    from oracle import packed_layout_model as M
    H, Wc = 12, 16
    for kind in ("v", "x"):
        idx = np.array([[M.tile_major_index(u, c, H, kind) for c in range(Wc)] for u in range(H)])
        assert sorted(idx.ravel().tolist()) == list(range(H * Wc))
        for k in range(H // 2):
            ra, rb = (2 * k, 2 * k + 1) if kind == "v" else ((2 * k - 1) % H, 2 * k)
            assert np.all(idx[rb] == idx[ra] + 1)                                   # (a)
            assert np.all(np.diff(idx[ra, :8]) == 2) and idx[ra, 0] % 16 == 0       # (b): 8 columns x 2 rows x 8 bytes
        assert idx[:, :8].max() < H * 8 and idx[:, 8:].min() >= H * 8               # (c)
    P = np.arange(H * Wc).reshape(H, Wc).astype(np.complex64)
    assert np.array_equal(np.sort(M.to_tile_major(P, "x").real), np.arange(H * Wc))


def test_cfg3_full_size_fixture_pins_oracle_to_reference():
    """BASELINE configs[2] at its real shape and length (1 x 3 x 2160 x 3840, 63 x 63 PSF, 200 iterations): the fixture of
    tests/golden/make_golden_cfg3.py holds summaries of the fp64 oracle AND of the unmodified reference (fp32: its fp64
    run needs 263 GB) on the same input; they must agree within the reference's own fp32 noise."""
    d = golden("cfg3_2160x3840_gauss63_n200")
    amax = float(d["ref32_absmax"])
    assert tuple(d["shape"]) == (1, 3, 2160, 3840) and int(d["maxit"]) == 200 and int(d["k"]) == 63
    assert float(np.abs(d["oracle64_crops"] - d["ref32_crops"]).max()) / amax < 1e-4
    assert float(np.abs(d["oracle64_rowmean"] - d["ref32_rowmean"]).max()) / amax < 1e-4
    assert float(np.abs(d["oracle64_colmean"] - d["ref32_colmean"]).max()) / amax < 1e-4
    assert abs(float(d["oracle64_sum"]) - float(d["ref32_sum"])) / abs(float(d["ref32_sum"])) < 1e-5
    assert float(d["oracle_vs_ref32"]) < 1e-4
