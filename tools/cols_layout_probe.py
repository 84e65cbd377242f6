"""Does the column kernel care about DRAM locality?  Same kernel (k_cols_pow2, COLS_FFT_FWD: load -> 3 passes -> store), same
bytes, same CTA count: planes of 512 x 512 (a 16-column tile = 128-byte runs at 2 KB stride) against planes of 512 x 32
(Wc = 16: a tile is one contiguous 64 KB block)."""
import ctypes, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_admm_deconv_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
p = lambda t: ctypes.c_void_p(t.data_ptr())
for H in (512, 256):
    for (P, W) in ((192 * (512 // H) ** 2, H), (192 * (512 // H) ** 2 * (H // 32), 32), (192 * (512 // H) ** 2 * (H // 64), 64)):
        Wc = W // 2
        n = lib.admm_query_workspace(P, H, W, 0, 0, 1)
        ws = torch.empty(n, dtype=torch.uint8, device=dev)
        a = torch.randn(P, H, Wc, 2, device=dev); b = torch.empty_like(a)
        for inv in (0, 1):
            for _ in range(3):
                _lib.check(lib.admm_dbg_cols_fft(p(a), p(b), P, H, W, inv, p(ws), ws.numel(), None), "cols")
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            _lib.set_option("profile", 1); _lib.profile_reset()
            for _ in range(20):
                _lib.check(lib.admm_dbg_cols_fft(p(a), p(b), P, H, W, inv, p(ws), ws.numel(), None), "cols")
            torch.cuda.synchronize()
            ms, cnt = _lib.profile_read(2)
            _lib.set_option("profile", 0)
            # the dbg entry also launches the twiddle kernels (kind 2 as well): subtract nothing, they are microseconds
            by = a.numel() * 4 * 2
            print("H=%d planes=%d W=%d (Wc=%d) inverse=%d: %.3f ms per call incl. twiddles, %.0f GB/s" % (H, P, W, Wc, inv, ms / 20, by / (ms / 20 * 1e-3) / 1e9), flush=True)
