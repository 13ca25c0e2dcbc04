"""Row pitch vs tcgen05 GEMM time: the same logical GEMM with operands at their natural pitch (H = 500 floats = 2000 B:
every 128-byte TMA row segment straddles two L2 lines) and in buffers whose rows are padded to a multiple of 32 floats."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from variational_mmt_b200 import ops
dev = "cuda"
shapes = [("gx      ", 1200, 2000, 500, 1, 1), ("dx      ", 1200, 500, 2000, 1, 0), ("dW_ih   ", 2000, 500, 1200, 0, 0),
          ("gen fwd ", 1240, 10000, 500, 1, 1), ("gen dX  ", 1240, 500, 10000, 1, 0), ("gen dW  ", 10000, 500, 1240, 0, 0),
          ("lin_out ", 1240, 500, 1000, 1, 1)]
flush = torch.empty(64 << 20, device=dev)
pad = lambda n: (n + 31) // 32 * 32


def view(rows, cols, padded):
    ld = pad(cols) if padded else cols
    return torch.randn(rows, ld, device=dev)[:, :cols]


for name, M, N, K, ak, bk in shapes:
    res = []
    for padded in (False, True):
        a = view(M, K, padded) if ak else view(K, M, padded)
        b = view(N, K, padded) if bk else view(K, N, padded)
        c = view(M, N, padded)
        fn = lambda: ops.gemm(a, b, c, M, N, K, a_kmajor=bool(ak), b_kmajor=bool(bk))
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res.append(sorted(ts)[len(ts) // 2] * 1e3)
    fl = 2.0 * M * N * K
    print(f"{name} M={M:6d} N={N:6d} K={K:6d} ak={ak} bk={bk}: natural pitch {res[0]:7.1f} us ({fl/res[0]/1e6:6.1f} TF/s) | "
          f"padded pitch {res[1]:7.1f} us ({fl/res[1]/1e6:6.1f} TF/s)")
