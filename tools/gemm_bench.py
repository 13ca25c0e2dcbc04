"""Device-time microbenchmark of vmmt_gemm (tcgen05 TF32) against cuBLAS TF32 (torch.matmul) per shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from variational_mmt_b200 import ops, _lib
torch.backends.cuda.matmul.allow_tf32 = True
dev = "cuda"
shapes = [(40, 500, 500, 1, 0), (500, 500, 40, 0, 0), (1000, 250, 1248, 0, 0), (40, 500, 3048, 1, 1), (2048, 2048, 40, 0, 0), (40, 2048, 2048, 1, 1), (1280, 1000, 500, 1, 1), (2000, 500, 1200, 0, 0),
          (1240, 2000, 500, 1, 1), (1240, 10000, 500, 1, 1), (1240, 500, 10000, 1, 0), (10000, 500, 1240, 0, 0),
          (2000, 500, 1240, 0, 0), (1240, 500, 2000, 1, 0), (4096, 4096, 4096, 1, 1), (8192, 8192, 1024, 1, 1),
          (40448, 32000, 1024, 1, 1)]
flush = torch.empty(64 << 20, device=dev)
for M, N, K, ak, bk in shapes:
    a = torch.randn((M, K) if ak else (K, M), device=dev)
    b = torch.randn((N, K) if bk else (K, N), device=dev)
    c = torch.empty(M, N, device=dev)
    def ours():
        ops.gemm(a, b, c, M, N, K, a_kmajor=bool(ak), b_kmajor=bool(bk))
    A = a if ak else a.t()
    B = b.t() if bk else b
    def cublas():
        torch.matmul(A, B, out=c)
    res = []
    for fn in (ours, cublas):
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res.append(sorted(ts)[len(ts) // 2])
    fl = 2.0 * M * N * K
    print(f"M={M:6d} N={N:6d} K={K:6d} ak={ak} bk={bk}: ours {res[0]*1e3:9.1f} us {fl/res[0]/1e9:8.1f} TF/s | cublas-tf32 {res[1]*1e3:9.1f} us {fl/res[1]/1e9:8.1f} TF/s")
