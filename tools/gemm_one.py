"""One vmmt_gemm shape (for VMMT_GEMM_TRACE=1 / ncu): M N K [a_kmajor b_kmajor accumulate act]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from variational_mmt_b200 import ops
M, N, K = [int(x) for x in sys.argv[1:4]]
ak, bk, acc, act = ([int(x) for x in sys.argv[4:8]] + [1, 1, 0, 0])[:4] if len(sys.argv) > 4 else (1, 1, 0, 0)
a = torch.randn((M, K) if ak else (K, M), device="cuda"); b = torch.randn((N, K) if bk else (K, N), device="cuda")
c = torch.zeros(M, N, device="cuda"); bias = torch.randn(N, device="cuda")
for _ in range(3):
    ops.gemm(a, b, c, M, N, K, a_kmajor=bool(ak), b_kmajor=bool(bk), accumulate=acc, act=act, bias=None if acc else bias)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(a, b, c, M, N, K, a_kmajor=bool(ak), b_kmajor=bool(bk), accumulate=acc, act=act, bias=None if acc else bias); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print("M=%d N=%d K=%d ak=%d bk=%d acc=%d act=%d: %.1f us (L2-warm, event-timed)" % (M, N, K, ak, bk, acc, act, sorted(ts)[5]))
