"""One fused generator+NLL forward (and backward) launch at a given shape, for `ncu --set full` captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from variational_mmt_b200 import _lib
from variational_mmt_b200.ops import fptr, ptr, stream
M, H, V = [int(x) for x in sys.argv[1:4]]
FLAGS = int(sys.argv[4]) if len(sys.argv) > 4 else 0      # 2 = bf16 operands
dev = "cuda"
x = torch.randn(M, H, device=dev) * 0.5
W = (torch.rand(V, H, device=dev) - 0.5) * 0.2
b = (torch.rand(V, device=dev) - 0.5) * 0.2
tgt = torch.randint(4, V, (M,), device=dev)
lse = torch.empty(M, device=dev)
stats = torch.zeros(3, device=dev)
wsb = _lib.lib.vmmt_generator_workspace_bytes(M, H, V)
ws = torch.empty(wsb // 4, device=dev)
for _ in range(3):
    _lib.call("vmmt_generator_nll_fwd", fptr(x), fptr(W), fptr(b), ptr(tgt), 1, M, H, V, fptr(lse), fptr(stats),
              fptr(ws), wsb, FLAGS, stream())
torch.cuda.synchronize()
print("ok", stats.tolist())
