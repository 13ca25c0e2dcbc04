"""Device time of one LSTM layer forward / backward (cfg1 encoder shape by default)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from variational_mmt_b200 import ops
T, N, In, H, ndir = [int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (30, 40, 500, 500, 1))]
dev = "cuda"
x = (torch.randn(T, N, In, device=dev) * 0.5).requires_grad_(True)
ws = []
for d in range(ndir):
    ws += [(torch.randn(4 * H, In, device=dev) * 0.1).requires_grad_(True), (torch.randn(4 * H, H, device=dev) * 0.1).requires_grad_(True),
           (torch.randn(4 * H, device=dev) * 0.1).requires_grad_(True), (torch.randn(4 * H, device=dev) * 0.1).requires_grad_(True)]
from variational_mmt_b200 import _lib
def run():
    prof = []
    _lib.set_profile(prof)
    o, hT, cT = ops.lstm_layer(x, None, None, None, None, {"save": True}, ws)
    o.sum().backward()
    torch.cuda.synchronize()
    _lib.set_profile(None)
    return {n: e0.elapsed_time(e1) for n, a, e0, e1 in prof if "lstm" in n}
for _ in range(3): run()
r = run()
print(f"T={T} N={N} H={H} ndir={ndir}: " + "  ".join(f"{k} {v*1e3:.0f} us ({v*1e3/T:.2f} us/step)" for k, v in r.items()))
