"""Attention core forward + backward at a given shape (for ncu captures / event timing): T B S H."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from variational_mmt_b200 import _lib
from variational_mmt_b200.ops import fptr, ptr, stream
T, B, S, H = [int(x) for x in sys.argv[1:5]] if len(sys.argv) > 4 else (31, 40, 30, 500)
dev = "cuda"
qp = torch.randn(T, B, H, device=dev) * 0.3
ctx = torch.randn(S, B, H, device=dev) * 0.3
lengths = torch.full((B,), S, device=dev, dtype=torch.int64)
align = torch.empty(T, B, S, device=dev)
cvec = torch.empty(T, B, H, device=dev)
dc = torch.randn(T, B, H, device=dev)
ds = torch.empty(T, B, S, device=dev)
dqp = torch.empty(T, B, H, device=dev)
dctx = torch.empty(S, B, H, device=dev)
flush = torch.empty(64 << 20, device=dev)


def fwd():
    _lib.call("vmmt_attention_fwd", fptr(qp), fptr(ctx), ptr(lengths), fptr(align), fptr(cvec), T, B, S, H, stream())


def bwd():
    _lib.call("vmmt_attention_bwd", fptr(dc), fptr(qp), fptr(ctx), fptr(align), ptr(lengths), fptr(ds), fptr(dqp), fptr(dctx),
              0, T, B, S, H, stream())


for fn, name in ((fwd, "fwd"), (bwd, "bwd")):
    for _ in range(3):
        fn()
    for hot in (0, 1):
        ts = []
        for _ in range(10):
            if not hot:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print("%s T=%d B=%d S=%d H=%d %s: %.1f us" % (name, T, B, S, H, "L2-warm" if hot else "L2-flushed", sorted(ts)[5]))
print("checksum", float(cvec.sum()), float(dctx.sum()))
