"""Device time of the exact-fp32 row-block linear kernel (csrc/rowlin.cu) per layer shape of config 1, alone on the GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from variational_mmt_b200 import ops
dev = "cuda"
flush = torch.empty(64 << 20, device=dev)
def bench(name, M, N, K, nprob, transposed=False, share_x=True):
    x = torch.randn(M, K, device=dev)
    ws = [torch.randn(*( (K, N) if transposed else (N, K)), device=dev) * 0.05 for _ in range(nprob)]
    bs = [torch.randn(N, device=dev) for _ in range(nprob)]
    outs = [torch.empty(M, N, device=dev) for _ in range(nprob)]
    from variational_mmt_b200 import _lib as L
    probs = [ops._rl_prob([x], w, None if transposed else b, o, 1) for w, b, o in zip(ws, bs, outs)]
    arr = (L.RowLin * len(probs))(*probs)
    st = L.stream()
    def run():
        L.lib.vmmt_rowlin(arr, len(probs), 0, int(transposed), M, N, K, st)
    for _ in range(3): run()
    ts = []
    REP = 20                                   # back-to-back launches: the host's launch latency is hidden behind the queue
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(REP): run()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / REP)
    ts.sort()
    by = nprob * N * K * 4
    print(f"{name:28s} M={M} N={N} K={K} x{nprob}: {ts[len(ts)//2]:7.1f} us (min {ts[0]:.1f})  weights {by/1e6:.1f} MB -> {by/ts[len(ts)//2]/1e3:.0f} GB/s")
bench("posterior L1", 40, 500, 3048, 2)
bench("prior L1 / L2", 40, 500, 500, 2)
bench("decoder z-bias", 40, 2000, 500, 1)
bench("image head fc1", 40, 2048, 500, 1)
bench("image head fc2", 40, 2048, 2048, 1)
bench("image head fc2 dgrad", 40, 2048, 2048, 1, transposed=True)
bench("posterior L1 dgrad(500)", 40, 500, 500, 2, transposed=True)
bench("prior decode M=250", 250, 500, 500, 2)
