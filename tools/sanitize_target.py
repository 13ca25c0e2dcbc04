"""Small-shape driver for compute-sanitizer (memcheck / racecheck / synccheck): the cluster LSTM recurrence forward and
BPTT (mbarrier + st.async / bulk-copy DSMEM hand-off, tcgen05), the row-MLP kernel (cluster split-K through DSMEM), the
tcgen05 GEMM (TMA + TMEM), the attention v2 kernels (cp.async staging) and the peer barrier / exchange kernels at world 1.
    compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import variational_mmt_b200 as vm
from variational_mmt_b200 import ops
dev = "cuda"
torch.manual_seed(0)
what = sys.argv[1:] or ["lstm", "rowmlp", "gemm", "attention", "peer"]
if "lstm" in what:
    for (T, N, In, H, ndir, masked) in [(4, 5, 32, 64, 1, True), (3, 12, 32, 96, 2, False)]:
        x = (torch.randn(T, N, In, device=dev) * 0.5).requires_grad_(True)
        ws = []
        for d in range(ndir):
            ws += [(torch.randn(4 * H, In, device=dev) * 0.1).requires_grad_(True), (torch.randn(4 * H, H, device=dev) * 0.1).requires_grad_(True),
                   (torch.randn(4 * H, device=dev) * 0.1).requires_grad_(True), (torch.randn(4 * H, device=dev) * 0.1).requires_grad_(True)]
        lengths = torch.tensor(sorted([T] + [max(1, T - i % T) for i in range(N - 1)], reverse=True), device=dev) if masked else None
        o, hT, cT = ops.lstm_layer(x, None, None, None, lengths, {"save": True}, ws)
        (o.sum() + hT.sum() + cT.sum()).backward()
        ops.join_side()
        torch.cuda.synchronize()
        print("lstm T=%d N=%d H=%d ndir=%d ok, out norm %.4f" % (T, N, H, ndir, float(o.norm())))
if "rowmlp" in what:
    xs = [torch.randn(7, 64, device=dev, requires_grad=True), torch.randn(7, 200, device=dev)]
    heads = [[torch.nn.Parameter(torch.randn(48, 264, device=dev) * 0.1), torch.nn.Parameter(torch.randn(48, device=dev)),
              torch.nn.Parameter(torch.randn(24, 48, device=dev) * 0.1), torch.nn.Parameter(torch.randn(24, device=dev))] for _ in range(2)]
    ys = ops.row_mlp(xs, heads, (ops.ACT_NONE, ops.ACT_SOFTPLUS))
    (ys[0].sum() + ys[1].sum()).backward()
    ops.join_side()
    torch.cuda.synchronize()
    print("rowmlp ok", float(ys[0].norm()))
if "gemm" in what:
    a, b = torch.randn(130, 200, device=dev), torch.randn(96, 200, device=dev)
    c = torch.empty(130, 96, device=dev)
    ops.gemm(a, b, c, 130, 96, 200, bias=torch.randn(96, device=dev), act=2)
    torch.cuda.synchronize()
    print("gemm ok", float(c.norm()))
if "attention" in what:
    # v2 kernels (cp.async staging, K split over warps): resident-context path, two source blocks, v1 fallback (H % 4 != 0),
    # and the strided masked mean
    for (T, B, S, H) in [(19, 3, 11, 64), (17, 2, 50, 96), (5, 2, 33, 30)]:
        qp = (torch.randn(T, B, H, device=dev) * 0.3).requires_grad_(True)
        cx = (torch.randn(S, B, H, device=dev) * 0.3).requires_grad_(True)
        ln = torch.tensor([S] + [max(1, S - 3 * i) for i in range(1, B)], device=dev)
        cv, al = ops.AttentionCoreFn.apply(qp, cx, ln)
        (cv * torch.randn_like(cv)).sum().backward()
        torch.cuda.synchronize()
        print("attention T=%d B=%d S=%d H=%d ok %.4f" % (T, B, S, H, float(cv.norm())))
    y = torch.randn(6, 9, 40, device=dev, requires_grad=True)
    m = ops.MaskedMeanFn.apply(y.transpose(0, 1), torch.tensor([9, 7, 7, 3, 2, 1], device=dev))
    m.sum().backward()
    torch.cuda.synchronize()
    print("masked mean (transposed view) ok %.4f" % float(m.norm()))
if "peer" in what:
    from variational_mmt_b200.flat import FlatParamsMixin

    class Toy(FlatParamsMixin, torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(n) * 0.1) for n in (1000, 37, 5003)])
    m = Toy().to(dev)
    m.flatten_parameters()
    o = vm.Optim("adam", 0.002, 5, exchange="peer")
    o.set_parameters(m.parameters())
    for it in range(2):
        o.gflat.normal_()
        o.step()
    torch.cuda.synchronize()
    print("peer step ok", float(o.flat.norm()))
