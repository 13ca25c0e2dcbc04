"""Per-call device time of the data-parallel optimiser step (torchrun --nproc-per-node N tools/dp_step_prof.py): the
peer / NVLS exchange kernels alone on the cfg1-sized flat buffers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import variational_mmt_b200 as vm
from variational_mmt_b200 import _lib
from variational_mmt_b200.flat import FlatParamsMixin

class Toy(FlatParamsMixin, torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(42781304 // 4, 4) * 0.1)])
m = Toy().to(dev)
m.flatten_parameters()
o = vm.Optim("adam", 0.002, 5)
o.set_parameters(m.parameters())
o.gflat.normal_()
for _ in range(3):
    o.step()
torch.cuda.synchronize(); dist.barrier()
res = {}
for it in range(10):
    prof = []
    dist.barrier(); torch.cuda.synchronize()
    _lib.set_profile(prof)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); o.step(); e1.record()
    torch.cuda.synchronize()
    _lib.set_profile(None)
    res.setdefault("step_total", []).append(e0.elapsed_time(e1) * 1e3)
    for n, a, a0, a1 in prof:
        res.setdefault(n, []).append(a0.elapsed_time(a1) * 1e3)
# the two halves separately (same kernels, one C call each)
from variational_mmt_b200._lib import fptr, stream
pe, n = o.peer, o.flat.numel()
for it in range(10):
    prof = []
    dist.barrier(); torch.cuda.synchronize()
    _lib.set_profile(prof)
    _lib.call("vmmt_peer_reduce_scatter", pe.segments, pe.mc_base, pe.grad_off, pe.rank, pe.world, 0, n, fptr(o._gsum), 0, fptr(o._pws), stream())
    _lib.call("vmmt_peer_adam_allgather", pe.segments, pe.mc_base, pe.param_off, pe.rank, pe.world, 0, n, fptr(o._gsum), fptr(o.exp_avg),
              fptr(o.exp_avg_sq), fptr(o._sq), 1, 5.0, 0.002, 0.9, 0.999, 1e-9, 20 + it, 1, 1, 0, stream())
    torch.cuda.synchronize()
    _lib.set_profile(None)
    for nme, a, a0, a1 in prof:
        res.setdefault(nme, []).append(a0.elapsed_time(a1) * 1e3)
if rank == 0:
    print("exchange:", o.exchange_in_use)
    for k, v in res.items():
        v = sorted(v)
        print(f"  {k:32s} median {v[len(v)//2]:8.1f} us  min {v[0]:8.1f}")
dist.barrier(); dist.destroy_process_group()
