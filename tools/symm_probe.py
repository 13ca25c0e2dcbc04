"""Probe: does torch's symmetric memory give an NVLS multicast mapping on this box?  torchrun --nproc-per-node N tools/symm_probe.py"""
import os, sys
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
t = sm.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", rank))
t.fill_(rank + 1)
h = sm.rendezvous(t, dist.group.WORLD)
print(f"rank {rank}: backend {sm.get_backend(torch.device('cuda', rank))} multicast_ptr {h.multicast_ptr:#x} buffer_ptrs {[hex(p) for p in h.buffer_ptrs]} "
      f"signal_pad_size {h.signal_pad_size} buffer_size {h.buffer_size}", flush=True)
try:
    print("has_multicast_support", type(h).has_multicast_support(torch.device("cuda").type, rank))
except Exception as e:
    print("has_multicast_support err", e)
h.barrier()
# peer read through the mapped pointer
peer = h.get_buffer((rank + 1) % world, (4,), torch.float32)
print(f"rank {rank}: peer value {peer.tolist()}", flush=True)
dist.barrier()
dist.destroy_process_group()
