"""Batch-sharded data parallelism: one process per GPU, one SUM all-reduce per step.

The reference has no multi-GPU path (train_mm_vi_model1.py:73-75 exits when more than one GPU id is
given); what an N-rank step must equal is the reference's own gradient accumulation over N batches
(onmt/TrainerMultimodal.py:342-346, 625-718, ``-accum_count N``): ``normalization`` is the sentence
count of ALL the batches, every batch back-propagates ``(NLL + image + KL) / normalization`` and the
gradients ADD, then one ``optim.step()``.  So: each rank divides its local loss by the GLOBAL sentence
count and the collective is SUM (not mean) over the flat gradient buffer (4 B per parameter), followed
by the same global-norm clip + Adam on every rank.  Whole batches are the unit of sharding because the
target encoder couples the examples inside a batch (SURVEY.md hazard H1).  Decode needs no collective:
sentences are independent and are simply dealt out to the ranks.

Works on any torch.distributed backend: "nccl" on the GPUs (NVLink 5 / NVSwitch), "gloo" in the CPU tests.

On one NVSwitch box the gradient exchange does not go through NCCL at all: ``PeerExchange`` places every
rank's flat parameter and gradient buffers in cudaIpc-exported memory and the optimiser step becomes
reduce-scatter (P2P loads) -> clip + Adam on the rank's 1/N slice -> all-gather (P2P stores), fused into
the update kernels themselves (csrc/peer.cu).  torch.distributed then only
carries the 64-byte IPC handles once and a few scalars.
"""
import ctypes
import os
import socket
import sys

import torch
import torch.distributed as dist


def is_active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def init_from_env(backend=None, device=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*) -> (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def rank_world():
    return (dist.get_rank(), dist.get_world_size()) if is_active() else (0, 1)


def batches_of_rank(n_batches, rank=None, world=None):
    """Indices of the (whole) batches rank r trains on within one global step sequence: r, r + N, r + 2N, ...
    Global step s consumes batches [s*N, (s+1)*N); a trailing partial group is dropped by every rank alike."""
    if rank is None:
        rank, world = rank_world()
    usable = (n_batches // world) * world
    return list(range(rank, usable, world))


def sentences_of_rank(n_sentences, rank=None, world=None):
    """Contiguous [start, stop) slice of a test set for replica decoding (no collective)."""
    if rank is None:
        rank, world = rank_world()
    per, rem = divmod(n_sentences, world)
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)


def global_normalization(local_sentences, device=None):
    """Sentence count of the global step = sum of the per-rank batch sizes (TrainerMultimodal.py:342-346)."""
    if not is_active():
        return int(local_sentences)
    t = torch.tensor([int(local_sentences)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def all_reduce_gradients(flat_grads):
    """SUM over ranks, in place, of the flat gradient buffer (one collective per step)."""
    if is_active():
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    return flat_grads


def early_final_begin(model):
    """Offset (floats) in the model's flat buffers from which on every parameter belongs to the latent / image networks or
    the generator -- the modules whose gradients are final BEFORE the encoders' backward pass starts (they hang off
    the loss, not off the recurrences).  They are registered last (onmt/Models.py:737-760 order; the generator is
    attached by the constructor), so they form the tail of the flat buffer.  None when that does not hold."""
    early = ("inf_net_global.", "gen_net_global.", "inf_net_image.", "generator.")
    off, begin, seen = 0, None, set()
    for name, p in model.named_parameters():
        if id(p) in seen:
            continue
        seen.add(id(p))
        if name.startswith(early):
            if begin is None:
                begin = off
        elif begin is not None:
            return None                       # a late-final parameter after the first early one: no contiguous tail
        off += ((p.numel() + 3) // 4) * 4
    return begin if begin else None


def reduce_statistics(vec):
    """SUM over ranks of a VIStatistics vector {nmt_loss, n_words, n_correct, kl, img, cos, kl_after, elbo}
    (reporting only; returns a new tensor)."""
    out = vec.clone()
    if is_active():
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


# ------------------------------------------------------------------------------------------------
class _RawDeviceBuffer(object):
    """__cuda_array_interface__ view of library-allocated device memory (zero copy into torch)."""

    def __init__(self, ptr, n_floats, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerExchange(object):
    """One cudaIpc segment per rank = [signal block | flat params | flat grads], opened by every peer.

    ``PeerExchange.create(n_floats, device)`` is collective over the default process group (or local when
    torch.distributed is not up: world = 1, used by the single-GPU kernel tests).  Returns None -- on every
    rank alike -- when the ranks are not on one host or a peer mapping fails; the caller then keeps the
    NCCL all-reduce path (still GPU; reported in bench.py's config)."""

    def __init__(self):
        self.rank, self.world = rank_world()
        self.segments = None           # ctypes array of `world` base pointers (index = rank)
        self.base = None
        self.flat = self.gflat = None
        self._opened = []
        self.mc_base = None            # multicast address of the segments (NVLS: multimem.ld_reduce / multimem.st), or None
        self._symm = None              # (tensor, handle) of torch's symmetric memory when the segment lives there

    @classmethod
    def _create_symmetric(cls, self, nbytes, device):
        """The segment in torch's symmetric memory (cuMem allocation + peer mappings + one cuMulticast object over all
        ranks' segments): gives the NVLS multicast address.  torch.distributed is plumbing here; every kernel that touches
        the memory is ours (csrc/peer.cu).  Returns False (on every rank alike) when unavailable."""
        ok = True
        try:
            import torch.distributed._symmetric_memory as sm
            t = sm.empty(nbytes // 4, dtype=torch.float32, device=device)
            t.zero_()
            torch.cuda.synchronize(device)
            h = sm.rendezvous(t, dist.group.WORLD)
            ptrs = [int(p) for p in h.buffer_ptrs]
            mc = int(h.multicast_ptr)
            ok = len(ptrs) == self.world and mc != 0
        except Exception as e:                                # noqa: BLE001
            ok, t, h, ptrs, mc = False, None, None, None, 0
            if self.rank == 0:
                sys.stderr.write("variational_mmt_b200: symmetric memory unavailable (%s)\n" % e)
        oks = [None] * self.world
        dist.all_gather_object(oks, ok)
        if not all(oks):
            return False
        self._symm = (t, h)
        self.base = ptrs[self.rank]
        self.mc_base = mc
        self.segments = (ctypes.c_void_p * self.world)(*ptrs)
        self.flat = t[self.param_off // 4: self.param_off // 4 + self.n]
        self.gflat = t[self.grad_off // 4: self.grad_off // 4 + self.n]
        self.device = device
        return True

    @classmethod
    def create(cls, n_floats, device):
        from . import _lib as L
        self = cls()
        sig = int(L.lib.vmmt_peer_signal_bytes())
        self.n = int(n_floats)
        self.param_off = sig
        self.grad_off = sig + ((self.n * 4 + 255) // 256) * 256
        nbytes = self.grad_off + ((self.n * 4 + 255) // 256) * 256
        # NVLS (multimem through the switch) or P2P loads / stores?  Measured on 8 x B200 (tools/dp_step_prof.py, one-phase
        # step over 42.8 M parameters): N=2 611 vs 338 us, N=4 540 vs 450 us, N=8 549 vs 587 us -- the reduce-scatter's
        # contributions still leave every GPU at (N-1)/N * 4 B/param and the all-gather still arrives at that rate, the
        # switch only thins the opposite direction, and one multimem access costs more than a peer access.  Default: NVLS
        # from 8 ranks on; VMMT_DP_NVLS=1 / 0 forces it on / off.
        want = os.environ.get("VMMT_DP_NVLS")
        use_nvls = (want == "1") if want in ("0", "1") else self.world >= 8
        if self.world > 1 and use_nvls and cls._create_symmetric(self, nbytes, device):
            return self
        hb = int(L.lib.vmmt_peer_handle_bytes())
        handle = ctypes.create_string_buffer(hb)
        base = ctypes.c_void_p()
        ok, why = True, ""
        with torch.cuda.device(device):
            rc = L.lib.vmmt_peer_alloc(nbytes, ctypes.byref(base), handle)
        if rc != 0:
            ok, why = False, "alloc: " + L.last_error()
        self.base = base.value
        ptrs = [None] * self.world
        ptrs[self.rank] = self.base
        if self.world > 1:
            infos = [None] * self.world
            dist.all_gather_object(infos, (socket.gethostname(), bytes(handle.raw), ok))
            if len({h for h, _b, _o in infos}) != 1:
                ok, why = False, "ranks span several hosts"
            if ok and all(o for _h, _b, o in infos):
                with torch.cuda.device(device):
                    for j, (_h, hbytes, _o) in enumerate(infos):
                        if j == self.rank:
                            continue
                        pp = ctypes.c_void_p()
                        if L.lib.vmmt_peer_open(hbytes, ctypes.byref(pp)) != 0:
                            ok, why = False, "open rank %d: %s" % (j, L.last_error())
                            break
                        ptrs[j] = pp.value
                        self._opened.append(pp.value)
            else:
                ok = False
            oks = [None] * self.world
            dist.all_gather_object(oks, (ok, why))
            if not all(o for o, _w in oks):
                if self.rank == 0:
                    sys.stderr.write("variational_mmt_b200: NVLink peer exchange unavailable (%s); using the NCCL "
                                     "all-reduce\n" % "; ".join(w for _o, w in oks if w))
                self.close()
                return None
        elif not ok:
            raise RuntimeError("PeerExchange: " + why)
        self.segments = (ctypes.c_void_p * self.world)(*ptrs)
        self.flat = torch.as_tensor(_RawDeviceBuffer(self.base + self.param_off, self.n, self), device=device)
        self.gflat = torch.as_tensor(_RawDeviceBuffer(self.base + self.grad_off, self.n, self), device=device)
        self.device = device
        return self

    def slice_bounds(self):
        """(lo, hi, capacity) in floats of the slice this rank reduces and updates."""
        from . import _lib as L
        lo, hi = ctypes.c_int64(), ctypes.c_int64()
        cap = L.lib.vmmt_peer_slice(self.n, self.world, self.rank, ctypes.byref(lo), ctypes.byref(hi))
        return int(lo.value), int(hi.value), int(cap)

    def barrier(self):
        """Device-side barrier over the peer segments on the current stream (no host synchronisation)."""
        from . import _lib as L
        L.call("vmmt_peer_barrier", self.segments, self.rank, self.world, 0, L.stream())

    def close(self):
        from . import _lib as L
        for p in self._opened:
            L.lib.vmmt_peer_close(p)
        self._opened = []
        if self.base is not None and self.flat is None:       # never handed out as tensors: safe to free
            L.lib.vmmt_peer_free(self.base)
            self.base = None
