"""torch.autograd.Function wrappers over the libvmmt C ABI.

Every arithmetic step of the hot path runs in libvmmt kernels; torch is used for device memory
(caching allocator), views and autograd bookkeeping only.  Parameter gradients are ACCUMULATED IN
PLACE into ``param.grad`` (views of one flat buffer, see ``flat.py``) by the backward kernels
themselves -- the Functions return ``None`` for parameters -- which is what makes the gradient
all-reduce and the fused clip+Adam single launches over one contiguous buffer.
"""
import ctypes as C
import os

import torch
from torch.autograd import Function

from . import _lib as L
from ._lib import ACT_NONE, ACT_RELU, ACT_TANH, ACT_SOFTPLUS, ACT_SIGMOID, fptr, ptr, stream

# ---- arithmetic mode: host-side state turned into the per-call `flags` of the C ABI (include/vmmt.h VMMT_F_*); the
# library itself has no global mode ---------------------------------------------------------------------------------
_mode = {"gemm": 1 if os.environ.get("VMMT_GEMM") == "simt" else 0, "background": False, "invariant": False}


def set_gemm_mode(mode):
    """0 = tensor cores, TF32 operands (fp32 storage; default); 1 = exact-fp32 SIMT contractions and recurrences
    everywhere (parity debugging; also env VMMT_GEMM=simt); 2 = tensor cores, bf16 operands (the bf16 variant)."""
    assert mode in (0, 1, 2)
    _mode["gemm"] = int(mode)


def get_gemm_mode():
    return _mode["gemm"]


def flags():
    """The per-call `flags` of the C ABI for the current host-side mode."""
    m = _mode["gemm"]
    return (L.F_EXACT if m == 1 else L.F_BF16 if m == 2 else 0) | (L.F_BACKGROUND if _mode["background"] else 0) | \
        (L.F_NO_SPLITK if _mode["invariant"] else 0)


class batch_invariant(object):
    """``with ops.batch_invariant():`` every contraction inside keeps one accumulation chain per output element in fixed K
    order (no split-K): a row's result does not depend on how many rows share the call.  Used by the decoder
    (translate/): a sentence decoded alone and the same sentence decoded in a batch give identical tokens."""

    def __enter__(self):
        self.prev = _mode["invariant"]
        _mode["invariant"] = True
        return self

    def __exit__(self, *exc):
        _mode["invariant"] = self.prev
        return False


_seed_state = {"seed": 0x5EED, "offset": 0, "base": None}
RNG_STEP_STRIDE = 1 << 16          # Philox offsets reserved per step (>> dropout / sample calls in one step)


def manual_seed(seed):
    """Seed of the in-kernel Philox streams (dropout masks, latent noise)."""
    _seed_state["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF
    _seed_state["offset"] = 0
    if _seed_state["base"] is not None:
        _seed_state["base"].zero_()


def _next_offset():
    _seed_state["offset"] += 1
    return _seed_state["offset"]


def rng_base(device=None):
    """Device-resident step counter added to every Philox offset.  It exists so that a captured CUDA graph
    (graph.py) draws fresh dropout masks / latent noise on each replay: the per-call offsets are baked into the
    graph, the counter is advanced by a kernel at the end of it."""
    if _seed_state["base"] is None:
        _seed_state["base"] = torch.zeros(1, dtype=torch.int64, device=device or torch.device("cuda", torch.cuda.current_device()))
    return _seed_state["base"]


def _base_ptr():
    b = _seed_state["base"]
    return None if b is None else b.data_ptr()


def begin_step():
    """Restart the per-step call index (call at the top of a step that ends with advance_rng())."""
    _seed_state["offset"] = 0
    _bf16["cache"].clear()


def advance_rng():
    L.call("vmmt_counter_add", rng_base().data_ptr(), RNG_STEP_STRIDE, stream())


class _P(object):
    """Hides a parameter from autograd's scan of Function inputs.

    Parameter gradients are written by the backward kernels straight into ``param.grad``; autograd never
    sees a gradient for them.  If the parameters were passed as tensor inputs, every Function node would
    hold an edge to the parameter's AccumulateGrad node, and that node remembers the stream it was created
    on: one created by an earlier eager forward (legacy default stream), kept alive by any live autograd
    graph, makes the engine record an event on that stream at the end of a backward that runs inside a
    CUDA-graph capture, which invalidates the capture.  Instead each Function takes a fresh zero-size
    ``anchor`` leaf (so that its output requires grad even when only parameters do)."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t


def _anchor(device):
    return torch.empty(0, device=device, requires_grad=True) if torch.is_grad_enabled() else None


# ---- side stream: weight / bias gradient GEMMs run off the critical path of the backward pass -------------------
# The backward's critical path is the chain of recurrences (latency bound, ~110 of the 148 SMs lightly used) plus the
# input-gradient GEMMs that feed the next layer down.  The weight-gradient GEMMs and bias column sums of every
# Linear / LSTM layer only feed the optimiser, so they are issued on lowest-priority side streams ("lanes": the launches
# of one block stay on one lane, hence serialised among themselves -- two of them may accumulate into slices of the same
# parameter) that fork from the main stream after the tensors they read are complete, and are joined before the
# gradients are consumed (end of the backward pass / Optim.step).  Under CUDA-graph capture the fork / join become
# graph edges.  (LSTM bias gradients are not among them any more: the BPTT kernel sums them itself.)
_side = {"streams": [], "keep": [], "dirty": False, "enabled": True, "next": 0,
         "lanes": int(os.environ.get("VMMT_SIDE_LANES", "4"))}


def set_side_stream_enabled(flag):
    _side["enabled"] = bool(flag)


def _side_stream(lane, device):
    st = _side["streams"]
    if not st or st[0].device != device:
        st[:] = [torch.cuda.Stream(device=device) for _ in range(_side["lanes"])]
    return st[lane % len(st)]


class _OnSide(object):
    """Issue the enclosed launches on a side stream, ordered after everything already on the current stream.
    Successive blocks alternate between ``lanes`` side streams: each block holds the weight / bias gradients of ONE
    module application (its launches stay ordered among themselves -- dW_hh receives two products), different
    modules own disjoint parameter memory, so blocks may overlap; what matters is the tail after the last
    recurrence of the backward pass, where the encoders' weight gradients would otherwise queue up on one stream."""

    def __init__(self, lane=None):
        self.lane = lane

    def __enter__(self):
        self.active = _side["enabled"]
        if not self.active:
            return self
        cur = torch.cuda.current_stream()
        if self.lane is None:
            side = _side_stream(_side["next"], cur.device)
            _side["next"] += 1
        else:
            side = _side_stream(self.lane, cur.device)
        side.wait_stream(cur)
        if not _side["dirty"]:
            # first side-stream block of this backward pass: join automatically when the pass ends, so that
            # whoever reads .grad afterwards (on the stream that called backward) sees complete gradients
            try:
                torch.autograd.Variable._execution_engine.queue_callback(join_side)
            except RuntimeError:
                pass                                   # not inside a backward pass: the caller joins explicitly
        self.ctx = torch.cuda.stream(side)
        self.ctx.__enter__()
        _mode["background"] = True
        _side["dirty"] = True
        return self

    def __exit__(self, *exc):
        if self.active:
            _mode["background"] = False
            self.ctx.__exit__(*exc)
        return False


def on_side(*keep, lane=None):
    """Context manager; ``keep`` are the tensors the side-stream launches read: they are held until join_side() so
    that the caching allocator cannot hand their memory to later main-stream work.  ``lane`` pins the block to one
    of the side streams (default: alternate)."""
    if _side["enabled"]:
        _side["keep"].extend(t for t in keep if t is not None)
    return _OnSide(lane)


def join_side():
    """The current stream waits for all side-stream work issued so far (call before the gradients are read)."""
    if _side["dirty"]:
        cur = torch.cuda.current_stream()
        for st in _side["streams"]:
            cur.wait_stream(st)
    _side["dirty"] = False
    _side["next"] = 0
    _side["keep"].clear()


# ---- branch stream: independent sub-graphs of the forward pass (and, through autograd's per-node streams, of the
# backward pass) run beside the main chain: prior network and image-feature head beside the decoder ----------------
_branch = {"streams": {}, "enabled": True}
LOW_LANE = 3          # branch lane for sub-graphs nothing on the critical chain waits for: lowest stream priority


class branch(object):
    """``with ops.branch(): ...`` issues the enclosed module calls on a branch stream, ordered after everything
    already on the current stream.  Autograd nodes created inside remember that stream, so their backward runs
    there too.  ``ops.join_branch(*tensors)`` makes the current stream wait and hands the tensors over.
    ``lane`` selects one of several branch streams (0: target encoder beside the source encoder, 1: the scale MLP of the
    posterior beside its location MLP, 2: the decoder's input projection, LOW_LANE: prior network and image head -- only
    the loss reads them); blocks may nest across lanes."""

    def __init__(self, lane=0):
        self.lane = lane

    def __enter__(self):
        self.active = _branch["enabled"] and torch.cuda.is_available()
        if not self.active:
            return self
        cur = torch.cuda.current_stream()
        st = _branch["streams"].get(self.lane)
        if st is None or st.device != cur.device:
            prio = int(os.environ.get("VMMT_BRANCH_LOW_PRIO", "0")) if self.lane == LOW_LANE \
                else int(os.environ.get("VMMT_BRANCH_PRIO", "-1"))
            st = _branch["streams"][self.lane] = torch.cuda.Stream(device=cur.device, priority=prio)
        st.wait_stream(cur)
        self.ctx = torch.cuda.stream(st)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.active:
            self.ctx.__exit__(*exc)
        return False


def join_branch(*tensors, lane=0):
    st = _branch["streams"].get(lane)
    if st is None or not _branch["enabled"]:
        return
    cur = torch.cuda.current_stream()
    cur.wait_stream(st)
    for t in tensors:
        if t is not None and t.is_cuda:
            t.record_stream(cur)


_early_cb = [None]


def arm_early_exchange(fn):
    """``fn`` runs once, at the start of the first encoder-stack backward node of the current backward pass."""
    _early_cb[0] = fn


def _fire_early_exchange(after_event=None):
    fn, _early_cb[0] = _early_cb[0], None
    if fn is not None:
        fn(after_event)


_gx = {}


def _gx_stream(device):
    """Second stream for the reverse direction's input projection of a bidirectional layer (same priority as the chain)."""
    st = _gx.get(device)
    if st is None:
        st = _gx[device] = torch.cuda.Stream(device=device, priority=int(os.environ.get("VMMT_BRANCH_PRIO", "-1")))
    return st


def aux_streams():
    """Every helper stream this module has created (side lanes, branch lanes, loss stream): what a consumer of ALL the
    gradients issued so far has to wait for."""
    out = list(_side["streams"]) + list(_branch["streams"].values()) + list(_gx.values()) + list(_attn_fork.values())
    if _loss_stream is not None:
        out.append(_loss_stream)
    return out


def set_branch_enabled(flag):
    _branch["enabled"] = bool(flag)


def grad_buf(p):
    """The tensor parameter gradients are accumulated into (allocated zero-filled on first use)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


# ---- bf16 variant (set_gemm_mode(2)): operands are cast to bf16 (pitch padded to 8 elements) and contracted by the same
# tcgen05 kernel with kind::f16 MMAs.  Casts of PARAMETERS (views of the flat parameter buffer) are cached until the
# weights change (Optim.step) or a new step begins (begin_step: a captured graph must contain its own casts).
_bf16 = {"version": 0, "cache": {}}


def weights_changed():
    _bf16["version"] += 1
    _bf16["cache"].clear()


def _as_bf16(t, rows, cols):
    """bf16 copy [rows, cols padded to 8] of the 2-D fp32 view `t` (unit inner stride)."""
    from .flat import _OWNERS
    is_param = t.untyped_storage().data_ptr() in _OWNERS              # a view of a flat parameter buffer
    key = (t.data_ptr(), rows, cols, t.stride(0)) if is_param else None
    if key is not None:
        hit = _bf16["cache"].get(key)
        if hit is not None:
            return hit
    ld = (cols + 7) // 8 * 8
    dst = torch.empty(rows, ld, device=t.device, dtype=torch.bfloat16)
    L.call("vmmt_cast_bf16", fptr(t), t.stride(0), dst.data_ptr(), ld, rows, cols, stream())
    if key is not None:
        _bf16["cache"][key] = dst
    return dst


def _bf16_ok(a, b, M, N, K):
    return (_mode["gemm"] == 2 and N >= 64 and K >= 16 and M * N * K >= 64 * 64 * 64 and a.dtype == torch.float32)


def gemm(a, b, out, M, N, K, a_kmajor=True, b_kmajor=True, bias=None, act=ACT_NONE, accumulate=0):
    """out[M,N] = act(op(a) op(b) + bias) (+out).  a, b, out: 2-D views with unit inner stride."""
    assert a.stride(-1) == 1 and b.stride(-1) == 1 and out.stride(-1) == 1
    if _bf16_ok(a, b, M, N, K):
        ab = _as_bf16(a, M if a_kmajor else K, K if a_kmajor else M)
        bb = _as_bf16(b, N if b_kmajor else K, K if b_kmajor else N)
        L.call("vmmt_gemm_bf16", ab.data_ptr(), ab.stride(0), int(a_kmajor), bb.data_ptr(), bb.stride(0), int(b_kmajor),
               fptr(out), out.stride(0), M, N, K, fptr(bias), act, int(accumulate), flags() & ~L.F_BF16, stream())
        return out
    L.call("vmmt_gemm", fptr(a), a.stride(0), int(a_kmajor), fptr(b), b.stride(0), int(b_kmajor),
           fptr(out), out.stride(0), M, N, K, fptr(bias), act, int(accumulate), flags(), stream())
    return out


def gemm_dual(a1, b1, a2, b2, out, M, N, K1, K2, bias=None, act=ACT_NONE):
    """out[M,N] = act(a1 b1^T + a2 b2^T + bias) in one launch (all operands [rows, K] with unit inner stride)."""
    assert a1.stride(-1) == 1 and b1.stride(-1) == 1 and a2.stride(-1) == 1 and b2.stride(-1) == 1 and out.stride(-1) == 1
    if _bf16_ok(a1, b1, M, N, K1) and _bf16_ok(a2, b2, M, N, K2):
        gemm(a1, b1, out, M, N, K1, bias=bias)                      # two accumulating launches in the bf16 variant
        return gemm(a2, b2, out, M, N, K2, act=act, accumulate=(1 if act == ACT_NONE else 2))
    L.call("vmmt_gemm_dual", fptr(a1), a1.stride(0), fptr(b1), b1.stride(0), K1, fptr(a2), a2.stride(0), fptr(b2),
           b2.stride(0), K2, fptr(out), out.stride(0), M, N, fptr(bias), act, flags(), stream())
    return out


def colsum_acc(a, M, N, out, out2=None):
    L.call("vmmt_colsum_acc", fptr(a), a.stride(0), M, N, fptr(out), fptr(out2), stream())


# --------------------------------------------------------------------------------------------
class EmbeddingFn(Function):
    """nn.Embedding lookup (onmt/modules/Embeddings.py:169-188)."""

    @staticmethod
    def forward(ctx, anchor, idx, weight, pad_idx):
        weight = weight.t
        idx = idx.contiguous()
        n, E = idx.numel(), weight.shape[1]
        out = torch.empty(*idx.shape, E, device=weight.device, dtype=torch.float32)
        L.call("vmmt_embedding_fwd", ptr(idx), n, fptr(weight), weight.shape[0], E, fptr(out), stream())
        ctx.save_for_backward(idx)
        ctx.weight, ctx.pad_idx = weight, pad_idx
        return out

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        w = ctx.weight
        dout = dout.contiguous()
        L.call("vmmt_embedding_bwd", ptr(idx), idx.numel(), fptr(dout), w.shape[1], ctx.pad_idx,
               fptr(grad_buf(w)), w.shape[0], stream())
        return None, None, None, None


def embedding(idx, weight, pad_idx):
    return EmbeddingFn.apply(_anchor(weight.device), idx, _P(weight), pad_idx)


class LinearFn(Function):
    """y = act(x W[:, c0:c1]^T + b), optionally y = act(x W^T + b + addend) (nn.Linear + activation)."""

    @staticmethod
    def forward(ctx, anchor, x, weight, bias, act, cols, addend, rowwise=False):
        weight, bias = weight.t, bias.t
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < x2.shape[1]):
            x2 = x2.contiguous()
        c0, c1 = cols if cols is not None else (0, weight.shape[1])
        wv = weight[:, c0:c1]
        M, K, N = x2.shape[0], c1 - c0, weight.shape[0]
        if rowwise:
            # batch-row operand (one row per example): exact fp32, the same arithmetic per row whatever M is (csrc/rowlin.cu)
            assert addend is None
            y = torch.empty(M, N, device=x.device, dtype=torch.float32)
            rowlin([_rl_prob([x2], wv, bias, y, act)], M, N, K)
        elif addend is not None:
            y = addend.reshape(M, N).clone()
            gemm(x2, wv, y, M, N, K, bias=bias, act=act, accumulate=2)
        else:
            y = torch.empty(M, N, device=x.device, dtype=torch.float32)
            gemm(x2, wv, y, M, N, K, bias=bias, act=act)
        ctx.save_for_backward(x2, y)
        ctx.weight, ctx.bias, ctx.act, ctx.cols, ctx.xshape = weight, bias, act, (c0, c1), x.shape
        ctx.has_addend = addend is not None
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, y = ctx.saved_tensors
        weight, bias, act = ctx.weight, ctx.bias, ctx.act
        c0, c1 = ctx.cols
        M, K, N = x2.shape[0], c1 - c0, weight.shape[0]
        dy = dy.reshape(M, N).contiguous()
        if act != ACT_NONE:
            dpre = torch.empty_like(dy)
            L.call("vmmt_act_bwd", fptr(dy), fptr(y), fptr(dpre), dy.numel(), act, stream())
        else:
            dpre = dy
        dx = None
        if ctx.needs_input_grad[1]:
            dx = torch.empty(M, K, device=dy.device, dtype=torch.float32)
            gemm(dpre, weight[:, c0:c1], dx, M, K, N, b_kmajor=False)                    # dx = dpre W
            dx = dx.view(ctx.xshape)
        with on_side(dpre, x2):                                                          # off the critical path
            if weight.requires_grad:
                gw = grad_buf(weight)[:, c0:c1]
                gemm(dpre, x2, gw, N, K, M, a_kmajor=False, b_kmajor=False, accumulate=1)   # dW += dpre^T x
            if bias is not None and bias.requires_grad:
                colsum_acc(dpre, M, N, grad_buf(bias))
        dadd = dpre.view(*ctx.xshape[:-1], N) if ctx.has_addend and ctx.needs_input_grad[6] else None
        return None, dx, None, None, None, None, dadd, None


class DualLinearFn(Function):
    """y = act([x1 ; x2] W^T) for W = [N, K1 + K2] without materialising the concatenation: ONE dual-operand GEMM forward
    (vmmt_gemm_dual), ONE input-gradient GEMM [dx1 ; dx2] = dpre W backward (GlobalAttention.py:187-190: linear_out over
    cat([c, q])), weight gradients per column block on a side stream."""

    @staticmethod
    def forward(ctx, anchor, x1, x2, weight, act):
        weight = weight.t
        K1, K2 = x1.shape[-1], x2.shape[-1]
        N = weight.shape[0]
        assert weight.shape[1] == K1 + K2
        a1 = x1.reshape(-1, K1)
        a2 = x2.reshape(-1, K2)
        a1 = a1 if a1.stride(-1) == 1 and (a1.shape[0] == 1 or a1.stride(0) >= K1) else a1.contiguous()
        a2 = a2 if a2.stride(-1) == 1 and (a2.shape[0] == 1 or a2.stride(0) >= K2) else a2.contiguous()
        M = a1.shape[0]
        y = torch.empty(M, N, device=x1.device, dtype=torch.float32)
        gemm_dual(a1, weight[:, :K1], a2, weight[:, K1:], y, M, N, K1, K2, act=act)
        ctx.save_for_backward(a1, a2, y)
        ctx.weight, ctx.act, ctx.shapes = weight, act, (x1.shape, x2.shape)
        return y.view(*x1.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        a1, a2, y = ctx.saved_tensors
        weight, act = ctx.weight, ctx.act
        M, K1 = a1.shape
        K2 = a2.shape[1]
        N = weight.shape[0]
        dy = dy.reshape(M, N).contiguous()
        if act != ACT_NONE:
            dpre = torch.empty_like(dy)
            L.call("vmmt_act_bwd", fptr(dy), fptr(y), fptr(dpre), dy.numel(), act, stream())
        else:
            dpre = dy
        dx = torch.empty(M, K1 + K2, device=dy.device, dtype=torch.float32)
        gemm(dpre, weight, dx, M, K1 + K2, N, b_kmajor=False)                      # [dx1 ; dx2] = dpre W, one launch
        with on_side(dpre, a1, a2):
            if weight.requires_grad:
                gw = grad_buf(weight)
                gemm(dpre, a1, gw[:, :K1], N, K1, M, a_kmajor=False, b_kmajor=False, accumulate=1)
                gemm(dpre, a2, gw[:, K1:], N, K2, M, a_kmajor=False, b_kmajor=False, accumulate=1)
        s1, s2 = ctx.shapes
        dx1 = dx[:, :K1].reshape(s1) if ctx.needs_input_grad[1] else None
        dx2 = dx[:, K1:].reshape(s2) if ctx.needs_input_grad[2] else None
        return None, dx1, dx2, None, None


def dual_linear(x1, x2, weight, act=ACT_NONE):
    return DualLinearFn.apply(_anchor(x1.device), x1, x2, _P(weight), act)


def linear(x, weight, bias=None, act=ACT_NONE, cols=None, addend=None, rowwise=False):
    """``rowwise``: x has one row per example (z, pooled encodings): exact-fp32 row kernel instead of the tensor-core GEMM."""
    return LinearFn.apply(_anchor(x.device), x, _P(weight), _P(bias), act, cols, addend, rowwise)


# --------------------------------------------------------------------------------------------
def _rl_prob(xsegs, w, bias, out, act=ACT_NONE, y=None, yact=ACT_NONE, xt_out=None):
    """One VmmtRowLin problem: x = [xsegs...] column-wise, forward (w [N,K]) or gradient (w [K,N]) form."""
    q = L.RowLin()
    k0 = 0
    assert 1 <= len(xsegs) <= 3
    for i, t in enumerate(xsegs):
        assert t.dim() == 2 and t.stride(1) == 1
        q.seg[i].p, q.seg[i].ld, q.seg[i].k0 = fptr(t), t.stride(0), k0
        k0 += t.shape[1]
    q.nseg, q.act = len(xsegs), act
    q.y, q.ldy, q.yact = fptr(y), (y.stride(0) if y is not None else 0), yact
    q.w, q.ldw, q.bias = fptr(w), w.stride(0), fptr(bias)
    q.out, q.ldo = fptr(out), out.stride(0)
    q.xt_out, q.ld_xt = fptr(xt_out), (xt_out.stride(0) if xt_out is not None else 0)
    return q


def rowlin(probs, M, N, K, sum_outputs=False, transposed=False):
    arr = (L.RowLin * len(probs))(*probs)
    L.call("vmmt_rowlin", arr, len(probs), int(sum_outputs), int(transposed), M, N, K, stream())


class RowMLPFn(Function):
    """One or two 2-layer MLP heads over the same batch rows, exact fp32 (csrc/rowlin.cu):
    y_p = act_p(relu(x W1_p^T + b1_p) W2_p^T + b2_p) -- LocationLayer / ScaleLayer pairs of the prior, posterior and image
    networks (onmt/modules/NormalVariationalEncoder.py:12-43).  x = [xsegs...] column-wise, never materialised.

    forward(anchor, params = _P((W1, b1, W2, b2) per head), acts, *xsegs) -> y_0 [, y_1]
    Two launches forward (layer 1 of all heads, layer 2 of all heads), two backward on the critical path (input
    gradients); weight / bias gradients go to the side stream through the tensor-core GEMM."""

    @staticmethod
    def forward(ctx, anchor, params, acts, *xsegs):
        heads = params.t
        nh = len(heads)
        xs = [x if (x.stride(1) == 1) else x.contiguous() for x in xsegs]
        M = xs[0].shape[0]
        K = sum(x.shape[1] for x in xs)
        N1, N2 = heads[0][0].shape[0], heads[0][2].shape[0]
        for W1, b1, W2, b2 in heads:
            assert W1.shape == (N1, K) and W2.shape == (N2, N1)
        dev = xs[0].device
        hs = [torch.empty(M, N1, device=dev, dtype=torch.float32) for _ in range(nh)]
        ys = [torch.empty(M, N2, device=dev, dtype=torch.float32) for _ in range(nh)]
        rowlin([_rl_prob(xs, W1, b1, h, ACT_RELU) for (W1, b1, _, _), h in zip(heads, hs)], M, N1, K)
        rowlin([_rl_prob([h], W2, b2, y, a) for (_, _, W2, b2), h, y, a in zip(heads, hs, ys, acts)], M, N2, N1)
        ctx.save_for_backward(*xs, *hs, *ys)
        ctx.heads, ctx.acts, ctx.nx = heads, acts, len(xs)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        heads, acts, nx = ctx.heads, ctx.acts, ctx.nx
        nh = len(heads)
        saved = ctx.saved_tensors
        xs, hs, ys = saved[:nx], saved[nx:nx + nh], saved[nx + nh:]
        M = xs[0].shape[0]
        K = sum(x.shape[1] for x in xs)
        N1, N2 = heads[0][0].shape[0], heads[0][2].shape[0]
        dev = xs[0].device
        dys = [torch.zeros_like(y) if d is None else d.contiguous() for d, y in zip(dys, ys)]
        # layer 2: dh_p = (dy_p * act'(y_p)) W2_p; the transformed dy_p (= d pre-activation) is kept for dW2 / db2
        dpre2 = [dy if a == ACT_NONE else torch.empty(M, N2, device=dev, dtype=torch.float32) for dy, a in zip(dys, acts)]
        dhs = [torch.empty(M, N1, device=dev, dtype=torch.float32) for _ in range(nh)]
        rowlin([_rl_prob([dy], W2, None, dh, y=(y if a != ACT_NONE else None), yact=a, xt_out=(dp if a != ACT_NONE else None))
                for (_, _, W2, _), dy, dh, y, a, dp in zip(heads, dys, dhs, ys, acts, dpre2)],
               M, N1, N2, transposed=True)
        # layer 1: dx = sum_p (dh_p * relu'(h_p)) W1_p over the column range some input needs
        need = [bool(ctx.needs_input_grad[3 + i]) for i in range(nx)]
        dpre1 = [torch.empty(M, N1, device=dev, dtype=torch.float32) for _ in range(nh)]
        dxs = [None] * nx
        if any(need):
            offs = [0]
            for x in xs:
                offs.append(offs[-1] + x.shape[1])
            c0 = min(offs[i] for i in range(nx) if need[i])
            c1 = max(offs[i + 1] for i in range(nx) if need[i])
            dx = torch.empty(M, c1 - c0, device=dev, dtype=torch.float32)
            rowlin([_rl_prob([dh], W1[:, c0:c1], None, dx, y=h, yact=ACT_RELU, xt_out=dp)
                    for (W1, _, _, _), dh, h, dp in zip(heads, dhs, hs, dpre1)],
                   M, c1 - c0, N1, sum_outputs=(nh == 2), transposed=True)
            for i in range(nx):
                if need[i]:
                    dxs[i] = dx[:, offs[i] - c0: offs[i + 1] - c0]
        else:
            for dh, h, dp in zip(dhs, hs, dpre1):
                L.call("vmmt_act_bwd", fptr(dh), fptr(h), fptr(dp), dh.numel(), ACT_RELU, stream())
        with on_side(*dpre1, *dpre2, *hs, *xs):                               # they only feed the optimiser
            for (W1, b1, W2, b2), h, d1, d2 in zip(heads, hs, dpre1, dpre2):
                if W2.requires_grad:
                    gemm(d2, h, grad_buf(W2), N2, N1, M, a_kmajor=False, b_kmajor=False, accumulate=1)
                if b2 is not None and b2.requires_grad:
                    colsum_acc(d2, M, N2, grad_buf(b2))
                if W1.requires_grad:
                    k0 = 0
                    for x in xs:
                        gemm(d1, x, grad_buf(W1)[:, k0:k0 + x.shape[1]], N1, x.shape[1], M, a_kmajor=False,
                             b_kmajor=False, accumulate=1)
                        k0 += x.shape[1]
                if b1 is not None and b1.requires_grad:
                    colsum_acc(d1, M, N1, grad_buf(b1))
        return (None, None, None, *dxs)


def row_mlp(xsegs, heads, acts):
    """heads: [(W1, b1, W2, b2), ...] (1 or 2), acts: output activation per head -> tuple of outputs."""
    xsegs = [x.reshape(-1, x.shape[-1]) for x in xsegs]
    return RowMLPFn.apply(_anchor(xsegs[0].device), _P(tuple(tuple(h) for h in heads)), tuple(acts), *xsegs)



# --------------------------------------------------------------------------------------------
class LSTMLayerFn(Function):
    """One nn.LSTM layer, one or two directions (onmt/Models.py:124-149,892-893; VI_Model1.py:106).

    forward(x [T,N,In], h0, c0 [ndir,N,Hd] | None, rowbias [N,4H] | None, lengths | None, cfg,
            *weights)  with weights = (w_ih, w_hh, b_ih, b_hh) per direction and
    cfg = dict(in_cols=(c0,c1) | None, save=bool)  ->  out [T,N,ndir*Hd], hT, cT [ndir,N,Hd]
    """

    @staticmethod
    def forward(ctx, anchor, x, h0, c0, rowbias, lengths, cfg, weights):
        weights = weights.t
        ndir = len(weights) // 4
        T, N, In = x.shape
        Hd = weights[1].shape[1]
        dev = x.device
        x = x.contiguous()
        save = cfg.get("save", True)
        gx_given = bool(cfg.get("gx_given"))     # x IS the input projection x W_ih^T (computed earlier, off the critical path)
        c0c, c1c = cfg.get("in_cols") or (0, weights[0].shape[1])
        if gx_given:
            assert ndir == 1 and In == 4 * Hd, "precomputed gate pre-activations: one direction, [T,N,4H]"
            gx = x.view(1, T, N, 4 * Hd)
        else:
            assert c1c - c0c == In
            gx = torch.empty(ndir, T, N, 4 * Hd, device=dev, dtype=torch.float32)
        out = torch.empty(T, N, ndir * Hd, device=dev, dtype=torch.float32)
        hT = torch.empty(ndir, N, Hd, device=dev, dtype=torch.float32)
        cT = torch.empty(ndir, N, Hd, device=dev, dtype=torch.float32)
        gates = torch.empty(ndir, T, N, 4 * Hd, device=dev, dtype=torch.float32) if save else None
        cs = torch.empty(ndir, T, N, Hd, device=dev, dtype=torch.float32) if save else None
        if h0 is not None:
            h0, c0 = h0.contiguous(), c0.contiguous()
        if rowbias is not None:
            rowbias = rowbias.contiguous()
        dirs = (L.LstmDir * ndir)()
        x2 = x.view(T * N, In)
        # input projections: the reverse direction's GEMM runs beside the forward direction's on a second stream (each is
        # ~80 tiles: together they fill the SMs; back to back they were 2 x 22 us at the head of the target-encoder chain)
        fork = None
        if ndir == 2 and not gx_given and x.is_cuda and _branch["enabled"]:
            cur = torch.cuda.current_stream(dev)
            fork = _gx_stream(dev)
            fork.wait_stream(cur)
            with torch.cuda.stream(fork):
                gemm(x2, weights[4][:, c0c:c1c], gx[1].view(T * N, 4 * Hd), T * N, 4 * Hd, In)
        for d in range(ndir):
            w_ih, w_hh, b_ih, b_hh = weights[4 * d: 4 * d + 4]
            if not gx_given and not (d == 1 and fork is not None):
                gemm(x2, w_ih[:, c0c:c1c], gx[d].view(T * N, 4 * Hd), T * N, 4 * Hd, In)
            D = dirs[d]
            D.gx, D.w_hh, D.b_ih, D.b_hh = fptr(gx[d]), fptr(w_hh), fptr(b_ih), fptr(b_hh)
            D.rowbias = fptr(rowbias)
            D.h0 = fptr(h0[d]) if h0 is not None else None
            D.c0 = fptr(c0[d]) if c0 is not None else None
            D.out = out.data_ptr() + 4 * d * Hd
            D.out_ld = ndir * Hd
            D.hT, D.cT = fptr(hT[d]), fptr(cT[d])
            D.gates = fptr(gates[d]) if save else None
            D.cs = fptr(cs[d]) if save else None
            D.reverse = 1 if d == 1 else 0
        if fork is not None:
            torch.cuda.current_stream(dev).wait_stream(fork)
        ws_bytes = L.lib.vmmt_lstm_workspace_bytes(ndir, N, Hd)
        ws = torch.empty(ws_bytes // 4, device=dev, dtype=torch.float32)
        L.call("vmmt_lstm_seq_fwd", dirs, ndir, ptr(lengths), T, N, Hd, flags(), int(cfg.get("cluster_budget") or 0),
               fptr(ws), ws_bytes, stream())
        if save:
            ctx.save_for_backward(None if gx_given else x, out, gates, cs, h0, c0, rowbias, lengths)
            ctx.weights, ctx.cfg = weights, (ndir, T, N, In, Hd, c0c, c1c)
            ctx.gx_given = gx_given
            ctx.cluster_budget = int(cfg.get("cluster_budget_bwd") or cfg.get("cluster_budget") or 0)
            ctx.side_lane = cfg.get("side_lane")
            ctx.fires_early_exchange = bool(cfg.get("fires_early_exchange"))
        ctx.set_materialize_grads(False)
        return out, hT, cT

    @staticmethod
    def backward(ctx, dout, dhT, dcT):
        x, out, gates, cs, h0, c0, rowbias, lengths = ctx.saved_tensors
        weights = ctx.weights
        ndir, T, N, In, Hd, c0c, c1c = ctx.cfg
        # encoder stacks (data parallel): everything issued before this node is what the early gradient exchange needs; the
        # exchange itself is launched AFTER this node's recurrence kernel, so that the recurrence's clusters are placed first
        early_ev = torch.cuda.current_stream().record_event() if (ctx.fires_early_exchange and _early_cb[0] is not None) \
            else None
        gx_given = ctx.gx_given
        dev = out.device
        dout = dout.contiguous() if dout is not None else None
        dhT = dhT.contiguous() if dhT is not None else None
        dcT = dcT.contiguous() if dcT is not None else None
        dg = torch.empty(ndir, T, N, 4 * Hd, device=dev, dtype=torch.float32)
        need_h0 = h0 is not None and ctx.needs_input_grad[2]
        dh0 = torch.empty(ndir, N, Hd, device=dev, dtype=torch.float32) if need_h0 else None
        dc0 = torch.empty(ndir, N, Hd, device=dev, dtype=torch.float32) if need_h0 else None
        dirs = (L.LstmDirBwd * ndir)()
        fused_bias = bool(L.lib.vmmt_lstm_seq_bwd_fuses_bias(ndir, N, Hd, flags()))
        need_drow = rowbias is not None and ctx.needs_input_grad[4]
        # (the per-example term enters every direction: only the one-direction case is summed inside the kernel)
        drow = torch.empty(N, 4 * Hd, device=dev, dtype=torch.float32) if (need_drow and fused_bias and ndir == 1) else None
        for d in range(ndir):
            D = dirs[d]
            D.w_hh, D.gates, D.cs = fptr(weights[4 * d + 1]), fptr(gates[d]), fptr(cs[d])
            D.c0 = fptr(c0[d]) if c0 is not None else None
            D.dout = (dout.data_ptr() + 4 * d * Hd) if dout is not None else None
            D.dout_ld = ndir * Hd
            D.dhT = fptr(dhT[d]) if dhT is not None else None
            D.dcT = fptr(dcT[d]) if dcT is not None else None
            D.dgates = fptr(dg[d])
            D.dh0 = fptr(dh0[d]) if need_h0 else None
            D.dc0 = fptr(dc0[d]) if need_h0 else None
            if fused_bias:                                # the recurrence kernel adds sum_{t,n} dG into the bias gradients
                b_ih, b_hh = weights[4 * d + 2], weights[4 * d + 3]
                D.db_ih = fptr(grad_buf(b_ih)) if b_ih.requires_grad else None
                D.db_hh = fptr(grad_buf(b_hh)) if b_hh.requires_grad else None
                D.drow = fptr(drow) if drow is not None else None
            D.reverse = 1 if d == 1 else 0
        ws_bytes = L.lib.vmmt_lstm_workspace_bytes(ndir, N, Hd)
        ws = torch.empty(ws_bytes // 4, device=dev, dtype=torch.float32)
        L.call("vmmt_lstm_seq_bwd", dirs, ndir, ptr(lengths), T, N, Hd, flags(), ctx.cluster_budget, fptr(ws), ws_bytes,
               stream())
        if early_ev is not None:
            _fire_early_exchange(early_ev)
        x2 = x.view(T * N, In) if not gx_given else None
        dx = torch.empty(T * N, In, device=dev, dtype=torch.float32) \
            if (ctx.needs_input_grad[1] and not gx_given) else None
        # critical path (main stream): what the layers below / the callers wait for
        for d in range(ndir):
            w_ih = weights[4 * d]
            if dx is not None:                                             # dx (+)= dG W_ih
                gemm(dg[d].view(T * N, 4 * Hd), w_ih[:, c0c:c1c], dx, T * N, In, 4 * Hd, b_kmajor=False,
                     accumulate=int(d > 0))
        if need_drow and drow is None:
            drow = torch.zeros(N, 4 * Hd, device=dev, dtype=torch.float32)
            for d in range(ndir):                                          # the term enters every direction
                colsum_acc(dg[d].view(T, N * 4 * Hd), T, N * 4 * Hd, drow.view(-1))
        # weight / bias gradients (side stream): they only feed the optimiser
        for d in range(ndir):                                                 # one block (= one lane) per direction
            with on_side(dg, x, out, h0, lane=None if ctx.side_lane is None else ctx.side_lane + d):
                w_ih, w_hh, b_ih, b_hh = weights[4 * d: 4 * d + 4]
                dg2 = dg[d].view(T * N, 4 * Hd)
                if w_ih.requires_grad and not gx_given:                    # dW_ih += dG^T x
                    gemm(dg2, x2, grad_buf(w_ih)[:, c0c:c1c], 4 * Hd, In, T * N, a_kmajor=False,
                         b_kmajor=False, accumulate=1)
                if w_hh.requires_grad and T > 1:                           # dW_hh += dG[t]^T h[t -/+ 1]
                    o_d = out[:, :, d * Hd:(d + 1) * Hd]
                    if d == 0:
                        a_, b_ = dg[d][1:], o_d[:-1]
                    else:
                        a_, b_ = dg[d][:-1], o_d[1:]
                    # rows (t,n) of the shifted views are contiguous blocks with row strides 4Hd / ndir*Hd
                    b2 = b_.reshape((T - 1) * N, Hd) if ndir == 1 else _rows(b_)
                    if b2.data_ptr() % 16 or b2.stride(0) % 4:
                        # the reverse direction's half of a bidirectional output starts Hd floats into each row: when
                        # that is not 16-byte aligned the TMA cannot address it -- stage it in a padded buffer (a 1 MB
                        # copy) instead of dropping to the SIMT kernel
                        pad = torch.empty((T - 1) * N, (Hd + 3) // 4 * 4, device=dev, dtype=torch.float32)
                        pad[:, :Hd].copy_(b2)
                        b2 = pad[:, :Hd]
                    gemm(a_.reshape(-1, 4 * Hd), b2, grad_buf(w_hh), 4 * Hd, Hd, (T - 1) * N, a_kmajor=False,
                         b_kmajor=False, accumulate=1)
                if w_hh.requires_grad and h0 is not None:                  # first step uses h0
                    t0 = 0 if d == 0 else T - 1
                    gemm(dg[d][t0], h0[d], grad_buf(w_hh), 4 * Hd, Hd, N, a_kmajor=False, b_kmajor=False,
                         accumulate=1)
                if fused_bias:
                    pass                                                   # done inside the recurrence kernel
                elif b_ih.requires_grad and b_hh.requires_grad:            # identical sums: one pass, two outputs
                    colsum_acc(dg2, T * N, 4 * Hd, grad_buf(b_ih), grad_buf(b_hh))
                elif b_ih.requires_grad or b_hh.requires_grad:
                    colsum_acc(dg2, T * N, 4 * Hd, grad_buf(b_ih if b_ih.requires_grad else b_hh))
        if dx is not None:
            dx = dx.view(T, N, In)
        if gx_given and ctx.needs_input_grad[1]:
            dx = dg[0]                 # d(gate pre-activations): the node that produced gx takes dW_ih / dx from it
        return None, dx, dh0, dc0, drow, None, None, None


def lstm_layer(x, h0, c0, rowbias, lengths, cfg, weights):
    """weights: (w_ih, w_hh, b_ih, b_hh) per direction."""
    return LSTMLayerFn.apply(_anchor(x.device), x, h0, c0, rowbias, lengths, cfg, _P(tuple(weights)))


def _rows(v):
    """[T,N,Hd] slice of a [T,N,ndir*Hd] buffer viewed as (T*N) rows with the parent's row stride."""
    T, N, Hd = v.shape
    return v.as_strided((T * N, Hd), (v.stride(1), 1), v.storage_offset())


class AttnJoinFn(Function):
    """Identity on the attention context.  Its backward runs right after AttentionCoreFn.backward (it consumes that node's
    context gradient) and makes the stream wait for the context-side kernel that AttentionCoreFn forked onto another
    stream: the decoder's backward chain (query side -> linear_in backward -> recurrence) does not queue behind a kernel
    only the encoders' backward needs."""

    @staticmethod
    def forward(ctx, context, token):
        ctx.token = token
        return context.view_as(context)

    @staticmethod
    def backward(ctx, g):
        ev = ctx.token.pop("event", None)
        if ev is not None:
            torch.cuda.current_stream(g.device).wait_event(ev)
        ctx.token.pop("keep", None)                       # the forked kernel's operands may be recycled from here on
        return g, None


def attention_context(context):
    """-> (context', token) for AttentionCoreFn.apply(qp, context', lengths, token); token is None when nothing is forked."""
    if context.is_cuda and context.requires_grad and torch.is_grad_enabled() and _branch["enabled"]:
        token = {}
        return AttnJoinFn.apply(context, token), token
    return context, None


_attn_fork = {}


class AttentionCoreFn(Function):
    """scores + masked softmax + context (onmt/modules/GlobalAttention.py:108-113,169-184)."""

    @staticmethod
    def forward(ctx, qp, context, lengths, token=None):
        qp, context = qp.contiguous(), context.contiguous()
        T, B, H = qp.shape
        S = context.shape[0]
        align = torch.empty(T, B, S, device=qp.device, dtype=torch.float32)
        cvec = torch.empty(T, B, H, device=qp.device, dtype=torch.float32)
        L.call("vmmt_attention_fwd", fptr(qp), fptr(context), ptr(lengths), fptr(align), fptr(cvec),
               T, B, S, H, stream())
        ctx.save_for_backward(qp, context, align, lengths)
        ctx.token = token
        ctx.mark_non_differentiable(align)
        return cvec, align

    @staticmethod
    def backward(ctx, dcvec, _dalign):
        qp, context, align, lengths = ctx.saved_tensors
        T, B, H = qp.shape
        S = context.shape[0]
        dev = qp.device
        dcvec = dcvec.contiguous()
        ds = torch.empty(T, B, S, device=dev, dtype=torch.float32)
        dqp = torch.empty_like(qp)
        dctx = torch.empty_like(context)
        token = ctx.token
        if token is None or not ctx.needs_input_grad[1]:
            L.call("vmmt_attention_bwd", fptr(dcvec), fptr(qp), fptr(context), fptr(align), ptr(lengths),
                   fptr(ds), fptr(dqp), fptr(dctx), 0, T, B, S, H, stream())
            return dqp, dctx, None, None
        # query side on this stream (the chain continues with it); context side on a forked stream, joined by AttnJoinFn
        L.call("vmmt_attention_bwd_query", fptr(dcvec), fptr(context), fptr(align), ptr(lengths), fptr(ds), fptr(dqp),
               T, B, S, H, stream())
        cur = torch.cuda.current_stream(dev)
        fork = _attn_fork.get(dev)
        if fork is None:
            fork = _attn_fork[dev] = torch.cuda.Stream(device=dev, priority=int(os.environ.get("VMMT_BRANCH_PRIO", "-1")))
        fork.wait_stream(cur)
        with torch.cuda.stream(fork):
            L.call("vmmt_attention_bwd_ctx", fptr(dcvec), fptr(qp), fptr(align), fptr(ds), fptr(dctx), 0, T, B, S, H, stream())
            ev = torch.cuda.Event()
            ev.record(fork)
        token["event"] = ev
        token["keep"] = (dcvec, qp, align, ds, dctx)       # read / written on the forked stream until the join
        return dqp, dctx, None, None


class MaskedMeanFn(Function):
    """GlobalInferenceNetwork.encode_seq (onmt/modules/NormalVariationalEncoder.py:65-84).  ``x`` may be time-major
    contiguous or the transposed view of a contiguous [B,T,H] tensor (the target encoder's output): the kernels take the
    two strides, and the gradient comes back in the layout of the input, so neither direction needs a transposing copy."""

    @staticmethod
    def _layout(x):
        T, B, H = x.shape
        if x.stride(2) == 1 and x.stride(0) == H and x.stride(1) == T * H and T > 1 and B > 1:
            return "bt"                                   # transposed view of [B,T,H]
        return "tb"

    @staticmethod
    def forward(ctx, x, lengths):
        lay = MaskedMeanFn._layout(x)
        if lay == "tb":
            x = x.contiguous()
        T, B, H = x.shape
        st, sb = (B * H, H) if lay == "tb" else (H, T * H)
        out = torch.empty(B, H, device=x.device, dtype=torch.float32)
        L.call("vmmt_masked_mean_fwd", fptr(x), st, sb, ptr(lengths), fptr(out), H, T, B, H, stream())
        ctx.save_for_backward(lengths)
        ctx.dims = (T, B, H, lay)
        return out

    @staticmethod
    def backward(ctx, dout):
        (lengths,) = ctx.saved_tensors
        T, B, H, lay = ctx.dims
        dout = dout.contiguous()
        if lay == "tb":
            dx = torch.empty(T, B, H, device=dout.device, dtype=torch.float32)
            st, sb = B * H, H
        else:
            base = torch.empty(B, T, H, device=dout.device, dtype=torch.float32)
            dx = base.transpose(0, 1)
            st, sb = H, T * H
        L.call("vmmt_masked_mean_bwd", fptr(dout), H, ptr(lengths), dx.data_ptr(), st, sb, 0, T, B, H, stream())
        return dx, None


# ---- phase stamps (measurement aid, VMMT_STAMPS=1): %globaltimer written by one-thread kernels at chosen points of the
# forward pass and, through StampFn's backward, of the backward pass -- phase boundaries of a replayed step with no profiler
# attached (tools/phase_stamps.py)
_stamps = {"on": os.environ.get("VMMT_STAMPS") == "1", "buf": None}


def stamp_buffer():
    if _stamps["buf"] is None:
        _stamps["buf"] = torch.zeros(64, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
    return _stamps["buf"]


def stamp(slot):
    if _stamps["on"]:
        L.call("vmmt_stamp", stamp_buffer().data_ptr(), int(slot), stream())


class StampFn(Function):
    @staticmethod
    def forward(ctx, x, slot_fwd, slot_bwd):
        stamp(slot_fwd)
        ctx.slot_bwd = slot_bwd
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        stamp(ctx.slot_bwd)
        return g, None, None


def stamped(x, slot_fwd, slot_bwd):
    """Identity; records when ``x`` is ready (forward) and when its gradient is ready (backward)."""
    return StampFn.apply(x, slot_fwd, slot_bwd) if (_stamps["on"] and x.requires_grad) else x


_dropout_log = [None]


def set_dropout_log(sink):
    """Parity tests: ``sink`` (a list) receives (shape, p, seed, offset) of every dropout call, in call order, so that the
    very masks the kernels drew can be regenerated (``dropout_mask``) and injected into the oracle."""
    _dropout_log[0] = sink


def dropout_mask(shape, p, seed, offset):
    """The scaled keep-mask (0 or 1/(1-p)) the dropout kernel draws for (seed, offset) at the current step base."""
    ones = torch.ones(*shape, device=torch.device("cuda", torch.cuda.current_device()), dtype=torch.float32)
    m = torch.empty_like(ones)
    L.call("vmmt_dropout", fptr(ones), fptr(m), ones.numel(), float(p), seed, offset, _base_ptr(), stream())
    return m


class DropoutFn(Function):
    """Inverted dropout with an in-kernel Philox mask, regenerated (not stored) in backward."""

    @staticmethod
    def forward(ctx, x, p):
        x = x.contiguous()
        y = torch.empty_like(x)
        ctx.p, ctx.seed, ctx.offset = p, _seed_state["seed"], _next_offset()
        if _dropout_log[0] is not None:
            _dropout_log[0].append((tuple(x.shape), p, ctx.seed, ctx.offset))
        L.call("vmmt_dropout", fptr(x), fptr(y), x.numel(), p, ctx.seed, ctx.offset, _base_ptr(), stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        L.call("vmmt_dropout", fptr(dy), fptr(dx), dy.numel(), ctx.p, ctx.seed, ctx.offset, _base_ptr(), stream())
        return dx, None


def dropout(x, p, training):
    if not training or p <= 0.0:
        return x
    return DropoutFn.apply(x, float(p))


def normal_sample(mu, sd, eps=None):
    """z = mu + sd*eps without a pathwise gradient (torch.normal semantics, onmt/modules/Dists.py:21-26)."""
    with torch.no_grad():
        mu_c, sd_c = mu.detach().contiguous(), sd.detach().contiguous()
        z = torch.empty_like(mu_c)
        L.call("vmmt_normal_sample", fptr(mu_c), fptr(sd_c), fptr(eps.contiguous()) if eps is not None else None,
               fptr(z), z.numel(), _seed_state["seed"], _next_offset(), _base_ptr(), stream())
    return z


class GateFn(Function):
    """gated = z * sigmoid(z.w + b); z carries no gradient (hazard H2)
    (onmt/modules/NormalVariationalEncoder.py:286-299)."""

    @staticmethod
    def forward(ctx, anchor, z, weight, bias):
        weight, bias = weight.t, bias.t
        z = z.contiguous()
        B, Z = z.shape
        gate = torch.empty(B, device=z.device, dtype=torch.float32)
        gated = torch.empty_like(z)
        L.call("vmmt_gate_fwd", fptr(z), fptr(weight), fptr(bias), fptr(gate), fptr(gated), B, Z, stream())
        ctx.save_for_backward(z, gate)
        ctx.weight, ctx.bias = weight, bias
        return gated

    @staticmethod
    def backward(ctx, dgated):
        z, gate = ctx.saved_tensors
        B, Z = z.shape
        dgated = dgated.contiguous()
        dpre = torch.empty(B, device=z.device, dtype=torch.float32)
        L.call("vmmt_gate_bwd", fptr(dgated), fptr(z), fptr(gate), fptr(dpre), fptr(grad_buf(ctx.weight)),
               fptr(grad_buf(ctx.bias)), B, Z, stream())
        return None, None, None, None


def gate(z, weight, bias):
    return GateFn.apply(_anchor(z.device), z, _P(weight), _P(bias))


_loss_stream = None


class VILossFn(Function):
    """NLL + image log-prob + KL in one node (onmt/VILoss.py:217-513).

    forward(out2d [M,H], target [M], gen_w, gen_b, mu_q, sd_q, mu_p, sd_p, img_loc, img_v, cfg) ->
      loss [1] = NLL - IMG + kl_weight*KL,  stats [8] = {nll, n_words, n_correct, kl, img_logprob,
      img_cos, 0, 0}.  cfg: pad_idx, kl_weight, legacy_image_grad.
    """

    @staticmethod
    def forward(ctx, out2d, target, gen_w, gen_b, mu_q, sd_q, mu_p, sd_p, img_loc, img_v, cfg):
        gen_w, gen_b = gen_w.t, gen_b.t
        dev = out2d.device
        out2d = out2d.contiguous()
        target = target.contiguous()
        M, H = out2d.shape
        V = gen_w.shape[0]
        stats = torch.zeros(8, device=dev, dtype=torch.float32)
        lse = torch.empty(M, device=dev, dtype=torch.float32)
        wsb = L.lib.vmmt_generator_workspace_bytes(M, H, V)
        ws = torch.empty(wsb // 4, device=dev, dtype=torch.float32)
        mu_q, sd_q = mu_q.contiguous(), sd_q.contiguous()
        B, Z = mu_q.shape
        if mu_p is not None:
            mu_p, sd_p = mu_p.contiguous(), sd_p.contiguous()
        img_loc, img_v = img_loc.contiguous(), img_v.contiguous()
        D = img_loc.shape[1]
        rowstats = torch.empty(B, 4, device=dev, dtype=torch.float32)
        # KL and image terms do not depend on the generator: they run on their own stream beside its GEMM
        cur = torch.cuda.current_stream(dev)
        global _loss_stream
        if _loss_stream is None or _loss_stream.device != dev:
            _loss_stream = torch.cuda.Stream(device=dev, priority=-1)
        _loss_stream.wait_stream(cur)
        with torch.cuda.stream(_loss_stream):
            L.call("vmmt_kl_fwd", fptr(mu_q), fptr(sd_q), fptr(mu_p), fptr(sd_p), fptr(stats[3:4]), B, Z, stream())
            L.call("vmmt_image_loss_fwd", fptr(img_loc), fptr(img_v), fptr(rowstats), fptr(stats[4:6]), B, D, stream())
        L.call("vmmt_generator_nll_fwd", fptr(out2d), fptr(gen_w), fptr(gen_b), ptr(target), cfg["pad_idx"],
               M, H, V, fptr(lse), fptr(stats[0:3]), fptr(ws), wsb, flags(), stream())
        cur.wait_stream(_loss_stream)
        kw = float(cfg["kl_weight"])
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        L.call("vmmt_loss_finalize", fptr(stats), kw, fptr(loss), stream())     # also stats[6] = kw * kl, stats[7] = loss
        ctx.save_for_backward(out2d, target, lse, mu_q, sd_q, mu_p, sd_p, img_loc, img_v, rowstats)
        ctx.gen_w, ctx.gen_b, ctx.cfg = gen_w, gen_b, cfg
        ctx.mark_non_differentiable(stats)
        return loss, stats

    @staticmethod
    def backward(ctx, dloss, _dstats):
        out2d, target, lse, mu_q, sd_q, mu_p, sd_p, img_loc, img_v, rowstats = ctx.saved_tensors
        cfg = ctx.cfg
        dev = out2d.device
        M, H = out2d.shape
        V = ctx.gen_w.shape[0]
        gs = dloss.reshape(1).contiguous()                     # device scalar: no host sync
        # KL / image gradients first, on the loss stream: they are tiny and independent of the generator's backward, and the
        # decoder-output gradient below must not queue behind them (it is the backward pass's critical path)
        B, Z = mu_q.shape
        dmq, dsq = torch.empty_like(mu_q), torch.empty_like(sd_q)
        dmp = torch.empty_like(mu_q) if mu_p is not None else None
        dsp = torch.empty_like(mu_q) if mu_p is not None else None
        D = img_loc.shape[1]
        dloc = torch.empty_like(img_loc)
        cur = torch.cuda.current_stream(dev)
        global _loss_stream
        if _loss_stream is None or _loss_stream.device != dev:
            _loss_stream = torch.cuda.Stream(device=dev, priority=-1)
        _loss_stream.wait_stream(cur)
        with torch.cuda.stream(_loss_stream):
            L.call("vmmt_kl_bwd", fptr(mu_q), fptr(sd_q), fptr(mu_p), fptr(sd_p), fptr(dmq), fptr(dsq), fptr(dmp),
                   fptr(dsp), fptr(gs), float(cfg["kl_weight"]), B, Z, stream())
            L.call("vmmt_image_loss_bwd", fptr(img_loc), fptr(img_v), fptr(rowstats), fptr(dloc), fptr(gs), 1.0,
                   int(cfg.get("legacy_image_grad", True)), B, D, stream())
        dx = torch.empty(M, H, device=dev, dtype=torch.float32)
        wsb = L.lib.vmmt_generator_workspace_bytes(M, H, V)
        ws = torch.empty(wsb // 4, device=dev, dtype=torch.float32)
        L.call("vmmt_generator_nll_bwd", fptr(out2d), fptr(ctx.gen_w), fptr(ctx.gen_b), ptr(target),
               cfg["pad_idx"], fptr(lse), fptr(gs), 1.0, M, H, V, fptr(dx), None, None, fptr(ws), wsb, flags(), stream())
        with on_side(ws, out2d):                               # generator weight gradient: off the critical path
            # the one weight-gradient GEMM with more tiles than SMs (V/128 x H/128): it starts when the decoder-output
            # gradient is done, i.e. together with the short kernels that lead to the decoder's backward recurrence
            # (measured: with every SM holding one of its CTAs those kernels waited 38 us for the first wave to retire)
            # (a cfg5-sized product, 2.6 TFLOP, keeps the one-tile-per-CTA background form: it would hold its share for ms)
            share = L.F_SHARE_SMS if (2.0 * M * H * V < 5e10 and os.environ.get("VMMT_GEN_WGRAD_SHARE", "1") != "0") else 0
            L.call("vmmt_generator_nll_wgrad", fptr(out2d), fptr(ws), M, H, V, fptr(grad_buf(ctx.gen_w)),
                   fptr(grad_buf(ctx.gen_b)), flags() | share, stream())
        cur.wait_stream(_loss_stream)
        return dx, None, None, None, dmq, dsq, dmp, dsp, dloc, None, None


def vi_loss(out2d, target, gen_w, gen_b, mu_q, sd_q, mu_p, sd_p, img_loc, img_v, cfg):
    return VILossFn.apply(out2d, target, _P(gen_w), _P(gen_b), mu_q, sd_q, mu_p, sd_p, img_loc, img_v, cfg)


def generator_logprobs(x2d, weight, bias):
    """log_softmax(x W^T + b) materialised [M,V] (decode: onmt/translate/TranslatorMultimodalVI.py:199)."""
    x2d = x2d.contiguous()
    M, H = x2d.shape
    V = weight.shape[0]
    out = torch.empty(M, V, device=x2d.device, dtype=torch.float32)
    lse = torch.empty(M, device=x2d.device, dtype=torch.float32)
    L.call("vmmt_generator_logprobs", fptr(x2d), fptr(weight), fptr(bias), M, H, V, fptr(out), fptr(lse), flags(), stream())
    return out
