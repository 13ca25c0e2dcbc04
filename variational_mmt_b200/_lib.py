"""ctypes binding of libvmmt.so (C ABI: include/vmmt.h).

The library is built in-tree by ``variational_mmt_b200/build.py`` (nvcc, sm_100a).  There is no
fallback: if the shared object is missing, importing this module raises; if a kernel call fails,
``RuntimeError(vmmt_last_error())`` is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvmmt.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first "
        "(python -m variational_mmt_b200.build, or __graft_entry__.build()). "
        "variational_mmt_b200 has no CPU / PyTorch fallback.")

lib = C.CDLL(LIB_PATH)

P, I, L, F, SZ, U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint64


class LstmDir(C.Structure):
    _fields_ = [("gx", P), ("w_hh", P), ("b_ih", P), ("b_hh", P), ("rowbias", P), ("h0", P), ("c0", P),
                ("out", P), ("out_ld", L), ("hT", P), ("cT", P), ("gates", P), ("cs", P),
                ("reverse", C.c_int32), ("pad_", C.c_int32)]


class LstmDirBwd(C.Structure):
    _fields_ = [("w_hh", P), ("gates", P), ("cs", P), ("c0", P), ("dout", P), ("dout_ld", L),
                ("dhT", P), ("dcT", P), ("dgates", P), ("dh0", P), ("dc0", P), ("db_ih", P), ("db_hh", P), ("drow", P),
                ("reverse", C.c_int32), ("pad_", C.c_int32)]


class RowLinSeg(C.Structure):
    _fields_ = [("p", P), ("ld", L), ("k0", C.c_int32), ("pad_", C.c_int32)]


class RowLin(C.Structure):
    _fields_ = [("seg", RowLinSeg * 3), ("nseg", C.c_int32), ("act", C.c_int32), ("y", P), ("ldy", L),
                ("yact", C.c_int32), ("pad_", C.c_int32), ("w", P), ("ldw", L), ("bias", P), ("out", P), ("ldo", L),
                ("xt_out", P), ("ld_xt", L)]


# name -> (restype, argtypes); must list every symbol declared in include/vmmt.h
SIGNATURES = {
    "vmmt_last_error": (C.c_char_p, []),
    "vmmt_version": (I, []),
    "vmmt_launch_count": (C.c_ulonglong, []),
    "vmmt_gemm": (I, [P, L, I, P, L, I, P, L, I, I, I, P, I, I, I, P]),
    "vmmt_gemm_dual": (I, [P, L, P, L, I, P, L, P, L, I, P, L, I, I, P, I, I, P]),
    "vmmt_cast_bf16": (I, [P, L, P, L, I, I, P]),
    "vmmt_gemm_bf16": (I, [P, L, I, P, L, I, P, L, I, I, I, P, I, I, I, P]),
    "vmmt_embedding_fwd": (I, [P, L, P, L, I, P, P]),
    "vmmt_embedding_bwd": (I, [P, L, P, I, L, P, L, P]),
    "vmmt_lstm_workspace_bytes": (SZ, [I, I, I]),
    "vmmt_lstm_seq_supported": (I, [I, I, I]),
    "vmmt_lstm_seq_fwd": (I, [C.POINTER(LstmDir), I, P, I, I, I, I, I, P, SZ, P]),
    "vmmt_lstm_seq_bwd_fuses_bias": (I, [I, I, I, I]),
    "vmmt_lstm_seq_bwd": (I, [C.POINTER(LstmDirBwd), I, P, I, I, I, I, I, P, SZ, P]),
    "vmmt_lstm_cell_fwd": (I, [P, P, P, P, P, P, P, I, I, P]),
    "vmmt_attention_fwd": (I, [P, P, P, P, P, I, I, I, I, P]),
    "vmmt_attention_bwd": (I, [P, P, P, P, P, P, P, P, I, I, I, I, I, P]),
    "vmmt_attention_bwd_query": (I, [P, P, P, P, P, P, I, I, I, I, P]),
    "vmmt_attention_bwd_ctx": (I, [P, P, P, P, P, I, I, I, I, I, P]),
    "vmmt_masked_mean_fwd": (I, [P, L, L, P, P, L, I, I, I, P]),
    "vmmt_masked_mean_bwd": (I, [P, L, P, P, L, L, I, I, I, I, P]),
    "vmmt_act_bwd": (I, [P, P, P, L, I, P]),
    "vmmt_loss_finalize": (I, [P, F, P, P]),
    "vmmt_colsum_acc": (I, [P, L, I, I, P, P, P]),
    "vmmt_axpy": (I, [P, P, F, L, P]),
    "vmmt_counter_add": (I, [P, U64, P]),
    "vmmt_stamp": (I, [P, I, P]),
    "vmmt_dropout": (I, [P, P, L, F, U64, U64, P, P]),
    "vmmt_normal_sample": (I, [P, P, P, P, L, U64, U64, P, P]),
    "vmmt_kl_fwd": (I, [P, P, P, P, P, I, I, P]),
    "vmmt_kl_bwd": (I, [P, P, P, P, P, P, P, P, P, F, I, I, P]),
    "vmmt_rowlin": (I, [C.POINTER(RowLin), I, I, I, I, I, I, P]),
    "vmmt_gate_fwd": (I, [P, P, P, P, P, I, I, P]),
    "vmmt_gate_bwd": (I, [P, P, P, P, P, P, I, I, P]),
    "vmmt_image_loss_fwd": (I, [P, P, P, P, I, I, P]),
    "vmmt_image_loss_bwd": (I, [P, P, P, P, P, F, I, I, I, P]),
    "vmmt_generator_workspace_bytes": (SZ, [I, I, I]),
    "vmmt_generator_nll_fwd": (I, [P, P, P, P, L, I, I, I, P, P, P, SZ, I, P]),
    "vmmt_generator_nll_bwd": (I, [P, P, P, P, L, P, P, F, I, I, I, P, P, P, P, SZ, I, P]),
    "vmmt_generator_nll_wgrad": (I, [P, P, I, I, I, P, P, I, P]),
    "vmmt_generator_logprobs": (I, [P, P, P, I, I, I, P, P, I, P]),
    "vmmt_generator_topk_workspace_bytes": (SZ, [I, I, I]),
    "vmmt_generator_topk_supported": (I, [P, P, I, I, I, I]),
    "vmmt_generator_topk": (I, [P, P, P, I, I, I, I, P, SZ, I, P]),
    "vmmt_fill_zero": (I, [P, L, I, P]),
    "vmmt_copy_list": (I, [P, P, P, I, P]),
    "vmmt_sqnorm_workspace_bytes": (SZ, []),
    "vmmt_sqnorm": (I, [P, L, P, I, P, P]),
    "vmmt_adam_clip_step": (I, [P, P, P, P, L, P, F, F, F, F, F, F, L, P]),
    "vmmt_peer_signal_bytes": (SZ, []),
    "vmmt_peer_handle_bytes": (I, []),
    "vmmt_peer_alloc": (I, [SZ, C.POINTER(P), P]),
    "vmmt_peer_open": (I, [P, C.POINTER(P)]),
    "vmmt_peer_close": (I, [P]),
    "vmmt_peer_free": (I, [P]),
    "vmmt_peer_barrier": (I, [P, I, I, I, P]),
    "vmmt_peer_adam_workspace_bytes": (SZ, []),
    "vmmt_peer_slice": (L, [L, I, I, C.POINTER(L), C.POINTER(L)]),
    "vmmt_peer_reduce_scatter": (I, [P, P, SZ, I, I, L, L, P, I, P, P]),
    "vmmt_peer_adam_allgather": (I, [P, P, SZ, I, I, L, L, P, P, P, P, I, F, F, F, F, F, L, I, I, I, P]),
    "vmmt_peer_adam_step": (I, [P, P, SZ, SZ, I, I, L, P, P, P, P, F, F, F, F, F, L, P, P]),
    "vmmt_beam_advance": (I, [P, I, I, I, I, P, P, P, L, P, P, P, P, P, P, P, P, P, P]),
    "vmmt_beam_advance_topk": (I, [P, I, I, I, I, P, P, P, L, P, P, P, P, P, P, P, P, P, P]),
    "vmmt_beam_record": (I, [P, P, P, L, P]),
    "vmmt_beam_reorder": (I, [P, P, P, P, I, I, I, I, P]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

ACT_NONE, ACT_RELU, ACT_TANH, ACT_SOFTPLUS, ACT_SIGMOID = range(5)
F_EXACT, F_BF16, F_BACKGROUND, F_NO_SPLITK, F_SHARE_SMS = 1, 2, 4, 8, 16          # include/vmmt.h VMMT_F_*


def last_error():
    return lib.vmmt_last_error().decode()


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"libvmmt {what} failed (status {rc}): {last_error()}")


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """data_ptr of a CUDA fp32/int64 tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "libvmmt operands must live on a CUDA device (no CPU fallback)"
    return t.data_ptr()


def fptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float32, f"expected a CUDA fp32 tensor, got {t.dtype} on {t.device}"
    return t.data_ptr()


_profile = None          # bench.py: list collecting (name, start_event, end_event) per C-ABI call


def set_profile(sink):
    """Enable (list) / disable (None) per-call CUDA-event timing on the current stream."""
    global _profile
    _profile = sink


def call(name, *args):
    if _profile is None:
        check(getattr(lib, name)(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib, name)(*args), name)
    e1.record()
    _profile.append((name, args, e0, e1))
