"""ctypes binding of ``libneunet_b200.so`` (C-ABI: ``include/neunet_b200.h``) for torch CUDA tensors.

This is the seam the reference's ``neunet/nn/experimental`` wrappers occupy
(``experimental/utils.py:64-85`` ``load_cuda_function`` / ``to_pointer`` / ``call_cuda_function``):
device pointers are taken from ``tensor.data_ptr()``, every output and scratch buffer is allocated
here (torch caching allocator) and the current torch stream is passed as the trailing
``cudaStream_t``. There is NO fallback: if the shared library is missing, was built for another
architecture, or a call returns a non-zero status, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get(
    "NEUNET_B200_LIB", os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libneunet_b200.so")))

PREC_BF16, PREC_BF16X3 = 0, 1
ACT_NONE, ACT_SWISH = 0, 1
OPT_ADAM_L2, OPT_ADAMW = 0, 1

_PREC_NAMES = {"bf16": PREC_BF16, "bf16x3": PREC_BF16X3}
_state = {"prec": _PREC_NAMES[os.environ.get("NEUNET_B200_PREC", "bf16x3").lower()], "weights_epoch": 0,
          "capture_epoch": 0}


def _cache_scope():
    """Staged-operand caches are only valid inside the scope that created them: a CUDA-graph capture
    must re-stage everything it reads (replays do not re-run this Python), so the key carries
    (capturing?, capture id)."""
    return (torch.cuda.is_current_stream_capturing(), _state["capture_epoch"])


def set_precision(name: str) -> None:
    """"bf16": one bf16 tensor-core product (throughput mode, ~2e-3 rel. vs fp32);
    "bf16x3": hi/lo split, three products (parity mode, ~1e-5 rel.)."""
    _state["prec"] = _PREC_NAMES[name.lower()]


def get_precision() -> str:
    return {v: k for k, v in _PREC_NAMES.items()}[_state["prec"]]


class precision:
    """Context manager: ``with b200.precision("bf16"): ...``"""

    def __init__(self, name):
        self.name, self.prev = name, None

    def __enter__(self):
        self.prev = get_precision()
        set_precision(self.name)

    def __exit__(self, *exc):
        set_precision(self.prev)


class nnb_conv2d_desc(ctypes.Structure):
    _fields_ = [("B", c_int64), ("Cin", c_int64), ("H", c_int64), ("W", c_int64), ("Cout", c_int64),
                ("kh", c_int64), ("kw", c_int64), ("stride", c_int32 * 2), ("pad", c_int32 * 4),
                ("dil", c_int32 * 2)]


_I64x4 = c_int64 * 4
_SIGNATURES = {
    # name: (restype, argtypes)
    "nnb_last_error": (c_char_p, []),
    "nnb_version": (c_int, []),
    "nnb_device_check": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "nnb_set_pdl": (c_int, [c_int]),
    "nnb_set_sm_budget": (c_int, [c_int]),
    "nnb_launch_count": (c_uint64, []),
    "nnb_launch_count_reset": (None, []),
    "nnb_linear_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int]),
    "nnb_linear_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                   c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nnb_linear_forward_staged": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                          c_int64, c_int, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "nnb_linear_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int64, c_int64, c_int64, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    "nnb_linear_backward_dropped": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_int64, c_int64, c_int64, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p,
                                            c_size_t, c_float, c_uint64, c_uint32, c_uint64, c_void_p, c_void_p]),
    "nnb_weight_staged_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "nnb_stage_weight": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "nnb_matmul_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int]),
    "nnb_matmul_staged_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int]),
    "nnb_matmul_uses_tensor_cores": (c_int, [c_int64, c_int64, c_int64, c_int64, c_int64]),
    "nnb_matmul_set_small_path": (c_int, [c_int]),
    "nnb_matmul_forward": (c_int, [c_void_p, POINTER(c_int64), c_void_p, POINTER(c_int64), c_void_p, c_int64,
                                   c_int64, c_int64, c_int64, c_int64, c_float, c_int, c_void_p, c_void_p,
                                   c_void_p, c_size_t, c_void_p]),
    "nnb_matmul_backward": (c_int, [c_void_p, POINTER(c_int64), c_void_p, POINTER(c_int64), c_void_p, c_void_p,
                                    c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_int,
                                    c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nnb_conv2d_out_shape": (c_int, [POINTER(nnb_conv2d_desc), POINTER(c_int64), POINTER(c_int64)]),
    "nnb_conv2d_workspace_bytes": (c_size_t, [POINTER(nnb_conv2d_desc), c_int, c_int]),
    "nnb_conv2d_forward": (c_int, [POINTER(nnb_conv2d_desc), c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                   c_void_p, c_size_t, c_void_p]),
    "nnb_conv2d_backward": (c_int, [POINTER(nnb_conv2d_desc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "nnb_conv2d_planes_bytes": (c_size_t, [POINTER(nnb_conv2d_desc), c_int]),
    "nnb_conv2d_forward_ex": (c_int, [POINTER(nnb_conv2d_desc), c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "nnb_conv2d_backward_ex": (c_int, [POINTER(nnb_conv2d_desc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nnb_conv_transpose2d_supported": (c_int, [POINTER(nnb_conv2d_desc), c_int, c_int]),
    "nnb_conv_transpose2d_out_shape": (c_int, [POINTER(nnb_conv2d_desc), c_int, c_int, POINTER(c_int64), POINTER(c_int64)]),
    "nnb_conv_transpose2d_planes_bytes": (c_size_t, [POINTER(nnb_conv2d_desc), c_int]),
    "nnb_conv_transpose2d_workspace_bytes": (c_size_t, [POINTER(nnb_conv2d_desc), c_int, c_int, c_int, c_int]),
    "nnb_conv_transpose2d_forward": (c_int, [POINTER(nnb_conv2d_desc), c_int, c_int, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nnb_conv_transpose2d_backward": (c_int, [POINTER(nnb_conv2d_desc), c_int, c_int, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                                              c_void_p]),
    "nnb_comm_available": (c_int, [POINTER(c_int)]),
    "nnb_comm_unique_id": (c_int, [c_void_p]),
    "nnb_comm_init": (c_int, [POINTER(c_void_p), c_void_p, c_int, c_int]),
    "nnb_comm_allreduce_sum": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "nnb_comm_destroy": (c_int, [c_void_p]),
    "nnb_embedding_forward": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "nnb_embedding_workspace_bytes": (c_size_t, [c_int64]),
    "nnb_embedding_backward": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_size_t,
                                       c_void_p]),
    "nnb_swish_forward": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "nnb_swish_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "nnb_softmax_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "nnb_softmax_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "nnb_rmsnorm_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                    c_float, c_void_p]),
    "nnb_rmsnorm_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "nnb_rmsnorm_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "nnb_probe_linear_gemm": (c_int, [c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int, POINTER(c_float),
                                      POINTER(c_int), c_void_p]),
    "nnb_dropout": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_uint64, c_uint32, c_uint64, c_void_p, c_void_p]),
    "nnb_rng_advance": (c_int, [c_void_p, c_void_p]),
    "nnb_dropout_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float, c_uint64, c_uint32, c_uint64,
                                  c_void_p, c_void_p, c_int, c_void_p]),
    "nnb_swish_dropout_fused": (c_int, [c_void_p, c_float, c_void_p, c_int64, c_int64, c_float, c_uint64, c_uint32, c_uint64,
                                        c_void_p, c_void_p, c_int, c_void_p]),
    "nnb_rmsnorm_forward_fused": (c_int, [c_void_p, c_void_p, c_float, c_uint64, c_uint32, c_uint64, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                          c_float, c_void_p]),
    "nnb_rmsnorm_backward_acc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "nnb_bn_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "nnb_bn_stats": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nnb_bn_finalize": (c_int, [c_void_p, c_double, c_int64, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "nnb_bn_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float,
                             c_void_p, c_void_p]),
    "nnb_bn_backward_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "nnb_bn_backward_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int64,
                                      c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nnb_attention_supported": (c_int, [c_int64, c_int64, c_int64]),
    "nnb_attention_forward": (c_int, [c_void_p, POINTER(c_int64), c_void_p, POINTER(c_int64), c_void_p, POINTER(c_int64),
                                      c_void_p, c_int, c_float, POINTER(c_int64), c_float, c_float, c_float, c_uint64,
                                      c_uint32, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                      c_int64, c_int64, c_int64, c_void_p]),
    "nnb_attention_backward": (c_int, [c_void_p, POINTER(c_int64), c_void_p, POINTER(c_int64), c_void_p, POINTER(c_int64),
                                       c_void_p, c_int, c_float, POINTER(c_int64), c_float, c_float, c_float, c_uint64,
                                       c_uint32, c_uint64, c_void_p, c_void_p, POINTER(c_int64), c_void_p, c_void_p,
                                       c_void_p, c_int64, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "nnb_cross_entropy_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p]),
    "nnb_cross_entropy_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                           c_int64, c_void_p, c_void_p]),
    "nnb_cross_entropy_staged_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "nnb_cross_entropy_backward_staged": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                                  c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nnb_linear_backward_staged": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int,
                                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "nnb_adamw_create": (c_int, [POINTER(c_void_p), c_int, POINTER(c_void_p), POINTER(c_void_p),
                                 POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int64), c_void_p]),
    "nnb_adamw_set_grads": (c_int, [c_void_p, POINTER(c_void_p), c_void_p]),
    "nnb_adamw_step": (c_int, [c_void_p, c_double, c_double, c_double, c_double, c_double, c_int64, c_int,
                               c_float, c_void_p]),
    "nnb_adamw_step_range": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_double, c_double, c_double, c_int64, c_int,
                                     c_float, c_int, c_void_p]),
    "nnb_adamw_set_step": (c_int, [c_void_p, c_int64, c_void_p]),
    "nnb_adamw_set_staging": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int64), c_int, c_void_p]),
    "nnb_adamw_destroy": (c_int, [c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_device_ok = set()


def lib():
    """Load (once) and return the ctypes library. Raises if it is missing -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libneunet_b200.so not found at {LIB_PATH}. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C numpy-nn-model_b200/csrc`. "
                'device="cuda" has no fallback path.')
        dll = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(dll, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = dll
    return _lib


def last_error() -> str:
    return lib().nnb_last_error().decode()


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed (status {rc}): {last_error()}")


def require_device():
    """The kernels are sm_100a-only; anything else is an error, not a slow path."""
    if not torch.cuda.is_available():
        raise RuntimeError('device="cuda" needs a CUDA device (B200); no CPU fallback exists')
    dev = torch.cuda.current_device()
    if dev not in _device_ok:
        sm, ma, mi = c_int(), c_int(), c_int()
        _check(lib().nnb_device_check(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi)), "nnb_device_check")
        _device_ok.add(dev)
    return dev


def launch_count() -> int:
    return int(lib().nnb_launch_count())


def reset_launch_count() -> None:
    lib().nnb_launch_count_reset()


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def _f32c(t):
    """fp32, C-contiguous (the reference wrappers call ascontiguousarray, softmax.py:59-62)."""
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t if t.is_contiguous() else t.contiguous()


# ---- workspace: one growing scratch buffer per (device, stream) -------------------------------
_workspaces = {}


def _workspace(nbytes: int):
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        # grow geometrically; stream-ordered reuse is safe because every kernel of a call is
        # enqueued on this same stream before the next call touches the buffer
        size = int(max(nbytes, 1 << 20) * (1.25 if buf is not None else 1.0))
        buf = torch.empty(size, dtype=torch.uint8, device="cuda")
        _workspaces[key] = buf
    return buf


# ---- weight staging cache ------------------------------------------------------------------------
class _StagedWeight:
    """bf16 planes of one weight. `persistent` entries are owned by an optimizer whose step kernel rewrites
    them with every update (nnb_adamw_set_staging): they stay valid across CUDA-graph captures/replays."""
    __slots__ = ("buf", "key", "persistent")

    def __init__(self):
        self.persistent = False


def weights_changed() -> None:
    """Called by optimizers that update parameters through this library (torch's version counter
    cannot see those writes)."""
    _state["weights_epoch"] += 1


def _staged_weight(owner, w2d: torch.Tensor, rows: int, cols: int):
    """bf16 planes of a weight matrix, converted once per optimizer step. `owner` is the Parameter
    (any object that can hold an attribute); None disables caching."""
    prec = _state["prec"]
    cached = getattr(owner, "_b200_staged", None) if owner is not None else None
    if owner is not None:
        try:
            owner._b200_wants_staging = True  # optimizers emit the planes of such weights from their step kernel
        except AttributeError:
            pass
    if cached is not None and cached.persistent and cached.key == (w2d.data_ptr(), w2d._version, _state["weights_epoch"],
                                                                     prec, rows, cols):
        return cached.buf
    key = (w2d.data_ptr(), w2d._version, _state["weights_epoch"], prec, rows, cols, _cache_scope())
    if cached is not None and not cached.persistent and cached.key == key:
        return cached.buf
    if cached is not None and cached.persistent:
        cached = None  # never scribble over an optimizer-owned buffer with another precision / shape
    L = lib()
    nbytes = L.nnb_weight_staged_bytes(rows, cols, prec)
    if cached is not None and cached.buf.numel() >= nbytes:
        buf = cached.buf
    else:
        buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    _check(L.nnb_stage_weight(_ptr(w2d), rows, cols, prec, _ptr(buf), _stream()), "nnb_stage_weight")
    if owner is not None:
        sw = _StagedWeight()
        sw.buf, sw.key = buf, key
        try:
            owner._b200_staged = sw
        except AttributeError:
            pass
    return buf


# ---- nn.Linear -------------------------------------------------------------------------------------
def linear_forward(x, w, bias=None, act=ACT_NONE, beta=1.0, save_z=False, owner=None, keep_x_staged=False,
                   x_owner=None):
    """O = act(x . w^T + bias); x: (..., K), w: (N, K), bias: (1, N) or None.
    Returns (O, Z, x_staged): Z is the pre-activation when save_z (else None); x_staged holds the
    bf16 planes of x for the backward pass when keep_x_staged (else None)."""
    require_device()
    L = lib()
    N, K = w.shape
    lead = tuple(x.shape[:-1])
    x2 = _f32c(x).reshape(-1, K)
    M = x2.shape[0]
    w = _f32c(w)
    out = torch.empty((M, N), dtype=torch.float32, device="cuda")
    z = torch.empty((M, N), dtype=torch.float32, device="cuda") if save_z else None
    b = _f32c(bias).reshape(-1) if bias is not None else None
    prec = _state["prec"]
    wst = _staged_weight(owner, w, N, K)
    wsb = L.nnb_linear_workspace_bytes(M, K, N, prec, 0)
    ws = _workspace(wsb)
    xst = None
    # several layers often read the same activation (q/k/v projections of one RMSNorm output): the bf16
    # planes are cached on the producing Tensor, keyed by storage + torch's version counter
    xkey = (x2.data_ptr(), x._version, prec, M, K, _cache_scope())
    cached = getattr(x_owner, "_b200_xst", None) if x_owner is not None else None
    if cached is not None and cached[0] == xkey:
        xst = cached[1]
        _check(L.nnb_linear_forward_staged(_ptr(xst), _ptr(wst), _ptr(b), _ptr(out), _ptr(z), M, K, N, act, float(beta),
                                           prec, _ptr(ws), ws.numel(), _stream()), "nnb_linear_forward_staged")
    else:
        if keep_x_staged:
            xst = torch.empty(L.nnb_weight_staged_bytes(M, K, prec), dtype=torch.uint8, device="cuda")
            xst._b200_prec = prec
            if x_owner is not None:
                try:
                    x_owner._b200_xst = (xkey, xst)
                except AttributeError:
                    pass
        _check(L.nnb_linear_forward(_ptr(x2), _ptr(w), _ptr(b), _ptr(out), _ptr(z), M, K, N, act, float(beta), prec,
                                    _ptr(wst), _ptr(xst), _ptr(ws), ws.numel(), _stream()), "nnb_linear_forward")
    out = out.reshape(lead + (N,))
    if z is not None:
        z = z.reshape(lead + (N,))
    return out, z, xst


def _usable_out(buf, shape):
    return (buf is not None and isinstance(buf, torch.Tensor) and buf.is_cuda and buf.dtype == torch.float32
            and buf.is_contiguous() and tuple(buf.shape) == tuple(shape))


def linear_backward(x, w, grad, z=None, act=ACT_NONE, beta=1.0, need_dx=True, need_db=True, owner=None,
                    x_staged=None, dw_out=None, db_out=None, grad_drop=None):
    """Returns (dX like x, dW (N,K), db (1,N)) -- dX / db None when not requested. dw_out / db_out: optional
    caller-owned destinations (e.g. slices of the data-parallel gradient bucket) the GEMM / column sums write
    into directly instead of fresh tensors. grad_drop = (p, ticket): `grad` still has to pass the backward of an
    nn.Dropout -- the mask is applied inside the staging pass over grad (nnb_linear_backward_dropped)."""
    require_device()
    L = lib()
    N, K = w.shape
    x2 = _f32c(x).reshape(-1, K)
    M = x2.shape[0]
    g2 = _f32c(grad).reshape(-1, N)
    w = _f32c(w)
    dx = torch.empty((M, K), dtype=torch.float32, device="cuda") if need_dx else None
    dw = dw_out if _usable_out(dw_out, (N, K)) else torch.empty((N, K), dtype=torch.float32, device="cuda")
    db = None
    if need_db:
        db = db_out if _usable_out(db_out, (1, N)) else torch.empty((1, N), dtype=torch.float32, device="cuda")
    z2 = _f32c(z).reshape(-1, N) if z is not None else None
    prec = _state["prec"]
    wst = _staged_weight(owner, w, N, K) if need_dx else None
    if x_staged is not None and getattr(x_staged, "_b200_prec", None) != prec:
        x_staged = None  # precision changed between forward and backward: convert again
    wsb = L.nnb_linear_workspace_bytes(M, K, N, prec, 1)
    ws = _workspace(wsb)
    if grad_drop is not None and N % 4 != 0:
        g2 = dropout_apply(g2, grad_drop[0], grad_drop[1])
        grad_drop = None
    if grad_drop is not None:
        p_drop, (seed, call_id, epoch, dev) = grad_drop
        _check(L.nnb_linear_backward_dropped(_ptr(x2), _ptr(w), _ptr(z2), _ptr(g2), _ptr(dx), _ptr(dw), _ptr(db), M, K, N, act,
                                             float(beta), prec, _ptr(wst), _ptr(x_staged), _ptr(ws), ws.numel(),
                                             float(p_drop), seed, call_id, epoch, _ptr(dev), _stream()),
               "nnb_linear_backward_dropped")
    else:
        _check(L.nnb_linear_backward(_ptr(x2), _ptr(w), _ptr(z2), _ptr(g2), _ptr(dx), _ptr(dw), _ptr(db), M, K, N, act,
                                     float(beta), prec, _ptr(wst), _ptr(x_staged), _ptr(ws), ws.numel(), _stream()),
               "nnb_linear_backward")
    if dx is not None:
        dx = dx.reshape(x.shape)
    return dx, dw, db


def probe_linear_gemm(M, K, N, form=0, with_bias=True, swish=False, rounds=3, sets=None):
    """GPU-paced microseconds of one GEMM launch of an nn.Linear form (0 fwd, 1 dgrad, 2 wgrad) in the current precision mode,
    over operand sets that together exceed the L2 (nnb_probe_linear_gemm). Returns (us, launches)."""
    require_device()
    per_set = 2 * (M * K + N * K + M * N) + 4 * max(M * N, M * K, N * K)  # upper bound, bytes
    if sets is None:
        sets = int(min(64, max(2, -(-(300 << 20) // per_set))))
    us, nl = c_float(0), c_int(0)
    flags = int(bool(with_bias)) | (2 if swish else 0) | (4 if _state["prec"] == PREC_BF16X3 else 0)
    _check(lib().nnb_probe_linear_gemm(M, K, N, form, flags, sets, rounds, ctypes.byref(us), ctypes.byref(nl),
                                       _stream()), "nnb_probe_linear_gemm")
    return float(us.value), int(nl.value)


# ---- Tensor.matmul -----------------------------------------------------------------------------------
def _as4(t, batch_shape):
    """View an (..., R, C) tensor as 4-D [b0, b1, R, C] element strides, broadcasting leading dims to
    `batch_shape` (stride 0). Copies only if more than two non-trivial batch dims cannot be merged."""
    R, C = t.shape[-2], t.shape[-1]
    lead = t.shape[:-2]
    full = tuple(batch_shape)
    exp = t.expand(full + (R, C)) if tuple(lead) != full else t
    if len(full) > 2:
        try:
            exp = exp.view((-1,) + (R, C))  # succeeds when the batch dims are mergeable
        except RuntimeError:
            exp = exp.contiguous().view((-1, R, C))
    while exp.ndim < 4:
        exp = exp.unsqueeze(0)
    return exp


def _strides4(t):
    return _I64x4(*[int(s) for s in t.stride()])


def _matmul_norm(a, b):
    """Normalise NumPy matmul operand ranks to 4-D batched views. Returns
    (a4, b4, (b0, b1, M, K, N), out_shape)."""
    a_vec, b_vec = a.ndim == 1, b.ndim == 1
    if a_vec:
        a = a.unsqueeze(0)
    if b_vec:
        b = b.unsqueeze(1)
    if a.shape[-1] != b.shape[-2]:
        raise ValueError(f"matmul: shapes {tuple(a.shape)} and {tuple(b.shape)} not aligned")
    batch = torch.broadcast_shapes(a.shape[:-2], b.shape[:-2])
    M, K, N = a.shape[-2], a.shape[-1], b.shape[-1]
    a4, b4 = _as4(a, batch), _as4(b, batch)
    out_shape = tuple(batch) + (M, N)
    if a_vec:
        out_shape = out_shape[:-2] + (N,)
    if b_vec:
        out_shape = out_shape[:-1] if not a_vec else out_shape[:-1]
    return a4, b4, (a4.shape[0], a4.shape[1], M, K, N), out_shape, tuple(batch)


def matmul(a, b, alpha=1.0, keep_staged=False):
    """``xp.matmul`` for device arrays (neunet/autograd.py:199), any NumPy-legal rank combination.
    With keep_staged the bf16 planes of both operands are kept and returned as a third value
    ``(a_staged, b_staged, prec)`` for ``matmul_backward(..., staged=...)``."""
    require_device()
    L = lib()
    a = a if a.dtype == torch.float32 else a.to(torch.float32)
    b = b if b.dtype == torch.float32 else b.to(torch.float32)
    if a.ndim == 0 or b.ndim == 0:
        raise ValueError("matmul: scalar operands are not allowed")
    a4, b4, (b0, b1, M, K, N), out_shape, _ = _matmul_norm(a, b)
    out = torch.empty((b0, b1, M, N), dtype=torch.float32, device="cuda")
    prec = _state["prec"]
    ws = _workspace(L.nnb_matmul_workspace_bytes(b0, b1, M, K, N, prec, 0))
    ast = bst = None
    if keep_staged and not L.nnb_matmul_uses_tensor_cores(b0, b1, M, K, N):
        keep_staged = False  # fp32 CUDA-core path for small batched products: nothing is staged
        out_only = True
    else:
        out_only = False
    if keep_staged:
        ast = torch.empty(L.nnb_matmul_staged_bytes(b0, b1, M, K, prec), dtype=torch.uint8, device="cuda")
        bst = torch.empty(L.nnb_matmul_staged_bytes(b0, b1, K, N, prec), dtype=torch.uint8, device="cuda")
    _check(L.nnb_matmul_forward(_ptr(a4), _strides4(a4), _ptr(b4), _strides4(b4), _ptr(out), b0, b1, M, K, N,
                                float(alpha), prec, _ptr(ast), _ptr(bst), _ptr(ws), ws.numel(), _stream()),
           "nnb_matmul_forward")
    if keep_staged:
        return out.reshape(out_shape), (ast, bst, prec, tuple(_strides4(a4)), tuple(_strides4(b4)))
    if out_only:
        return out.reshape(out_shape), None
    return out.reshape(out_shape)


def matmul_backward(a, b, grad, need_da=True, need_db=True, alpha=1.0, staged=None):
    """(dA, dB) for matrix x matrix operands, shaped like the BROADCAST operands (the caller's
    apply_grad un-broadcasts, neunet/autograd.py:85-93, 948-962)."""
    require_device()
    L = lib()
    a4, b4, (b0, b1, M, K, N), _, batch = _matmul_norm(a, b)
    g4 = _f32c(grad).reshape(b0, b1, M, N)
    da = torch.empty((b0, b1, M, K), dtype=torch.float32, device="cuda") if need_da else None
    db = torch.empty((b0, b1, K, N), dtype=torch.float32, device="cuda") if need_db else None
    prec = _state["prec"]
    ws = _workspace(L.nnb_matmul_workspace_bytes(b0, b1, M, K, N, prec, 1))
    ast = bst = None
    if staged is not None and staged[2] == prec and staged[3] == tuple(_strides4(a4)) and staged[4] == tuple(_strides4(b4)):
        ast, bst = staged[0], staged[1]  # planes converted by the forward call: same views, same precision
    _check(L.nnb_matmul_backward(_ptr(a4), _strides4(a4), _ptr(b4), _strides4(b4), _ptr(g4), _ptr(da), _ptr(db),
                                 b0, b1, M, K, N, float(alpha), prec, _ptr(ast), _ptr(bst), _ptr(ws), ws.numel(),
                                 _stream()), "nnb_matmul_backward")
    if da is not None:
        da = da.reshape(batch + (M, K))
    if db is not None:
        db = db.reshape(batch + (K, N))
    return da, db


# ---- nn.Conv2d -------------------------------------------------------------------------------------
def _conv_desc(x_shape, w_shape, stride, pad4, dil):
    d = nnb_conv2d_desc()
    d.B, d.Cin, d.H, d.W = [int(v) for v in x_shape]
    d.Cout, _, d.kh, d.kw = [int(v) for v in w_shape]
    d.stride[0], d.stride[1] = int(stride[0]), int(stride[1])
    for i in range(4):
        d.pad[i] = int(pad4[i])
    d.dil[0], d.dil[1] = int(dil[0]), int(dil[1])
    return d


def conv2d_out_hw(x_shape, w_shape, stride, pad4, dil):
    d = _conv_desc(x_shape, w_shape, stride, pad4, dil)
    ho, wo = c_int64(), c_int64()
    _check(lib().nnb_conv2d_out_shape(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)), "nnb_conv2d_out_shape")
    return int(ho.value), int(wo.value)


def conv2d_forward(x, w, bias, stride, pad4, dil, keep_planes=False):
    """Returns the NCHW output; with keep_planes a pair (output, planes) where planes is (buffer, prec) holding the
    channels-last bf16 form of x for `conv2d_backward(..., x_planes=...)` (None when the geometry does not use planes)."""
    require_device()
    L = lib()
    x, w = _f32c(x), _f32c(w)
    if x.shape[1] != w.shape[1]:
        raise ValueError(f"conv2d: input has {x.shape[1]} channels, weight expects {w.shape[1]}")
    d = _conv_desc(x.shape, w.shape, stride, pad4, dil)
    ho, wo = conv2d_out_hw(x.shape, w.shape, stride, pad4, dil)
    out = torch.empty((x.shape[0], w.shape[0], ho, wo), dtype=torch.float32, device="cuda")
    b = _f32c(bias).reshape(-1) if bias is not None else None
    prec = _state["prec"]
    ws = _workspace(L.nnb_conv2d_workspace_bytes(ctypes.byref(d), prec, 0))
    planes = None
    if keep_planes:
        nb = L.nnb_conv2d_planes_bytes(ctypes.byref(d), prec)
        if nb:
            planes = (torch.empty(nb, dtype=torch.uint8, device="cuda"), prec)
    _check(L.nnb_conv2d_forward_ex(ctypes.byref(d), _ptr(x), _ptr(w), _ptr(b), _ptr(out), prec, None,
                                   _ptr(planes[0]) if planes else None, _ptr(ws), ws.numel(), _stream()),
           "nnb_conv2d_forward")
    return (out, planes) if keep_planes else out


def conv2d_backward(x, w, grad, stride, pad4, dil, need_dx=True, need_db=True, x_planes=None):
    require_device()
    L = lib()
    x, w, grad = _f32c(x), _f32c(w), _f32c(grad)
    d = _conv_desc(x.shape, w.shape, stride, pad4, dil)
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty_like(w)
    db = torch.empty((w.shape[0],), dtype=torch.float32, device="cuda") if need_db else None
    prec = _state["prec"]
    ws = _workspace(L.nnb_conv2d_workspace_bytes(ctypes.byref(d), prec, 1))
    xp_buf = x_planes[0] if (x_planes is not None and x_planes[1] == prec) else None
    _check(L.nnb_conv2d_backward_ex(ctypes.byref(d), _ptr(x), _ptr(w), _ptr(grad), _ptr(dx), _ptr(dw), _ptr(db), prec,
                                    _ptr(xp_buf), _ptr(ws), ws.numel(), _stream()), "nnb_conv2d_backward")
    return dx, dw, db


# ---- native nn.ConvTranspose2d -------------------------------------------------------------------------------
def conv_transpose2d_supported(x_shape, w_shape, stride, pad4, dil, out_pad):
    d = _conv_desc(x_shape, w_shape, stride, pad4, dil)
    return bool(lib().nnb_conv_transpose2d_supported(ctypes.byref(d), int(out_pad[0]), int(out_pad[1])))


def conv_transpose2d_forward(x, w, bias, stride, pad4, dil, out_pad):
    """Gather-form transposed convolution (real taps only). Returns (out, planes-of-x for backward)."""
    require_device()
    L = lib()
    x, w = _f32c(x), _f32c(w)
    d = _conv_desc(x.shape, w.shape, stride, pad4, dil)
    op0, op1 = int(out_pad[0]), int(out_pad[1])
    ho, wo = c_int64(), c_int64()
    _check(L.nnb_conv_transpose2d_out_shape(ctypes.byref(d), op0, op1, ctypes.byref(ho), ctypes.byref(wo)),
           "nnb_conv_transpose2d_out_shape")
    out = torch.empty((x.shape[0], w.shape[0], int(ho.value), int(wo.value)), dtype=torch.float32, device="cuda")
    b = _f32c(bias).reshape(-1) if bias is not None else None
    prec = _state["prec"]
    planes = (torch.empty(L.nnb_conv_transpose2d_planes_bytes(ctypes.byref(d), prec), dtype=torch.uint8, device="cuda"), prec)
    ws = _workspace(L.nnb_conv_transpose2d_workspace_bytes(ctypes.byref(d), op0, op1, prec, 0))
    _check(L.nnb_conv_transpose2d_forward(ctypes.byref(d), op0, op1, _ptr(x), _ptr(w), _ptr(b), _ptr(out), prec,
                                          _ptr(planes[0]), _ptr(ws), ws.numel(), _stream()), "nnb_conv_transpose2d_forward")
    return out, planes


def conv_transpose2d_backward(x, w, grad, stride, pad4, dil, out_pad, need_dx=True, need_db=True, x_planes=None):
    require_device()
    L = lib()
    x, w, grad = _f32c(x), _f32c(w), _f32c(grad)
    d = _conv_desc(x.shape, w.shape, stride, pad4, dil)
    op0, op1 = int(out_pad[0]), int(out_pad[1])
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty_like(w)
    db = torch.empty((w.shape[0],), dtype=torch.float32, device="cuda") if need_db else None
    prec = _state["prec"]
    ws = _workspace(L.nnb_conv_transpose2d_workspace_bytes(ctypes.byref(d), op0, op1, prec, 1))
    xp_buf = x_planes[0] if (x_planes is not None and x_planes[1] == prec) else None
    _check(L.nnb_conv_transpose2d_backward(ctypes.byref(d), op0, op1, _ptr(x), _ptr(w), _ptr(grad), _ptr(dx), _ptr(dw),
                                           _ptr(db), prec, _ptr(xp_buf), _ptr(ws), ws.numel(), _stream()),
           "nnb_conv_transpose2d_backward")
    return dx, dw, db


# ---- stand-alone epilogue ops ---------------------------------------------------------------------------
def swish_forward(x, beta=1.0):
    require_device()
    x = _f32c(x)
    y = torch.empty_like(x)
    if x.numel():
        _check(lib().nnb_swish_forward(_ptr(x), _ptr(y), x.numel(), float(beta), _stream()), "nnb_swish_forward")
    return y


def swish_backward(x, grad, beta=1.0):
    require_device()
    x, grad = _f32c(x), _f32c(grad)
    dx = torch.empty_like(x)
    if x.numel():
        _check(lib().nnb_swish_backward(_ptr(x), _ptr(grad), _ptr(dx), x.numel(), float(beta), _stream()),
               "nnb_swish_backward")
    return dx


def _softmax_dims(shape, axis):
    axis = axis % len(shape)
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
    return outer, int(shape[axis]), inner


def softmax_forward(x, axis=-1):
    require_device()
    x = _f32c(x)
    y = torch.empty_like(x)
    outer, n, inner = _softmax_dims(x.shape, axis)
    _check(lib().nnb_softmax_forward(_ptr(x), _ptr(y), outer, n, inner, _stream()), "nnb_softmax_forward")
    return y


def softmax_backward(y, grad, axis=-1):
    require_device()
    y, grad = _f32c(y), _f32c(grad)
    dx = torch.empty_like(y)
    outer, n, inner = _softmax_dims(y.shape, axis)
    _check(lib().nnb_softmax_backward(_ptr(y), _ptr(grad), _ptr(dx), outer, n, inner, _stream()),
           "nnb_softmax_backward")
    return dx


def planes_key(t2d, prec=None):
    """Cache key under which `linear_forward` looks for ready-made bf16 planes of a [M, K] activation."""
    M, K = t2d.shape
    return (t2d.data_ptr(), t2d._version, _state["prec"] if prec is None else prec, int(M), int(K), _cache_scope())


def _new_planes(rows, cols):
    """(buffer, prec) for the bf16 planes of a [rows, cols] activation, or (None, prec) when TMA cannot tile it."""
    prec = _state["prec"]
    if cols % 8 != 0:
        return None, prec
    buf = torch.empty(lib().nnb_weight_staged_bytes(rows, cols, prec), dtype=torch.uint8, device="cuda")
    buf._b200_prec = prec
    return buf, prec


def rmsnorm_forward(x, w, b=None, eps=1e-6, add_dropout=None, want_planes=False):
    """Returns (Y, X_std, S, planes): X_std shaped (..., 1); X_norm is not materialised (the backward recomputes
    X / X_std). add_dropout = (a, p, ticket): the norm runs on S = x + dropout(a) and S is returned (else None).
    planes: ((key, buffer) for `linear_forward`'s operand cache) when want_planes and the fused kernel applies."""
    require_device()
    x = _f32c(x)
    cols = x.shape[-1]
    rows = x.numel() // cols
    y = torch.empty_like(x)
    std = torch.empty(x.shape[:-1] + (1,), dtype=torch.float32, device="cuda")
    w, b = _f32c(w), (_f32c(b) if b is not None else None)
    fusable = cols <= 1024 and cols % 4 == 0 and x.data_ptr() % 16 == 0
    if not fusable:
        if add_dropout is not None:
            a, p, ticket = add_dropout
            x = dropout_apply(a, p, ticket, residual=x)
            s_out = x
        else:
            s_out = None
        _check(lib().nnb_rmsnorm_forward(_ptr(x), _ptr(w), _ptr(b), _ptr(y), _ptr(std), None, rows, cols, float(eps),
                                         _stream()), "nnb_rmsnorm_forward")
        return y, std, s_out, None
    buf, prec = _new_planes(rows, cols) if want_planes else (None, _state["prec"])
    a_t, s_out, p, seed, call_id, epoch, dev = None, None, 0.0, 0, 0, 0, None
    if add_dropout is not None:
        a_t, p, (seed, call_id, epoch, dev) = add_dropout
        a_t = _f32c(a_t)
        s_out = torch.empty_like(x)
    _check(lib().nnb_rmsnorm_forward_fused(_ptr(x), _ptr(a_t), float(p), seed, call_id, epoch, _ptr(dev), _ptr(s_out), _ptr(w),
                                           _ptr(b), _ptr(y), _ptr(std), _ptr(buf), prec, rows, cols, float(eps), _stream()),
           "nnb_rmsnorm_forward_fused")
    planes = (planes_key(y.reshape(rows, cols), prec), buf) if buf is not None else None
    return y, std, s_out, planes


def rmsnorm_backward(grad, x, w, std, need_db=False, dx_add=None):
    """dx_add: gradient the input already holds (same shape); the kernel returns dx_add + dX in one pass."""
    require_device()
    L = lib()
    x, grad = _f32c(x), _f32c(grad)
    cols = x.shape[-1]
    rows = x.numel() // cols
    dx = torch.empty_like(x)
    dw = torch.empty((cols,), dtype=torch.float32, device="cuda")
    db = torch.empty((cols,), dtype=torch.float32, device="cuda") if need_db else None
    ws = _workspace(L.nnb_rmsnorm_workspace_bytes(rows, cols))
    if dx_add is not None and not (cols <= 1024 and cols % 4 == 0 and tuple(dx_add.shape) == tuple(x.shape)):
        dx_add = None
    if dx_add is not None:
        dx_add = _f32c(dx_add)
    _check(L.nnb_rmsnorm_backward_acc(_ptr(grad), _ptr(x), _ptr(_f32c(w)), _ptr(_f32c(std)), None, _ptr(dx_add), _ptr(dx),
                                      _ptr(dw), _ptr(db), rows, cols, _ptr(ws), ws.numel(), _stream()),
           "nnb_rmsnorm_backward")
    return dx, dw, db, dx_add is not None


# ---- fused CrossEntropyLoss ---------------------------------------------------------------------------------------
_REDUCTION = {"none": 0, "mean": 1, "sum": 2}


# ---- Dropout (device RNG) ----------------------------------------------------------------------------
_rng = {"seed": 0x5EED5EED, "epoch": 0, "dev": None, "capture_calls": 0, "graph_used": False}


def manual_seed(seed: int) -> None:
    """Seed of the device dropout RNG (Philox key). Masks are a function of (seed, call, epoch, index)."""
    _rng["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF
    _rng["epoch"] = 0


def _rng_epoch_dev():
    if _rng["dev"] is None:
        if torch.cuda.is_current_stream_capturing():
            # allocated (and zero-filled) inside a capture it would belong to the graph and be reset by every replay
            raise RuntimeError("dropout inside a CUDA-graph capture needs the RNG epoch buffer created beforehand "
                               "(use neunet.b200.GraphedStep, or call neunet.b200.rng_prepare_capture() first)")
        _rng["dev"] = torch.zeros(1, dtype=torch.int64, device="cuda")
    return _rng["dev"]


def rng_prepare_capture():
    """Create the device-resident RNG epoch before a CUDA-graph capture that contains dropout."""
    require_device()
    return _rng_epoch_dev()


def dropout_ticket():
    """Identity of one dropout mask: (seed, call_id, host epoch, device-epoch tensor or None).
    Eager calls take a fresh host epoch each; calls recorded into a CUDA graph take a per-capture call
    index and read the device-resident epoch, which GraphedStep.replay() advances before every replay."""
    if torch.cuda.is_current_stream_capturing():
        _rng["capture_calls"] += 1
        _rng["graph_used"] = True
        return (_rng["seed"], _rng["capture_calls"], 0, _rng_epoch_dev())
    _rng["epoch"] += 1
    return (_rng["seed"], 0, (1 << 62) + _rng["epoch"], None)


def dropout_apply(x, p, ticket, residual=None, want_planes=False):
    """y = (residual +) x * mask(ticket) / (1 - p): forward on activations, backward on the upstream gradient.
    With want_planes returns (y, planes) where planes = (key, buffer) for `linear_forward`'s operand cache."""
    require_device()
    x = _f32c(x)
    y = torch.empty_like(x)
    seed, call_id, epoch, dev = ticket
    cols = x.shape[-1] if x.ndim >= 1 and x.numel() else 1
    rows = x.numel() // max(cols, 1)
    buf, prec = (None, _state["prec"])
    if want_planes and x.ndim >= 2:
        buf, prec = _new_planes(rows, cols)
    if residual is not None:
        residual = _f32c(residual)
        if tuple(residual.shape) != tuple(x.shape):
            raise ValueError("dropout_apply: residual must have the shape of x")
    if x.numel():
        _check(lib().nnb_dropout_fused(_ptr(x), _ptr(residual), _ptr(y), rows, cols, float(p), seed, call_id, epoch, _ptr(dev),
                                       _ptr(buf), prec, _stream()), "nnb_dropout_fused")
    if want_planes:
        return y, ((planes_key(y.reshape(rows, cols), prec), buf) if buf is not None else None)
    return y


def swish_dropout_apply(z, beta, p, ticket, want_planes=True):
    """dropout(swish(z)) in one pass over the pre-activation (+ the bf16 planes of the result for the next nn.Linear).
    Same mask and (to ~2 ulp) the same values as dropout_apply(swish_forward(z, beta), p, ticket)."""
    require_device()
    z = _f32c(z)
    y = torch.empty_like(z)
    seed, call_id, epoch, dev = ticket
    cols = z.shape[-1]
    rows = z.numel() // max(cols, 1)
    buf, prec = _new_planes(rows, cols) if want_planes else (None, _state["prec"])
    if z.numel():
        _check(lib().nnb_swish_dropout_fused(_ptr(z), float(beta), _ptr(y), rows, cols, float(p), seed, call_id, epoch,
                                             _ptr(dev), _ptr(buf), prec, _stream()), "nnb_swish_dropout_fused")
    return y, ((planes_key(y.reshape(rows, cols), prec), buf) if buf is not None else None)


# ---- nn.Embedding ---------------------------------------------------------------------------------------------------
def _ids(ids):
    if ids.dtype not in (torch.int32, torch.int64):
        ids = ids.to(torch.int64)
    return ids.contiguous()


def embedding_forward(weight, ids):
    """weight[ids] for an integer device tensor of any shape -> (*ids.shape, D)."""
    require_device()
    weight = _f32c(weight)
    V, D = weight.shape
    ids = _ids(ids)
    out = torch.empty(tuple(ids.shape) + (D,), dtype=torch.float32, device="cuda")
    if ids.numel():
        _check(lib().nnb_embedding_forward(_ptr(weight), _ptr(ids), int(ids.dtype == torch.int64), ids.numel(), V, D, _ptr(out),
                                           _stream()), "nnb_embedding_forward")
    return out


def embedding_backward(ids, grad, V, out=None):
    """Gradient of weight[ids] with the reference's assignment semantics (the LAST duplicate wins). `out`: optional
    caller-owned [V, D] destination (e.g. a slice of the data-parallel gradient bucket)."""
    require_device()
    L = lib()
    grad = _f32c(grad)
    D = grad.shape[-1]
    ids = _ids(ids)
    dw = out if _usable_out(out, (V, D)) else torch.empty((V, D), dtype=torch.float32, device="cuda")
    ws = _workspace(L.nnb_embedding_workspace_bytes(V))
    _check(L.nnb_embedding_backward(_ptr(ids), int(ids.dtype == torch.int64), _ptr(grad), ids.numel(), V, D, _ptr(dw), _ptr(ws),
                                    ws.numel(), _stream()), "nnb_embedding_backward")
    return dw


# ---- gradient all-reduce inside the C-ABI (NCCL resolved at run time) -------------------------------------------------
class NativeComm:
    """``nnb_comm_*``: the library's own NCCL communicator for this rank. The 128-byte unique id travels through
    `share`, a callable (bytes-or-None) -> bytes that returns rank 0's bytes on every rank; by default a
    ``torch.distributed`` broadcast of a CPU/GPU byte tensor (any initialised backend, e.g. gloo)."""

    def __init__(self, rank, world, share=None):
        require_device()
        L = lib()
        if not L.nnb_comm_available(None):
            raise RuntimeError("nnb_comm: NCCL (libnccl.so.2) is not available in this process")
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _check(L.nnb_comm_unique_id(buf), "nnb_comm_unique_id")
        raw = bytes(buf.raw)
        if share is None:
            import torch.distributed as dist
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone().to(dev)
            dist.broadcast(t, 0)
            raw = bytes(t.cpu().numpy().tobytes())
        else:
            raw = share(raw if rank == 0 else None)
        self._h = c_void_p()
        self.rank, self.world = rank, world
        _check(L.nnb_comm_init(ctypes.byref(self._h), ctypes.create_string_buffer(raw, 128), world, rank), "nnb_comm_init")
        self.stream = torch.cuda.Stream()  # the collective runs beside the backward kernels

    def all_reduce_(self, flat, async_op=False):
        """In-place sum of a contiguous fp32 tensor over all ranks, on the communicator's own stream, ordered after the
        work already queued on the current stream. Returns an event-like handle with ``wait()`` when async_op."""
        assert flat.dtype == torch.float32 and flat.is_contiguous()
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        _check(lib().nnb_comm_allreduce_sum(self._h, _ptr(flat), flat.numel(), c_void_p(self.stream.cuda_stream)),
               "nnb_comm_allreduce_sum")
        flat.record_stream(self.stream)
        comm_stream = self.stream

        class _Work:
            def wait(self_inner):
                torch.cuda.current_stream().wait_stream(comm_stream)
        w = _Work()
        if not async_op:
            w.wait()
            return None
        return w

    def destroy(self):
        if self._h:
            lib().nnb_comm_destroy(self._h)
            self._h = c_void_p()


# ---- fused LeakyReLU + BatchNorm2d ------------------------------------------------------------------------------
_bn = {"sync": False}


def set_sync_batchnorm(on: bool) -> bool:
    """SyncBN: batch statistics (and their gradients) are all-reduced over the data-parallel ranks, so an N-rank run
    normalises exactly like one process on the global batch. Off by default: per-shard
    statistics, like running the reference once per shard."""
    prev = _bn["sync"]
    _bn["sync"] = bool(on)
    return prev


def _bn_world():
    if not _bn["sync"]:
        return 1
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def bn_forward(x, w, b, alpha, eps, momentum, running_mean=None, running_var=None, stats=None):
    """y = BatchNorm2d(leaky_relu(x, alpha)) over NCHW x. Training (stats None): batch statistics, running stats updated
    in place; eval: stats = (mean, inv_std). Returns (y, mean, inv_std)."""
    require_device()
    L = lib()
    x = _f32c(x)
    B, C, H, W = x.shape
    HW = H * W
    y = torch.empty_like(x)
    if stats is None:
        mean = torch.empty(C, dtype=torch.float32, device="cuda")
        inv = torch.empty(C, dtype=torch.float32, device="cuda")
        sums = torch.empty(2 * C, dtype=torch.float64, device="cuda")
        ws = _workspace(L.nnb_bn_workspace_bytes(B, C))
        _check(L.nnb_bn_stats(_ptr(x), B, C, HW, float(alpha), _ptr(sums), _ptr(ws), ws.numel(), _stream()), "nnb_bn_stats")
        world = _bn_world()
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(sums)
        _check(L.nnb_bn_finalize(_ptr(sums), float(B * HW * world), C, float(eps), float(momentum), _ptr(mean), _ptr(inv),
                                 _ptr(running_mean), _ptr(running_var), _stream()), "nnb_bn_finalize")
    else:
        mean, inv = stats
    _check(L.nnb_bn_apply(_ptr(x), _ptr(mean), _ptr(inv), _ptr(w), _ptr(b), B, C, HW, float(alpha), _ptr(y), _stream()),
           "nnb_bn_apply")
    return y, mean, inv


def bn_backward(x, grad, mean, inv, w, alpha, need_dx=True, need_dw=True):
    """Returns (dx, dw, db) for the fused LeakyReLU + BatchNorm2d (training-mode statistics)."""
    require_device()
    L = lib()
    x, grad = _f32c(x), _f32c(grad)
    B, C, H, W = x.shape
    HW = H * W
    sums = torch.empty(2 * C, dtype=torch.float64, device="cuda")
    ws = _workspace(L.nnb_bn_workspace_bytes(B, C))
    _check(L.nnb_bn_backward_stats(_ptr(x), _ptr(grad), _ptr(mean), _ptr(inv), B, C, HW, float(alpha), _ptr(sums), _ptr(ws),
                                   ws.numel(), _stream()), "nnb_bn_backward_stats")
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty(C, dtype=torch.float32, device="cuda") if need_dw else None
    db = torch.empty(C, dtype=torch.float32, device="cuda") if need_dw else None
    world = _bn_world()
    count = float(B * HW * world)
    if world > 1:
        import torch.distributed as dist
        # parameter gradients from the LOCAL sums (the gradient all-reduce adds the ranks later); dx needs the global ones
        if need_dw:
            _check(L.nnb_bn_backward_apply(_ptr(x), _ptr(grad), _ptr(mean), _ptr(inv), _ptr(w), _ptr(sums), count, B, C, HW,
                                           float(alpha), None, _ptr(dw), _ptr(db), _stream()), "nnb_bn_backward_apply")
        dist.all_reduce(sums)
        if need_dx:
            _check(L.nnb_bn_backward_apply(_ptr(x), _ptr(grad), _ptr(mean), _ptr(inv), _ptr(w), _ptr(sums), count, B, C, HW,
                                           float(alpha), _ptr(dx), None, None, _stream()), "nnb_bn_backward_apply")
    else:
        _check(L.nnb_bn_backward_apply(_ptr(x), _ptr(grad), _ptr(mean), _ptr(inv), _ptr(w), _ptr(sums), count, B, C, HW,
                                       float(alpha), _ptr(dx), _ptr(dw), _ptr(db), _stream()), "nnb_bn_backward_apply")
    return dx, dw, db


# ---- fused attention (short sequences) -------------------------------------------------------------------------
def attention_supported(Tq, Tk, D):
    return bool(lib().nnb_attention_supported(int(Tq), int(Tk), int(D)))


def _mask_args(mask, shape4):
    """mask = None or (tensor, kind, cmp): broadcast strides over (B, H, Tq, Tk)."""
    if mask is None:
        return None, 0, 0.0, None, None
    t, kind, cmp = mask
    want = torch.int32 if kind == 2 else torch.float32
    if t.dtype != want:
        t = t.to(want)
    t = t.expand(shape4)
    return t, kind, float(cmp), _I64x4(*[int(v) for v in t.stride()]), t


def attention_forward(q, kT, v, mask, fill, scale, p, ticket, want_planes=False):
    """q (B,H,Tq,D), kT (B,H,D,Tk), v (B,H,Tk,D): any strides. Returns (out, attn, planes): out is a (B,H,Tq,D)
    VIEW of a (B,Tq,H,D)-contiguous buffer (so `.transpose(0,2,1,3).reshape(B,Tq,H*D)` is free), attn is the
    post-dropout (B,H,Tq,Tk) tensor the example returns."""
    require_device()
    B, H, Tq, D = q.shape
    Tk = kT.shape[3]
    out = torch.empty((B, Tq, H, D), dtype=torch.float32, device="cuda")
    attn = torch.empty((B, H, Tq, Tk), dtype=torch.float32, device="cuda")
    buf, prec = _new_planes(B * Tq, H * D) if want_planes else (None, _state["prec"])
    mt, kind, cmp, ms, _keep = _mask_args(mask, (B, H, Tq, Tk))
    seed, call_id, epoch, dev = ticket if ticket is not None else (0, 0, 0, None)
    _check(lib().nnb_attention_forward(_ptr(q), _strides4(q), _ptr(kT), _strides4(kT), _ptr(v), _strides4(v), _ptr(mt), kind,
                                       cmp, ms, float(fill), float(scale), float(p), seed, call_id, epoch, _ptr(dev),
                                       _ptr(attn), _ptr(out), _ptr(buf), prec, B, H, Tq, Tk, D, _stream()),
           "nnb_attention_forward")
    planes = (planes_key(out.reshape(B * Tq, H * D), prec), buf) if buf is not None else None
    return out.permute(0, 2, 1, 3), attn, planes


def attention_backward(q, kT, v, mask, fill, scale, p, ticket, grad):
    """Returns (dq, dkT, dv) shaped like q, kT, v: views of (B,T,H,D)-contiguous buffers."""
    require_device()
    B, H, Tq, D = q.shape
    Tk = kT.shape[3]
    grad = grad if grad.dtype == torch.float32 else grad.to(torch.float32)
    pitch = 0
    if Tq == Tk:
        # dq | dk | dv as the column blocks of ONE [B*T, 3*H*D] matrix: a fused q/k/v projection (nn/layers/linear.py:
        # _SiblingGroup) takes it as its upstream gradient without a gather
        packed = torch.empty((B, Tq, 3, H, D), dtype=torch.float32, device="cuda")
        dq, dk, dv = packed[:, :, 0], packed[:, :, 1], packed[:, :, 2]
        pitch = 3 * H * D
    else:
        dq = torch.empty((B, Tq, H, D), dtype=torch.float32, device="cuda")
        dk = torch.empty((B, Tk, H, D), dtype=torch.float32, device="cuda")
        dv = torch.empty((B, Tk, H, D), dtype=torch.float32, device="cuda")
    mt, kind, cmp, ms, _keep = _mask_args(mask, (B, H, Tq, Tk))
    seed, call_id, epoch, dev = ticket if ticket is not None else (0, 0, 0, None)
    _check(lib().nnb_attention_backward(_ptr(q), _strides4(q), _ptr(kT), _strides4(kT), _ptr(v), _strides4(v), _ptr(mt), kind,
                                        cmp, ms, float(fill), float(scale), float(p), seed, call_id, epoch, _ptr(dev),
                                        _ptr(grad), _strides4(grad), _ptr(dq), _ptr(dk), _ptr(dv), pitch, _state["prec"], B, H, Tq,
                                        Tk, D,
                                        _stream()), "nnb_attention_backward")
    return dq.permute(0, 2, 1, 3), dk.permute(0, 2, 3, 1), dv.permute(0, 2, 1, 3)


def cross_entropy_forward(logits, targets, ignore_index=-100, reduction="mean"):
    """Returns (loss, saved) -- loss is a 0-d tensor (mean/sum) or [rows]; saved feeds the backward."""
    require_device()
    logits = _f32c(logits)
    rows, C = logits.shape
    tgt = targets.reshape(-1)
    if tgt.dtype != torch.int32:
        tgt = tgt.to(torch.int32)
    tgt = tgt.contiguous()
    row_loss = torch.empty(rows, dtype=torch.float32, device="cuda")
    lse = torch.empty(rows, dtype=torch.float32, device="cuda")
    red = _REDUCTION[reduction]
    scal = torch.empty(2, dtype=torch.float32, device="cuda") if red else None
    _check(lib().nnb_cross_entropy_forward(_ptr(logits), _ptr(tgt), rows, C, int(ignore_index), red, _ptr(row_loss),
                                           _ptr(lse), _ptr(scal[0:1]) if red else None,
                                           _ptr(scal[1:2]) if red else None, _stream()), "nnb_cross_entropy_forward")
    loss = scal[0] if red else row_loss
    return loss, (logits, tgt, lse, scal[1:2] if red else None, int(ignore_index))


def cross_entropy_backward(saved, upstream):
    logits, tgt, lse, inv, ignore_index = saved
    rows, C = logits.shape
    up = _f32c(upstream).reshape(-1)
    per_row = 1 if up.numel() == rows and rows > 1 and inv is None else 0
    d = torch.empty_like(logits)
    _check(lib().nnb_cross_entropy_backward(_ptr(logits), _ptr(tgt), _ptr(lse), _ptr(inv), _ptr(up), per_row, rows, C,
                                            ignore_index, _ptr(d), _stream()), "nnb_cross_entropy_backward")
    return d


def _force(obj):
    """Materialise deferred results (neunet/autograd.py: _Deferred) nested in lists / tuples / dicts."""
    if isinstance(obj, (list, tuple)):
        for o in obj:
            _force(o)
    elif isinstance(obj, dict):
        for o in obj.values():
            _force(o)
    elif hasattr(obj, "grad_fn") and hasattr(obj, "data"):
        obj.data  # noqa: B018


def cross_entropy_linear_backward(saved, upstream, x, w, need_dx=True, need_db=True, owner=None, x_staged=None,
                                  dw_out=None, db_out=None):
    """Backward of CrossEntropy(Linear(x)) in one go (row N3): dlogits leave the loss kernel as the bf16 operand planes of the
    Linear's dgrad / wgrad (never as a fp32 tensor), db as their column sums. Returns (dX like x, dW (N, K), db (1, N))."""
    require_device()
    L = lib()
    logits, tgt, lse, inv, ignore_index = saved
    rows, C = logits.shape
    N, K = w.shape
    assert C == N
    x2 = _f32c(x).reshape(-1, K)
    M = x2.shape[0]
    assert M == rows
    w = _f32c(w)
    prec = _state["prec"]
    up = _f32c(upstream).reshape(-1)
    dz = torch.empty(L.nnb_weight_staged_bytes(rows, C, prec), dtype=torch.uint8, device="cuda")
    db = None
    if need_db:
        db = db_out if _usable_out(db_out, (1, N)) else torch.empty((1, N), dtype=torch.float32, device="cuda")
    ws = _workspace(max(L.nnb_cross_entropy_staged_workspace_bytes(rows, C), L.nnb_linear_workspace_bytes(M, K, N, prec, 1)))
    _check(L.nnb_cross_entropy_backward_staged(_ptr(logits), _ptr(tgt), _ptr(lse), _ptr(inv), _ptr(up), rows, C, ignore_index,
                                               _ptr(dz), prec, _ptr(db), _ptr(ws), ws.numel(), _stream()),
           "nnb_cross_entropy_backward_staged")
    dx = torch.empty((M, K), dtype=torch.float32, device="cuda") if need_dx else None
    dw = dw_out if _usable_out(dw_out, (N, K)) else torch.empty((N, K), dtype=torch.float32, device="cuda")
    wst = _staged_weight(owner, w, N, K) if need_dx else None
    if x_staged is not None and getattr(x_staged, "_b200_prec", None) != prec:
        x_staged = None
    _check(L.nnb_linear_backward_staged(_ptr(x2), _ptr(w), _ptr(dz), _ptr(dx), _ptr(dw), M, K, N, prec, _ptr(wst), _ptr(x_staged),
                                        _ptr(ws), ws.numel(), _stream()), "nnb_linear_backward_staged")
    if dx is not None:
        dx = dx.reshape(x.shape)
    return dx, dw, db


# ---- whole-step CUDA graphs ---------------------------------------------------------------------------------
class GraphedStep:
    """Capture ``fn(*inputs)`` -- typically one whole training step (forward, loss, backward,
    optimizer.step) -- into a CUDA graph and replay it with new input VALUES.

    B200 runs the named small configs (MLP 784-128-10, conv classifier) in a few microseconds per
    kernel; the Python tape would otherwise be 10-100x slower than the GPU. ``inputs`` are neunet
    Tensors on "cuda" whose storage becomes the graph's static input buffers; ``__call__`` copies
    new host/device data into them and replays. Everything the step launches goes to the capturing
    stream because every C-ABI call takes torch's current stream. Optimizers must be created (and
    have taken their first eager step during warm-up) before capture.
    """

    def __init__(self, fn, inputs, optimizer=None, warmup=3, pdl_retry=True):
        """pdl_retry=False: a failed capture raises at once (multi-rank callers must agree on any retry
        collectively -- see bench.py -- because `fn` contains collectives)."""
        require_device()
        self.inputs = list(inputs)
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                _force(fn(*self.inputs))
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if optimizer is not None and getattr(optimizer, "_fused", None) is not None:
            optimizer._fused.set_step(optimizer.t)
        self.optimizer = optimizer
        _state["capture_epoch"] += 1  # invalidates every staged-operand cache made outside this capture
        rng_prepare_capture()
        t_before = optimizer.t if optimizer is not None else 0
        try:
            self._capture(fn)
        except Exception:
            if optimizer is not None:
                optimizer.t = t_before
            # programmatic-dependent-launch edges are the one capture ingredient older drivers may refuse:
            # retry once with plain stream serialization before giving up
            if not pdl_retry:
                raise
            prev = lib().nnb_set_pdl(0)
            if not prev:
                raise
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
            _state["capture_epoch"] += 1
            self._capture(fn)
        _state["capture_epoch"] += 1
        if optimizer is not None:
            optimizer.t -= 1  # the capture pass itself launched nothing
        self._dev_t = optimizer.t if optimizer is not None else 0  # value of the device step counter

    def _capture(self, fn):
        self.graph = torch.cuda.CUDAGraph()
        _rng["graph_used"] = False
        with torch.cuda.graph(self.graph):
            self.outputs = fn(*self.inputs)
            _force(self.outputs)  # pending (deferred) results must be launched INSIDE the capture
        self._uses_rng = _rng["graph_used"]  # dropout inside: bump the device epoch before every replay

    def load(self, *arrays, non_blocking=True):
        """Copy new values (pinned torch tensors / device tensors / NumPy arrays) into the static inputs."""
        for dst, src in zip(self.inputs, arrays):
            if isinstance(src, np.ndarray):
                src = torch.from_numpy(src)
            dst.data.copy_(src, non_blocking=non_blocking)

    def replay(self):
        opt = self.optimizer
        if opt is not None and opt.t != self._dev_t and getattr(opt, "_fused", None) is not None:
            opt._fused.set_step(opt.t)  # eager steps were interleaved: re-sync the device counter
        if self._uses_rng:
            _check(lib().nnb_rng_advance(_ptr(_rng_epoch_dev()), _stream()), "nnb_rng_advance")
        self.graph.replay()
        if opt is not None:
            opt.t += 1
            self._dev_t = opt.t
        return self.outputs

    def __call__(self, *arrays):
        if arrays:
            self.load(*arrays)
        return self.replay()


# ---- multi-tensor Adam / AdamW ---------------------------------------------------------------------------
class FusedAdam:
    """Owns an ``nnb_adamw`` handle for a fixed list of fp32 parameter tensors."""

    def __init__(self, params, ms, vs):
        require_device()
        self.n = len(params)
        self._keep = (list(params), list(ms), list(vs))  # keep storage alive
        arr = c_void_p * self.n
        self._p = arr(*[t.data_ptr() for t in params])
        self._m = arr(*[t.data_ptr() for t in ms])
        self._v = arr(*[t.data_ptr() for t in vs])
        self._sizes = (c_int64 * self.n)(*[t.numel() for t in params])
        self._h = c_void_p()
        for t in list(params) + list(ms) + list(vs):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("FusedAdam: parameters and moments must be contiguous fp32")
        _check(lib().nnb_adamw_create(ctypes.byref(self._h), self.n, self._p, None, self._m, self._v, self._sizes,
                                      _stream()), "nnb_adamw_create")

    def step(self, grads, lr, betas, eps, weight_decay, t, mode, grad_scale=1.0):
        """grads: list of tensors or None (None = skipped, optim.py:21-22)."""
        self._grads_keep = grads
        arr = (c_void_p * self.n)(*[(g.data_ptr() if g is not None else None) for g in grads])
        L = lib()
        _check(L.nnb_adamw_set_grads(self._h, arr, _stream()), "nnb_adamw_set_grads")
        if torch.cuda.is_current_stream_capturing():
            t = 0  # captured step: use the device-resident counter so replays keep advancing it
        _check(L.nnb_adamw_step(self._h, float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                int(t), mode, float(grad_scale), _stream()), "nnb_adamw_step")
        weights_changed()

    def set_grads(self, grads):
        """Upload the gradient pointer table (None = skipped) for the range steps that follow."""
        self._grads_keep = grads
        arr = (c_void_p * self.n)(*[(g.data_ptr() if g is not None else None) for g in grads])
        _check(lib().nnb_adamw_set_grads(self._h, arr, _stream()), "nnb_adamw_set_grads")

    def step_range(self, first, count, lr, betas, eps, weight_decay, t, mode, grad_scale=1.0, advance=False):
        """The update for parameters [first, first + count) only (after set_grads); see nnb_adamw_step_range."""
        if torch.cuda.is_current_stream_capturing():
            t = 0
        _check(lib().nnb_adamw_step_range(self._h, int(first), int(count), float(lr), float(betas[0]), float(betas[1]),
                                          float(eps), float(weight_decay), int(t), mode, float(grad_scale), int(bool(advance)),
                                          _stream()), "nnb_adamw_step_range")

    def sync_staging(self, owners):
        """Fused weight staging: for every parameter that a Linear layer has asked bf16 planes for
        (``_b200_wants_staging``), the step kernel also emits those planes, and the parameter's staging
        cache entry is stamped valid for the new weights. Called by the optimizer around each step."""
        prec = _state["prec"]
        want = tuple(bool(getattr(o, "_b200_wants_staging", False)) and o.data.ndim == 2 for o in owners)
        if torch.cuda.is_current_stream_capturing():
            if getattr(self, "_stage_key", None) != (want, prec):
                return  # tables cannot be re-uploaded inside a capture; the layers keep converting on their own
        elif getattr(self, "_stage_key", None) != (want, prec):
            bufs, cols = [], []
            for o, w in zip(owners, want):
                if w:
                    r, c = o.data.shape
                    bufs.append(torch.empty(lib().nnb_weight_staged_bytes(r, c, prec), dtype=torch.uint8, device="cuda"))
                    cols.append(c)
                else:
                    bufs.append(None)
                    cols.append(1)
            arr = (c_void_p * self.n)(*[(b.data_ptr() if b is not None else None) for b in bufs])
            carr = (c_int64 * self.n)(*cols)
            _check(lib().nnb_adamw_set_staging(self._h, arr if any(want) else None, carr, prec, _stream()),
                   "nnb_adamw_set_staging")
            self._stage_bufs, self._stage_key = bufs, (want, prec)
        self._stage_live = True

    def stamp_staging(self, owners, grads):
        """After a step: mark the emitted planes as the staged form of the CURRENT weights."""
        if not getattr(self, "_stage_live", False):
            return
        want, prec = self._stage_key
        for o, w, b, g in zip(owners, want, self._stage_bufs, grads):
            if not w or b is None:
                continue
            prev = getattr(o, "_b200_staged", None)
            if g is None and not (prev is not None and prev.persistent and prev.buf is b):
                continue  # parameter skipped by the kernel and never emitted: nothing valid to stamp
            sw = _StagedWeight()
            sw.buf, sw.persistent = b, True
            sw.key = (o.data.data_ptr(), o.data._version, _state["weights_epoch"], prec) + tuple(o.data.shape)
            o._b200_staged = sw

    def set_step(self, t):
        """Seed the device-resident step counter (steps already taken) before capturing a graph."""
        _check(lib().nnb_adamw_set_step(self._h, int(t), _stream()), "nnb_adamw_set_step")

    def __del__(self):
        try:
            if self._h:
                lib().nnb_adamw_destroy(self._h)
                self._h = c_void_p()
        except Exception:
            pass
