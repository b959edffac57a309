"""ctypes binding of libx2gnn.so (include/x2gnn.h).  The library handle is module-global --
never stored on an nn.Module -- so modules stay deep-copyable / picklable
(train_ema.py:47 deep-copies the model for EMA).

There is NO CPU fallback: if the shared library is missing, or a tensor is not on a
CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("X2GNN_LIB") or os.path.join(_HERE, "lib", "libx2gnn.so")   # X2GNN_LIB: A/B builds

_lock = threading.Lock()
_lib = None
_checked_devices = set()

c_f32p = C.c_void_p
c_i32p = C.c_void_p
c_i64p = C.c_void_p
c_stream = C.c_void_p


class WgradJob(C.Structure):      # x2_wgrad_job
    _fields_ = [("Y", C.c_void_p), ("ldy", C.c_int64), ("X", C.c_void_p), ("ldx", C.c_int64),
                ("dW", C.c_void_p), ("lddw", C.c_int64), ("db", C.c_void_p)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("E", C.c_int64), ("T", C.c_int64),
        ("D", C.c_int32), ("H", C.c_int32), ("C", C.c_int32), ("S", C.c_int32),
        ("R", C.c_int32), ("A", C.c_int32),
        ("fuse_skip", C.c_int32), ("mode", C.c_int32),
        ("dropout_p", C.c_float), ("tgt_sorted", C.c_int32), ("seed", C.c_uint64),
        ("x", c_f32p), ("rbf", c_f32p), ("sbf", c_f32p), ("edge_attr", c_f32p),
        ("src", c_i32p), ("tgt", c_i32p), ("rowptr_tgt", c_i32p), ("order_tgt", c_i32p),
        ("rowptr_src", c_i32p), ("order_src", c_i32p),
        ("w_rbf", c_f32p), ("w_q", c_f32p), ("b_q", c_f32p), ("w_k", c_f32p), ("b_k", c_f32p),
        ("w_v", c_f32p), ("b_v", c_f32p), ("w_edge", c_f32p), ("w_sbf", c_f32p), ("b_sbf", c_f32p),
        ("w_skip", c_f32p), ("b_skip", c_f32p),
        ("ea_rows", C.c_int64), ("ea_index", c_i32p), ("ea_rowptr", c_i32p), ("ea_order", c_i32p),
        ("items", c_i32p), ("itemptr", c_i32p), ("items_bound", C.c_int64),
        ("nblk", C.c_int64), ("blk_max_src", C.c_int32), ("sbf_L", C.c_int32), ("sbf_R", C.c_int32),
        ("blk_max_trip", C.c_int32), ("blk_max_tgt", C.c_int32), ("pad0_", C.c_int32),
        ("blk_sptr", c_i32p), ("blk_tptr", c_i32p), ("blk_tord", c_i32p), ("blk_tpos", c_i32p),
        ("sbf_tab", c_f32p), ("angles", c_f32p),
    ]


class ConvSaved(C.Structure):
    _fields_ = [("qkvs", c_f32p), ("attn", c_f32p), ("lse", c_f32p), ("ea", c_f32p), ("sg", c_f32p),
                ("xs", c_f32p)]


class ConvGrads(C.Structure):
    _fields_ = [(n, c_f32p) for n in (
        "dx", "drbf", "dsbf", "dedge_attr", "dw_rbf", "dw_q", "db_q", "dw_k", "db_k", "dw_v", "db_v",
        "dw_edge", "dw_sbf", "db_sbf", "dw_skip", "db_skip")]


# name -> (restype, argtypes); mirrors include/x2gnn.h one to one
_I64, _I32, _F, _SZ, _P = C.c_int64, C.c_int32, C.c_float, C.c_size_t, C.c_void_p
SIGNATURES = {
    "x2_version": (C.c_int, []),
    "x2_last_error": (C.c_char_p, []),
    "x2_device_check": (C.c_int, [C.c_int]),
    "x2_launch_count": (C.c_int64, []),
    "x2_timing_enable": (C.c_int, [C.c_int]),
    "x2_timing_read": (C.c_int, [_P, _P, C.c_int]),
    "x2_timing_phase_name": (C.c_char_p, [C.c_int]),
    "x2_dij": (C.c_int, [_P, _I64, _P, _P]),
    "x2_bonds_count": (C.c_int, [_P, _I64, _F, _P, _P, _SZ, _P]),
    "x2_bonds_fill": (C.c_int, [_P, _I64, _F, _P, _P, _I64, _P]),
    "x2_radius_graph_count": (C.c_int, [_P, _P, _P, _I64, _F, _P, _P, _SZ, _P]),
    "x2_radius_graph_fill": (C.c_int, [_P, _P, _P, _I64, _F, _P, _P, _I64, _P]),
    "x2_scan_workspace_bytes": (_SZ, [_I64]),
    "x2_triplets_workspace_bytes": (_SZ, [_I64, _I64]),
    "x2_triplets_count": (C.c_int, [_P, _I64, _I64, _P, _P, _P, _SZ, _P]),
    "x2_triplets_fill": (C.c_int, [_P, _I64, _I64, _P, _I64, _P, _P, _P, _P, _P, _SZ, _P]),
    "x2_meta_workspace_bytes": (_SZ, [_I64, _I64]),
    "x2_meta_build": (C.c_int, [_P, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "x2_items_bound": (C.c_int64, [_I64, _I64]),
    "x2_items_workspace_bytes": (_SZ, [_I64]),
    "x2_items_build": (C.c_int, [_P, _I64, _I64, _P, _P, _P, _SZ, _P]),
    "x2_blocks_workspace_bytes": (_SZ, [_I64, _I64]),
    "x2_blocks_build": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "x2_envelope_fwd": (C.c_int, [_P, _I64, _F, _I32, _F, _F, _F, _P, _P]),
    "x2_radial_fwd": (C.c_int, [_P, _P, _P, _I64, _I32, _F, _P, _P]),
    "x2_radial_bwd_workspace_bytes": (_SZ, [_I64, _I32]),
    "x2_radial_bwd": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _F, _P, _P, _P, _SZ, _P]),
    "x2_sbf_table": (C.c_int, [_P, _I64, _I32, _I32, _P, _P, _F, _F, _I32, _F, _F, _F, _P, _P]),
    "x2_sbf_fwd": (C.c_int, [_P, _P, _P, _I64, _I64, _I32, _I32, _P, _P]),
    "x2_angular_fwd": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "x2_bond_lengths": (C.c_int, [_P, _P, _P, _I64, _P, _P]),
    "x2_triplet_angles": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P]),
    "x2_envelope_bwd": (C.c_int, [_P, _P, _I64, _F, _I32, _F, _F, _F, _P, _P]),
    "x2_angular_bwd": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "x2_sbf_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _I32, _P, _P, _F, _F, _I32, _F, _F, _F,
                             _P, _P, _P]),
    "x2_tc_gemm_workspace_bytes": (_SZ, [_I32, _I32]),
    "x2_tc_gemm": (C.c_int, [_P, _I64, _I64, _I32, _P, _I64, _I64, _I32, _P, _P, _I64, _I32, _P, _SZ, _P]),
    "x2_tc_wgrad_workspace_bytes": (_SZ, [_I64, _I32]),
    "x2_tc_wgrad": (C.c_int, [_P, _I64, _P, _I64, _I64, _I32, _P, _I64, _P, _P, _SZ, _P]),
    "x2_tc_wgrad_batch": (C.c_int, [_P, _I32, _I64, _I32, _P, _SZ, _P]),
    "x2_sbfconv_plan": (C.c_int, [C.POINTER(ConvDesc)]),
    "x2_sbfconv_fwd_workspace_bytes": (_SZ, [C.POINTER(ConvDesc)]),
    "x2_sbfconv_bwd_workspace_bytes": (_SZ, [C.POINTER(ConvDesc)]),
    "x2_sbfconv_fwd": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(ConvSaved), _P, _P, _P, _SZ, _P]),
    "x2_sbfconv_bwd": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(ConvSaved), _P, C.POINTER(ConvGrads),
                                 _P, _SZ, _P]),
    "x2_collate_workspace_bytes": (_SZ, [_I64]),
    "x2_collate_sizes": (C.c_int, [_P, _I64, _P, _P, _I64, _P, _P, _P, _P, _SZ, _P]),
    "x2_collate_fill": (C.c_int, [_P, _I64, _I64, _P, _P, _P, _P, _P, _I32, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _I64, _P]),
    "x2_optim_workspace_bytes": (_SZ, [_I64]),
    "x2_optim_tail": (C.c_int, [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _F, _F, _P, _P, _P, _SZ, _P]),
    "x2_graph_layernorm_fwd": (C.c_int, [_P, _P, _I64, _I32, _F, _P, _P, _P]),
    "x2_graph_layernorm_bwd": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _P]),
    "x2_rbf_readout_fwd": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I32, _I32, _P, _P]),
    "x2_rbf_readout_bwd_workspace_bytes": (_SZ, [_I64, _I64, _I32, _I32]),
    "x2_rbf_readout_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _P, _SZ, _P]),
}


class X2Error(RuntimeError):
    pass


_tls = threading.local()      # .dev: index of the device of the tensors of the call in progress (require_cuda)


class _DeviceBound:
    """The library launches on the calling thread's CURRENT device and the stream it is handed.  Every
    entry point that takes a stream is wrapped so that it runs with the device of the call's tensors
    current (`require_cuda` records it; `stream()` returns that device's current stream): a model on cuda:1
    works while cuda:0 is the process's current device.  Entering the guard is a no-op when the device is
    already current (the one-process-per-GPU case)."""

    def __init__(self, handle):
        self._h = handle

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if SIGNATURES.get(name, (None, []))[1][-1:] != [_P] or name in ("x2_timing_read",):
            wrapped = fn                                    # no stream argument: nothing is launched
        else:
            def wrapped(*args, _fn=fn):
                dev = getattr(_tls, "dev", None)
                if dev is None or dev == torch.cuda.current_device():
                    return _fn(*args)
                with torch.cuda.device(dev):
                    return _fn(*args)
        setattr(self, name, wrapped)
        return wrapped


def lib():
    """Load (once) and return the library handle.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise X2Error(
                    f"{LIB_PATH} not found: build it with `python x2-gnn_b200/build.py` "
                    "(x2gnn_b200 has no CPU / PyTorch fallback)")
            h = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(h, name)      # AttributeError => header/library mismatch
                fn.restype = res
                fn.argtypes = args
            _lib = _DeviceBound(h)
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().x2_last_error().decode(errors="replace")
        raise X2Error(f"{what or 'libx2gnn'} failed (code {rc}): {msg}")


def require_cuda(*tensors, what: str):
    """All given tensors (None allowed) must live on the same CUDA device."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise X2Error(f"{what}: expected CUDA tensors, got a tensor on '{t.device}' "
                          "(x2gnn_b200 has no CPU path)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise X2Error(f"{what}: tensors on different devices ({dev} vs {t.device})")
    if dev is None:
        raise X2Error(f"{what}: no tensors given")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    _tls.dev = idx
    if idx not in _checked_devices:
        check(lib().x2_device_check(idx), "x2_device_check")
        _checked_devices.add(idx)
    return dev


def f32(t, what: str):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype} (the sm_100a kernels are fp32-I/O)")
    return t.contiguous()


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    """Current stream of the device of the call in progress (see _DeviceBound)."""
    dev = getattr(_tls, "dev", None)
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


NUM_PHASES = 8


def timing_enable(on: bool):
    check(lib().x2_timing_enable(1 if on else 0), "x2_timing_enable")


def timing_read():
    """-> {phase_name: (total_ms, calls)} accumulated since the last read."""
    ms = (C.c_double * NUM_PHASES)()
    calls = (C.c_int64 * NUM_PHASES)()
    check(lib().x2_timing_read(ms, calls, NUM_PHASES), "x2_timing_read")
    return {lib().x2_timing_phase_name(i).decode(): (ms[i], calls[i]) for i in range(NUM_PHASES)}


def launch_count() -> int:
    return int(lib().x2_launch_count())


def workspace(nbytes: int, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
