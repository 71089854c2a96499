"""
ctypes binding of libzopt_b200.so (include/zopt_b200.h) and the tensor plumbing shared by the API
modules.  PyTorch tensors are used only as device buffers: the library receives raw device
pointers, element strides and the current CUDA stream.

There is no CPU fallback: importing this module without the built library, or calling an entry
point without a CUDA device, raises.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzopt_b200.so")

ZB_F32, ZB_F64 = 0, 1
ZB_MAX_N, ZB_MAX_M = 16, 8
MODEL_LINEAR, MODEL_QUADCOPTER = 0, 1


class ZbArr(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride_b", C.c_int64), ("stride_t", C.c_int64)]


class ZbModel(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("m", C.c_int32), ("has_wind", C.c_int32), ("dt", C.c_double),
                ("wind", C.c_double * 3), ("A", ZbArr), ("B", ZbArr)]


class ZbCost(C.Structure):
    _fields_ = [("Q", ZbArr), ("R", ZbArr), ("Qf", ZbArr)]


class ZbAdmmOpts(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("check_every", C.c_int32), ("rho", C.c_double), ("sigma", C.c_double),
                ("alpha", C.c_double), ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double)]


_P = C.c_void_p
_AP = C.POINTER(ZbArr)
_i32, _i64, _f64, _sz = C.c_int32, C.c_int64, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/zopt_b200.h declaration by declaration
SIGNATURES = {
    "zb_version": (_i32, []),
    "zb_last_error": (_i32, [C.c_char_p, _sz]),
    "zb_device_info": (_i32, [_i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_sz)]),
    "zb_lqr_dfh": (_i32, [_i32, _i32, _P, _i64, _i32, _i32, _i32, _i32, _AP, _AP, _AP, _AP, _P, _P]),
    "zb_lqr_dfh_flags": (_i32, [_i32, _i32, _P, _i64, _i32, _i32, _i32, _i32, _AP, _AP, _AP, _AP, _i32, _P, _P]),
    "zb_lqr_care_rk4": (_i32, [_i32, _i32, _P, _i64, _i32, _i32, _i32, _i32, _f64, _AP, _AP, _AP, _AP, _AP, _P]),
    "zb_lqr_bilinear": (_i32, [_i32, _i32, _P, _i64, _i32, _i32, _i32, _i32] + [_AP] * 9 + [_P, _P]),
    "zb_lqr_bilinear_flags": (_i32, [_i32, _i32, _P, _i64, _i32, _i32, _i32, _i32] + [_AP] * 9 + [_i32, _P, _P]),
    "zb_quad_dynamics": (_i32, [_i32, _i32, _P, _i64, _P, _P, C.POINTER(_f64), _P]),
    "zb_quad_linearize": (_i32, [_i32, _i32, _P, _i64, _P, _P, C.POINTER(_f64), _f64, _P, _P]),
    "zb_quad_hess_contract": (_i32, [_i32, _i32, _P, _i64, _P, _P, C.POINTER(_f64), _f64, _P, _P]),
    "zb_ilqr_rollout": (_i32, [_i32, _i32, _P, _i64, _i32, C.POINTER(ZbModel), C.POINTER(ZbCost), _P, _P, _P, _P, _P,
                               _f64, _P, _P, _P]),
    "zb_ilqr_forward_pass": (_i32, [_i32, _i32, _P, _i64, _i32, C.POINTER(ZbModel), C.POINTER(ZbCost), _P, _P, _P, _P,
                                    _P, _P, _P, _P, _P, _P]),
    "zb_ilqr_backward": (_i32, [_i32, _i32, _P, _i64, _i32, _i32, _i32, _i32] + [_AP] * 14 + [_P] * 5),
    "zb_pd_clamp": (_i32, [_i32, _i32, _P, _i64, _i32, _f64, _P, _P]),
    "zb_ilqr_workspace_bytes": (_sz, [_i32, _i64, _i32, _i32, _i32]),
    "zb_ilqr_solve": (_i32, [_i32, _i32, _P, _i64, _i32, _i32, C.POINTER(ZbModel), C.POINTER(ZbCost), _P, _P, _i32,
                             _f64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _sz]),
    "zb_mpc_workspace_bytes": (_sz, [_i32, _i64, _i32, _i32, _i32]),
    "zb_mpc_lqr_solve": (_i32, [_i32, _i32, _P, _i64, _i32, _i32, _i32] + [_AP] * 9 + [_i32, _P,
                                C.POINTER(ZbAdmmOpts), _P, _P, _P, _P, _P, _P, _sz]),
    "zb_mpc_box_tables_bytes": (_sz, [_i32, _i32]),
    "zb_mpc_box_workspace_bytes": (_sz, [_i32, _i64, _i32]),
    "zb_mpc_box_build_tables": (_i32, [_i32, _i32, _P, _i32] + [C.POINTER(_f64)] * 5 + [_f64, _P, _sz]),
    "zb_mpc_box_solve": (_i32, [_i32, _i32, _P, _i64, _i32] + [C.POINTER(_f64)] * 6 + [_P, _P, C.POINTER(ZbAdmmOpts), _i32,
                                _P, _P, _P, _P, _P, _P, _sz]),
    "zb_mpc_box_closed_loop_workspace_bytes": (_sz, [_i32, _i64, _i32]),
    "zb_mpc_box_closed_loop": (_i32, [_i32, _i32, _P, _i64, _i32, _i32] + [C.POINTER(_f64)] * 6 + [_P, _P, C.POINTER(ZbAdmmOpts),
                                      _i32, _f64, _P, _P, _P, _P, _P, _sz]),
    "zb_mpc_closed_loop_quad": (_i32, [_i32, _i32, _P, _i64, _i32, _i32, _f64, C.POINTER(_f64), _AP, _AP, _AP, _i32, _P, _P, _P]),
    "zb_peak_fma": (_i32, [_i32, _i32, C.POINTER(_f64), C.POINTER(_f64)]),
}


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise RuntimeError(
            f"zopt_b200: CUDA library not found at {path}. Build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C zopt_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load_library()


def last_error():
    buf = C.create_string_buffer(512)
    lib.zb_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc):
    """0 OK; <0 argument error -> ValueError/TypeError; >0 CUDA error -> RuntimeError."""
    if rc == 0:
        return
    msg = last_error()
    if rc == -2:
        raise TypeError(f"zopt_b200: {msg}")
    if rc == -3:
        raise NotImplementedError(f"zopt_b200: {msg}")
    if rc < 0:
        raise ValueError(f"zopt_b200: {msg}")
    raise RuntimeError(f"zopt_b200: CUDA error {rc}: {msg}")


# ------------------------------------------------------------------------------------------------ tensors
def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("zopt_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def pick_device(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def pick_dtype(*xs):
    """fp32 only if every floating input is fp32 (numpy / python numbers default to fp64, like the
    reference, which always runs with jax_enable_x64 -- zopt/quadcopter.py:7)."""
    seen = []
    for x in xs:
        if isinstance(x, torch.Tensor):
            if x.is_floating_point():
                seen.append(x.dtype)
        elif isinstance(x, np.ndarray):
            if x.dtype.kind == "f":
                seen.append(torch.float32 if x.dtype == np.float32 else torch.float64)
    if seen and all(d == torch.float32 for d in seen):
        return torch.float32
    return torch.float64


def to_dev(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=device)


def dcode(dtype):
    return ZB_F32 if dtype == torch.float32 else ZB_F64


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _inner_contig(t, k):
    """last k dims are dense row-major"""
    exp = 1
    for d in range(t.ndim - 1, t.ndim - 1 - k, -1):
        if t.shape[d] != 1 and t.stride(d) != exp:
            return False
        exp *= t.shape[d]
    return True


class View:
    """A tensor described for the library: block = last `k` dims; optional time axis before it;
    optional batch axis before that.  Missing axes get stride 0, so reference-shaped (un-batched)
    or expanded (stride-0) operands are never materialised."""

    def __init__(self, t, block_ndim, has_time, batched):
        if not _inner_contig(t, block_ndim):
            t = t.contiguous()
        self.t = t  # keep alive
        lead = t.ndim - block_ndim
        sb = st = 0
        if has_time:
            st = t.stride(lead - 1) if t.shape[lead - 1] > 1 else 0
            if batched:
                sb = t.stride(lead - 2) if t.shape[lead - 2] > 1 else 0
        elif batched:
            sb = t.stride(lead - 1) if t.shape[lead - 1] > 1 else 0
        self.arr = ZbArr(t.data_ptr(), sb, st)

    def ref(self):
        return C.byref(self.arr)


_FLAG_CACHE = {}


def cached_matrix_flag(t, name, fn):
    """A boolean property of a batch of matrices that needs a device->host read (exact diagonality, symmetry), cached by
    tensor OBJECT identity (weak reference) and in-place version counter -- never by address, so a recycled allocation can
    not produce a stale answer.  Re-building a problem around the same weights every MPC step costs nothing."""
    import weakref
    ent = _FLAG_CACHE.get((id(t), name))
    if ent is not None and ent[0]() is t and ent[1] == t._version:
        return ent[2]
    if len(_FLAG_CACHE) > 512:
        _FLAG_CACHE.clear()
    res = bool(fn(t))
    _FLAG_CACHE[(id(t), name)] = (weakref.ref(t), t._version, res)
    return res


def is_symmetric(t):
    """exact symmetry of the trailing two axes (the (12,4) fast kernels read the lower triangle of the weights only)"""
    return cached_matrix_flag(t, "sym", lambda a: bool((a - a.transpose(-1, -2)).abs().max() == 0) if a.numel() else True)


def null_arr():
    return C.byref(ZbArr(None, 0, 0))
