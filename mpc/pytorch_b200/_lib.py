"""ctypes binding of libmpcb200.so (the C ABI in include/mpcb200.h).

The product path has NO CPU fallback: if the library is missing, or a tensor is
not a CUDA tensor, the calls here raise.  torch is used only for device memory
and the current stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPCB200_LIB", os.path.join(_HERE, "libmpcb200.so"))   # override: developer experiments


class Dims(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in (
        "B", "T", "n", "m", "F_T", "has_f", "bounds_kind", "has_zero_mask",
        "has_delta_u", "max_ls_iter", "pnqp_max_iter", "do_rollout", "dynamics_kind", "reserved0")] + \
        [(k, ctypes.c_int64) for k in ("C_tstride", "c_tstride", "F_tstride", "f_tstride")]



class Params(ctypes.Structure):
    _fields_ = [(k, ctypes.c_double) for k in ("u_lo", "u_hi", "delta_u", "ls_decay")] + [("dyn", ctypes.c_double * 8)]


class MpcB200Error(RuntimeError):
    pass


_lib = None

# every symbol include/mpcb200.h declares
EXPORTED_SYMBOLS = (
    "mpcb200_lqr_step_f32", "mpcb200_lqr_step_f64", "mpcb200_lqr_grad_f32", "mpcb200_lqr_grad_f64",
    "mpcb200_rollout_f32", "mpcb200_rollout_f64", "mpcb200_pnqp_f32", "mpcb200_pnqp_f64",
    "mpcb200_lqr_adjoint_f32", "mpcb200_lqr_adjoint_f64", "mpcb200_adjoint_workspace_bytes",
    "mpcb200_dyn_rollout_f32", "mpcb200_dyn_rollout_f64", "mpcb200_dyn_linearize_f32", "mpcb200_dyn_linearize_f64",
    "mpcb200_supported", "mpcb200_supported_list", "mpcb200_launch_count",
    "mpcb200_step_smem_bytes", "mpcb200_step_prefers_workspace", "mpcb200_version", "mpcb200_strerror",
)


def lib():
    """Load (once) and return the shared library; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpcB200Error(
            f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C mpc/pytorch_b200/csrc`). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    step_args = [ctypes.POINTER(Dims), ctypes.POINTER(Params)] + [vp] * 22
    for name in ("mpcb200_lqr_step_f32", "mpcb200_lqr_step_f64"):
        fn = getattr(L, name)
        fn.argtypes = step_args
        fn.restype = ctypes.c_int
    grad_args = [ctypes.POINTER(Dims)] + [vp] * 15
    for name in ("mpcb200_lqr_grad_f32", "mpcb200_lqr_grad_f64"):
        fn = getattr(L, name)
        fn.argtypes = grad_args
        fn.restype = ctypes.c_int
    for name in ("mpcb200_rollout_f32", "mpcb200_rollout_f64"):
        fn = getattr(L, name)
        fn.argtypes = [ctypes.POINTER(Dims)] + [vp] * 6
        fn.restype = ctypes.c_int
    for name in ("mpcb200_pnqp_f32", "mpcb200_pnqp_f64"):
        fn = getattr(L, name)
        fn.argtypes = [ctypes.c_int32, ctypes.c_int32] + [vp] * 5 + [ctypes.c_int32] + [vp] * 6
        fn.restype = ctypes.c_int
    for name in ("mpcb200_lqr_adjoint_f32", "mpcb200_lqr_adjoint_f64"):
        fn = getattr(L, name)
        fn.argtypes = [ctypes.POINTER(Dims), ctypes.POINTER(Params)] + [vp] * 15 + [ctypes.c_size_t, vp]
        fn.restype = ctypes.c_int
    L.mpcb200_adjoint_workspace_bytes.argtypes = [ctypes.POINTER(Dims), ctypes.c_int32]
    L.mpcb200_adjoint_workspace_bytes.restype = ctypes.c_size_t
    for name in ("mpcb200_dyn_rollout_f32", "mpcb200_dyn_rollout_f64"):
        fn = getattr(L, name)
        fn.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_double), ctypes.c_int32, ctypes.c_int32] + [vp] * 4
        fn.restype = ctypes.c_int
    for name in ("mpcb200_dyn_linearize_f32", "mpcb200_dyn_linearize_f64"):
        fn = getattr(L, name)
        fn.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_double), ctypes.c_int32, ctypes.c_int32] + [vp] * 5
        fn.restype = ctypes.c_int
    L.mpcb200_supported.argtypes = [ctypes.c_int32, ctypes.c_int32]
    L.mpcb200_supported.restype = ctypes.c_int
    L.mpcb200_supported_list.argtypes = [ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]
    L.mpcb200_supported_list.restype = ctypes.c_int
    L.mpcb200_launch_count.argtypes = []
    L.mpcb200_launch_count.restype = ctypes.c_uint64
    L.mpcb200_step_smem_bytes.argtypes = [ctypes.POINTER(Dims), ctypes.c_int32]
    L.mpcb200_step_smem_bytes.restype = ctypes.c_size_t
    L.mpcb200_step_prefers_workspace.argtypes = [ctypes.POINTER(Dims), ctypes.c_int32]
    L.mpcb200_step_prefers_workspace.restype = ctypes.c_int
    L.mpcb200_version.argtypes = []
    L.mpcb200_version.restype = ctypes.c_int
    L.mpcb200_strerror.argtypes = [ctypes.c_int]
    L.mpcb200_strerror.restype = ctypes.c_char_p
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise MpcB200Error(f"{what} failed: [{rc}] {lib().mpcb200_strerror(rc).decode()}")


def supported_pairs():
    L = lib()
    n = L.mpcb200_supported_list(None, 0)
    buf = (ctypes.c_int32 * (2 * n))()
    L.mpcb200_supported_list(buf, n)
    return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]


def launch_count():
    return int(lib().mpcb200_launch_count())


def ptr(t):
    """Device pointer of a dense CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise MpcB200Error("mpc.pytorch_b200 runs on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise MpcB200Error("internal error: non-contiguous tensor reached the C ABI")
    return ctypes.c_void_p(t.data_ptr())


def ptr_view(t):
    """Device pointer of the first element of a CUDA tensor whose layout the caller has already validated
    (time-strided [T, B, ...] inputs: contiguous [B, ...] slices, any time stride)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise MpcB200Error("mpc.pytorch_b200 runs on CUDA tensors only (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


def stream_handle(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
