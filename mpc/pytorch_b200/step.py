"""LQRStep: one box-constrained LQR step as a torch.autograd.Function over the C ABI.

Mirrors the reference factory ``LQRStep(...)`` (mpc/lqr_step.py:22-38, 409): same
keyword names, defaults, call signature ``(x_init, C, c, F, f)``, return arity and
shapes, and the same gradient tuple ``(dx_init, dC, dc, dF, df)`` (:407).  The
arithmetic runs in csrc/lqr_step.cuh and csrc/lqr_grad.cuh.
"""
import ctypes
import threading

import torch
from torch.autograd import Function
from torch.nn import Module

from . import _lib
from ._lib import Dims, Params, MpcB200Error, check, ptr, ptr_view, stream_handle

PNQP_MAX_ITER = 20  # reference passes n_iter=20 (mpc/lqr_step.py:137)

# MPC's loop sets `_host_reads.defer` to a dict around its LQRStep calls: the per-step pnqp counters then stay
# on the device (no .item() per step) and MPC reads them with its own stop-test scalars.  Thread local; the
# LQRStep(...) signature itself stays the reference's.
_host_reads = threading.local()


def _is_empty(t):
    return t is None or t.nelement() == 0


def _dense(t, dtype=None):
    """Dense, detached, right-dtype view of `t` (no copy and no new tensor object when it already is)."""
    if t is None:
        return None
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.requires_grad:
        t = t.detach()
    return t if t.is_contiguous() else t.contiguous()


def _expect(name, t, shape, dev):
    """The kernels read raw device pointers: a wrong shape or a tensor on another GPU would be an
    out-of-bounds read, so fail here like the reference's indexing / eclamp size asserts would."""
    if t is None:
        return
    if tuple(t.shape) != tuple(shape):
        raise MpcB200Error(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    if t.device != dev:
        raise MpcB200Error(f"{name}: expected a tensor on {dev}, got {t.device}")


def _time_strided(t, dtype):
    """(tensor, mpcb200_dims time-stride field) for a [T, B, ...] input WITHOUT materialising views whose
    [B, ...] slices are contiguous: dense -> 0; stride-0 over time (`x.unsqueeze(0).expand(T, ...)`, an LTI
    `F`, reference mpc/mpc.py:205-226) -> MPCB200_TIME_INVARIANT; any other 16-byte aligned time stride -> it.
    Everything else (batch-expanded, transposed, ...) is copied to a dense tensor, as before."""
    if t is None:
        return None, 0
    if t.dtype != dtype:
        t = t.to(dtype)
    if t.requires_grad:
        t = t.detach()
    if t.is_contiguous():
        return t, 0
    if t.dim() >= 2 and t.shape[0] > 0 and t[0].is_contiguous():
        st0 = t.stride(0)
        if st0 == 0:
            return t, -1
        if st0 > 0 and (st0 * t.element_size()) % 16 == 0:
            return t, st0
    return t.contiguous(), 0


# ----------------------------------------------------------------------------------------------
# padding to a compiled (n,m) instance (compatibility path for shapes without an exact kernel)
# ----------------------------------------------------------------------------------------------
_pairs_cache = None
_pick_cache = {}
_smem_fits_cache = {}


def _pick_instance(n, m):
    hit = _pick_cache.get((n, m))
    if hit is not None:
        return hit
    hit = _pick_instance_uncached(n, m)
    _pick_cache[(n, m)] = hit
    return hit


def _pick_instance_uncached(n, m):
    global _pairs_cache
    if _pairs_cache is None:
        _pairs_cache = _lib.supported_pairs()
    if (n, m) in _pairs_cache:
        return n, m
    cands = [(N + M, N, M) for (N, M) in _pairs_cache if N >= n and M >= m]
    if not cands:
        raise MpcB200Error(
            f"(n_state={n}, n_ctrl={m}) exceeds every compiled kernel instance {_pairs_cache}; "
            "add it to mpc/pytorch_b200/csrc/instances.def and rebuild")
    _, N, M = min(cands)
    return N, M


class _Pad:
    """Embeds an (n,m) problem in an (N,M) one: padded states are 0 with zero dynamics/cost;
    padded controls get unit cost, zero linear term, no effect on the dynamics and bounds
    [-1,1], so they stay at 0, free, and contribute nothing to any output."""

    def __init__(self, n, m, N, M, device):
        self.n, self.m, self.N, self.M = n, m, N, M
        self.active = (N, M) != (n, m)
        self.idx = (torch.cat((torch.arange(n, device=device), N + torch.arange(m, device=device)))
                    if self.active else None)

    def mat_pp(self, C):          # [..., p, p] -> [..., P, P]
        out = C.new_zeros(*C.shape[:-2], self.N + self.M, self.N + self.M)
        out[..., self.idx[:, None], self.idx[None, :]] = C
        if self.M > self.m:
            d = torch.arange(self.N + self.m, self.N + self.M, device=C.device)
            out[..., d, d] = 1.0
        return out

    def vec_p(self, c):           # [..., p] -> [..., P]
        out = c.new_zeros(*c.shape[:-1], self.N + self.M)
        out[..., self.idx] = c
        return out

    def mat_np(self, F):          # [..., n, p] -> [..., N, P]
        out = F.new_zeros(*F.shape[:-2], self.N, self.N + self.M)
        out[..., : self.n, self.idx] = F
        return out

    def vec_n(self, x):
        out = x.new_zeros(*x.shape[:-1], self.N)
        out[..., : self.n] = x
        return out

    def vec_m(self, u, fill=0.0):
        out = u.new_full((*u.shape[:-1], self.M), fill)
        out[..., : self.m] = u
        return out


class _on_device:
    """`torch.cuda.device(dev)` only when `dev` is not already current (the guard costs microseconds)."""

    def __init__(self, dev):
        self.guard = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)


# ----------------------------------------------------------------------------------------------
# raw calls (dense CUDA tensors in, fresh CUDA tensors out)
# ----------------------------------------------------------------------------------------------
def lqr_step_raw(n_state, n_ctrl, T, x_init, C, c, F, f, cur_x, cur_u,
                 u_lower=None, u_upper=None, u_zero_I=None, delta_u=None,
                 linesearch_decay=0.2, max_linesearch_iter=10, do_rollout=True,
                 want_gains=False, want_stats=True, want_du_first=False, dyn=None):
    """Run the step kernel.  Returns a dict of device tensors:
    new_x,new_u,costs,full_du_norm,alphas (do_rollout) and, on request, Ks,ks,qp_iters,
    free_mask,status."""
    if not C.is_cuda:
        raise MpcB200Error("mpc.pytorch_b200 runs on CUDA tensors only (no CPU fallback)")
    dtype, dev = C.dtype, C.device
    if dtype not in (torch.float32, torch.float64):
        raise MpcB200Error(f"unsupported dtype {dtype}")
    n, m = n_state, n_ctrl
    if C.dim() != 4:
        raise MpcB200Error(f"C: expected [T,B,n+m,n+m], got {tuple(C.shape)}")
    B = C.shape[1]
    p = n + m
    _expect("C", C, (T, B, p, p), dev)
    _expect("c", c, (T, B, p), dev)
    if not _is_empty(F):
        if F.dim() != 4 or F.shape[0] not in (T - 1, T):
            raise MpcB200Error(f"F: expected [T-1|T,B,n,n+m], got {tuple(F.shape)}")
        _expect("F", F, (F.shape[0], B, n, p), dev)
    elif T > 1:
        raise MpcB200Error("F is required for T > 1")
    if not _is_empty(f):
        if f.dim() != 3 or f.shape[0] not in (T - 1, T):     # util.get_traj wants f.shape == F.shape[:3]
            raise MpcB200Error(f"f: expected [T-1|T,B,n], got {tuple(f.shape)}")
        _expect("f", f, (f.shape[0], B, n), dev)
    _expect("x_init", x_init, (B, n), dev)
    _expect("current_x", cur_x, (T, B, n), dev)
    _expect("current_u", cur_u, (T, B, m), dev)
    if (u_lower is None) != (u_upper is None):
        raise MpcB200Error("u_lower and u_upper must be given together")
    for nm, bnd in (("u_lower", u_lower), ("u_upper", u_upper)):
        if torch.is_tensor(bnd):
            _expect(nm, bnd, (T, B, m), dev)
    if u_zero_I is not None:
        _expect("u_zero_I", u_zero_I, (T, B, m), dev)
    N, M = _pick_instance(n, m)
    pad = _Pad(n, m, N, M, dev)

    ts = dict(C=0, c=0, F=0, f=0)
    if pad.active:
        C_, c_ = _dense(C, dtype), _dense(c, dtype)
        F_ = _dense(F, dtype) if not _is_empty(F) else None
        f_ = _dense(f, dtype) if not _is_empty(f) else None
    else:       # honour time strides (time-invariant cost / LTI dynamics are read once, not T times)
        C_, ts["C"] = _time_strided(C, dtype)
        c_, ts["c"] = _time_strided(c, dtype)
        F_, ts["F"] = _time_strided(F, dtype) if not _is_empty(F) else (None, 0)
        f_, ts["f"] = _time_strided(f, dtype) if not _is_empty(f) else (None, 0)
    x0_ = _dense(x_init, dtype) if x_init is not None else None
    cx_, cu_ = _dense(cur_x, dtype), _dense(cur_u, dtype)
    F_T = F_.shape[0] if F_ is not None else T - 1
    bounds_kind = 0
    lo_t = hi_t = None
    s_lo = s_hi = 0.0
    if u_lower is not None:
        if isinstance(u_lower, float) and isinstance(u_upper, float) and not pad.active:
            bounds_kind, s_lo, s_hi = 1, u_lower, u_upper
        else:
            bounds_kind = 2
            lo_t = (torch.full((T, B, m), u_lower, dtype=dtype, device=dev)
                    if isinstance(u_lower, float) else _dense(u_lower, dtype))
            hi_t = (torch.full((T, B, m), u_upper, dtype=dtype, device=dev)
                    if isinstance(u_upper, float) else _dense(u_upper, dtype))
    zmask = None
    if u_zero_I is not None:
        zmask = (u_zero_I != 0).to(torch.uint8).contiguous()

    if pad.active:
        C_, c_ = pad.mat_pp(C_), pad.vec_p(c_)
        F_ = pad.mat_np(F_) if F_ is not None else None
        f_ = pad.vec_n(f_) if f_ is not None else None
        x0_ = pad.vec_n(x0_) if x0_ is not None else None
        cx_, cu_ = pad.vec_n(cx_), pad.vec_m(cu_)
        if lo_t is not None:
            lo_t, hi_t = pad.vec_m(lo_t, -1.0), pad.vec_m(hi_t, 1.0)
        if zmask is not None:
            zmask = pad.vec_m(zmask, 0)

    out = {}
    new_x = new_u = costs = fdn = alphas = None
    if do_rollout:
        new_x = torch.empty(T, B, N, dtype=dtype, device=dev)
        new_u = torch.empty(T, B, M, dtype=dtype, device=dev)
        costs = torch.empty(B, dtype=dtype, device=dev)
        fdn = torch.empty(B, dtype=dtype, device=dev)
        alphas = torch.empty(B, dtype=dtype, device=dev)
    du_first = None
    if do_rollout and want_du_first:
        du_first = torch.empty(T, B, M, dtype=dtype, device=dev)
    qp_iters = free_mask = status = None
    if want_stats:
        qp_iters = torch.zeros(T, B, dtype=torch.int32, device=dev) if bounds_kind else None
        free_mask = torch.empty(T, B, M, dtype=torch.uint8, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
    Ks = ks = None
    dims = Dims(B=B, T=T, n=N, m=M, F_T=F_T, has_f=int(f_ is not None), bounds_kind=bounds_kind,
                has_zero_mask=int(zmask is not None), has_delta_u=int(delta_u is not None),
                max_ls_iter=int(max_linesearch_iter), pnqp_max_iter=PNQP_MAX_ITER,
                do_rollout=int(bool(do_rollout)), dynamics_kind=int(dyn[0]) if dyn is not None else 0,
                C_tstride=ts["C"], c_tstride=ts["c"], F_tstride=ts["F"], f_tstride=ts["f"])
    if dyn is not None and pad.active:
        raise MpcB200Error("in-kernel dynamics need an exact (n_state, n_ctrl) kernel instance")
    L = _lib.lib()
    need_gains = want_gains or not do_rollout
    if not need_gains:
        # long horizons do not fit shared memory: the kernel then keeps gains in a caller buffer
        key = (N, M, T, C_.element_size())
        fits = _smem_fits_cache.get(key)
        if fits is None:
            fits = not L.mpcb200_step_prefers_workspace(ctypes.byref(dims), C_.element_size())
            _smem_fits_cache[key] = fits
        need_gains = not fits
    if need_gains:
        Ks = torch.empty(T, B, M, N, dtype=dtype, device=dev)
        ks = torch.empty(T, B, M, dtype=dtype, device=dev)
    params = Params(u_lo=float(s_lo), u_hi=float(s_hi),
                    delta_u=float(delta_u) if delta_u is not None else 0.0,
                    ls_decay=float(linesearch_decay))
    if dyn is not None:
        for i, v in enumerate(dyn[1]):
            params.dyn[i] = float(v)
    fn = L.mpcb200_lqr_step_f32 if dtype == torch.float32 else L.mpcb200_lqr_step_f64
    with _on_device(dev):
        rc = fn(ctypes.byref(dims), ctypes.byref(params), ptr_view(C_), ptr_view(c_), ptr_view(F_), ptr_view(f_),
                ptr(x0_), ptr(cx_), ptr(cu_), ptr(lo_t), ptr(hi_t), ptr(zmask), ptr(new_x), ptr(new_u),
                ptr(costs), ptr(fdn), ptr(alphas), ptr(du_first), ptr(qp_iters), ptr(free_mask), ptr(status),
                ptr(Ks), ptr(ks), stream_handle(dev))
    check(rc, "mpcb200_lqr_step")
    if do_rollout:
        out.update(new_x=new_x[..., :n] if pad.active else new_x,
                   new_u=new_u[..., :m] if pad.active else new_u,
                   costs=costs, full_du_norm=fdn, alphas=alphas)
        if du_first is not None:
            out["du_first"] = du_first[..., :m] if pad.active else du_first
    if Ks is not None:
        out.update(Ks=Ks[..., :m, :n] if pad.active else Ks, ks=ks[..., :m] if pad.active else ks)
    if want_stats:
        out.update(qp_iters=qp_iters, free_mask=free_mask[..., :m] if pad.active else free_mask,
                   status=status)
    return out


def lqr_grad_raw(n_state, n_ctrl, T, C, c, F, new_x, new_u, dx, du, dl_dx, want_df, f_T=None):
    """Run the gradient-assembly kernel; returns (dx_init, dC, dc, dF, df|None)."""
    dtype, dev = C.dtype, C.device
    n, m = n_state, n_ctrl
    B = C.shape[1]
    p = n + m
    _expect("C", C, (T, B, p, p), dev)
    _expect("c", c, (T, B, p), dev)
    if not _is_empty(F):
        _expect("F", F, (F.shape[0], B, n, p), dev)
        if F.shape[0] not in (T - 1, T):
            raise MpcB200Error(f"F: expected T-1 or T time slices, got {F.shape[0]}")
    for nm, t_, sh in (("new_x", new_x, (T, B, n)), ("new_u", new_u, (T, B, m)), ("dx", dx, (T, B, n)),
                       ("du", du, (T, B, m)), ("dl_dx", dl_dx, (T, B, n))):
        _expect(nm, t_, sh, dev)
    N, M = _pick_instance(n, m)
    pad = _Pad(n, m, N, M, dev)
    C_, c_ = _dense(C, dtype), _dense(c, dtype)
    F_ = _dense(F, dtype) if not _is_empty(F) else None
    nx_, nu_ = _dense(new_x, dtype), _dense(new_u, dtype)
    dx_, du_, r_ = _dense(dx, dtype), _dense(du, dtype), _dense(dl_dx, dtype)
    F_T = F_.shape[0] if F_ is not None else 0
    if pad.active:
        C_, c_ = pad.mat_pp(C_), pad.vec_p(c_)
        F_ = pad.mat_np(F_) if F_ is not None else None
        nx_, nu_, dx_, du_, r_ = pad.vec_n(nx_), pad.vec_m(nu_), pad.vec_n(dx_), pad.vec_m(du_), pad.vec_n(r_)
    P = N + M
    dx_init = torch.empty(B, N, dtype=dtype, device=dev)
    dC = torch.empty(T, B, P, P, dtype=dtype, device=dev)
    dc = torch.empty(T, B, P, dtype=dtype, device=dev)
    dF = torch.empty(F_T, B, N, P, dtype=dtype, device=dev) if F_ is not None else None
    # df has f's leading dimension; the kernel writes slices < T-1, a T-th slice (full-length f) is zero
    f_T = T - 1 if f_T is None else f_T
    df = torch.empty(f_T, B, N, dtype=dtype, device=dev) if want_df else None
    if want_df and f_T == T:
        df[T - 1].zero_()
    dims = Dims(B=B, T=T, n=N, m=M, F_T=F_T if F_ is not None else T - 1, has_f=int(want_df),
                bounds_kind=0, has_zero_mask=0, has_delta_u=0, max_ls_iter=1, pnqp_max_iter=1,
                do_rollout=0)
    L = _lib.lib()
    fn = L.mpcb200_lqr_grad_f32 if dtype == torch.float32 else L.mpcb200_lqr_grad_f64
    ws = torch.empty(2 * T * B * N, dtype=dtype, device=dev)     # costates: enables the two-kernel path
    with _on_device(dev):
        rc = fn(ctypes.byref(dims), ptr(C_), ptr(c_), ptr(F_), ptr(nx_), ptr(nu_), ptr(dx_), ptr(du_),
                ptr(r_), ptr(dx_init), ptr(dC), ptr(dc), ptr(dF), ptr(df), ptr(ws), stream_handle(dev))
    check(rc, "mpcb200_lqr_grad")
    if pad.active:
        i = pad.idx
        dx_init = dx_init[:, :n]
        dC = dC[:, :, i[:, None], i[None, :]]
        dc = dc[:, :, i]
        dF = dF[:, :, :n][..., i] if dF is not None else None
        df = df[..., :n] if df is not None else None
    return dx_init, dC, dc, dF, df


_ws_cache = threading.local()


def _workspace(nbytes, dev):
    """Scratch buffer for the one-call adjoint, reused per (thread, device, stream): the library call is
    stream ordered, so consecutive backward passes on one stream can share it."""
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    store = getattr(_ws_cache, "store", None)
    if store is None:
        store = _ws_cache.store = {}
    buf = store.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = store[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    return buf


_adj_plans = {}


def lqr_adjoint_raw(n_state, n_ctrl, T, C, c, F, new_x, new_u, dl_dx, dl_du, u_lower, u_upper, want_df, f_T=None,
                    validated=False):
    """LQRStepFn.backward in ONE library call (prep + nested masked step + costates + outer products);
    returns (dx_init, dC, dc, dF, df|None), or None when this shape needs the general multi-call path
    (zero-padded instance, horizon too long for the shared-memory gain store).  `validated`: the tensors are the
    ones LQRStepFn.forward checked and saved plus autograd's gradients of its outputs (shapes follow), so the
    shape / device checks are not repeated on the backward path."""
    dtype, dev = C.dtype, C.device
    n, m = n_state, n_ctrl
    if _pick_instance(n, m) != (n, m) or _is_empty(F):
        return None
    B = C.shape[1]
    p = n + m
    F_T = F.shape[0]
    if not validated:
        _expect("C", C, (T, B, p, p), dev)
        _expect("c", c, (T, B, p), dev)
        _expect("F", F, (F_T, B, n, p), dev)
        for nm, t_, sh in (("new_x", new_x, (T, B, n)), ("new_u", new_u, (T, B, m)), ("dl_dx", dl_dx, (T, B, n)),
                           ("dl_du", dl_du, (T, B, m))):
            _expect(nm, t_, sh, dev)
    kind, s_lo, s_hi, lo_t, hi_t = 0, 0.0, 0.0, None, None
    if u_lower is not None:
        if isinstance(u_lower, float) and isinstance(u_upper, float):
            kind, s_lo, s_hi = 1, u_lower, u_upper
        else:
            kind = 2
            lo_t = (torch.full((T, B, m), u_lower, dtype=dtype, device=dev) if isinstance(u_lower, float)
                    else _dense(u_lower, dtype))
            hi_t = (torch.full((T, B, m), u_upper, dtype=dtype, device=dev) if isinstance(u_upper, float)
                    else _dense(u_upper, dtype))
            _expect("u_lower", lo_t, (T, B, m), dev)
            _expect("u_upper", hi_t, (T, B, m), dev)
    (C_, tsC), (c_, tsc), (F_, tsF) = _time_strided(C, dtype), _time_strided(c, dtype), _time_strided(F, dtype)
    L = _lib.lib()
    esz = C.element_size()
    # the ctypes structs, the shared-memory fit and the workspace size depend only on this key: build them once
    key = (n, m, T, B, F_T, esz, kind, s_lo, s_hi, bool(want_df), tsC, tsc, tsF)
    plan = _adj_plans.get(key)
    if plan is None:
        dims = Dims(B=B, T=T, n=n, m=m, F_T=F_T, has_f=int(want_df), bounds_kind=kind, has_zero_mask=0,
                    has_delta_u=0, max_ls_iter=10, pnqp_max_iter=PNQP_MAX_ITER, do_rollout=1,
                    C_tstride=tsC, c_tstride=tsc, F_tstride=tsF)
        fits = not L.mpcb200_step_prefers_workspace(ctypes.byref(dims), esz)
        params = Params(u_lo=float(s_lo), u_hi=float(s_hi), delta_u=0.0, ls_decay=0.2)
        nbytes = L.mpcb200_adjoint_workspace_bytes(ctypes.byref(dims), esz) if fits else 0
        if len(_adj_plans) > 256:
            _adj_plans.clear()
        plan = _adj_plans[key] = (fits, dims, params, ctypes.byref(dims), ctypes.byref(params), nbytes)
    fits, dims, params, dims_ref, params_ref, nbytes = plan
    if not fits:
        return None
    nx_, nu_, gx_, gu_ = _dense(new_x, dtype), _dense(new_u, dtype), _dense(dl_dx, dtype), _dense(dl_du, dtype)
    dx_init = torch.empty(B, n, dtype=dtype, device=dev)
    dC = torch.empty(T, B, p, p, dtype=dtype, device=dev)
    dc = torch.empty(T, B, p, dtype=dtype, device=dev)
    dF = torch.empty(F_T, B, n, p, dtype=dtype, device=dev)
    f_T = T - 1 if f_T is None else f_T
    df = torch.empty(f_T, B, n, dtype=dtype, device=dev) if want_df else None
    if want_df and f_T == T:
        df[T - 1].zero_()
    ws = _workspace(nbytes, dev)
    fn = L.mpcb200_lqr_adjoint_f32 if dtype == torch.float32 else L.mpcb200_lqr_adjoint_f64
    with _on_device(dev):
        rc = fn(dims_ref, params_ref, ptr_view(C_), ptr_view(c_), ptr_view(F_), ptr(nx_), ptr(nu_), ptr(gx_),
                ptr(gu_), ptr(lo_t), ptr(hi_t), ptr(dx_init), ptr(dC), ptr(dc), ptr(dF), ptr(df), ptr(ws),
                nbytes, stream_handle(dev))
    check(rc, "mpcb200_lqr_adjoint")
    return dx_init, dC, dc, dF, df


def rollout_raw(n_state, n_ctrl, T, x_init, u, F, f=None):
    """x = get_traj(T, u, x_init, LinDx(F, f)) in ONE kernel (reference mpc/util.py:102-126)."""
    dtype, dev = x_init.dtype, x_init.device
    if not x_init.is_cuda:
        raise MpcB200Error("mpc.pytorch_b200 runs on CUDA tensors only (no CPU fallback)")
    n, m = n_state, n_ctrl
    B = x_init.shape[0]
    _expect("x_init", x_init, (B, n), dev)
    _expect("u", u, (T, B, m), dev)
    if not _is_empty(F):
        if F.dim() != 4 or F.shape[0] not in (T - 1, T):
            raise MpcB200Error(f"F: expected [T-1|T,B,n,n+m], got {tuple(F.shape)}")
        _expect("F", F, (F.shape[0], B, n, n + m), dev)
    elif T > 1:
        raise MpcB200Error("F is required for T > 1")
    if not _is_empty(f):
        if f.shape[0] not in (T - 1, T):
            raise MpcB200Error(f"f: expected [T-1|T,B,n], got {tuple(f.shape)}")
        _expect("f", f, (f.shape[0], B, n), dev)
    N, M = _pick_instance(n, m)
    pad = _Pad(n, m, N, M, dev)
    x0_, u_ = _dense(x_init, dtype), _dense(u, dtype)
    tsF = tsf = 0
    if pad.active:
        F_ = _dense(F, dtype) if not _is_empty(F) else None
        f_ = _dense(f, dtype) if not _is_empty(f) else None
    else:
        F_, tsF = _time_strided(F, dtype) if not _is_empty(F) else (None, 0)
        f_, tsf = _time_strided(f, dtype) if not _is_empty(f) else (None, 0)
    if pad.active:
        x0_, u_ = pad.vec_n(x0_), pad.vec_m(u_)
        F_ = pad.mat_np(F_) if F_ is not None else None
        f_ = pad.vec_n(f_) if f_ is not None else None
    x = torch.empty(T, B, N, dtype=dtype, device=dev)
    dims = Dims(B=B, T=T, n=N, m=M, F_T=F_.shape[0] if F_ is not None else T - 1, has_f=int(f_ is not None),
                bounds_kind=0, has_zero_mask=0, has_delta_u=0, max_ls_iter=1, pnqp_max_iter=1, do_rollout=1,
                F_tstride=tsF, f_tstride=tsf)
    L = _lib.lib()
    fn = L.mpcb200_rollout_f32 if dtype == torch.float32 else L.mpcb200_rollout_f64
    with _on_device(dev):
        rc = fn(ctypes.byref(dims), ptr_view(F_), ptr_view(f_), ptr(x0_), ptr(u_), ptr(x), stream_handle(dev))
    check(rc, "mpcb200_rollout")
    return x[..., :n] if pad.active else x


# ----------------------------------------------------------------------------------------------
# split-mode rollout for nn.Module dynamics / costs (cannot run inside the kernel)
# ----------------------------------------------------------------------------------------------
def _bound_at(v, t):
    return v if isinstance(v, float) else v[t]


def _clamp_assign(x, lo, hi):
    lo = torch.as_tensor(lo, dtype=x.dtype, device=x.device).expand_as(x)
    hi = torch.as_tensor(hi, dtype=x.dtype, device=x.device).expand_as(x)
    x = torch.where(x < lo, lo, x)          # util.eclamp order (reference mpc/util.py:64-68): lower, then upper
    return torch.where(x > hi, hi, x)


def _stage_cost(true_cost, tau, t):
    from .solver import QuadCost
    if isinstance(true_cost, QuadCost):
        Ct, ct = true_cost.C[t], true_cost.c[t]
        return 0.5 * (tau * torch.einsum("bij,bj->bi", Ct, tau)).sum(1) + (tau * ct).sum(1)
    return true_cost(tau)


def _step_dynamics(true_dynamics, x, u, t):
    from .solver import LinDx
    if isinstance(true_dynamics, LinDx):
        nx = torch.einsum("bij,bj->bi", true_dynamics.F[t], torch.cat((x, u), 1))
        if not _is_empty(true_dynamics.f):
            nx = nx + true_dynamics.f[t]
        return nx
    with torch.no_grad():
        return true_dynamics(x, u)


def rollout_split(T, x_init, cur_x, cur_u, Ks, ks, true_cost, true_dynamics,
                  u_lower, u_upper, u_zero_I, delta_u, decay, max_iter):
    """lqr_forward (reference mpc/lqr_step.py:164-261) with gains from the Riccati kernel and an
    arbitrary true model, as batched torch ops on the device."""
    B = x_init.shape[0]
    with torch.no_grad():
        old = 0
        for t in range(T):
            old = old + _stage_cost(true_cost, torch.cat((cur_x[t], cur_u[t]), 1), t)
        alphas = torch.ones(B, dtype=x_init.dtype, device=x_init.device)
        full_du_norm = None
        cost = None
        it = 0
        while it < max_iter and (cost is None or bool((cost > old).any())):
            xs, us = [x_init], []
            cost = 0
            for t in range(T):
                u = torch.einsum("bij,bj->bi", Ks[t], xs[t] - cur_x[t]) + cur_u[t] \
                    + alphas.unsqueeze(1) * ks[t]
                if u_zero_I is not None:
                    u = torch.where(u_zero_I[t].bool(), torch.zeros_like(u), u)
                if u_lower is not None:
                    lo, hi = _bound_at(u_lower, t), _bound_at(u_upper, t)
                    if delta_u is not None:
                        lo = torch.maximum(cur_u[t] - delta_u,
                                           torch.as_tensor(lo, dtype=u.dtype, device=u.device).expand_as(u))
                        hi = torch.minimum(cur_u[t] + delta_u,
                                           torch.as_tensor(hi, dtype=u.dtype, device=u.device).expand_as(u))
                    u = _clamp_assign(u, lo, hi)
                us.append(u)
                if t < T - 1:
                    xs.append(_step_dynamics(true_dynamics, xs[t], u, t))
                cost = cost + _stage_cost(true_cost, torch.cat((xs[t], u), 1), t)
            new_x, new_u = torch.stack(xs), torch.stack(us)
            if full_du_norm is None:
                full_du_norm = (cur_u - new_u).transpose(1, 2).reshape(B, -1).norm(2, 1)
            worse = cost > old
            alphas = torch.where(worse, alphas * decay, alphas)
            it += 1
        alphas = torch.where(cost > old, alphas / decay, alphas)
    return new_x, new_u, cost, full_du_norm, alphas


def reference_full_du_norm(du_first):
    """full_du_norm exactly as the reference forms it (mpc/lqr_step.py:244-245): the [T,B,m]
    difference is transposed to [T,m,B] and VIEWED as [B, T*m] before the row norm, so for
    B > 1 each entry mixes batch elements.  MPC's stop test / printed table depend on it."""
    B = du_first.shape[1]
    return du_first.transpose(1, 2).reshape(B, -1).norm(2, 1)


def _same_storage(a, b):
    if a is None or b is None:
        return _is_empty(a) and _is_empty(b)
    return a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride()


# ----------------------------------------------------------------------------------------------
# the autograd node
# ----------------------------------------------------------------------------------------------
class LQRStepFn(Function):
    """The autograd node behind LQRStep(...).  ONE class for every call (the reference builds a new Function
    class per call, mpc/lqr_step.py:275; the Python class creation alone costs more than the kernels at
    config 3): the closure arguments travel as the first, non-tensor argument `o`."""
    @staticmethod
    def forward(ctx, o, x_init, C, c, F, f=None):
        from .solver import QuadCost, LinDx
        from .dynamics import known_kind
        ctx.o = o
        if o.no_op_forward:                                   # reference :278-282
            # nothing is computed here, but backward hands these tensors to the kernels as raw pointers:
            # check shapes / devices now, at the call site, once
            n, m, T = o.n_state, o.n_ctrl, o.T
            dev, B = C.device, C.shape[1] if C.dim() == 4 else -1
            _expect("C", C, (T, B, n + m, n + m), dev)
            _expect("c", c, (T, B, n + m), dev)
            _expect("x_init", x_init, (B, n), dev)
            if not _is_empty(F):
                if F.dim() != 4 or F.shape[0] not in (T - 1, T):
                    raise MpcB200Error(f"F: expected [T-1|T,B,n,n+m], got {tuple(F.shape)}")
                _expect("F", F, (F.shape[0], B, n, n + m), dev)
            if not _is_empty(f):
                if f.shape[0] not in (T - 1, T):
                    raise MpcB200Error(f"f: expected [T-1|T,B,n], got {tuple(f.shape)}")
                _expect("f", f, (f.shape[0], B, n), dev)
            _expect("current_x", o.current_x, (T, B, n), dev)
            _expect("current_u", o.current_u, (T, B, m), dev)
            ctx.save_for_backward(x_init, C, c, F, f, o.current_x, o.current_u)
            return o.current_x, o.current_u
        assert o.delta_space                                  # reference :284,298
        assert o.current_x is not None and o.current_u is not None
        assert not (o.delta_u is not None and o.u_lower is None)   # reference :195

        quad_same = (isinstance(o.true_cost, QuadCost) and _same_storage(o.true_cost.C, C)
                     and _same_storage(o.true_cost.c, c))
        fused = (quad_same and isinstance(o.true_dynamics, LinDx)
                 and _same_storage(o.true_dynamics.F, F)
                 and (_same_storage(o.true_dynamics.f, f)
                      or (_is_empty(o.true_dynamics.f) and _is_empty(f))))
        dyn = None
        if quad_same and not fused and isinstance(o.true_dynamics, Module):
            kind, kparams = known_kind(o.true_dynamics, o.n_state, o.n_ctrl, C)
            if kind:                       # a known system: its step function runs inside the kernel
                dyn, fused = (kind, kparams), True
        if fused:
            res = lqr_step_raw(o.n_state, o.n_ctrl, o.T, x_init, C, c, F, f, o.current_x, o.current_u,
                               u_lower=o.u_lower, u_upper=o.u_upper, u_zero_I=o.u_zero_I, delta_u=o.delta_u,
                               linesearch_decay=o.linesearch_decay,
                               max_linesearch_iter=o.max_linesearch_iter, do_rollout=True,
                               want_du_first=True, dyn=dyn)
            new_x, new_u = res["new_x"], res["new_u"]
            costs, alphas = res["costs"], res["alphas"]
            fdn = reference_full_du_norm(res["du_first"])
        else:
            assert o.true_cost is not None and o.true_dynamics is not None
            res = lqr_step_raw(o.n_state, o.n_ctrl, o.T, x_init, C, c, F, f, o.current_x, o.current_u,
                               u_lower=o.u_lower, u_upper=o.u_upper, u_zero_I=o.u_zero_I, delta_u=o.delta_u,
                               do_rollout=False)
            new_x, new_u, costs, fdn, alphas = rollout_split(
                o.T, x_init.detach(), o.current_x.detach(), o.current_u.detach(), res["Ks"], res["ks"],
                o.true_cost, o.true_dynamics, o.u_lower, o.u_upper, o.u_zero_I, o.delta_u,
                o.linesearch_decay, o.max_linesearch_iter)
        if o.u_lower is not None and o._defer_host is not None:
            # MPC's loop: no host read per step; the counters stay on the device and MPC reads them
            # together with its own stop-test scalars (one sync per iteration)
            o._defer_host["n_qp"] = (1 + res["qp_iters"].max(dim=1).values).sum()
            o._defer_host["unconverged"] = (res["status"] & 1).any()
            n_qp = float("nan")
        elif o.u_lower is not None:
            # reference: sum_t (1 + i_t) with one batched pnqp per step (:140)
            per_t = 1 + res["qp_iters"].max(dim=1).values
            if o.verbose > 1:                                   # reference :138-139, one line per time step
                for v in reversed(per_t.tolist()):              # the sweep runs t = T-1 .. 0
                    print("  + n_qp_iter: ", v)
            n_qp = float(per_t.sum().item())
            if o.verbose >= 0 and bool((res["status"] & 1).any()):
                print("[WARNING] pnqp warning: Did not converge")   # reference pnqp.py:81
        else:
            n_qp = 0.0
        ctx.save_for_backward(x_init, C, c, F, f, new_x, new_u)
        return new_x, new_u, torch.Tensor([n_qp]), costs, fdn, alphas.mean()

    @staticmethod
    def backward(ctx, dl_dx, dl_du, temp=None, temp2=None, temp3=None, temp4=None):
        o = ctx.o
        x_init, C, c, F, f, new_x, new_u = ctx.saved_tensors
        B = C.size(1)
        if dl_dx is None:
            dl_dx = torch.zeros_like(new_x)
        if dl_du is None:
            dl_du = torch.zeros_like(new_u)
        want_df = not _is_empty(f)
        fast = lqr_adjoint_raw(o.n_state, o.n_ctrl, o.T, C, c, F, new_x, new_u, dl_dx, dl_du, o.u_lower, o.u_upper,
                               want_df, f_T=f.shape[0] if want_df else None, validated=True)
        if fast is not None:                                # the whole backward in one library call
            dx_init, dC, dc, dF, df = fast
            if df is None:
                df = torch.zeros_like(f) if f is not None else None
            return None, dx_init, dC, dc, dF, df
        r = torch.cat((dl_dx, dl_du), 2)                     # reference :316-320
        if o.u_lower is None:
            I = None
        else:                                               # reference :325-326
            I = (torch.abs(new_u - o.u_lower) <= 1e-8) | (torch.abs(new_u - o.u_upper) <= 1e-8)
        zx = torch.zeros(o.T, B, o.n_state, dtype=C.dtype, device=C.device)
        zu = torch.zeros(o.T, B, o.n_ctrl, dtype=C.dtype, device=C.device)
        # nested MPC(lqr_iter=1, u_zero_I=I)(0, QuadCost(C,-r), LinDx(F,None)) (reference :328-340):
        # one masked LQR step from the zero trajectory with the reference's default line search.
        res = lqr_step_raw(o.n_state, o.n_ctrl, o.T, torch.zeros_like(x_init), C, -r, F, None, zx, zu,
                           u_zero_I=I, linesearch_decay=0.2, max_linesearch_iter=10,
                           do_rollout=True, want_stats=False)
        want_df = not _is_empty(f)
        dx_init, dC, dc, dF, df = lqr_grad_raw(o.n_state, o.n_ctrl, o.T, C, c, F, new_x, new_u,
                                               res["new_x"], res["new_u"], dl_dx, want_df,
                                               f_T=f.shape[0] if want_df else None)
        if dF is None:
            dF = torch.zeros_like(F)
        if df is None:                                       # reference :402 (empty tensor)
            df = torch.zeros_like(f) if f is not None else None
        return None, dx_init, dC, dc, dF, df


# ----------------------------------------------------------------------------------------------
# the factory (reference mpc/lqr_step.py:22-38)
# ----------------------------------------------------------------------------------------------
def LQRStep(n_state,
            n_ctrl,
            T,
            u_lower=None,
            u_upper=None,
            u_zero_I=None,
            delta_u=None,
            linesearch_decay=0.2,
            max_linesearch_iter=10,
            true_cost=None,
            true_dynamics=None,
            delta_space=True,
            current_x=None,
            current_u=None,
            verbose=0,
            back_eps=1e-3,
            no_op_forward=False):
    """A single step of the box-constrained iLQR solver (drop-in for the reference factory).

    Returns a callable ``(x_init, C, c, F, f=None)`` giving
    ``(new_x[T,B,n], new_u[T,B,m], n_total_qp_iter (CPU float [1]), costs[B],
    full_du_norm[B], mean_alphas (0-d))`` - or ``(current_x, current_u)`` when
    ``no_op_forward`` - differentiable w.r.t. ``x_init, C, c, F, f``.
    """
    from types import SimpleNamespace
    o = SimpleNamespace(n_state=n_state, n_ctrl=n_ctrl, T=T, u_lower=u_lower, u_upper=u_upper, u_zero_I=u_zero_I,
                        delta_u=delta_u, linesearch_decay=linesearch_decay, max_linesearch_iter=max_linesearch_iter,
                        true_cost=true_cost, true_dynamics=true_dynamics, delta_space=delta_space,
                        current_x=current_x, current_u=current_u, verbose=verbose, back_eps=back_eps,
                        no_op_forward=no_op_forward, _defer_host=getattr(_host_reads, "defer", None))

    def apply(x_init, C, c, F, f=None):
        return LQRStepFn.apply(o, x_init, C, c, F, f)
    return apply

