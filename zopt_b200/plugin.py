"""
User-defined dynamics for the iLQR / DDP solvers and trajectoryRollout.

The reference's solvers take an arbitrary Python callable `dynamics(x, u)` and differentiate it with JAX
(zopt/ilqrUtils.py:260-268, zopt/pytrees.py:138-194).  A CUDA kernel cannot trace a lambda, so a model outside the
registered ones (zopt_b200.models) is defined SYMBOLICALLY:

    import sympy as sp
    from zopt_b200.plugin import SymbolicDynamics
    pend = SymbolicDynamics(lambda x, u: [x[0] + 0.1 * x[1], x[1] + 0.1 * (u[0] - sp.sin(x[0]))], n=2, m=1)
    traj, L, J, converged = iterativeLqr(pend, QuadraticCost(Q, R), QuadraticTerminalCost(Qf), x0, uGuess)

`f` is called once with lists of sympy symbols and returns the n expressions of x+.  sympy differentiates them (Jacobians
f_x, f_u and the costate-contracted Hessian sum_i lam_i d2f_i/dz2, z = [x; u], with common subexpressions shared), the
result is printed as CUDA into a header, and `csrc/zb_user_model.cu` -- the library's generic solver kernels compiled
around that header -- is built with nvcc for sm_100a into its own shared library (cached by content hash under
`zopt_b200/_plugin_cache/`, in-tree so that a library built on a machine without a GPU travels with the repo).
The object stays callable like the lambda it replaces (`dyn(x, u)` on torch tensors, differentiable with torch.func: that
is what the test oracle does), and there is no CPU fallback: solving needs the compiled plug-in and a CUDA device.
Limits of the generic kernels: n <= 16, m <= 8, one thread per problem (16 per problem in the line search).
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import torch

from ._lib import ZbCost, check as _check_main, dcode, pick_device, ptr, stream_ptr, to_dev

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
CACHE_DIR = os.path.join(_HERE, "_plugin_cache")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
MAX_N, MAX_M = 16, 8


def _cuda_printer():
    from sympy.printing.c import C99CodePrinter

    class Printer(C99CodePrinter):
        """C with every literal typed T(...) (the kernels are templates over float / double) and small integer powers
        written as products."""

        def _print_Float(self, e):
            return "T(%s)" % repr(float(e))

        def _print_Integer(self, e):
            return "T(%d)" % int(e)

        def _print_Rational(self, e):
            return "(T(%d)/T(%d))" % (int(e.p), int(e.q))

        def _print_Pow(self, e):
            b, ex = e.base, e.exp
            if ex.is_Integer and 2 <= int(ex) <= 4:
                return "(" + "*".join(["(%s)" % self._print(b)] * int(ex)) + ")"
            if ex.is_Integer and -4 <= int(ex) <= -1:
                return "(T(1)/(" + "*".join(["(%s)" % self._print(b)] * (-int(ex))) + "))"
            if ex == 0.5 or (ex.is_Rational and ex.p == 1 and ex.q == 2):
                return "sqrt(%s)" % self._print(b)
            return "pow((T)(%s), (T)(%s))" % (self._print(b), self._print(ex))

    return Printer()


def _emit(name, args, outputs, sizes):
    """One templated ZB_HD function `name(args...)` writing the flat output arrays in `outputs` ({array name: [exprs]})."""
    import sympy as sp
    pr = _cuda_printer()
    flat = [e for k in outputs for e in outputs[k]]
    repl, red = sp.cse(flat, symbols=sp.numbered_symbols("t_"), optimizations="basic")
    lines = ["template <typename T>", "ZB_HD void %s(%s) {" % (name, args)]
    for s, e in repl:
        lines.append("    const T %s = %s;" % (s, pr.doprint(e)))
    i = 0
    for k in outputs:
        for j in range(len(outputs[k])):
            lines.append("    %s[%d] = %s;" % (k, j, pr.doprint(red[i])))
            i += 1
    lines.append("}")
    return "\n".join(lines)


def _emit_scalar(name, args, expr):
    """templated ZB_HD function returning one scalar expression"""
    import sympy as sp
    pr = _cuda_printer()
    repl, red = sp.cse([expr], symbols=sp.numbered_symbols("t_"), optimizations="basic")
    lines = ["template <typename T>", "ZB_HD T %s(%s) {" % (name, args)]
    for sym, e in repl:
        lines.append("    const T %s = %s;" % (sym, pr.doprint(e)))
    lines.append("    return %s;" % pr.doprint(red[0]))
    lines.append("}")
    return "\n".join(lines)


def _torch_modules():
    tm = {k: getattr(torch, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "tanh", "sinh", "cosh", "atan", "asin", "acos")}
    tm["Abs"] = torch.abs
    return tm


class SymbolicCost:
    """
    User-defined running and terminal costs `c(x, u)`, `cf(x)` given as sympy expressions -- the counterpart, for costs, of
    SymbolicDynamics: the reference takes arbitrary callables `runningCost(x, u)`, `terminalCost(x)` and differentiates
    them with JAX (zopt/ilqrUtils.py:260-268, pytrees.py:71-81, 99-115).  `SymbolicCost(c, cf, n, m)` calls `c` and `cf`
    once with sympy symbols, derives gradient and Hessian, and generates CUDA for `c`, `(c_x, c_u)`, the stacked Hessian
    `[[c_xx, c_ux'], [c_ux, c_uu]]`, `cf`, `cf_x`, `cf_xx`.  Pass `cost.running` and `cost.terminal` where the reference
    takes the two callables; both stay callable on torch tensors (that is how the test oracle differentiates them).
    Used together with a SymbolicDynamics model: the pair is compiled into one solver plug-in; the eigen-clamp of the cost
    Hessian (ilqrUtils.py:222-234) then runs per time step inside the backward pass, as in the reference.
    """

    def __init__(self, c, cf, n, m):
        import sympy as sp
        n, m = int(n), int(m)
        self.n, self.m = n, m
        xs, us = list(sp.symbols(f"x0:{n}", real=True)), list(sp.symbols(f"u0:{m}", real=True))
        ce, cfe = sp.sympify(c(xs, us)), sp.sympify(cf(xs))
        for e, allowed, what in ((ce, set(xs) | set(us), "c(x, u)"), (cfe, set(xs), "cf(x)")):
            free = e.free_symbols - allowed
            if free:
                raise ValueError(f"{what} may only depend on its arguments; free symbols: {sorted(map(str, free))}")
        self._xs, self._us, self._c, self._cf = xs, us, ce, cfe
        z = xs + us
        sub = {xs[i]: sp.Symbol(f"x[{i}]") for i in range(n)}
        sub.update({us[i]: sp.Symbol(f"u[{i}]") for i in range(m)})
        S = lambda e: sp.sympify(e).xreplace(sub)
        p = n + m
        Hc, Hf = sp.hessian(ce, z), sp.hessian(cfe, xs)
        XU = "const T* __restrict__ x, const T* __restrict__ u"
        self.source = "\n\n".join([
            "// user-defined cost (zopt_b200/plugin.py::SymbolicCost)", "#define ZB_USER_COST 1",
            _emit_scalar("user_cost", XU, S(ce)),
            _emit("user_cost_grad", XU + ", T* __restrict__ cx, T* __restrict__ cu",
                  {"cx": [S(sp.diff(ce, v)) for v in xs], "cu": [S(sp.diff(ce, v)) for v in us]}, None),
            _emit("user_cost_hess", XU + ", T* __restrict__ Z", {"Z": [S(Hc[i, j]) for i in range(p) for j in range(p)]}, None),
            _emit_scalar("user_tcost", "const T* __restrict__ x", S(cfe)),
            _emit("user_tcost_grad", "const T* __restrict__ x, T* __restrict__ vx", {"vx": [S(sp.diff(cfe, v)) for v in xs]}, None),
            _emit("user_tcost_hess", "const T* __restrict__ x, T* __restrict__ Vxx", {"Vxx": [S(Hf[i, j]) for i in range(n) for j in range(n)]}, None),
        ]) + "\n"
        self.running, self.terminal = _SymbolicRunningCost(self), _SymbolicTerminalCost(self)


class _SymbolicRunningCost:
    is_symbolic_cost = True

    def __init__(self, owner):
        self.owner, self._fn = owner, None

    def __call__(self, x, u):
        import sympy as sp
        if self._fn is None:
            self._fn = sp.lambdify(self.owner._xs + self.owner._us, self.owner._c, modules=[_torch_modules()])
        x, u = torch.as_tensor(x), torch.as_tensor(u)
        return x[..., 0] * 0 + self._fn(*x.unbind(-1), *u.unbind(-1))


class _SymbolicTerminalCost:
    is_symbolic_cost = True

    def __init__(self, owner):
        self.owner, self._fn = owner, None

    def __call__(self, x):
        import sympy as sp
        if self._fn is None:
            self._fn = sp.lambdify(self.owner._xs, self.owner._cf, modules=[_torch_modules()])
        x = torch.as_tensor(x)
        return x[..., 0] * 0 + self._fn(*x.unbind(-1))


class SymbolicDynamics:
    """x+ = f(x, u) given as sympy expressions; compiled into a solver plug-in (see the module docstring)."""
    is_plugin = True

    def __init__(self, f, n, m, build=True):
        import sympy as sp
        n, m = int(n), int(m)
        if not (1 <= n <= MAX_N and 1 <= m <= MAX_M):
            raise ValueError(f"user models are limited to n <= {MAX_N}, m <= {MAX_M} (got {n}, {m})")
        self.n, self.m = n, m
        xs, us = list(sp.symbols(f"x0:{n}", real=True)), list(sp.symbols(f"u0:{m}", real=True))
        exprs = [sp.sympify(e) for e in f(xs, us)]
        if len(exprs) != n:
            raise ValueError(f"f must return {n} expressions (got {len(exprs)})")
        free = set().union(*[e.free_symbols for e in exprs]) - set(xs) - set(us)
        if free:
            raise ValueError(f"f may only depend on its arguments; free symbols: {sorted(map(str, free))}")
        self._xs, self._us, self._exprs = xs, us, exprs
        z = xs + us
        lam = list(sp.symbols(f"lam0:{n}", real=True))
        F = sp.Matrix(exprs)
        Jx, Ju = F.jacobian(xs), F.jacobian(us)
        contracted = sum((lam[i] * exprs[i] for i in range(n)), sp.Integer(0))
        H = sp.hessian(contracted, z)
        # symbols -> array accesses
        sub = {xs[i]: sp.Symbol(f"x[{i}]") for i in range(n)}
        sub.update({us[i]: sp.Symbol(f"u[{i}]") for i in range(m)})
        sub.update({lam[i]: sp.Symbol(f"lam[{i}]") for i in range(n)})
        S = lambda e: sp.sympify(e).xreplace(sub)
        p = n + m
        parts = ["// GENERATED by zopt_b200/plugin.py -- do not edit.  x+ = f(x, u), n = %d, m = %d" % (n, m), "#pragma once",
                 "#define ZB_USER_N %d" % n, "#define ZB_USER_M %d" % m,
                 _emit("user_step", "const T* __restrict__ x, const T* __restrict__ u, T* __restrict__ xn", {"xn": [S(e) for e in exprs]}, None),
                 _emit("user_lin", "const T* __restrict__ x, const T* __restrict__ u, T* __restrict__ fx, T* __restrict__ fu",
                       {"fx": [S(Jx[i, j]) for i in range(n) for j in range(n)], "fu": [S(Ju[i, j]) for i in range(n) for j in range(m)]}, None),
                 _emit("user_hess", "const T* __restrict__ x, const T* __restrict__ u, const T* __restrict__ lam, T* __restrict__ H",
                       {"H": [S(H[i, j]) for i in range(p) for j in range(p)]}, None)]
        self.source = "\n\n".join(parts) + "\n"
        self._torch_fn = None
        self._lib = None
        if build:
            self.build()

    # ------------------------------------------------------------------------------------------ callable like the lambda
    def __call__(self, x, u):
        """Evaluate x+ = f(x, u) on torch tensors (leading batch axes allowed); differentiable with torch autograd."""
        import sympy as sp
        if self._torch_fn is None:
            tm = {k: getattr(torch, k) for k in ("sin", "cos", "tan", "exp", "log", "sqrt", "tanh", "sinh", "cosh", "atan", "asin", "acos")}
            tm["Abs"] = torch.abs
            self._torch_fn = sp.lambdify(self._xs + self._us, self._exprs, modules=[tm])
        x, u = torch.as_tensor(x), torch.as_tensor(u)
        out = self._torch_fn(*x.unbind(-1), *u.unbind(-1))
        zero = x[..., 0] * 0
        return torch.stack([zero + o for o in out], dim=-1)

    def batch(self):
        return 1

    def with_cost(self, cost):
        """The same model compiled together with a SymbolicCost into one solver plug-in (cached on the model)."""
        if (cost.n, cost.m) != (self.n, self.m):
            raise ValueError(f"cost dimensions ({cost.n},{cost.m}) do not match the model ({self.n},{self.m})")
        cache = self.__dict__.setdefault("_with_cost", {})
        if id(cost) not in cache:
            import copy
            bound = copy.copy(self)
            bound.__dict__.pop("_with_cost", None)
            bound.source = self.source + "\n" + cost.source
            bound._lib, bound.cost = None, cost
            bound.build()
            cache[id(cost)] = bound
        return cache[id(cost)]

    # ------------------------------------------------------------------------------------------ build / load
    _NVCC_FLAGS = ["-cudart", "static", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                   "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]

    def _key(self):
        """Content hash of everything the plug-in is compiled from: the generated model, EVERY header of csrc/ (the solver
        kernels include them transitively), the public headers and the compiler command line."""
        h = hashlib.sha1(self.source.encode())
        files = [os.path.join(_CSRC, "zb_user_model.cu")] + sorted(
            os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh"))
        inc = os.path.join(os.path.dirname(os.path.dirname(_CSRC)), "include")
        files += sorted(os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".h"))
        for fn in files:
            with open(fn, "rb") as fh:
                h.update(os.path.basename(fn).encode())
                h.update(fh.read())
        h.update(" ".join([NVCC] + self._NVCC_FLAGS + ["ZB_PD_MAX=n+m"]).encode())
        return h.hexdigest()[:16]

    def build(self):
        """Generate the model header and compile the plug-in for sm_100a (nvcc cross-compiles without a GPU); cached.
        Safe when several ranks build the same model at once: each compiles to its own temporary name and publishes the
        library with an atomic rename (the losers of the race replace it with identical bytes)."""
        os.makedirs(CACHE_DIR, exist_ok=True)
        key = self._key()
        hdr, so = os.path.join(CACHE_DIR, f"model_{key}.cuh"), os.path.join(CACHE_DIR, f"libzb_model_{key}.so")
        if not os.path.exists(so):
            import tempfile
            fd, tmp_hdr = tempfile.mkstemp(prefix=f"model_{key}.", suffix=".cuh", dir=CACHE_DIR)
            with os.fdopen(fd, "w") as fh:
                fh.write(self.source)
            fd, tmp_so = tempfile.mkstemp(prefix=f"libzb_model_{key}.", suffix=".so.tmp", dir=CACHE_DIR)
            os.close(fd)
            try:
                # the eigen-clamp scratch is sized for this model's stacked Hessian (n + m), not for the library's largest (24)
                cmd = [NVCC] + self._NVCC_FLAGS + [f"-DZB_PD_MAX={self.n + self.m}", f'-DZB_USER_MODEL_HEADER="{tmp_hdr}"', "-I", _CSRC, "-shared", "-o", tmp_so,
                                                   os.path.join(_CSRC, "zb_user_model.cu")]
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for the user model:\n" + r.stderr[-4000:])
                os.replace(tmp_hdr, hdr)
                os.replace(tmp_so, so)
            finally:
                for t in (tmp_hdr, tmp_so):
                    if os.path.exists(t):
                        os.unlink(t)
        self.so_path = so
        L = C.CDLL(so)
        L.zb_user_ilqr_workspace_bytes.restype = C.c_size_t
        L.zb_user_ilqr_workspace_bytes.argtypes = [C.c_int32, C.c_int64, C.c_int32]
        for name in ("zb_user_dims", "zb_user_last_error", "zb_user_step", "zb_user_hess", "zb_user_rollout", "zb_user_ilqr_solve"):
            getattr(L, name).restype = C.c_int32
        L.zb_user_step.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int64] + [C.c_void_p] * 5
        L.zb_user_hess.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int64] + [C.c_void_p] * 4
        L.zb_user_rollout.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(ZbCost)] + [C.c_void_p] * 5 + \
            [C.c_double] + [C.c_void_p] * 3
        L.zb_user_ilqr_solve.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.POINTER(ZbCost),
                                         C.c_void_p, C.c_void_p, C.c_int32, C.c_double] + [C.c_void_p] * 9 + [C.c_size_t]
        nn, mm = C.c_int32(0), C.c_int32(0)
        L.zb_user_dims(C.byref(nn), C.byref(mm))
        assert (nn.value, mm.value) == (self.n, self.m)
        self._lib = L
        return self

    def _check(self, rc):
        if rc == 0:
            return
        buf = C.create_string_buffer(512)
        self._lib.zb_user_last_error(buf, 512)
        msg = buf.value.decode(errors="replace")
        if rc < 0:
            raise ValueError(msg)
        raise RuntimeError(f"CUDA error {rc}: {msg}")

    def _need(self):
        if self._lib is None:
            raise RuntimeError("the model's plug-in library is not built (SymbolicDynamics(..., build=False)); call .build()")
        if not torch.cuda.is_available():
            raise RuntimeError("zopt_b200 needs a CUDA device; there is no CPU fallback")

    # ------------------------------------------------------------------------------------------ device entry points
    def step(self, x, u, linearize=False):
        """x+ (and optionally f_x, f_u) at a batch of points, evaluated by the generated CUDA code."""
        self._need()
        device = pick_device(x, u)
        dtype = x.dtype if isinstance(x, torch.Tensor) and x.dtype in (torch.float32, torch.float64) else torch.float64
        x, u = to_dev(x, dtype, device), to_dev(u, dtype, device)
        batched = x.ndim == 2
        x, u = (x if batched else x[None]).contiguous(), (u if batched else u[None]).contiguous()
        Bsz = x.shape[0]
        xn = torch.empty_like(x)
        fx = torch.empty((Bsz, self.n, self.n), dtype=dtype, device=device) if linearize else None
        fu = torch.empty((Bsz, self.n, self.m), dtype=dtype, device=device) if linearize else None
        self._check(self._lib.zb_user_step(dcode(dtype), device.index, stream_ptr(device), Bsz, ptr(x), ptr(u), ptr(xn), ptr(fx), ptr(fu)))
        if not batched:
            xn, fx, fu = xn[0], (fx[0] if linearize else None), (fu[0] if linearize else None)
        return (xn, fx, fu) if linearize else xn

    def second_order(self, x, u):
        """f_xx (P,n,n,n), f_ux (P,n,m,n), f_uu (P,n,m,m) at P points (QuadraticDynamics, pytrees.py:179-194): the generated
        contracted Hessian evaluated with lam = e_i."""
        self._need()
        n, m, p = self.n, self.m, self.n + self.m
        P = x.shape[0]
        eye = torch.eye(n, dtype=x.dtype, device=x.device)
        Z = torch.empty((n, P, p, p), dtype=x.dtype, device=x.device)
        for i in range(n):
            lam = eye[i].expand(P, n).contiguous()
            self._check(self._lib.zb_user_hess(dcode(x.dtype), x.device.index, stream_ptr(x.device), P, ptr(x), ptr(u), ptr(lam), ptr(Z[i])))
        Z = Z.permute(1, 0, 2, 3)
        return Z[:, :, :n, :n].contiguous(), Z[:, :, n:, :n].contiguous(), Z[:, :, n:, n:].contiguous()

    def solve(self, cspec, dtype, device, Bsz, N, x0, uGuess, maxIter, tol, second_order, return_log):
        """iterativeLqr / differentialDynamicProgramming for this model (called by zopt_b200.ilqrUtils._solve)."""
        self._need()
        n, m = self.n, self.m
        xTraj = torch.empty((Bsz, N + 1, n), dtype=dtype, device=device)
        uTraj = torch.empty((Bsz, N, m), dtype=dtype, device=device)
        L = torch.empty((Bsz, N, m, n), dtype=dtype, device=device)
        J = torch.empty((Bsz,), dtype=dtype, device=device)
        conv = torch.empty((Bsz,), dtype=torch.uint8, device=device)
        iters = torch.empty((Bsz,), dtype=torch.int32, device=device)
        alog = torch.empty((Bsz, max(maxIter, 1)), dtype=torch.int32, device=device) if return_log else None
        Jlog = torch.empty((Bsz, maxIter + 1), dtype=dtype, device=device) if return_log else None
        wsb = self._lib.zb_user_ilqr_workspace_bytes(dcode(dtype), Bsz, N)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=device)
        self._check(self._lib.zb_user_ilqr_solve(dcode(dtype), device.index, stream_ptr(device), Bsz, N, 1 if second_order else 0,
                                                 C.byref(cspec), ptr(x0), ptr(uGuess), maxIter, float(tol), ptr(xTraj), ptr(uTraj),
                                                 ptr(L), ptr(J), ptr(conv), ptr(iters), ptr(alog), ptr(Jlog), ptr(ws), wsb))
        return xTraj, uTraj, L, J, conv, iters, alog, Jlog
