#!/usr/bin/env python
"""
Offline code generator: analytic quadcopter dynamics, Jacobian and Hessian
contraction as straight-line CUDA device code.

The model is the reference's `Quadcopter.inertialDynamics`
(zopt/quadcopter.py:116-144 -> rigidBodyDynamics :70-113, rotation matrices
:23-48, aero :51-67), written here symbolically with sympy from the equations of
motion (NOT traced from the reference code).  Emitted functions, all for the
continuous-time right-hand side xdot = F(x,u,wind_ned):

  quad_xdot(tr, x, u, w, xd)            xdot (12)
  quad_jac_x(tr, x, u, w, J)            dF/dx as 12x12 row-major (zeros written too)
  quad_hess_contract(tr, x, u, w, lam, H)   sum_i lam_i d2F_i/dx2, 12x12 symmetric
                                            (only the leading 9x9 block is non-zero)

dF/du is constant: [2,0] = -1, [3,1] = [4,2] = [5,3] = +1 ; d2F/dudx = d2F/du2 = 0.
`tr` carries sin/cos of (phi,theta,psi), tan(theta), 1/cos(theta).
The caller forms the forward-Euler map x + dt*xdot (demos/iterativeLqr.py:35).

Run:  python scripts/gen_quad_model.py > zopt_b200/csrc/quad_model_gen.cuh
The output is committed; tests check it against torch autodiff of the oracle.
"""
import sympy as sp
from sympy.printing.c import C99CodePrinter

x = sp.symbols('x0:12', real=True)
u = sp.symbols('u0:4', real=True)
w = sp.symbols('w0:3', real=True)
lam = sp.symbols('lam0:12', real=True)
uu, vv, ww, p, q, r, phi, th, psi = x[:9]
g, mass = sp.Float('9.807'), sp.Float('2.5')

sph, cph, sth, cth, sps, cps = (sp.sin(phi), sp.cos(phi), sp.sin(th), sp.cos(th), sp.sin(psi), sp.cos(psi))
R = sp.Matrix([
    [cth * cps, sph * sth * cps - cph * sps, cph * sth * cps - sph * sps],
    [cth * sps, sph * sth * sps + cph * cps, cph * sth * sps - sph * cps],
    [-sth, sph * cth, cph * cth],
])
tth = sth / cth
E = sp.Matrix([[1, sph * tth, cph * tth], [0, cph, -sph], [0, sph / cth, cph / cth]])

uvw = sp.Matrix([uu, vv, ww])
pqr = sp.Matrix([p, q, r])
wind_body = R.T * sp.Matrix(w)
ua = uvw - wind_body
flin = [sp.Rational(-2, 10), sp.Rational(-2, 10), sp.Rational(-3, 10)]
fquad = [sp.Rational(-5, 100), sp.Rational(-5, 100), sp.Rational(-1, 10)]
mlin = [sp.Rational(-1, 10), sp.Rational(-1, 10), sp.Rational(-5, 100)]
Fa = sp.Matrix([flin[i] * ua[i] + fquad[i] * ua[i]**2 for i in range(3)])
Fc = mass * sp.Matrix([0, 0, -u[0]])
Fg = mass * g * sp.Matrix([-sth, sph * cth, cph * cth])
uvwDot = (1 / mass) * (-pqr.cross(uvw) + Fa + Fc + Fg)
pqrDot = sp.Matrix([u[1 + i] + mlin[i] * pqr[i] for i in range(3)])  # I = eye(3): pqr x (I pqr) == 0
eulDot = E * pqr
xyzDot = R * uvw
F = sp.Matrix(list(uvwDot) + list(pqrDot) + list(eulDot) + list(xyzDot))

# trig symbols
S = sp.symbols('tr_sph tr_cph tr_sth tr_cth tr_sps tr_cps tr_tth tr_sec', real=True)
Ssph, Scph, Ssth, Scth, Ssps, Scps, Stth, Ssec = S


def trigsub(e):
    e = sp.expand_trig(e)
    e = e.subs({sp.tan(th): Stth})
    e = e.subs({sp.sin(phi): Ssph, sp.cos(phi): Scph, sp.sin(th): Ssth, sp.cos(th): Scth, sp.sin(psi): Ssps,
                sp.cos(psi): Scps})
    # negative powers of cos(theta) -> powers of sec; sin(theta)*sec -> tan(theta)
    e = e.replace(lambda t: t.is_Pow and t.base == Scth and t.exp.is_negative, lambda t: Ssec**(-t.exp))
    return e


def emit(name, args_sig, outs, doc):
    """outs: list of (lhs_string, expr)"""
    exprs = [trigsub(e) for _, e in outs]
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols('t'), optimizations='basic')
    lines = [f"// {doc}", "template <typename T>", f"ZB_HD void {name}({args_sig}) {{"]
    for s, e in repl:
        lines.append(f"    const T {s} = {ccode(e)};")
    for (lhs, _), e in zip(outs, red):
        lines.append(f"    {lhs} = {ccode(e)};")
    lines.append("}")
    return "\n".join(lines)


class _P(C99CodePrinter):
    def _print_Float(self, e):
        return f"T({C99CodePrinter._print_Float(self, e)})"

    def _print_Integer(self, e):
        return f"T({int(e)})"

    def _print_Rational(self, e):
        return f"T({float(e)!r})"

    def _print_Pow(self, e):
        b, ex = e.as_base_exp()
        if ex.is_Integer and 2 <= int(ex) <= 4:
            return "(" + "*".join([self._print(b) if b.is_Symbol else f"({self._print(b)})"] * int(ex)) + ")"
        if ex == -1:
            return f"(T(1)/({self._print(b)}))"
        return super()._print_Pow(e)


def ccode(e):
    s = _P().doprint(e)
    for i in range(12):
        s = s.replace(f"lam{11 - i}", f"lam[{11 - i}]")
    import re
    s = re.sub(r'\bx(\d+)\b', r'x[\1]', s)
    s = re.sub(r'\bu(\d+)\b', r'u[\1]', s)
    s = re.sub(r'\bw(\d+)\b', r'w[\1]', s)
    s = re.sub(r'\btr_(\w+)\b', r'tr.\1', s)
    return s


xs = sp.Matrix(x)
J = F.jacobian(xs)
h = sum(lam[i] * F[i] for i in range(12))
H = sp.hessian(h, xs)

# sanity on structure: f_u constant 4 nnz, Hessian only touches states 0..8
Ju = F.jacobian(sp.Matrix(u))
assert [(i, j, Ju[i, j]) for i in range(12) for j in range(4) if Ju[i, j] != 0] == \
    [(2, 0, -1.0), (3, 1, 1), (4, 2, 1), (5, 3, 1)], Ju
assert all(H[i, j] == 0 for i in range(12) for j in range(9, 12))

out = []
out.append("""// GENERATED by scripts/gen_quad_model.py -- do not edit by hand.
// Quadcopter 12-state NED model (reference: zopt/quadcopter.py:23-144), continuous-time
// right-hand side, its state Jacobian and the costate-contracted state Hessian.
#pragma once
#include "zb_math.cuh"  // ZB_HD

template <typename T>
struct QuadTrig {
    T sph, cph, sth, cth, sps, cps, tth, sec;
};
""")
for suffix, wsub, wsig in (("", {wi: 0 for wi in w}, ""), ("_wind", {}, ", const T* __restrict__ w")):
    sig = "const QuadTrig<T>& tr, const T* __restrict__ x, const T* __restrict__ u" + wsig
    note = " (wind_ned = 0)" if not suffix else " (general wind_ned)"
    out.append(emit("quad_xdot" + suffix, sig + ", T* __restrict__ xd",
                    [(f"xd[{i}]", F[i].subs(wsub)) for i in range(12)], "xdot = F(x,u,wind_ned)" + note))
    out.append("")
    jouts = [(f"J[{i * 12 + j}]", J[i, j].subs(wsub)) for i in range(12) for j in range(12)]
    out.append(emit("quad_jac_x" + suffix, sig + ", T* __restrict__ J", jouts,
                    "J[i*12+j] = dF_i/dx_j (row-major 12x12)" + note))
    out.append("")
    houts = [(f"H[{i * 9 + j}]", H[i, j].subs(wsub)) for i in range(9) for j in range(i + 1)]
    out.append(emit("quad_hess_contract" + suffix, sig + ", const T* __restrict__ lam, T* __restrict__ H", houts,
                    "H[i*9+j] = sum_k lam[k] d2F_k/dx_i dx_j for j <= i < 9 (LOWER triangle of the symmetric 9x9; "
                    "all other entries of the 12x12 are 0)" + note))
    out.append("")
print("\n".join(out))
