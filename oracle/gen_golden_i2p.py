"""TEST INFRASTRUCTURE ONLY -- golden accelerations for the analytic inverted double pendulum.

Run in the build container:  ``python oracle/gen_golden_i2p.py``  (about a minute of sympy) -> tests/golden/i2p_dynamics.npz

Two derivations, both evaluated with sympy on the same random states:

* ``acc_script``: the reference's own script EXECUTED unmodified,
  ``emei/envs/classic_control/auxiliary/lagrange_eqs.py:12-60`` ``cartpole(2)`` (under an ``IPython.display`` stub),
  its three Lagrange equations solved for the accelerations numerically per state.
* ``acc_physical``: an independent derivation written here with the complete potential energy.  The reference
  script's potential energy of pole i is ``m_i g l_i cos(vertical_i)`` (lagrange_eqs.py:45): for i >= 1 it omits
  the height of the hinge the pole hangs from (``sum_{j<i} 2 l_j cos(vertical_j)``), so for n = 2 its theta_0 equation
  lacks the gravity torque ``2 m_1 g l_0 sin(theta_0)`` of pole 1 on pole 0.  n = 1 (the cart-pole / inverted
  pendulum) is unaffected.  The reference's I2P env gets its accelerations from MuJoCo (complete physics), so the
  engine implements ``acc_physical``; the oracle restates both and is pinned against both.
"""
import os
import sys
import types

import numpy as np
import sympy as sp
from sympy.physics.mechanics import dynamicsymbols

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "i2p_dynamics.npz")
REF = "/root/reference/emei/envs/classic_control/auxiliary/lagrange_eqs.py"


def reference_script_equations():
    ip = types.ModuleType("IPython")
    ipd = types.ModuleType("IPython.display")
    ipd.display, ipd.Latex = print, str
    ip.display = ipd
    sys.modules.setdefault("IPython", ip)
    sys.modules.setdefault("IPython.display", ipd)
    import importlib.util

    spec = importlib.util.spec_from_file_location("lagrange_eqs", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.cartpole(2)  # (equations == 0, (x, theta_0, theta_1))


def physical_equations():
    t = sp.symbols("t")
    g, F, M = sp.symbols("g F M", real=True)
    m0, m1, l0, l1 = sp.symbols("m0 m1 l0 l1", real=True)
    x, th0, th1 = dynamicsymbols("x"), dynamicsymbols(r"\theta_0"), dynamicsymbols(r"\theta_1")
    v0, v1 = th0, th0 + th1  # angles from the vertical
    p0 = (x + l0 * sp.sin(v0), l0 * sp.cos(v0))
    p1 = (x + 2 * l0 * sp.sin(v0) + l1 * sp.sin(v1), 2 * l0 * sp.cos(v0) + l1 * sp.cos(v1))
    T = sp.Rational(1, 2) * M * sp.diff(x, t) ** 2
    V = -F * x
    for m, l, p, v in ((m0, l0, p0, v0), (m1, l1, p1, v1)):
        T += sp.Rational(1, 2) * m * (sp.diff(p[0], t) ** 2 + sp.diff(p[1], t) ** 2) + sp.Rational(1, 2) * (sp.Rational(1, 3) * m * l**2) * sp.diff(v, t) ** 2
        V += m * g * p[1]
    L = T - V
    eqs = [sp.diff(sp.diff(L, sp.diff(q, t)), t) - sp.diff(L, q) for q in (x, th0, th1)]
    return eqs, (x, th0, th1)


def numeric_solver(eqs, syms):
    """equations (== 0), linear in the second derivatives -> f(params, q, qd) -> accelerations"""
    t = sp.symbols("t")
    acc = [sp.diff(s, t, 2) for s in syms]
    a_sym = sp.symbols("a0:3")
    vel = [sp.diff(s, t) for s in syms]
    v_sym, q_sym = sp.symbols("v0:3"), sp.symbols("q0:3")
    sub = {}
    for a, s in zip(acc, a_sym):
        sub[a] = s
    eqs = [e.subs(sub) for e in eqs]
    eqs = [e.subs({v: s for v, s in zip(vel, v_sym)}) for e in eqs]
    eqs = [e.subs({q: s for q, s in zip(syms, q_sym)}) for e in eqs]
    A, b = sp.linear_eq_to_matrix(eqs, list(a_sym))
    names = sp.symbols("g F M m0 m1 l0 l1", real=True)
    fA = sp.lambdify([names, q_sym, v_sym], A, "numpy")
    fb = sp.lambdify([names, q_sym, v_sym], b, "numpy")

    def solve(params, q, qd):
        out = np.empty_like(q)
        for i in range(q.shape[0]):
            p = list(params[:1]) + [params[1][i]] + list(params[2:])
            out[i] = np.linalg.solve(np.array(fA(p, q[i], qd[i]), dtype=np.float64), np.array(fb(p, q[i], qd[i]), dtype=np.float64).reshape(3))
        return out

    return solve


def main():
    sys.path.insert(0, ROOT)
    from oracle import emei_oracle as O

    p = O.I2PParams()
    rng = np.random.default_rng(2002)
    n = 512
    q = rng.uniform(-1, 1, size=(n, 3)) * np.array([2.5, np.pi, np.pi])
    qd = rng.uniform(-1, 1, size=(n, 3)) * np.array([4.0, 8.0, 10.0])
    force = rng.uniform(-1, 1, size=n) * p.gear
    params = (p.gravity, force, p.mass_cart, p.mass_pole0, p.mass_pole1, p.length0, p.length1)
    acc_script = numeric_solver(*reference_script_equations())(params, q, qd)
    acc_physical = numeric_solver(*physical_equations())(params, q, qd)
    np.savez(OUT, q=q, qd=qd, force=force, acc_script=acc_script, acc_physical=acc_physical,
             params=np.array([p.gravity, p.mass_cart, p.mass_pole0, p.mass_pole1, p.length0, p.length1]))
    d = np.abs(acc_script - acc_physical).max(axis=0)
    print("wrote", OUT, "max |script - physical| per coordinate:", d)


if __name__ == "__main__":
    main()
