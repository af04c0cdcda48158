"""Cases of tests/golden/make_reference_fixtures.py (fixtures computed by the reference's own
source through tests/taichi_shim) applied to the oracle and to the CUDA solver."""
import os
import sys

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_reference_fixtures as mk  # noqa: E402

FACE = {"x0": 0, "x1": 1, "y0": 2, "y1": 3, "z0": 4, "z1": 5}
NAMES = sorted(mk.CASES)


def fixture(name):
    return np.load(os.path.join(GOLD, "ref_sp_%s.npz" % name))


def make_oracle(cls, name, **kw):
    shape, _, _, setup, _ = mk.CASES[name]
    o = cls(*shape, **kw)
    o.set_solid(fixture(name)["solid"])
    for fn, arg in setup:
        if fn.startswith("set_bc_rho_"):
            o.set_bc_rho(FACE[fn[-2:]], arg)
        elif fn.startswith("set_bc_vel_"):
            o.set_bc_vel(FACE[fn[-2:]], arg)
        else:
            getattr(o, fn)(arg)
    o.init_simulation()
    return o


def make_solver(name, sparse=False, strict=True, force_field=None, **kw):
    """the CUDA class, driven with the reference's own method names (kw: tau_mode / guo_mode of the
    class's other copy; force_field: the array form of a cal_local_force override, which replaces
    the case's set_force)"""
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    shape, _, _, setup, _ = mk.CASES[name]
    lb = LB3D_Solver_Single_Phase(*shape, sparse_storage=sparse in (True, "aa"), strict=strict,
                                  in_place=sparse in ("aa", "daa"), **kw)
    lb.solid.from_numpy(fixture(name)["solid"])
    for fn, arg in setup:
        if force_field is None or fn != "set_force":
            getattr(lb, fn)(arg)
    if force_field is not None:
        lb.set_force_field(force_field)
    lb.init_simulation()
    return lb


# ---- two-phase script (2phase/lbm_solver_3d_2phase.py through the shim) -------------------------
NAMES2P = sorted(mk.CASES2P)
FIELDS2P = ("F", "rho", "v", "psi", "rho_r", "rho_b")
_SPEC2P = {
    "drainage": dict(flow_bc=(), psi_bc=((0, -1.0),), force=(5e-5, -2e-5, 0.0), niu_l=0.05, niu_g=0.2),
    "pressure_and_psi_faces": dict(flow_bc=[(0, 1, 1.0), (1, 1, 0.995), (5, 2, 1.0)], psi_bc=[(0, -1.0), (3, 1.0)],
                                   force=(0.0, 0.0, 0.0), niu_l=0.1, niu_g=0.1),
    "periodic_bubble": dict(flow_bc=(), psi_bc=(), force=(1e-5, 0.0, 0.0), niu_l=0.05, niu_g=0.2),
    "drainage10": dict(flow_bc=(), psi_bc=((0, -1.0),), force=(5e-5, -2e-5, 0.0), niu_l=0.05, niu_g=0.2),
}


def fixture2p(name):
    return np.load(os.path.join(GOLD, "ref_tp_%s.npz" % name))


def case2p(name):
    """the fixture's inputs as a tests/cases2p.Case2P (the parameter lines the generator replaced)"""
    from tests import cases2p
    g = fixture2p(name)
    return cases2p.Case2P(name, g["solid"], g["psi0"], **_SPEC2P[name])


def check2p(get, g, what, production=False):
    """The reference accumulates rho_r / rho_b with float atomics (order undefined, here the shim's
    loop order) and the shim keeps some f32 locals in f64, so the comparison is to fp32 round-off:
    5e-7 relative on every field, v to 3e-7 of the population scale (it is a difference of
    populations divided by rho).  Production arithmetic (factored transforms, FMA): the
    north_star bar of 1e-5 relative, v to 3e-6 of the population scale."""
    fl = g["solid"] == 0
    rel, vrel = (1e-5, 3e-6) if production else (5e-7, 3e-7)
    if not production and int(g["steps"]) > 10:      # summation-order round-off accumulates with the horizon
        rel, vrel = rel * int(g["steps"]) / 10.0, vrel * int(g["steps"]) / 10.0
    for n in FIELDS2P:
        a, b = np.asarray(get(n))[fl].astype(np.float64), g[n][fl].astype(np.float64)
        tol = vrel * float(np.abs(g["F"][fl]).max()) if n == "v" else rel * float(np.abs(b).max())
        assert float(np.abs(a - b).max()) <= tol, (what, n, float(np.abs(a - b).max()), tol)


# ---- single-phase script copy (Single_phase/lbm_solver_3d.py through the shim) --------------------
NAMES_SCRIPT = sorted(mk.CASES_SCRIPT)


def fixture_script(name):
    return np.load(os.path.join(GOLD, "ref_script_%s.npz" % name))


def _script_setup(name):
    """the parameter lines the generator replaced, as setter calls on a class-style object"""
    _, _, _, lines, _ = mk.CASES_SCRIPT[name]
    calls = [("set_force", [float(t) for t in lines["fx,fy,fz"].split(",")]), ("set_viscosity", float(lines["niu"]))]
    for key, side in (("bc_x_left, rho_bcxl, vx_bcxl, vy_bcxl, vz_bcxl", "x0"),
                      ("bc_x_right, rho_bcxr, vx_bcxr, vy_bcxr, vz_bcxr", "x1")):
        t, rho, vx, vy, vz = [float(q) for q in lines[key].split(",")]
        if int(t) == 1:
            calls.append(("set_bc_rho_" + side, rho))
        elif int(t) == 2:
            calls.append(("set_bc_vel_" + side, [vx, vy, vz]))
    return calls


def make_oracle_script(cls, name, **kw):
    """the script copy's physics: tau = 3 niu + 1/2, un-scaled Guo term, in-place velocity faces"""
    shape = mk.CASES_SCRIPT[name][0]
    o = cls(*shape, tau_mode="textbook", guo_mode="unscaled", vel_bc_mode="script", **kw)
    o.set_solid(fixture_script(name)["solid"])
    for fn, arg in _script_setup(name):
        if fn.startswith("set_bc_rho_"):
            o.set_bc_rho(FACE[fn[-2:]], arg)
        elif fn.startswith("set_bc_vel_"):
            o.set_bc_vel(FACE[fn[-2:]], arg)
        else:
            getattr(o, fn)(arg)
    o.init_simulation()
    return o


def make_solver_script(name, sparse=False, strict=True):
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    shape = mk.CASES_SCRIPT[name][0]
    lb = LB3D_Solver_Single_Phase(*shape, sparse_storage=sparse in (True, "aa"), strict=strict,
                                  in_place=sparse in ("aa", "daa"), tau_mode="textbook", guo_mode="unscaled",
                                  vel_bc_mode="script")
    lb.solid.from_numpy(fixture_script(name)["solid"])
    for fn, arg in _script_setup(name):
        getattr(lb, fn)(arg)
    lb.init_simulation()
    return lb


# ---- grey-scale script (Grey_Scale/lbm_solver_3d_Macro_Sukop.py through the shim) ------------------
NAMES_GREY = sorted(mk.CASES_GREY)
_GREY_PHYSICS = dict(tau_mode="textbook", guo_mode="unscaled", vel_bc_mode="script")


def fixture_grey(name):
    return np.load(os.path.join(GOLD, "ref_grey_%s.npz" % name))


def _grey_setup(name):
    _, _, lines, _ = mk.CASES_GREY[name]
    calls = [("set_force", [float(t) for t in lines["fx,fy,fz"].split(",")]), ("set_viscosity", float(lines["niu"]))]
    for key, side in (("bc_x_left, rho_bcxl, vx_bcxl, vy_bcxl, vz_bcxl", "x0"),
                      ("bc_x_right, rho_bcxr, vx_bcxr, vy_bcxr, vz_bcxr", "x1")):
        if key in lines and int(float(lines[key].split(",")[0])) == 1:
            calls.append(("set_bc_rho_" + side, float(lines[key].split(",")[1])))
    return calls


def make_oracle_grey(cls, name, ns=None, **kw):
    o = cls(*mk.CASES_GREY[name][0], **_GREY_PHYSICS, **kw)
    o.set_grey_scale(fixture_grey(name)["ns"] if ns is None else ns)
    for fn, arg in _grey_setup(name):
        if fn.startswith("set_bc_rho_"):
            o.set_bc_rho(FACE[fn[-2:]], arg)
        else:
            getattr(o, fn)(arg)
    o.init_simulation()
    return o


def make_solver_grey(name, strict=True, ns=None):
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    lb = LB3D_Solver_Single_Phase(*mk.CASES_GREY[name][0], strict=strict, **_GREY_PHYSICS)
    lb.ns.from_numpy(fixture_grey(name)["ns"] if ns is None else ns)
    for fn, arg in _grey_setup(name):
        getattr(lb, fn)(arg)
    lb.init_simulation()
    return lb
