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


def make_solver(name, sparse=False, strict=True):
    """the CUDA class, driven with the reference's own method names"""
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    shape, _, _, setup, _ = mk.CASES[name]
    lb = LB3D_Solver_Single_Phase(*shape, sparse_storage=sparse in (True, "aa"), strict=strict,
                                  in_place=sparse in ("aa", "daa"))
    lb.solid.from_numpy(fixture(name)["solid"])
    for fn, arg in setup:
        getattr(lb, fn)(arg)
    lb.init_simulation()
    return lb
