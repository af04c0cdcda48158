"""Fixtures computed by THE REFERENCE'S OWN SOURCE, run here through tests/taichi_shim.

    python tests/golden/make_reference_fixtures.py          # needs /root/reference

Taichi is not installable in this image, so /root/reference/Single_phase/
LBM_3D_SinglePhase_Solver.py is imported unmodified with `taichi` and `pyevtk` resolved to the
pure-Python stand-ins under tests/taichi_shim (what they implement, and what they cannot know
about real Taichi -- reassociation under fast_math -- is stated in their header).  Each case
drives the class exactly like the reference's example scripts (ctor, solid.from_numpy, set_bc_*,
set_force, set_viscosity, init_simulation, step) on a lattice small enough for interpreted
loops, and stores inputs and outputs in tests/golden/ref_sp_<case>.npz.  tests/test_reference_pin.py
holds the oracle (and, on a GPU, the CUDA path) to these files.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/Single_phase"
sys.path.insert(0, ROOT)

# name -> (shape, solid fraction, seed, setup calls, steps)
CASES = {
    "pressure_x": ((6, 5, 4), 0.25, 2, [("set_bc_rho_x0", 1.0), ("set_bc_rho_x1", 0.99)], 4),
    "lid_and_force": ((5, 5, 5), 0.2, 5, [("set_bc_vel_z1", [0.05, 0.0, 0.0]), ("set_force", [1e-4, -5e-5, 2e-5]),
                                         ("set_viscosity", 0.1)], 3),
    "all_faces": ((5, 4, 6), 0.15, 9, [("set_bc_vel_x0", [0.02, 0.0, 0.0]), ("set_bc_rho_x1", 0.98),
                                      ("set_bc_rho_y0", 1.0), ("set_bc_vel_y1", [0.0, -0.01, 0.01]),
                                      ("set_bc_rho_z0", 1.02), ("set_bc_vel_z1", [0.0, 0.0, 0.03])], 3),
    "periodic_force": ((4, 6, 5), 0.3, 11, [("set_force", [1e-5, 2e-5, -1e-5])], 4),
}


def load_reference():
    """the reference module with taichi / pyevtk resolved to the shim"""
    shim = os.path.join(ROOT, "tests", "taichi_shim")
    sys.path.insert(0, shim)
    sys.path.insert(0, REF)
    try:
        for m in ("taichi", "pyevtk", "pyevtk.hl", "LBM_3D_SinglePhase_Solver"):
            sys.modules.pop(m, None)
        mod = importlib.import_module("LBM_3D_SinglePhase_Solver")
        assert mod.__file__.startswith(REF), mod.__file__
        return mod
    finally:
        sys.path.remove(REF)
        sys.path.remove(shim)


def case_solid(name):
    shape, frac, seed, _, _ = CASES[name]
    return (np.random.default_rng(seed).random(shape) < frac).astype(np.int8)


def run_reference(mod, name):
    shape, _, _, setup, steps = CASES[name]
    solid = case_solid(name)
    lb = mod.LB3D_Solver_Single_Phase(nx=shape[0], ny=shape[1], nz=shape[2])
    lb.solid.from_numpy(solid)
    for fn, arg in setup:
        getattr(lb, fn)(arg)
    lb.init_simulation()
    out = {"solid": solid, "steps": steps}
    for it in range(steps):
        lb.step()
        if it == 0:
            out.update(F1=lb.F.to_numpy(), rho1=lb.rho.to_numpy(), v1=lb.v.to_numpy())
    out.update(F=lb.F.to_numpy(), rho=lb.rho.to_numpy(), v=lb.v.to_numpy(), f=lb.f.to_numpy(),
               max_v=np.float32(lb.get_max_v()), S=np.asarray(lb.S_dig[None]), inv_M=np.asarray(lb.inv_M[None]))
    return out


def main():
    mod = load_reference()
    for name in CASES:
        out = run_reference(mod, name)
        np.savez_compressed(os.path.join(HERE, "ref_sp_%s.npz" % name), **out)
        print(name, "steps", out["steps"], "max_v", float(out["max_v"]), "F dtype", out["F"].dtype)


if __name__ == "__main__":
    main()
