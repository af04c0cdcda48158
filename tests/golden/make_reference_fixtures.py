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
    # example_cavity.py in small (the reference's geo_cavity generator: walls on x0, y0, y1, z0, z1; lid on x1)
    "cavity8": ((8, 8, 8), 0.0, 0, [("set_bc_vel_x1", [0.0, 0.0, 0.1])], 25),
    # one whole 3^3 block solid: never activated in the reference's sparse mode (reads give 0)
    "solid_block": ((6, 6, 6), 0.15, 13, [("set_bc_rho_x0", 1.0), ("set_bc_rho_x1", 0.995),
                                         ("set_force", [0.0, 1e-5, 0.0])], 2),
    # example_porous_medium.py in small, at a horizon where round-off has had time to grow: 16^3
    # porous block, pressure faces in x plus a transverse body force, 100 steps (20 minutes of
    # interpreted loops for the dense and the sparse run; the fixture is committed)
    "porous16": ((16, 16, 16), 0.2, 17, [("set_bc_rho_x0", 1.0), ("set_bc_rho_x1", 0.99),
                                         ("set_force", [0.0, 1e-5, 0.0])], 100),
    "porous12": ((12, 12, 12), 0.2, 19, [("set_bc_rho_x0", 1.0), ("set_bc_rho_x1", 0.99),
                                         ("set_force", [0.0, 1e-5, 0.0])], 60),
}


REF_OTHER = "/root/reference/Phase_change"      # the class's other copy: tau = 3 niu + 1/2, un-scaled Guo term


def load_reference(directory=REF):
    """the reference module (from `directory`) with taichi / pyevtk resolved to the shim"""
    shim = os.path.join(ROOT, "tests", "taichi_shim")
    sys.path.insert(0, shim)
    sys.path.insert(0, directory)
    try:
        for m in ("taichi", "pyevtk", "pyevtk.hl", "LBM_3D_SinglePhase_Solver"):
            sys.modules.pop(m, None)
        mod = importlib.import_module("LBM_3D_SinglePhase_Solver")
        assert mod.__file__.startswith(directory), mod.__file__
        return mod
    finally:
        sys.path.remove(directory)
        sys.path.remove(shim)


def local_force_field(shape):
    """a per-node force for the cal_local_force hook (:217-220)"""
    rng = np.random.default_rng(29)
    return (1e-5 * rng.standard_normal(shape + (3,))).astype(np.float32)


def run_reference_local_force(mod, name):
    """the reference's extension point as its users exercise it (Phase_change/
    LBM_3D_SinglePhase_Solute_Solver.py:185-190): a SUBCLASS overriding the @ti.func
    cal_local_force(i,j,k); here it returns a per-node vector.  fx is set non-zero only to switch
    the class's force_flag on (:137-140), the hook ignores it."""
    import taichi as ti          # the shim (load_reference left it in sys.modules)
    shape, _, _, setup, steps = CASES[name]
    ff = local_force_field(shape)

    class WithLocalForce(mod.LB3D_Solver_Single_Phase):
        @ti.func
        def cal_local_force(self, i, j, k):
            return ti.Vector([ff[i, j, k, 0], ff[i, j, k, 1], ff[i, j, k, 2]])

    solid = case_solid(name)
    lb = WithLocalForce(nx=shape[0], ny=shape[1], nz=shape[2])
    lb.solid.from_numpy(solid)
    for fn, arg in setup:
        if fn != "set_force":
            getattr(lb, fn)(arg)
    lb.set_force([1e-30, 0.0, 0.0])
    lb.init_simulation()
    for _ in range(steps):
        lb.step()
    return {"solid": solid, "steps": steps, "force_field": ff, "F": lb.F.to_numpy(), "rho": lb.rho.to_numpy(),
            "v": lb.v.to_numpy()}


def case_solid(name):
    shape, frac, seed, _, _ = CASES[name]
    solid = (np.random.default_rng(seed).random(shape) < frac).astype(np.int8)
    if name == "solid_block":
        solid[3:6, 3:6, 3:6] = 1
    if name == "cavity8":
        from taichi_lbm3d_b200.geometry import cavity
        solid = cavity(*shape)
    return solid


def run_reference(mod, name, sparse_storage=False):
    shape, _, _, setup, steps = CASES[name]
    solid = case_solid(name)
    lb = mod.LB3D_Solver_Single_Phase(nx=shape[0], ny=shape[1], nz=shape[2], sparse_storage=sparse_storage)
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
    if sparse_storage:
        # pointer/dense SNode fields have the extent 3*(n//3+1) (:43-44); the reference itself
        # slices them back to the lattice (export_VTK :469-474)
        nx, ny, nz = shape
        for k in ("F", "rho", "v", "f", "F1", "rho1", "v1"):
            out[k] = out[k][:nx, :ny, :nz]
        out["active_blocks"] = lb.rho.active.copy()
    return out


# ---- two-phase script ---------------------------------------------------------------------------
# 2phase/lbm_solver_3d_2phase.py is a flat script: module-level parameters that its users edit by
# hand, kernels, then a driver that reads the (missing) 131^3 input files and loops 80001 times.
# Here the kernel part (everything above the driver) is executed through the shim with ONLY the
# hand-edited parameter lines replaced -- lattice extents and, per case, fluid / BC parameters --
# and the driver's loop body (:626-632) is called from here.
REF2P = "/root/reference/2phase/lbm_solver_3d_2phase.py"
REF2P_SPARSE = "/root/reference/2phase/lbm_solver_3d_2phase_sparse.py"     # same kernels, pointer SNode fields

CASES2P = {
    # name -> (shape, solid fraction, seed, parameter-line overrides, steps)
    # (the dense script's default force; the sparse script's own default differs: fy = 0)
    "drainage": ((6, 5, 4), 0.25, 3, {"niu_l": "0.05", "niu_g": "0.2", "fx,fy,fz": "5.0e-5,-2e-5,0.0"}, 4),
    "pressure_and_psi_faces": ((5, 5, 5), 0.2, 8, {
        "fx,fy,fz": "0.0,0.0,0.0",
        "bc_x_left, rho_bcxl, vx_bcxl, vy_bcxl, vz_bcxl": "1, 1.0, 0.0e-5, 0.0, 0.0",
        "bc_x_right, rho_bcxr, vx_bcxr, vy_bcxr, vz_bcxr": "1, 0.995, 0.0, 0.0, 0.0",
        "bc_z_right, rho_bczr, vx_bczr, vy_bczr, vz_bczr": "2, 1.0, 0.0, 0.0, 0.0",
        "bc_psi_y_right, psi_y_right": "1, 1.0"}, 3),
    "periodic_bubble": ((6, 6, 6), 0.0, 1, {"bc_psi_x_left, psi_x_left": "0, -1.0", "fx,fy,fz": "1e-5,0.0,0.0",
                                            "niu_l": "0.05", "niu_g": "0.2"}, 3),
    # config 4 in small at a longer horizon: 10^3 porous block, psi = -1 slab entering from x0, 50 steps
    "drainage10": ((10, 10, 10), 0.2, 4, {"niu_l": "0.05", "niu_g": "0.2", "fx,fy,fz": "5.0e-5,-2e-5,0.0"}, 50),
}


def case2p_inputs(name):
    shape, frac, seed, _, _ = CASES2P[name]
    solid = (np.random.default_rng(seed).random(shape) < frac).astype(np.int8)
    if name == "periodic_bubble":
        x, y, z = np.meshgrid(*[np.arange(n) for n in shape], indexing='ij')
        psi = np.where((x - 2.5) ** 2 + (y - 2.5) ** 2 + (z - 2.5) ** 2 < 4.5, -1.0, 1.0).astype(np.float32)
    else:
        psi = np.ones(shape, np.float32)
        psi[:2] = -1.0
    return solid, psi


def run_reference_two_phase(name, script=REF2P):
    import re
    shape, _, _, overrides, steps = CASES2P[name]
    src = open(script).read()
    head = src[:src.index("time_init = time.time()")]
    lines = dict(overrides)
    lines["nx,ny,nz"] = "%d,%d,%d" % shape
    for lhs, rhs in lines.items():
        head, n = re.subn(r"^%s\s*=.*$" % re.escape(lhs), "%s = %s" % (lhs, rhs), head, count=1, flags=re.M)
        assert n == 1, lhs
    shim = os.path.join(ROOT, "tests", "taichi_shim")
    sys.path.insert(0, shim)
    try:
        for m in ("taichi", "pyevtk", "pyevtk.hl"):
            sys.modules.pop(m, None)
        ns = {"__name__": "lbm_solver_3d_2phase"}
        exec(compile(head, script, "exec"), ns)
    finally:
        sys.path.remove(shim)
    # Taichi embeds the Python-scope floats a kernel reads (wl, wg, lg0, l1, ..., psi_solid, CapA) as f32
    # constants; the script computes them in Python floats first (:100-108), so: one rounding, here
    for k, val in list(ns.items()):
        if isinstance(val, float):
            ns[k] = np.float32(val)
    solid, psi = case2p_inputs(name)
    ns["solid"].from_numpy(solid)
    ns["psi"].from_numpy(psi)
    ns["static_init"]()
    ns["init"]()
    for _ in range(steps):                  # the driver's loop body, :626-632
        ns["colission"]()
        ns["streaming1"]()
        ns["Boundary_condition"]()
        ns["streaming3"]()
        ns["Boundary_condition_psi"]()
    out = {"solid": solid, "psi0": psi, "steps": steps}
    nx, ny, nz = shape
    for n in ("F", "rho", "v", "psi", "rho_r", "rho_b"):
        out[n] = ns[n].to_numpy()[:nx, :ny, :nz]        # SNode-placed fields have the extent 3*(n//3+1)
    for n in ("fx", "fy", "fz", "niu_l", "niu_g", "psi_solid", "CapA"):
        out[n] = np.float32(ns[n])
    return out


# ---- single-phase script copy --------------------------------------------------------------------
# Single_phase/lbm_solver_3d.py: the flat-script copy of the single-phase solver (x faces only).  It
# differs from the class in three places: tau = 3 niu + 1/2 (:47), the Guo term without the /3, /9
# (:175-180) and the fixed-velocity face, F[s] = feq(LR[s],1,u) - F[LR[s]] + feq(s,1,u) in place
# (:253, :268).  Executed like the two-phase script: kernel part through the shim with only the
# hand-edited parameter lines replaced, driver loop body (:304-308) called from here.
REFSP_SCRIPT = "/root/reference/Single_phase/lbm_solver_3d.py"
CASES_SCRIPT = {
    # name -> (shape, solid fraction, seed, parameter-line overrides, steps)
    "vel_x0_rho_x1": ((6, 5, 4), 0.2, 6, {
        "fx,fy,fz": "1.0e-5,0.0,-2.0e-6", "niu": "0.1",
        "bc_x_left, rho_bcxl, vx_bcxl, vy_bcxl, vz_bcxl": "2, 1.0, 0.02, 0.0, 0.005",
        "bc_x_right, rho_bcxr, vx_bcxr, vy_bcxr, vz_bcxr": "1, 0.99, 0.0, 0.0, 0.0"}, 5),
    "vel_both": ((5, 4, 5), 0.15, 12, {
        "fx,fy,fz": "0.0e-6,0.0,0.0", "niu": "0.16",
        "bc_x_left, rho_bcxl, vx_bcxl, vy_bcxl, vz_bcxl": "2, 1.0, 0.03, 0.0, 0.0",
        "bc_x_right, rho_bcxr, vx_bcxr, vy_bcxr, vz_bcxr": "2, 1.0, 0.03, 0.01, 0.0"}, 4),
}


def case_script_solid(name):
    shape, frac, seed, _, _ = CASES_SCRIPT[name]
    return (np.random.default_rng(seed).random(shape) < frac).astype(np.int8)


def run_reference_sp_script(name):
    import re
    shape, _, _, overrides, steps = CASES_SCRIPT[name]
    src = open(REFSP_SCRIPT).read()
    head = src[:src.index("time_init = time.time()")]
    lines = dict(overrides)
    lines["nx,ny,nz"] = "%d,%d,%d" % shape
    for lhs, rhs in lines.items():
        head, n = re.subn(r"^%s\s*=.*$" % re.escape(lhs), "%s = %s" % (lhs, rhs), head, count=1, flags=re.M)
        assert n == 1, lhs
    shim = os.path.join(ROOT, "tests", "taichi_shim")
    sys.path.insert(0, shim)
    try:
        for m in ("taichi", "pyevtk", "pyevtk.hl"):
            sys.modules.pop(m, None)
        ns = {"__name__": "lbm_solver_3d"}
        exec(compile(head, REFSP_SCRIPT, "exec"), ns)
    finally:
        sys.path.remove(shim)
    for k, val in list(ns.items()):      # Python-scope floats read by kernels are embedded as f32
        if isinstance(val, float):
            ns[k] = np.float32(val)
    solid = case_script_solid(name)
    ns["solid"].from_numpy(solid.astype(np.int32))
    ns["static_init"]()
    ns["init"]()
    for _ in range(steps):                  # the driver's loop body, :304-308
        ns["colission"]()
        ns["streaming1"]()
        ns["Boundary_condition"]()
        ns["streaming3"]()
    out = {"solid": solid, "steps": steps}
    for n in ("F", "rho", "v"):
        out[n] = ns[n].to_numpy()
    out["S"] = np.asarray(ns["S_dig"].to_numpy())
    return out


# ---- grey-scale script ---------------------------------------------------------------------------
# Grey_Scale/lbm_solver_3d_Macro_Sukop.py: the script copy's physics (tau = 3 niu + 1/2, un-scaled Guo
# term, in-place velocity faces) plus a per-node solid fraction ns: streaming0 (:233-239) blends every
# post-collision population with the opposite one of the node it is about to move to,
# f2[i,s] = f[i,s] + ns[i] (f[i+e_s, LR[s]] - f[i,s]), and streaming1 (:242-247) pushes f2 to ALL
# neighbours -- no bounce-back; a node with ns = 1 is solid (solid = int(ns), :339) and is never
# collided, so it keeps f = w for ever, and the links that leave it are never written.  Executed like
# the other scripts: kernel part through the shim with the parameter lines replaced, ns assigned
# directly instead of np.loadtxt('./BC.dat'), loop body :345-350 called from here.
REFGS_SCRIPT = "/root/reference/Grey_Scale/lbm_solver_3d_Macro_Sukop.py"
CASES_GREY = {
    # name -> (shape, seed, parameter-line overrides, steps)
    "periodic_force": ((6, 5, 4), 21, {"fx,fy,fz": "1.0e-5,-4.0e-6,2.0e-6", "niu": "0.1"}, 6),
    "pressure_x": ((7, 5, 5), 23, {
        "fx,fy,fz": "2.0e-6,0.0,0.0", "niu": "0.16",
        "bc_x_left, rho_bcxl, vx_bcxl, vy_bcxl, vz_bcxl": "1, 1.0, 0.0, 0.0, 0.0",
        "bc_x_right, rho_bcxr, vx_bcxr, vy_bcxr, vz_bcxr": "1, 0.99, 0.0, 0.0, 0.0"}, 6),
}


def case_grey_ns(name):
    """solid fraction per node: a quarter of the nodes fully solid (1.0), a quarter open (0.0), the
    rest grey with fractions in (0, 1)"""
    shape, seed = CASES_GREY[name][0], CASES_GREY[name][1]
    rng = np.random.default_rng(seed)
    kind = rng.random(shape)
    frac = rng.random(shape)
    return np.where(kind < 0.25, 1.0, np.where(kind < 0.5, 0.0, frac)).astype(np.float32)


def run_reference_grey_script(name):
    import re
    shape, _, overrides, steps = CASES_GREY[name]
    src = open(REFGS_SCRIPT).read()
    head = src[:src.index("time_init = time.time()")]
    lines = dict(overrides)
    lines["nx,ny,nz"] = "%d,%d,%d" % shape
    for lhs, rhs in lines.items():
        head, n = re.subn(r"^%s\s*=.*$" % re.escape(lhs), "%s = %s" % (lhs, rhs), head, count=1, flags=re.M)
        assert n == 1, lhs
    shim = os.path.join(ROOT, "tests", "taichi_shim")
    sys.path.insert(0, shim)
    try:
        for m in ("taichi", "pyevtk", "pyevtk.hl", "evtk", "evtk.hl"):
            sys.modules.pop(m, None)
        g = {"__name__": "lbm_solver_3d_Macro_Sukop"}
        exec(compile(head, REFGS_SCRIPT, "exec"), g)
    finally:
        sys.path.remove(shim)
    for k, val in list(g.items()):       # Python-scope floats read by kernels are embedded as f32
        if isinstance(val, float):
            g[k] = np.float32(val)
    ns_np = case_grey_ns(name)
    g["solid"].from_numpy(ns_np.astype(int))       # :339
    g["ns"].from_numpy(ns_np)                      # :343
    g["static_init"]()
    g["init"]()
    for _ in range(steps):                         # :345-350
        g["colission"]()
        g["streaming0"]()
        g["streaming1"]()
        g["Boundary_condition"]()
        g["streaming3"]()
    out = {"ns": ns_np, "solid": ns_np.astype(int).astype(np.int8), "steps": steps}
    for n in ("F", "rho", "v"):
        out[n] = g[n].to_numpy()
    out["S"] = np.asarray(g["S_dig"].to_numpy())
    return out


def write_grey(name):
    out = run_reference_grey_script(name)
    np.savez_compressed(os.path.join(HERE, "ref_grey_%s.npz" % name), **out)
    print("grey-scale script", name, "steps", out["steps"], "max |v|", float(np.abs(out["v"]).max()),
          "rho range", float(out["rho"][out["solid"] == 0].min()), float(out["rho"][out["solid"] == 0].max()))


def write_script(name):
    out = run_reference_sp_script(name)
    np.savez_compressed(os.path.join(HERE, "ref_script_%s.npz" % name), **out)
    print("single-phase script", name, "steps", out["steps"], "max |v|", float(np.abs(out["v"]).max()))


def main():
    only = [a for a in sys.argv[1:] if not a.startswith("--")]        # case names: just those
    if only:
        mod = load_reference()
        for name in only:
            if name in CASES:
                write_single(mod, name)
            elif name in CASES_SCRIPT:
                write_script(name)
            elif name.startswith("grey_") and name[5:] in CASES_GREY:
                write_grey(name[5:])
            else:
                write_two_phase(name)
        return
    if "--two-phase-only" not in sys.argv:
        main_single()
    for name in CASES2P:
        write_two_phase(name)
    for name in CASES_SCRIPT:
        write_script(name)
    for name in CASES_GREY:
        write_grey(name)


def write_two_phase(name):
    if True:
        out = run_reference_two_phase(name)
        sp = run_reference_two_phase(name, REF2P_SPARSE)
        for n in ("F", "rho", "v", "psi", "rho_r", "rho_b"):
            out[n + "_sparse"] = sp[n]
        np.savez_compressed(os.path.join(HERE, "ref_tp_%s.npz" % name), **out)
        print("two-phase", name, "steps", out["steps"], "psi range", float(out["psi"].min()), float(out["psi"].max()))


def main_single():
    other = load_reference(REF_OTHER)
    for name in ("lid_and_force", "periodic_force"):
        out = run_reference(other, name)
        np.savez_compressed(os.path.join(HERE, "ref_sp_other_copy_%s.npz" % name), **out)
        print("other copy", name, "max_v", float(out["max_v"]))
    mod = load_reference()
    for name in ("pressure_x", "periodic_force"):
        out = run_reference_local_force(mod, name)
        np.savez_compressed(os.path.join(HERE, "ref_sp_local_force_%s.npz" % name), **out)
        print("cal_local_force override", name, "max |v|", float(np.abs(out["v"]).max()))
    for name in CASES:
        write_single(mod, name)


def write_single(mod, name):
    if True:
        out = run_reference(mod, name)
        # the same case with sparse_storage=True (pointer SNode tree of 3^3 blocks, :36-44)
        sp = run_reference(mod, name, sparse_storage=True)
        for k in ("F", "rho", "v"):
            out[k + "_sparse"] = sp[k]
        out["active_blocks"] = sp["active_blocks"]
        np.savez_compressed(os.path.join(HERE, "ref_sp_%s.npz" % name), **out)
        print(name, "steps", out["steps"], "max_v", float(out["max_v"]), "F dtype", out["F"].dtype,
              "active blocks", int(sp["active_blocks"].sum()), "of", sp["active_blocks"].size)


if __name__ == "__main__":
    main()
