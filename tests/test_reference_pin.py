"""The oracle against THE REFERENCE'S OWN SOURCE (CPU, no GPU).

tests/golden/ref_sp_*.npz were computed by /root/reference/Single_phase/
LBM_3D_SinglePhase_Solver.py, imported unmodified and executed through the pure-Python Taichi
stand-in tests/taichi_shim (tests/golden/make_reference_fixtures.py).  Both oracle forms must
reproduce them BIT FOR BIT: F, rho, v after the first step and after the last, the relaxation
rates, the fp32 inverse matrix and get_max_v.  Where /root/reference is mounted, the reference
is run again and must reproduce the committed files."""
import os

import numpy as np
import pytest

from oracle.cref import RefSinglePhaseC
from oracle.ref_single_phase import RefSinglePhase
from tests import refpin


@pytest.mark.parametrize("name", refpin.NAMES)
@pytest.mark.parametrize("cls", [RefSinglePhase, RefSinglePhaseC])
def test_oracle_reproduces_reference_source(name, cls):
    g = refpin.fixture(name)
    o = refpin.make_oracle(cls, name)
    assert np.array_equal(o.S, g["S"])                                   # init_simulation :126-131
    assert np.array_equal(np.asarray(o.inv_M, np.float32), g["inv_M"])   # :83, :110
    fl = g["solid"] == 0
    o.step()
    for n in ("F", "rho", "v"):
        assert np.array_equal(getattr(o, n)[fl], g[n + "1"][fl]), (n, "after step 1")
    for _ in range(int(g["steps"]) - 1):
        o.step()
    for n in ("F", "rho", "v", "f"):
        assert np.array_equal(getattr(o, n)[fl], g[n][fl]), n
    # solid nodes: rho = 1, v = 0 (:390-392), F keeps its initial w (:164-169)
    assert np.array_equal(o.rho[~fl], g["rho"][~fl]) and np.array_equal(o.v[~fl], g["v"][~fl])
    assert np.array_equal(o.F[~fl], g["F"][~fl])
    vmax = np.sqrt((o.v.astype(np.float32) ** 2).sum(-1, dtype=np.float32)).max()
    assert abs(float(vmax) - float(g["max_v"])) <= 1e-9 + 2e-7 * float(g["max_v"])


@pytest.mark.parametrize("name", ["lid_and_force", "periodic_force"])
@pytest.mark.parametrize("cls", [RefSinglePhase, RefSinglePhaseC])
def test_other_copy_of_the_class(name, cls):
    """Phase_change/LBM_3D_SinglePhase_Solver.py (tau = 3 niu + 1/2 :126, Guo term not divided :235),
    run unmodified through the shim, is what tau_mode="textbook", guo_mode="unscaled" compute"""
    g = np.load(os.path.join(refpin.GOLD, "ref_sp_other_copy_%s.npz" % name))
    o = refpin.make_oracle(cls, name, tau_mode="textbook", guo_mode="unscaled")
    assert np.array_equal(o.S, g["S"])
    for _ in range(int(g["steps"])):
        o.step()
    fl = g["solid"] == 0
    for n in ("F", "rho", "v"):
        assert np.array_equal(getattr(o, n)[fl], g[n][fl]), n


@pytest.mark.parametrize("name", ["pressure_x", "periodic_force"])
@pytest.mark.parametrize("cls", [RefSinglePhase, RefSinglePhaseC])
def test_cal_local_force_override_is_the_force_array(name, cls):
    """the reference's extension point, a subclass overriding the @ti.func cal_local_force(i,j,k)
    (:217-220; used by colission :231 and streaming3 :385), returning a per-node vector: bit for
    bit what set_force_field computes (lbm_set_force_field in the C ABI)"""
    g = np.load(os.path.join(refpin.GOLD, "ref_sp_local_force_%s.npz" % name))
    shape, _, _, setup, _ = refpin.mk.CASES[name]
    o = cls(*shape)
    o.set_solid(g["solid"])
    for fn, arg in setup:
        if fn.startswith("set_bc_rho_"):
            o.set_bc_rho(refpin.FACE[fn[-2:]], arg)
        elif fn.startswith("set_bc_vel_"):
            o.set_bc_vel(refpin.FACE[fn[-2:]], arg)
        elif fn != "set_force":
            getattr(o, fn)(arg)
    o.set_force_field(g["force_field"])
    o.init_simulation()
    for _ in range(int(g["steps"])):
        o.step()
    fl = g["solid"] == 0
    for n in ("F", "rho", "v"):
        assert np.array_equal(getattr(o, n)[fl], g[n][fl]), n


@pytest.mark.parametrize("name", refpin.NAMES_SCRIPT)
@pytest.mark.parametrize("cls", [RefSinglePhase, RefSinglePhaseC])
def test_script_copy_of_the_single_phase_solver(name, cls):
    """Single_phase/lbm_solver_3d.py, the flat-script copy (tau = 3 niu + 1/2 :47, Guo term not
    divided :175-180, fixed-velocity faces F[s] = feq(LR[s],1,u) - F[LR[s]] + feq(s,1,u) in place
    :253,:268), its kernels run through the shim with only the parameter lines replaced: bit for bit
    what tau_mode="textbook", guo_mode="unscaled", vel_bc_mode="script" compute"""
    g = refpin.fixture_script(name)
    o = refpin.make_oracle_script(cls, name)
    assert np.array_equal(o.S, g["S"])
    for _ in range(int(g["steps"])):
        o.step()
    fl = g["solid"] == 0
    for n in ("F", "rho", "v"):
        assert np.array_equal(getattr(o, n)[fl], g[n][fl]), n


@pytest.mark.parametrize("name", refpin.NAMES_GREY)
@pytest.mark.parametrize("cls", [RefSinglePhase, RefSinglePhaseC])
def test_grey_scale_script(name, cls):
    """Grey_Scale/lbm_solver_3d_Macro_Sukop.py through the shim (parameter lines replaced, ns assigned
    instead of read from BC.dat): the script copy's physics plus the partial bounce-back of
    streaming0/streaming1 (:233-247) on a lattice with open, grey and fully solid nodes -- bit for bit
    what set_grey_scale(ns) computes.  Also pins the side effect the fused kernels rely on: a link that
    leaves a solid node is never written by the script, so it holds w[s] of init() for ever."""
    from oracle.ref_single_phase import E, W64
    g = refpin.fixture_grey(name)
    o = refpin.make_oracle_grey(cls, name)
    assert np.array_equal(o.S, g["S"]) and np.array_equal(o.solid, g["solid"])
    for _ in range(int(g["steps"])):
        o.step()
    fl = g["solid"] == 0
    for n in ("F", "rho", "v"):
        assert np.array_equal(getattr(o, n)[fl], g[n][fl]), n
    interior = np.zeros(fl.shape, bool)
    interior[1:-1] = True                      # x faces may carry a boundary condition that overwrites F
    for s in range(1, 19):
        src_solid = np.roll(g["solid"], tuple(int(c) for c in E[s]), axis=(0, 1, 2)) > 0
        m = fl & src_solid & interior
        assert m.any() and np.all(g["F"][..., s][m] == np.float32(W64[s])), s


@pytest.mark.parametrize("name", refpin.NAMES)
def test_reference_sparse_storage_semantics(name):
    """sparse_storage=True of the reference (pointer SNode tree of 3^3 blocks, :36-44, modelled by
    the shim): on fluid nodes it computes bit for bit what its dense mode computes; solid cells
    inside an activated block read rho = 1, v = 0 (:390-392) and keep F = 0 (init :164 skips them).
    The CUDA build returns the dense convention (rho = 1, v = 0, F = w) on solid nodes in both
    modes, so parity is asserted on fluid nodes (SURVEY 8a': 6)."""
    g = refpin.fixture(name)
    fl = g["solid"] == 0
    for k in ("F", "rho", "v"):
        assert np.array_equal(g[k + "_sparse"][fl], g[k][fl]), k
    nx, ny, nz = g["solid"].shape
    blk = np.repeat(np.repeat(np.repeat(g["active_blocks"], 3, 0), 3, 1), 3, 2)[:nx, :ny, :nz]
    assert blk[fl].all()                                   # every fluid cell lives in an active block
    solid_active = ~fl & blk
    assert np.all(g["rho_sparse"][solid_active] == 1.0) and np.all(g["v_sparse"][solid_active] == 0.0)
    assert np.all(g["F_sparse"][~fl] == 0.0) and np.all(g["rho_sparse"][~fl & ~blk] == 0.0)


@pytest.mark.skipif(not os.path.exists(refpin.mk.REF), reason="/root/reference is not mounted here")
def test_fixtures_are_what_the_reference_computes():
    mod = refpin.mk.load_reference()
    for name in ("pressure_x", "lid_and_force"):
        out = refpin.mk.run_reference(mod, name)
        g = refpin.fixture(name)
        for k in ("F", "rho", "v", "F1", "S", "inv_M"):
            assert np.array_equal(out[k], g[k]), (name, k)


@pytest.mark.parametrize("name", refpin.NAMES2P)
def test_two_phase_oracle_matches_reference_script(name):
    """tests/golden/ref_tp_*.npz: 2phase/lbm_solver_3d_2phase.py, its kernels executed through the
    shim with only the hand-edited parameter lines replaced (make_reference_fixtures.py).

    The script adds g_r, g_b into rhor / rhob with float atomics whose order Taichi leaves open
    (:365-372).  With the additions done one after the other in node order -- what the shim's
    sequential run does -- the oracle reproduces the script BIT FOR BIT on all six fields; with its
    own order (ascending direction at the destination, the order the CUDA path uses) it differs
    from it by that summation order only: fp32 round-off."""
    from oracle.cref import RefTwoPhaseC
    from oracle.ref_two_phase import RefTwoPhase

    class InPushOrder(RefTwoPhase):
        _accumulate_colour = RefTwoPhase.accumulate_colour_in_push_order

    g = refpin.fixture2p(name)
    fl = g["solid"] == 0
    o = refpin.case2p(name).make_oracle(InPushOrder)
    for _ in range(int(g["steps"])):
        o.step()
    for n in refpin.FIELDS2P:
        assert np.array_equal(getattr(o, n)[fl], g[n][fl]), n
    for cls in (RefTwoPhase, RefTwoPhaseC):
        o = refpin.case2p(name).make_oracle(cls)
        for _ in range(int(g["steps"])):
            o.step()
        refpin.check2p(lambda n: getattr(o, n), g, cls.__name__)


@pytest.mark.parametrize("name", refpin.NAMES2P)
def test_two_phase_sparse_script_equals_dense_script(name):
    """2phase/lbm_solver_3d_2phase_sparse.py (pointer SNode fields, same kernels) computes on fluid
    nodes what the dense script computes -- the premise of serving both with one CUDA path"""
    g = refpin.fixture2p(name)
    fl = g["solid"] == 0
    for n in refpin.FIELDS2P:
        assert np.array_equal(g[n + "_sparse"][fl], g[n][fl]), n


@pytest.mark.skipif(not os.path.exists(refpin.mk.REF), reason="/root/reference is not mounted here")
def test_init_geo_of_the_reference_reads_what_our_loader_reads(tmp_path):
    """init_geo :173-177 (np.loadtxt, >0 -> 1, Fortran-order reshape) through the shim, against
    taichi_lbm3d_b200.geometry.load_geometry on the same text file, and on .npy / .raw copies"""
    from taichi_lbm3d_b200 import geometry
    rng = np.random.default_rng(3)
    shape = (4, 3, 5)
    vals = rng.integers(0, 3, size=shape[0] * shape[1] * shape[2])          # 0, 1 and a stray 2
    path = str(tmp_path / "geo.dat")
    np.savetxt(path, vals, fmt="%d")
    mod = refpin.mk.load_reference()
    lb = mod.LB3D_Solver_Single_Phase(nx=shape[0], ny=shape[1], nz=shape[2])
    lb.init_geo(path)
    want = lb.solid.to_numpy()
    got = geometry.load_geometry(path, *shape)
    assert got.dtype == np.int8 and np.array_equal(got, want)
    raw = str(tmp_path / "geo.raw")
    vals.astype(np.uint8).tofile(raw)
    assert np.array_equal(geometry.load_geometry(raw, *shape), want)
    npy = str(tmp_path / "geo.npy")
    np.save(npy, want)
    assert np.array_equal(geometry.load_geometry(npy, *shape), want)
    txt = str(tmp_path / "geo2.dat")
    geometry.save_geometry_text(txt, want)                                    # the generator's format
    lb.init_geo(txt)
    assert np.array_equal(lb.solid.to_numpy(), want)


@pytest.mark.skipif(not os.path.exists(refpin.mk.REF2P), reason="/root/reference is not mounted here")
def test_two_phase_init_geo_reads_what_our_class_reads(tmp_path):
    """init_geo(filename, filename2) of the two-phase script (:194-202: geometry >0 -> 1, phase as
    floats, both reshaped in Fortran order) against LB3D_Solver_Two_Phase.init_geo"""
    import re
    import sys
    from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase
    shape = (4, 3, 5)
    rng = np.random.default_rng(5)
    n = shape[0] * shape[1] * shape[2]
    geo, pha = str(tmp_path / "img.txt"), str(tmp_path / "phase.dat")
    np.savetxt(geo, rng.integers(0, 3, size=n), fmt="%d")
    np.savetxt(pha, rng.choice([-1.0, 1.0], size=n), fmt="%.1f")
    src = open(refpin.mk.REF2P).read()
    head = src[:src.index("time_init = time.time()")]
    head = re.sub(r"^nx,ny,nz\s*=.*$", "nx,ny,nz = %d,%d,%d" % shape, head, count=1, flags=re.M)
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "taichi_shim")
    sys.path.insert(0, shim)
    try:
        for m in ("taichi", "pyevtk", "pyevtk.hl"):
            sys.modules.pop(m, None)
        ns = {"__name__": "lbm_solver_3d_2phase"}
        exec(compile(head, refpin.mk.REF2P, "exec"), ns)
    finally:
        sys.path.remove(shim)
        sys.modules.pop("taichi", None)
    solid_ref, phase_ref = ns["init_geo"](geo, pha)
    lb = LB3D_Solver_Two_Phase(*shape)
    solid, phase = lb.init_geo(geo, pha)
    assert np.array_equal(solid, solid_ref) and np.array_equal(phase, phase_ref.astype(np.float32))
    assert np.array_equal(lb.solid.to_numpy(), solid_ref) and np.array_equal(lb.psi.to_numpy(), phase_ref.astype(np.float32))


def test_shim_is_not_reachable_from_the_product():
    """the stand-in lives under tests/ and no product module imports taichi"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hits = subprocess.run(["grep", "-rlE", r"^\s*(import|from) taichi\b", os.path.join(root, "taichi_lbm3d_b200"),
                           os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")],
                          stdout=subprocess.PIPE).stdout.decode().split()
    assert hits == []
