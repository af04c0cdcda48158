"""GPU parity tests: the CUDA path (through the ctypes/C-ABI binding) against the oracle.

Bars (BASELINE.json north_star): compaction and neighbour tables bit-exact; populations,
rho and v within relative L-inf 1e-5 after 1000 steps in fp32 (``TOL``).  In verification
mode (strict=True: oracle evaluation order, no FMA contraction) the CUDA kernels must be
BIT-IDENTICAL to the oracle, which pins all the plumbing (pull streaming, bounce-back,
face BCs with their ordering, the pipeline state machine) exactly.
"""
import numpy as np
import pytest

from tests import cases
from tests.cases import rel_linf

pytestmark = pytest.mark.gpu

TOL = cases.TOL  # relative L-infinity, populations / rho / v (north_star)


_V_TOL = {}


def _v_tol(case, steps, o32):
    """see cases.v_abs_tolerance: max(1e-5 max|v|, 2 x the oracle's own fp32 round-off)"""
    key = (case.name, case.shape, steps)
    if key not in _V_TOL:
        o64, _ = _oracle(case, steps, dtype=np.float64)
        _V_TOL[key] = cases.v_abs_tolerance(o32, o64)
    return _V_TOL[key]


def _oracle(case, steps, **kw):
    from oracle.cref import RefSinglePhaseC
    o = case.make_oracle(RefSinglePhaseC, **kw)
    o0 = None
    if case.perturb:
        class _S:       # start state snapshot
            pass
        o0 = _S()
        o0.F, o0.rho, o0.v = o.F.copy(), o.rho.copy(), o.v.copy()
    o.run(steps)
    return o, o0


def _run_solver(case, steps, sparse, strict, start):
    lb = case.make_solver(sparse=sparse, strict=strict)
    if start is not None:
        case.apply_start(lb, start)
    lb.run(steps)
    return lb


def _compare(lb, o, exact, case=None, steps=None):
    fl = o.solid == 0
    F, rho, v = lb.F.to_numpy(), lb.rho.to_numpy(), lb.v.to_numpy()
    if exact:
        assert np.array_equal(F[fl], o.F[fl])
        assert np.array_equal(rho[fl], o.rho[fl])
        assert np.array_equal(v[fl], o.v[fl])
    else:
        errs = {"F": rel_linf(F[fl], o.F[fl]), "rho": rel_linf(rho[fl], o.rho[fl])}
        assert max(errs.values()) <= TOL, errs
        dv = float(np.abs(v[fl].astype(np.float64) - o.v[fl]).max())
        vtol = TOL * float(np.abs(o.v[fl]).max())
        if dv > vtol:                    # creeping flow: fall back to the fp64 yardstick
            vtol = _v_tol(case, steps, o)
        assert dv <= vtol, ("v", dv, vtol, float(np.abs(o.v[fl]).max()))
    # solid nodes keep the dense convention rho=1, v=0, F=w (reference :164-169, :390-392)
    from taichi_lbm3d_b200.constants import W
    sd = ~fl
    if sd.any():
        assert np.all(rho[sd] == 1.0) and np.all(v[sd] == 0.0)
        assert np.array_equal(F[sd], np.broadcast_to(W, F[sd].shape))
    return F, rho, v


CASES = [cases.case_mixed_bc, cases.case_periodic_force, cases.case_all_faces, cases.case_force_field,
         cases.case_other_copy]


@pytest.mark.parametrize("make", CASES)
@pytest.mark.parametrize("sparse", [False, True, "aa", "daa"])
@pytest.mark.parametrize("steps", [1, 2, 25])
def test_strict_bit_identical(cuda, make, sparse, steps):
    case = make()
    o, o0 = _oracle(case, steps)
    lb = _run_solver(case, steps, sparse, True, o0)
    _compare(lb, o, exact=True)


@pytest.mark.parametrize("make", CASES)
@pytest.mark.parametrize("sparse", [False, True, "aa", "daa"])
def test_fast_parity_small(cuda, make, sparse):
    case = make()
    o, o0 = _oracle(case, 200)
    lb = _run_solver(case, 200, sparse, False, o0)
    _compare(lb, o, False, case, 200)


GOLDEN = {"mixed_bc": cases.case_mixed_bc, "all_faces": cases.case_all_faces,
          "periodic_force": cases.case_periodic_force, "force_field": cases.case_force_field,
          "other_copy": cases.case_other_copy}


@pytest.mark.parametrize("name", sorted(GOLDEN))
@pytest.mark.parametrize("sparse", [False, True, "aa", "daa"])
def test_committed_golden_vectors(cuda, name, sparse):
    """verification arithmetic against the committed fixtures (tests/golden/sp_*.npz): start state
    and result after `steps` steps, bit for bit, without running the oracle"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sp_%s.npz" % name))
    case = GOLDEN[name]()
    assert np.array_equal(case.solid, g["solid"])
    lb = case.make_solver(sparse=sparse, strict=True)
    lb.F.from_numpy(g["F0"])
    lb.rho.from_numpy(g["rho0"])
    lb.v.from_numpy(g["v0"])
    lb.run(int(g["steps"]))
    fl = case.solid == 0
    assert np.array_equal(lb.F.to_numpy()[fl], g["F"][fl])
    assert np.array_equal(lb.rho.to_numpy()[fl], g["rho"][fl])
    assert np.array_equal(lb.v.to_numpy()[fl], g["v"][fl])


@pytest.mark.parametrize("sparse", [False, True, "aa", "daa"])
def test_reference_source_fixtures(cuda, sparse):
    """tests/golden/ref_sp_*.npz: what the reference's own source computes (run through
    tests/taichi_shim, tests/golden/make_reference_fixtures.py).  Verification arithmetic bit for
    bit, production arithmetic within 1e-5, through the reference's own method names."""
    from tests import refpin
    for name in refpin.NAMES:
        g = refpin.fixture(name)
        fl = g["solid"] == 0
        lb = refpin.make_solver(name, sparse=sparse, strict=True)
        lb.step()
        assert np.array_equal(lb.F.to_numpy()[fl], g["F1"][fl]) and np.array_equal(lb.v.to_numpy()[fl], g["v1"][fl])
        for _ in range(int(g["steps"]) - 1):
            lb.step()
        for n in ("F", "rho", "v"):
            assert np.array_equal(getattr(lb, n).to_numpy()[fl], g[n][fl]), (name, n)
        assert abs(lb.get_max_v() - float(g["max_v"])) <= 2e-7 * float(g["max_v"]) + 1e-12
        lbf = refpin.make_solver(name, sparse=sparse, strict=False)
        lbf.run(int(g["steps"]))
        assert rel_linf(lbf.F.to_numpy()[fl], g["F"][fl]) <= TOL and rel_linf(lbf.rho.to_numpy()[fl], g["rho"][fl]) <= TOL
        dv = np.abs(lbf.v.to_numpy()[fl] - g["v"][fl]).max()
        assert dv <= max(TOL * np.abs(g["v"][fl]).max(), 2e-7), (name, dv)


@pytest.mark.parametrize("sparse", [False, True])
def test_reference_other_copy_and_force_hook_fixtures(cuda, sparse):
    """tests/golden/ref_sp_other_copy_*.npz (Phase_change/LBM_3D_SinglePhase_Solver.py through the shim:
    tau = 3 niu + 1/2, un-scaled Guo term) and ref_sp_local_force_*.npz (a subclass overriding the
    cal_local_force hook): verification arithmetic bit for bit"""
    import os
    from tests import refpin
    for name in ("lid_and_force", "periodic_force"):
        g = np.load(os.path.join(refpin.GOLD, "ref_sp_other_copy_%s.npz" % name))
        fl = g["solid"] == 0
        lb = refpin.make_solver(name, sparse=sparse, strict=True, tau_mode="textbook", guo_mode="unscaled")
        lb.run(int(g["steps"]))
        for n in ("F", "rho", "v"):
            assert np.array_equal(getattr(lb, n).to_numpy()[fl], g[n][fl]), ("other copy", name, n)
    for name in ("pressure_x", "periodic_force"):
        g = np.load(os.path.join(refpin.GOLD, "ref_sp_local_force_%s.npz" % name))
        fl = g["solid"] == 0
        lb = refpin.make_solver(name, sparse=sparse, strict=True, force_field=g["force_field"])
        lb.run(int(g["steps"]))
        for n in ("F", "rho", "v"):
            assert np.array_equal(getattr(lb, n).to_numpy()[fl], g[n][fl]), ("force hook", name, n)


@pytest.mark.parametrize("sparse", [False, True, "aa", "daa"])
def test_reference_script_copy_fixtures(cuda, sparse):
    """tests/golden/ref_script_*.npz: what Single_phase/lbm_solver_3d.py -- the flat-script copy with
    tau = 3 niu + 1/2, the un-scaled Guo term and the in-place fixed-velocity faces (:253,:268) --
    computes through the shim; tau_mode="textbook", guo_mode="unscaled", vel_bc_mode="script"."""
    from tests import refpin
    for name in refpin.NAMES_SCRIPT:
        g = refpin.fixture_script(name)
        fl = g["solid"] == 0
        lb = refpin.make_solver_script(name, sparse=sparse, strict=True)
        lb.run(int(g["steps"]))
        for n in ("F", "rho", "v"):
            assert np.array_equal(getattr(lb, n).to_numpy()[fl], g[n][fl]), (name, n)
        lbf = refpin.make_solver_script(name, sparse=sparse, strict=False)
        lbf.run(int(g["steps"]))
        assert rel_linf(lbf.F.to_numpy()[fl], g["F"][fl]) <= TOL and rel_linf(lbf.rho.to_numpy()[fl], g["rho"][fl]) <= TOL
        assert np.abs(lbf.v.to_numpy()[fl] - g["v"][fl]).max() <= 3e-7


# ---- grey-scale lattice (Grey_Scale/lbm_solver_3d_Macro_Sukop.py) -----------------------------------
def test_reference_grey_scale_fixtures(cuda):
    """tests/golden/ref_grey_*.npz: what the grey-scale script computes through the shim on lattices
    with open, grey and fully solid nodes (periodic + body force; fixed-pressure x faces): verification
    arithmetic bit for bit, production arithmetic to the north_star's 1e-5"""
    from tests import refpin
    for name in refpin.NAMES_GREY:
        g = refpin.fixture_grey(name)
        fl = g["solid"] == 0
        lb = refpin.make_solver_grey(name, strict=True)
        assert np.array_equal(lb.solid.to_numpy(), g["solid"]) and np.array_equal(lb.ns.to_numpy(), g["ns"])
        lb.run(int(g["steps"]))
        for n in ("F", "rho", "v"):
            assert np.array_equal(getattr(lb, n).to_numpy()[fl], g[n][fl]), (name, n)
        lbf = refpin.make_solver_grey(name, strict=False)
        for _ in range(int(g["steps"])):
            lbf.step()
        assert rel_linf(lbf.F.to_numpy()[fl], g["F"][fl]) <= TOL and rel_linf(lbf.rho.to_numpy()[fl], g["rho"][fl]) <= TOL
        assert np.abs(lbf.v.to_numpy()[fl] - g["v"][fl]).max() <= 3e-7


def _grey_pair(shape, seed, strict, steps, faces=()):
    from oracle.cref import RefSinglePhaseC
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    rng = np.random.default_rng(seed)
    kind, frac = rng.random(shape), rng.random(shape)
    ns = np.where(kind < 0.2, 1.0, np.where(kind < 0.5, 0.0, frac)).astype(np.float32)
    phys = dict(tau_mode="textbook", guo_mode="unscaled")
    o = RefSinglePhaseC(*shape, **phys)
    o.set_grey_scale(ns)
    lb = LB3D_Solver_Single_Phase(*shape, strict=strict, **phys)
    lb.ns.from_numpy(ns)
    for face, rho in faces:
        o.set_bc_rho(face, rho)
        getattr(lb, cases.FACE_SETTERS_RHO[face])(rho)
    for s in (o, lb):
        s.set_force([2e-5, -1e-5, 5e-6])
        s.set_viscosity(0.12)
        s.init_simulation()
    o.run(steps)
    return o, lb


@pytest.mark.parametrize("faces", [(), ((2, 1.0), (3, 0.99)), ((0, 1.0), (1, 0.995), (4, 1.01), (5, 1.0))])
def test_grey_scale_against_the_oracle(cuda, faces):
    """a 20 x 18 x 33 grey lattice (periodic wrap on every axis, z rows longer than a warp), 40 steps:
    verification arithmetic bit-identical to the C oracle, step() and run() alike; production
    arithmetic within 1e-5"""
    o, lb = _grey_pair((20, 18, 33), 5, True, 40, faces)
    lb.run(17)
    assert np.isfinite(lb.rho.to_numpy()).all()            # a read in between closes and re-opens the pipeline
    for _ in range(3):
        lb.step()
    lb.run(20)
    _compare(lb, o, exact=True)
    o, lbf = _grey_pair((20, 18, 33), 5, False, 40, faces)
    lbf.run(40)
    fl = o.solid == 0
    assert rel_linf(lbf.F.to_numpy()[fl], o.F[fl]) <= TOL and rel_linf(lbf.rho.to_numpy()[fl], o.rho[fl]) <= TOL
    assert np.abs(lbf.v.to_numpy()[fl] - o.v[fl]).max() <= max(TOL * float(np.abs(o.v[fl]).max()), 3e-7)


def test_grey_scale_reference_case(cuda):
    """the case the grey-scale script ships (BC.dat: 60 x 50 x 5 channel with a grey layer, fx = 1e-6,
    niu = 0.1, all faces periodic; examples/example_grey_scale.py), 300 steps against the C oracle:
    verification arithmetic bit for bit; the flow in the grey layer is slower than in the open part"""
    from oracle.cref import RefSinglePhaseC
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase, geometry
    ns = geometry.grey_channel()
    phys = dict(tau_mode="textbook", guo_mode="unscaled")
    o = RefSinglePhaseC(*ns.shape, **phys)
    o.set_grey_scale(ns)
    lb = LB3D_Solver_Single_Phase(*ns.shape, strict=True, **phys)
    lb.ns.from_numpy(ns)
    for s in (o, lb):
        s.set_force([1.0e-6, 0.0, 0.0])
        s.set_viscosity(0.1)
        s.init_simulation()
    o.run(300)
    lb.run(300)
    _compare(lb, o, exact=True)
    v = lb.v.to_numpy()
    assert 0 < v[:, 10, :, 0].mean() < 0.5 * v[:, 35, :, 0].mean()


def test_grey_scale_with_zero_fraction_is_the_plain_lattice_without_walls(cuda):
    """ns = 0 everywhere: the blend vanishes, the step is the ordinary periodic one (no solid nodes)"""
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    shape = (9, 8, 7)
    a = LB3D_Solver_Single_Phase(*shape, strict=True)
    b = LB3D_Solver_Single_Phase(*shape, strict=True)
    b.ns.from_numpy(np.zeros(shape, np.float32))
    for s in (a, b):
        s.set_force([1e-5, 2e-5, -1e-5])
        s.init_simulation()
        s.run(12)
    assert np.array_equal(a.F.to_numpy(), b.F.to_numpy()) and np.array_equal(a.v.to_numpy(), b.v.to_numpy())


def test_grey_scale_refuses_what_it_cannot_reproduce(cuda):
    """sparse / in-place storage and fixed-velocity faces are refused loudly, not approximated"""
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    from taichi_lbm3d_b200._lib import LbmError
    ns = np.full((6, 5, 4), 0.3, np.float32)
    for kw in (dict(sparse_storage=True), dict(in_place=True)):
        lb = LB3D_Solver_Single_Phase(6, 5, 4, **kw)
        lb.ns.from_numpy(ns)
        with pytest.raises(LbmError, match="grey-scale"):
            lb.init_simulation()
    lb = LB3D_Solver_Single_Phase(6, 5, 4)
    lb.ns.from_numpy(ns)
    lb.set_bc_vel_x0([0.01, 0.0, 0.0])
    with pytest.raises(LbmError, match="grey-scale"):
        lb.init_simulation()


def test_grey_scale_checkpoint(cuda, tmp_path):
    o, lb = _grey_pair((8, 7, 6), 9, True, 10)
    lb.run(6)
    path = str(tmp_path / "grey.npz")
    lb.save_checkpoint(path)
    _o2, lb2 = _grey_pair((8, 7, 6), 9, True, 0)
    lb2.load_checkpoint(path)
    lb2.run(4)
    _compare(lb2, o, exact=True)
    _o3, lb3 = _grey_pair((8, 7, 6), 10, True, 0)          # another ns on the same lattice size
    with pytest.raises(ValueError):
        lb3.load_checkpoint(path)


@pytest.mark.parametrize("sparse", [False, True, "aa", "daa"])
def test_script_velocity_faces_on_every_axis(cuda, sparse):
    """the in-place velocity form next to pressure faces on all six faces (the script copy has x
    faces only; the class API has six): a later pressure face overwrites, every velocity face after
    the last pressure face is applied in order -- against the oracle, bit for bit, 30 steps"""
    from oracle.cref import RefSinglePhaseC
    solid = cases.random_porous((10, 9, 8), 0.2, 77)
    bcs = [(0, "vel", [0.02, 0.0, 0.0]), (1, "rho", 0.98), (2, "vel", [0.0, 0.01, 0.0]), (3, "vel", [0.0, -0.01, 0.01]),
           (4, "rho", 1.01), (5, "vel", [0.0, 0.0, 0.03])]
    o = RefSinglePhaseC(*solid.shape, vel_bc_mode="script")
    o.set_solid(solid)
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    lb = LB3D_Solver_Single_Phase(*solid.shape, sparse_storage=sparse in (True, "aa"), in_place=sparse in ("aa", "daa"),
                                  strict=True, vel_bc_mode="script")
    lb.solid.from_numpy(solid)
    for face, kind, val in bcs:
        if kind == "rho":
            o.set_bc_rho(face, val)
            getattr(lb, cases.FACE_SETTERS_RHO[face])(val)
        else:
            o.set_bc_vel(face, val)
            getattr(lb, cases.FACE_SETTERS_VEL[face])(val)
    o.set_force([1e-5, 0.0, -1e-5])
    lb.set_force([1e-5, 0.0, -1e-5])
    o.init_simulation()
    lb.init_simulation()
    o.run(30)
    lb.run(30)
    _compare(lb, o, exact=True)


def test_buffer_cache_reuses_dirty_buffers_without_changing_results(cuda):
    """csrc/lbm_devpool.cuh: a solver set up after another one of the same size was closed gets that
    one's device buffers back, contents and all (a 128 x 64 x 64 lattice: every large buffer is above
    the cache's 1 MiB threshold) -- its results must not depend on what they held; lbm_pool_trim()
    returns the memory to the driver"""
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase, _lib
    shape = (128, 64, 64)
    solid = cases.random_porous(shape, 0.1, 3)

    def run(lid, steps, sparse=False):
        lb = LB3D_Solver_Single_Phase(*shape, strict=True, sparse_storage=sparse)
        lb.solid.from_numpy(solid)
        lb.set_bc_vel_x1(lid)
        lb.set_force([1e-5, 0.0, 0.0])
        lb.init_simulation()
        lb.run(steps)
        out = lb.F.to_numpy(), lb.rho.to_numpy(), lb.v.to_numpy()
        lb.close()
        return out

    lib = _lib.load()
    lib.lbm_pool_trim()
    clean = run([0.0, 0.0, 0.05], 6)
    assert lib.lbm_pool_trim() > 0                      # the closed solver's buffers were cached
    run([0.0, 0.08, -0.03], 9)                          # another flow leaves its state in the cache ...
    run([0.0, 0.08, -0.03], 4, sparse=True)
    again = run([0.0, 0.0, 0.05], 6)                    # ... and this solver starts on those buffers
    for a, b in zip(clean, again):
        assert np.array_equal(a, b)
    assert lib.lbm_pool_trim() > 0 and lib.lbm_pool_trim() == 0


def test_step_by_step_equals_run(cuda):
    """step() x n, with field reads in between, equals run(n) (state machine, :477-481)."""
    case = cases.case_mixed_bc()
    o, o0 = _oracle(case, 7)
    lb = case.make_solver(strict=True)
    case.apply_start(lb, o0)
    for i in range(7):
        lb.step()
        if i in (2, 3):
            lb.rho.to_numpy()          # extraction pass must not disturb the pipeline
            lb.get_max_v()
        if i == 4:
            lb.F.to_numpy()
    _compare(lb, o, exact=True)


def test_from_numpy_restart(cuda):
    """F/rho/v.from_numpy mid-run restarts the pipeline from the user-visible state."""
    case = cases.case_periodic_force()
    o, o0 = _oracle(case, 5)
    lb = case.make_solver(strict=True, sparse=True)
    case.apply_start(lb, o0)
    lb.run(3)
    F, rho, v = lb.F.to_numpy(), lb.rho.to_numpy(), lb.v.to_numpy()
    lb.F.from_numpy(F)
    lb.rho.from_numpy(rho)
    lb.v.from_numpy(v)
    lb.run(2)
    _compare(lb, o, exact=True)


def test_cavity50_1000_steps(cuda):
    """reference example_cavity.py shape: 50^3 geo_cavity, lid vz=0.1 on x1, 1000 steps."""
    case = cases.case_cavity(50)
    o, _ = _oracle(case, 1000)
    for sparse in (False, True):
        lb = _run_solver(case, 1000, sparse, False, None)
        _, _, v = _compare(lb, o, False, case, 1000)
        assert abs(lb.get_max_v() - o.get_max_v()) <= 1e-6
        # the cavity geometry is mirror-symmetric in y and the lid moves along z: so is the flow
        assert rel_linf(v[:, ::-1, :, 2], v[..., 2]) < 1e-4
    lbs = _run_solver(case, 1000, False, True, None)
    _compare(lbs, o, exact=True)


def test_poiseuille_reference_example(cuda):
    """example_poiseuille_flow.py: plates at z=0 and z=15, force fy=1e-4, niu=0.1667."""
    case = cases.case_poiseuille()
    steps = 3000
    o, _ = _oracle(case, steps)
    lb = _run_solver(case, steps, False, False, None)
    _compare(lb, o, False, case, steps)
    lbs = _run_solver(case, steps, True, True, None)
    _compare(lbs, o, exact=True)


def test_porous_pressure_bc_1000_steps(cuda):
    """config-1 shape (example_porous_medium.py): sphere-pack stand-in, rho 1.0 -> 0.99 in x."""
    case = cases.case_porous(64)
    o, _ = _oracle(case, 1000)
    for sparse in (False, True):
        lb = _run_solver(case, 1000, sparse, False, None)
        _compare(lb, o, False, case, 1000)
    # pressure faces: v stays exactly 0 without a force (SURVEY section 4 KAT)
    v = lb.v.to_numpy()
    fl = case.solid == 0
    assert np.all(v[0][fl[0]] == 0.0) and np.all(v[-1][fl[-1]] == 0.0)


def test_sparse_tables_bit_exact(cuda):
    """fluid-node compaction and 18-neighbour pull table vs a NumPy construction from solid,
    periodic_index (:247-257) and e (:183-187)."""
    from taichi_lbm3d_b200.constants import E
    # the last lattice holds > 65535 fluid nodes per x plane pair: the periodic x wrap makes the
    # neighbour-row ranks of some 256-node table blocks span more than 16 bits (exception path)
    for shape, frac, seed in [((12, 10, 9), 0.35, 3), ((33, 17, 40), 0.8, 9), ((7, 7, 7), 0.0, 1),
                              ((40, 48, 48), 0.1, 5)]:
        solid = cases.random_porous(shape, frac, seed)
        case = cases.Case("t", solid, bc=[(0, "rho", 1.0), (5, "vel", [0, 0, 0.01])])
        lb = case.make_solver(sparse=True)
        fluid = solid == 0
        lin = np.flatnonzero(fluid.reshape(-1))
        assert lb.num_fluid() == lin.size
        assert np.array_equal(lb.fluid_index(), lin)
        rank = np.full(solid.size, -1, np.int64)
        rank[lin] = np.arange(lin.size)
        rank3 = rank.reshape(shape)
        want = np.empty((18, lin.size), np.int32)
        for s in range(1, 19):
            src = np.roll(rank3, tuple(int(c) for c in E[s]), axis=(0, 1, 2))   # value at i - e_s
            want[s - 1] = src[fluid]
        assert np.array_equal(lb.neighbor_table(), want)
        # dense link words carry the same information
        lbd = case.make_solver(sparse=False)
        fl = lbd.link_flags()
        for s in range(1, 19):
            assert np.array_equal(((fl[fluid] >> s) & 1).astype(bool), want[s - 1] < 0)
        assert np.array_equal((fl >> 19) & 1, solid.astype(np.uint32))


@pytest.mark.parametrize("sparse", [True, "aa"])
def test_wide_table_blocks_bit_identical(cuda, sparse):
    """periodic lattice large enough that table blocks straddling the x wrap overflow the
    16-bit rank offsets (all their nodes become exceptions): still bit-identical to the oracle"""
    solid = cases.random_porous((40, 48, 48), 0.1, 5)
    case = cases.Case("wide", solid, force=[1e-5, 2e-6, -3e-6], perturb=1e-3)
    o, o0 = _oracle(case, 3)
    lb = _run_solver(case, 3, sparse, True, o0)
    _compare(lb, o, exact=True)


@pytest.mark.parametrize("sparse", [False, True, "aa", "daa"])
def test_force_field_replaced_between_steps(cuda, sparse):
    """the per-node force array (cal_local_force override point :217-220) may change between
    steps (buoyancy follows the temperature field in the solute solver) and be switched off"""
    case = cases.case_force_field()
    o, o0 = _oracle(case, 3)
    lb = _run_solver(case, 3, sparse, True, o0)
    ff2 = (case.force_field * np.float32(-0.5)).astype(np.float32)
    o.set_force_field(ff2)
    lb.set_force_field(ff2)
    o.run(3)
    lb.run(3)
    _compare(lb, o, exact=True)
    o.set_force_field(None)              # back to the uniform force (zero here)
    o._p.force_flag = 0
    lb.set_force_field(None)
    o.run(2)
    lb.run(2)
    _compare(lb, o, exact=True)


def test_rest_state_and_mass(cuda):
    """rest state is a fixed point to fp32 round-off; closed periodic box conserves mass."""
    from taichi_lbm3d_b200.constants import W
    solid = cases.random_porous((20, 18, 16), 0.3, 4)
    lb = cases.Case("rest", solid).make_solver()
    lb.run(50)
    F = lb.F.to_numpy()
    fl = solid == 0
    assert np.abs(F[fl] - W).max() < 5e-7
    assert np.abs(lb.v.to_numpy()).max() < 1e-6
    case = cases.Case("mass", solid, perturb=1e-3)
    o, o0 = _oracle(case, 1)
    lb = case.make_solver(sparse=True)
    case.apply_start(lb, o0)
    m0 = lb.F.to_numpy()[fl].astype(np.float64).sum()
    lb.run(200)
    m1 = lb.F.to_numpy()[fl].astype(np.float64).sum()
    assert abs(m1 - m0) / m0 < 5e-6


def test_full_size_properties_256(cuda):
    """BASELINE config 2 at full size (256^3 cavity): properties that need no oracle run --
    y-mirror symmetry, closed-cavity mass drift, dense == sparse, lid nodes at the BC value."""
    n = 256
    case = cases.case_cavity(n)
    lb = case.make_solver()
    lb.run(100)
    rho, v = lb.rho.to_numpy(), lb.v.to_numpy()
    fl = case.solid == 0
    assert np.isfinite(rho).all() and np.isfinite(v).all()
    assert rel_linf(v[:, ::-1, :, 2], v[..., 2]) < 1e-4
    assert rel_linf(v[:, ::-1, :, 1], -v[..., 1]) < 1e-4
    lid = fl[n - 1]
    assert np.abs(v[n - 1][lid] - np.array([0, 0, 0.1], np.float32)).max() < 2e-7
    lbs = case.make_solver(sparse=True)
    lbs.run(100)
    assert rel_linf(lbs.rho.to_numpy(), rho) <= 1e-6
    assert rel_linf(lbs.v.to_numpy(), v) <= TOL
    assert lbs.num_fluid() == 255 * 254 * 254
    lba = case.make_solver(sparse="aa")                  # in place == two buffers, bit for bit
    lba.run(100)
    assert np.array_equal(lba.v.to_numpy(), lbs.v.to_numpy())
    lbda = case.make_solver(sparse="daa")                # dense in place == dense two buffers
    lbda.run(100)
    assert np.array_equal(lbda.v.to_numpy(), v) and np.array_equal(lbda.rho.to_numpy(), rho)


def test_full_size_properties_config3(cuda):
    """BASELINE config 3 at full size (512^3 periodic sphere pack, porosity 0.20, fx = 1e-6,
    sparse storage): properties that need no oracle run -- two-buffer == in-place bit for bit,
    sparse == dense, flow along the force, mass drift bounded by the reference's own Guo mass
    source (moment 0 gets (-8/27) v.f per step)."""
    from taichi_lbm3d_b200.geometry import sphere_pack
    n = 512
    solid = sphere_pack(n, n, n, 0.80, 8.0, 16.0, seed=n, periodic=True)
    case = cases.Case("cfg3", solid, force=[1e-6, 0.0, 0.0])
    fl = solid == 0
    lb = case.make_solver(sparse=True)
    assert lb.num_fluid() == int(fl.sum()) == 26809316
    lb.run(60)
    rho, v = lb.rho.to_numpy(), lb.v.to_numpy()
    assert np.isfinite(rho).all() and np.isfinite(v).all()
    assert v[fl][:, 0].mean() > 0 and abs(v[fl][:, 1].mean()) < 0.05 * v[fl][:, 0].mean()
    assert abs(rho[fl].astype(np.float64).mean() - 1.0) < 1e-6
    del lb
    lba = case.make_solver(sparse="aa")
    lba.run(60)
    assert np.array_equal(lba.v.to_numpy(), v) and np.array_equal(lba.rho.to_numpy(), rho)
    del lba
    lbd = case.make_solver(sparse=False)
    lbd.run(60)
    assert rel_linf(lbd.rho.to_numpy()[fl], rho[fl]) <= 1e-6
    assert np.abs(lbd.v.to_numpy()[fl] - v[fl]).max() <= TOL * np.abs(v[fl]).max()


def test_launches_are_counted(cuda):
    lb = cases.case_poiseuille().make_solver()
    n0 = lb.launch_count
    lb.run(10)
    assert lb.launch_count - n0 == 10


def test_config1_ftb131_1000_steps(cuda):
    """BASELINE config 1 as example_porous_medium.py runs it: 131^3, rho(x0)=1.0, rho(x1)=0.99,
    default viscosity, no force, dense storage -- on img_ftb131.txt when the user has it
    (LBM3D_FTB131 or ./img_ftb131.txt), else on the seeded stand-in; populations,
    rho, v after 1000 steps against the oracle."""
    from taichi_lbm3d_b200.geometry import ftb131_geometry
    solid, _source = ftb131_geometry()
    case = cases.Case("cfg1_ftb131", solid, bc=[(1, "rho", 0.99), (0, "rho", 1.0)])
    o, _ = _oracle(case, 1000)
    lb = _run_solver(case, 1000, False, False, None)
    _compare(lb, o, False, case, 1000)
    lbs = _run_solver(case, 1000, True, False, None)          # same case on the sparse storage
    _compare(lbs, o, False, case, 1000)
    assert abs(lb.get_max_v() - o.get_max_v()) <= 1e-6


def test_checkpoint_roundtrip(cuda, tmp_path):
    case = cases.case_mixed_bc()
    o, o0 = _oracle(case, 8)
    lb = case.make_solver(strict=True)
    case.apply_start(lb, o0)
    lb.run(5)
    path = str(tmp_path / "ckpt.npz")
    lb.save_checkpoint(path)
    lb2 = case.make_solver(strict=True, sparse=True)           # restart on the other storage mode
    lb2.load_checkpoint(path)
    lb2.run(3)
    _compare(lb2, o, exact=True)


EDGE_SHAPES = [(1, 4, 5), (2, 2, 2), (3, 1, 40), (5, 3, 33), (4, 6, 1)]


@pytest.mark.parametrize("shape", EDGE_SHAPES)
@pytest.mark.parametrize("sparse", [False, True, "aa", "daa"])
def test_degenerate_extents(cuda, shape, sparse):
    """extents of 1 or 2 make periodic_index wrap a node onto itself or onto the same neighbour
    twice (:247-257); rows shorter / longer than a warp; all against the oracle, bit for bit"""
    solid = cases.random_porous(shape, 0.25, 5)
    solid.flat[0] = 0
    case = cases.Case("edge", solid, force=[1e-5, -2e-5, 3e-5], perturb=1e-3)
    o, o0 = _oracle(case, 6)
    lb = case.make_solver(sparse=sparse, strict=True)
    case.apply_start(lb, o0)
    lb.run(6)
    _compare(lb, o, exact=True)


@pytest.mark.parametrize("sparse", [False, True, "aa", "daa"])
def test_all_solid_and_single_fluid_node(cuda, sparse):
    """empty fluid set, and one fluid node enclosed by solid (every link bounces back)"""
    from taichi_lbm3d_b200.constants import W
    solid = np.ones((4, 5, 6), np.int8)
    lb = cases.Case("solid", solid).make_solver(sparse=sparse)
    lb.run(3)
    assert lb.num_fluid() == 0
    assert np.all(lb.rho.to_numpy() == 1.0) and np.all(lb.v.to_numpy() == 0.0)
    assert lb.get_max_v() == 0.0
    solid[2, 2, 3] = 0
    case = cases.Case("one", solid, force=[1e-4, 0.0, 0.0], perturb=1e-3)
    o, o0 = _oracle(case, 5)
    lb = case.make_solver(sparse=sparse, strict=True)
    case.apply_start(lb, o0)
    lb.run(5)
    _compare(lb, o, exact=True)
    assert lb.num_fluid() == 1
    if sparse in (True, "aa"):
        assert np.all(lb.neighbor_table() == -1)
    # mass of the enclosed node is conserved up to the reference's Guo mass sink (moment 0)
    assert abs(lb.F.to_numpy()[2, 2, 3].sum() - o.F[2, 2, 3].sum()) == 0.0
    assert np.array_equal(lb.F.to_numpy()[0, 0, 0], W)
