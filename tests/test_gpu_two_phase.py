"""GPU parity of the two-phase colour-gradient path against the oracle (same bars as the
single-phase path: verification mode bit-identical, production arithmetic within 1e-5
relative L-inf; v against the oracle's own fp32 round-off in creeping flow)."""
import numpy as np
import pytest

from tests import cases2p
from tests.cases import TOL, rel_linf

pytestmark = pytest.mark.gpu

FIELDS = ("F", "rho", "v", "psi", "rho_r", "rho_b")
CASES = [cases2p.case_drainage, cases2p.case_bcs, cases2p.case_periodic_bubble]


def _oracle(case, steps, **kw):
    from oracle.cref import RefTwoPhaseC
    o = case.make_oracle(RefTwoPhaseC, **kw)
    o.run(steps)
    return o


@pytest.mark.parametrize("make", CASES)
@pytest.mark.parametrize("steps", [1, 2, 20])
@pytest.mark.parametrize("sparse", [False, True])
def test_strict_bit_identical(cuda, make, steps, sparse):
    case = make()
    o = _oracle(case, steps)
    lb = case.make_solver(strict=True, sparse=sparse)
    lb.run(steps)
    fl = case.solid == 0
    for n in FIELDS:
        got = getattr(lb, n).to_numpy()
        assert np.array_equal(got[fl], getattr(o, n)[fl]), n
    # solid nodes keep the input phase value
    assert np.array_equal(lb.psi.to_numpy()[~fl], case.psi[~fl])


@pytest.mark.parametrize("make", CASES)
@pytest.mark.parametrize("sparse", [False, True])
def test_committed_golden_vectors(cuda, make, sparse):
    """verification arithmetic against the committed fixtures (tests/golden/tp_*.npz)"""
    import os
    case = make()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "tp_%s.npz" % case.name))
    assert np.array_equal(case.solid, g["solid"]) and np.array_equal(case.psi, g["psi0"])
    lb = case.make_solver(strict=True, sparse=sparse)
    lb.run(int(g["steps"]))
    fl = case.solid == 0
    for n in FIELDS:
        assert np.array_equal(getattr(lb, n).to_numpy()[fl], g[n][fl]), n


@pytest.mark.parametrize("sparse", [False, True])
@pytest.mark.parametrize("strict", [True, False])
def test_reference_script_fixtures(cuda, sparse, strict):
    """tests/golden/ref_tp_*.npz: what the reference's own two-phase script computes (its kernels run
    through tests/taichi_shim); fp32 round-off tolerance because the script's colour accumulation
    is order-dependent (tests/refpin.py: check2p)"""
    from tests import refpin
    for name in refpin.NAMES2P:
        g = refpin.fixture2p(name)
        lb = refpin.case2p(name).make_solver(strict=strict, sparse=sparse)
        lb.run(int(g["steps"]))
        refpin.check2p(lambda n: getattr(lb, n).to_numpy(), g, name, production=not strict)


def _yardstick(o32, o64, name, fl):
    """max(1e-5 relative, 2 x the fp32 oracle's own distance to its fp64 form): interface
    dynamics amplify fp32 round-off (the oracle built with -ffast-math drifts 1e-4 in psi from
    the strict build after 300 steps of the droplet case), see tests/cases.py v_abs_tolerance"""
    a, b = getattr(o32, name)[fl], getattr(o64, name)[fl]
    return max(TOL * float(np.abs(a).max()), 2.0 * float(np.abs(a.astype(np.float64) - b).max()))


@pytest.mark.parametrize("make", CASES)
@pytest.mark.parametrize("sparse", [False, True])
def test_fast_parity(cuda, make, sparse):
    case = make()
    fl = case.solid == 0
    # short horizon: plain 1e-5 relative L-inf on everything but v
    o = _oracle(case, 50)
    o64 = _oracle(case, 50, dtype=np.float64)
    lb = case.make_solver(sparse=sparse)
    lb.run(50)
    for n in ("F", "rho", "psi", "rho_r", "rho_b"):
        assert rel_linf(getattr(lb, n).to_numpy()[fl], getattr(o, n)[fl]) <= TOL, n
    dv = float(np.abs(lb.v.to_numpy()[fl].astype(np.float64) - o.v[fl]).max())
    assert dv <= _yardstick(o, o64, "v", fl), dv
    # long horizon: every field against the oracle's own fp32 round-off
    o.run(250)
    o64.run(250)
    lb.run(250)
    for n in FIELDS:
        d = float(np.abs(getattr(lb, n).to_numpy()[fl].astype(np.float64) - getattr(o, n)[fl]).max())
        assert d <= _yardstick(o, o64, n, fl), (n, d)


@pytest.mark.parametrize("sparse", [False, True])
def test_step_equals_run_and_restart(cuda, sparse):
    case = cases2p.case_drainage()
    o = _oracle(case, 9)
    lb = case.make_solver(strict=True, sparse=sparse)
    for i in range(5):
        lb.step()
        if i == 2:
            lb.psi.to_numpy()
            lb.get_max_v()
    state = [getattr(lb, n).to_numpy() for n in ("F", "rho", "v", "psi", "rho_r", "rho_b")]
    lb.set_state(*state)
    lb.run(4)
    fl = case.solid == 0
    for n in FIELDS:
        assert np.array_equal(getattr(lb, n).to_numpy()[fl], getattr(o, n)[fl]), n
    assert lb.launch_count > 0


def test_drainage_131_properties(cuda):
    """BASELINE config 4 at full size (131^3 stand-in, README parameters): finite, colour
    conserved up to what the constant-psi inlet injects, interface kept sharp."""
    from taichi_lbm3d_b200.geometry import ftb131_standin
    solid = ftb131_standin()
    psi = np.ones(solid.shape, np.float32)
    psi[:13] = -1.0
    case = cases2p.Case2P("cfg4", solid, psi, niu_l=0.05, niu_g=0.2, CapA=0.005, psi_solid=0.7)
    lb = case.make_solver()
    lb.run(200)
    fl = solid == 0
    p, rr, rb = lb.psi.to_numpy(), lb.rho_r.to_numpy(), lb.rho_b.to_numpy()
    assert np.isfinite(p[fl]).all() and np.isfinite(lb.F.to_numpy()[fl]).all()
    assert p[fl].min() > -1.2 and p[fl].max() < 1.2
    assert abs((rr + rb)[fl].mean() - 1.0) < 1e-3


@pytest.mark.parametrize("sparse", [False, True])
def test_config4_131_against_oracle(cuda, sparse):
    """BASELINE config 4 at full size (131^3 stand-in, README parameters niu_l=0.05, niu_g=0.2,
    CapA=0.005, psi_solid=0.7, constant psi=-1 on x0, force (5e-5,-2e-5,0)): 40 steps against
    the C oracle, production arithmetic 1e-5, verification arithmetic bit-identical."""
    from taichi_lbm3d_b200.geometry import ftb131_standin
    solid = ftb131_standin()
    psi = np.ones(solid.shape, np.float32)
    psi[:13] = -1.0
    case = cases2p.Case2P("cfg4", solid, psi, niu_l=0.05, niu_g=0.2, CapA=0.005, psi_solid=0.7)
    o = _oracle(case, 40)
    fl = solid == 0
    lb = case.make_solver(strict=True, sparse=sparse)
    lb.run(40)
    for n in FIELDS:
        assert np.array_equal(getattr(lb, n).to_numpy()[fl], getattr(o, n)[fl]), n
    lbf = case.make_solver(sparse=sparse)
    lbf.run(40)
    for n in ("F", "rho", "psi", "rho_r", "rho_b"):
        assert rel_linf(getattr(lbf, n).to_numpy()[fl], getattr(o, n)[fl]) <= TOL, n
