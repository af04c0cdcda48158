"""CPU tests of the two-phase oracle: NumPy and C forms bit-identical, the pull colour
accumulation equals the reference's literal push to round-off, colour mass is conserved."""
import numpy as np
import pytest

from oracle.cref import RefTwoPhaseC
from oracle.ref_two_phase import RefTwoPhase
from tests import cases2p


@pytest.mark.parametrize("make", [cases2p.case_drainage, cases2p.case_bcs, cases2p.case_periodic_bubble])
def test_numpy_and_c_forms_bit_identical(make):
    case = make()
    a = case.make_oracle(RefTwoPhase)
    b = case.make_oracle(RefTwoPhaseC)
    for _ in range(6):
        a.step()
    b.run(6)
    for n in ("F", "rho", "v", "psi", "rho_r", "rho_b"):
        assert np.array_equal(getattr(a, n), getattr(b, n)), n
    assert np.isfinite(a.F).all()


@pytest.mark.parametrize("make", [cases2p.case_drainage, cases2p.case_bcs, cases2p.case_periodic_bubble])
def test_golden_vectors(make):
    """committed fixtures (tests/golden/make_golden.py, NumPy oracle) pin both oracle forms"""
    import os
    case = make()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "tp_%s.npz" % case.name))
    assert np.array_equal(case.solid, g["solid"]) and np.array_equal(case.psi, g["psi0"])
    o = case.make_oracle(RefTwoPhaseC)
    o.run(int(g["steps"]))
    for n in ("F", "rho", "v", "psi", "rho_r", "rho_b"):
        assert np.array_equal(getattr(o, n), g[n]), n


def test_pull_accumulation_matches_literal_push():
    case = cases2p.case_drainage((8, 7, 6))
    o = case.make_oracle(RefTwoPhase)
    for _ in range(3):
        o.step()
    o.colission()
    rr, rb = o.push_colour_loops()            # 2phase/lbm_solver_3d_2phase.py:365-372 as written
    fl = o.solid == 0
    assert np.abs(rr[fl] - o.rhor[fl]).max() < 5e-7 and np.abs(rb[fl] - o.rhob[fl]).max() < 5e-7


def test_colour_mass_and_interface():
    case = cases2p.case_periodic_bubble()
    o = case.make_oracle(RefTwoPhaseC)
    fl = o.solid == 0
    m_r0 = o.rho_r[fl].astype(np.float64).sum()
    o.run(100)
    # recolouring only moves colour between the two opposite directions of a pair: each colour is conserved
    assert abs(o.rho_r[fl].astype(np.float64).sum() - m_r0) / m_r0 < 1e-5
    assert o.psi[fl].min() < -0.9 and o.psi[fl].max() > 0.9          # the droplet is still there
    # Laplace: pressure (rho/3) is higher inside the droplet
    assert o.rho[8, 8, 8] > o.rho[0, 0, 0]


def test_psi_clamp_and_kill_rule():
    case = cases2p.case_bcs()
    o = case.make_oracle(RefTwoPhase)
    C = o.Compute_C()
    # nodes whose stencil touches a solid and that are pure phase (|rho_r-rho_b| > 0.9) have C = 0 (:271-273)
    from oracle.ref_single_phase import E
    near = np.zeros(o.solid.shape, bool)
    for s in range(19):
        near |= o._psi_shift(o.solid, s) != 0
    kill = near & (np.abs(o.rho_r - o.rho_b) > 0.9) & (o.solid == 0)
    assert kill.any() and np.all(C[kill] == 0)
