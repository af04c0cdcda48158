"""CPU tests of the oracle (no GPU): the NumPy, pure-Python-loop and C forms agree bit for
bit, reproduce the committed golden vectors, and satisfy the known-answer properties that
follow from the reference source."""
import os

import numpy as np
import pytest

from oracle import ref_single_phase as ref
from oracle.cref import RefSinglePhaseC
from tests import cases
from tests.cases import rel_linf

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_constants_match_reference_tables():
    # opposite directions, weights, M rows (reference :85, :183-197, :64-82)
    assert np.array_equal(ref.E[ref.LR], -ref.E)
    assert abs(ref.W64.sum() - 1.0) < 1e-15
    assert np.array_equal(ref.M_INT[0], np.ones(19, int))
    assert np.array_equal(ref.M_INT[3], ref.E[:, 0]) and np.array_equal(ref.M_INT[5], ref.E[:, 1]) \
        and np.array_equal(ref.M_INT[7], ref.E[:, 2])
    assert np.allclose(ref.inv_M64() @ ref.M_INT, np.eye(19), atol=1e-14)
    S = ref.relaxation_rates(0.16667)
    assert abs(S[1] - 1.0 / (0.16667 / 3 + 0.5)) < 1e-15 and S[0] == S[3] == S[5] == S[7] == 0
    assert abs(ref.relaxation_rates(0.1, "textbook")[1] - 1.0 / 0.8) < 1e-15


def test_meq_is_M_feq():
    """meq_vec(rho,u) == M @ feq(1,u) with m0 replaced by rho."""
    o = ref.RefSinglePhase(1, 1, 1, dtype=np.float64)
    rng = np.random.default_rng(0)
    for _ in range(5):
        u = rng.normal(0, 0.05, 3)
        feq = np.array([o._feq(k, 1.0, u) for k in range(19)])
        m = ref.M_INT @ feq
        meq = o._meq(np.array([1.0]), u[None, :])[0]
        assert np.allclose(m, meq, atol=1e-15)


def test_force_field_generalises_uniform_force():
    """a constant force ARRAY gives bit for bit what set_force gives (cal_local_force :217-220
    returns the same vector at every node), in the NumPy and the C form"""
    base = cases.case_periodic_force()
    ff = np.broadcast_to(np.asarray(base.force, np.float32), base.shape + (3,)).copy()
    arr = cases.Case("ff", base.solid, force_field=ff, perturb=base.perturb)
    for cls in (ref.RefSinglePhase, RefSinglePhaseC):
        a, b = base.make_oracle(cls), arr.make_oracle(cls)
        for _ in range(4):
            a.step()
            b.step()
        assert np.array_equal(a.F, b.F) and np.array_equal(a.v, b.v)


@pytest.mark.parametrize("make", [cases.case_mixed_bc, cases.case_all_faces, cases.case_periodic_force,
                                  cases.case_force_field, cases.case_other_copy])
def test_numpy_and_c_forms_bit_identical(make):
    case = make()
    a = case.make_oracle(ref.RefSinglePhase)
    b = case.make_oracle(RefSinglePhaseC)
    for _ in range(6):
        a.step()
    b.run(6)
    for n in ("f", "F", "rho", "v"):
        assert np.array_equal(getattr(a, n), getattr(b, n)), n
    a64 = case.make_oracle(ref.RefSinglePhase, dtype=np.float64)
    b64 = case.make_oracle(RefSinglePhaseC, dtype=np.float64)
    a64.step()
    b64.run(1)
    assert np.array_equal(a64.F, b64.F) and np.array_equal(a64.v, b64.v)


@pytest.mark.parametrize("name", ["mixed_bc", "all_faces", "periodic_force", "force_field", "other_copy"])
def test_golden_vectors(name):
    g = np.load(os.path.join(GOLD, "sp_%s.npz" % name))
    case = {"mixed_bc": cases.case_mixed_bc, "all_faces": cases.case_all_faces,
            "periodic_force": cases.case_periodic_force, "force_field": cases.case_force_field,
            "other_copy": cases.case_other_copy}[name]()
    assert np.array_equal(case.solid, g["solid"])
    o = case.make_oracle(RefSinglePhaseC)
    assert np.array_equal(o.F, g["F0"]) and np.array_equal(o.v, g["v0"])
    o.run(int(g["steps"]))
    assert np.array_equal(o.F, g["F"]) and np.array_equal(o.rho, g["rho"]) and np.array_equal(o.v, g["v"])


def test_push_equals_pull_and_loop_form():
    case = cases.case_mixed_bc()
    a = case.make_oracle(ref.RefSinglePhase)
    a.colission()
    b = case.make_oracle(ref.RefSinglePhase)
    b.colission()
    a.streaming1()
    b.streaming1_loops()
    assert np.array_equal(a.F, b.F)
    fl = a.solid == 0
    assert np.array_equal(ref.pull_stream(a.f, a.solid)[fl], a.F[fl])


def test_rest_state_fixed_point_and_mass_conservation():
    solid = cases.random_porous((10, 9, 8), 0.3, 5)
    o = cases.Case("rest", solid).make_oracle(RefSinglePhaseC)
    o.run(30)
    fl = solid == 0
    assert np.abs(o.F[fl] - o.w).max() < 5e-7 and np.abs(o.v).max() < 1e-6
    p = cases.Case("mass", solid, perturb=1e-3).make_oracle(RefSinglePhaseC)
    m0 = p.F[fl].astype(np.float64).sum()
    p.run(50)
    assert abs(p.F[fl].astype(np.float64).sum() - m0) / m0 < 2e-6


def test_pressure_faces_keep_zero_velocity():
    solid = cases.random_porous((10, 9, 8), 0.3, 6)
    o = cases.Case("p", solid, bc=[(0, "rho", 1.0), (1, "rho", 0.99)]).make_oracle(RefSinglePhaseC)
    o.run(30)
    fl = solid == 0
    assert np.all(o.v[0][fl[0]] == 0) and np.all(o.v[-1][fl[-1]] == 0)
    assert np.abs(o.rho[-1][fl[-1]] - 0.99).max() < 1e-6


def test_poiseuille_known_answer():
    """Plates with half-way walls, body force fy: the steady momentum profile is the
    parabola g/(2 nu) (z-z0)(z1-z) with the reference's effective force g = fy/9 (Guo term
    /9, :236) and nu = (tau-1/2)/3 = niu/9 (tau = niu/3+1/2, :127).  Momentum rather than v
    because the s=0 Guo moment is a small mass sink."""
    g = np.zeros((3, 4, 8), np.int8)
    g[:, :, 0] = 1
    g[:, :, -1] = 1
    fy, niu = 1e-4, 0.1667
    o = cases.Case("p", g, force=[0.0, fy, 0.0], niu=niu).make_oracle(RefSinglePhaseC, dtype=np.float64)
    o.run(12000)
    z = np.arange(1, 7)
    ana = (fy / 9.0) / (2 * niu / 9.0) * (z - 0.5) * (6.5 - z)
    j = o.rho[1, 2, 1:7] * o.v[1, 2, 1:7, 1] - fy / 2 + (fy / 9.0) / 2
    assert rel_linf(j, ana) < 1e-5


def test_poiseuille_known_answer_other_copy():
    """same plates with the physics of the solver's other copy (tau = 3 niu + 1/2 :126 of
    Phase_change/LBM_3D_SinglePhase_Solver.py, Guo term not divided :235): effective force
    g = fy/3 (the 1/cs^2 factor of Guo's scheme is still missing), nu = niu."""
    g = np.zeros((3, 4, 8), np.int8)
    g[:, :, 0] = 1
    g[:, :, -1] = 1
    fy, niu = 1e-5, 0.1667
    o = cases.Case("p", g, force=[0.0, fy, 0.0], niu=niu, tau_mode="textbook",
                   guo_mode="unscaled").make_oracle(RefSinglePhaseC, dtype=np.float64)
    o.run(4000)
    z = np.arange(1, 7)
    ana = (fy / 3.0) / (2 * niu) * (z - 0.5) * (6.5 - z)
    j = o.rho[1, 2, 1:7] * o.v[1, 2, 1:7, 1] - fy / 2 + (fy / 3.0) / 2
    assert rel_linf(j, ana) < 1e-5


def test_cavity_fixture_and_symmetry():
    from taichi_lbm3d_b200.geometry import cavity
    g = np.load(os.path.join(GOLD, "geo_cavity_50.npz"))
    fixture = np.unpackbits(g["packed"])[:125000].reshape(50, 50, 50)
    assert np.array_equal(fixture, cavity(50, 50, 50))       # generator == reference's geo_cavity.dat
    assert int(fixture.sum()) == 12104
    case = cases.case_cavity(20)
    o = case.make_oracle(RefSinglePhaseC)
    o.run(100)
    assert rel_linf(o.v[:, ::-1, :, 2], o.v[..., 2]) < 1e-5
    assert abs(o.get_max_v() - 0.1) < 1e-6


def test_fp32_roundoff_yardstick():
    """documents the fp32 noise floor the parity tolerance on v is built on"""
    case = cases.case_periodic_force()
    a = case.make_oracle(RefSinglePhaseC)
    b = case.make_oracle(RefSinglePhaseC, dtype=np.float64)
    c = case.make_oracle(RefSinglePhaseC, kind="fast")       # same algorithm, -ffast-math
    for o in (a, b, c):
        o.run(100)
    fl = a.solid == 0
    assert rel_linf(a.F[fl], b.F[fl]) < 1e-5 and rel_linf(c.F[fl], a.F[fl]) < 1e-5
    assert np.abs(c.v[fl] - a.v[fl]).max() <= cases.v_abs_tolerance(a, b)
