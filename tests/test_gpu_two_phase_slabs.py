"""GPU tests of the x-slab decomposition of the two-phase solver: the colour pass
reads the neighbours' colour records and the collision their psi, so a step has two halo
exchanges; P slabs must stay BIT-IDENTICAL to the single-domain run in verification mode
(every node evaluates the same expression on the same inputs).  P > 1 ranks are emulated in
one process on one GPU through the pack / unpack / stage entry points the NCCL loop uses; the
multi-process path itself is exercised by scripts/multi_gpu_check_2p.py under torchrun."""
import ctypes

import numpy as np
import pytest

from tests import cases2p

pytestmark = pytest.mark.gpu

FIELDS = ("F", "rho", "v", "psi", "rho_r", "rho_b")
_FACE = ("x_left", "x_right", "y_left", "y_right", "z_left", "z_right")
_SFX = ("xl", "xr", "yl", "yr", "zl", "zr")
CASES = [cases2p.case_drainage, cases2p.case_bcs, cases2p.case_periodic_bubble]


def _configure(lb, case):
    lb.set_force(case.force)
    lb.niu_l, lb.niu_g, lb.CapA, lb.psi_solid = case.niu_l, case.niu_g, case.CapA, case.psi_solid
    for face in range(6):
        setattr(lb, "bc_" + _FACE[face], 0)
        setattr(lb, "bc_psi_" + _FACE[face], 0)
    for face, t, rho in case.flow_bc:
        setattr(lb, "bc_" + _FACE[face], t)
        setattr(lb, "rho_bc" + _SFX[face], rho)
    for face, val in case.psi_bc:
        lb.set_bc_psi(face, val)


def _single(case, steps, strict, sparse=False):
    lb = case.make_solver(strict=strict, sparse=sparse)
    lb.run(steps)
    return {n: getattr(lb, n).to_numpy() for n in FIELDS}


@pytest.mark.parametrize("make", CASES)
@pytest.mark.parametrize("transport", ["native", "torch"])
@pytest.mark.parametrize("sparse", [False, True])
def test_single_slab_ring_equals_plain_solver(cuda, make, transport, sparse):
    """world = 1: the ghost planes are fed by the slab's own opposite faces (periodic ring);
    x-face flow and psi BCs live on the first / last owned plane."""
    from taichi_lbm3d_b200.multi_gpu import TwoPhaseSlabSolver
    case = make()
    steps = 9
    want = _single(case, steps, True)
    ss = TwoPhaseSlabSolver(*case.shape, strict=True, transport=transport, sparse_storage=sparse)
    ss.set_fields(case.solid, case.psi)
    _configure(ss.local, case)
    ss.init_simulation()
    ss.run(4)
    ss.local_field("psi")                # mid-run read must not disturb the pipeline
    ss.run(steps - 4)
    fl = case.solid == 0
    for n in FIELDS:
        assert np.array_equal(ss.local_field(n)[fl], want[n][fl]), n


@pytest.mark.parametrize("make", CASES)
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("sparse", [False, True])
def test_emulated_ranks_equal_plain_solver(cuda, make, world, strict, sparse):
    import torch
    from taichi_lbm3d_b200.multi_gpu import SlabPartition, _two_phase_slab_class
    case = make()
    steps = 7
    want = _single(case, steps, strict, sparse)       # sparse slabs against the sparse single-domain solver
    parts = [SlabPartition(case.shape[0], world, r) for r in range(world)]
    Slab = _two_phase_slab_class()
    slabs = []
    for p in parts:
        s = Slab(p, case.shape[1], case.shape[2], strict=strict, sparse_storage=sparse)
        s.solid.from_numpy(p.local_solid(case.solid))
        s.psi.from_numpy(np.ascontiguousarray(np.take(case.psi, p.local_planes(), axis=0)))
        _configure(s, case)
        s.init_simulation()
        slabs.append(s)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib = slabs[0]._lib
    n0 = max(int(lib.lbm2p_halo_floats(s._ctx, 0)) for s in slabs)      # sparse slabs: planes differ in size

    def exchange(stage):
        packed = []
        for s in slabs:
            bufs = [torch.empty(n0, dtype=torch.float32, device="cuda") for _ in range(2)]
            for side in (0, 1):
                s._ck(lib.lbm2p_halo_pack(s._ctx, stage, side, ctypes.c_void_p(bufs[side].data_ptr()), st), "pack")
            packed.append(bufs)
        for r, s in enumerate(slabs):
            s._ck(lib.lbm2p_halo_unpack(s._ctx, stage, 0, ctypes.c_void_p(packed[parts[r].left][1].data_ptr()), st), "unpack")
            s._ck(lib.lbm2p_halo_unpack(s._ctx, stage, 1, ctypes.c_void_p(packed[parts[r].right][0].data_ptr()), st), "unpack")

    def stage(k):
        for s in slabs:
            s._ck(lib.lbm2p_slab_stage(s._ctx, k, st), "stage")

    stage(0)
    exchange(0)
    for _ in range(steps - 1):
        stage(1)
        exchange(1)
        stage(2)
        exchange(0)
    fl = case.solid == 0
    for n in FIELDS:
        got = np.concatenate([p.owned(getattr(s, n).to_numpy()) for s, p in zip(slabs, parts)], axis=0)
        assert np.array_equal(got[fl], want[n][fl]), n


def test_plain_step_refused_on_a_slab(cuda):
    from taichi_lbm3d_b200 import _lib
    from taichi_lbm3d_b200.multi_gpu import TwoPhaseSlabSolver
    case = cases2p.case_drainage()
    ss = TwoPhaseSlabSolver(*case.shape)
    ss.set_fields(case.solid, case.psi)
    _configure(ss.local, case)
    ss.init_simulation()
    with pytest.raises(_lib.LbmError):
        ss.local.step()
