"""GPU tests of the x-slab decomposition: P slabs must be BIT-IDENTICAL to the
single-domain run, because streaming is pure copies and collision is node-local.  P > 1 ranks
are emulated in one process on one GPU (one halo_x context per slab, lockstep exchange);
the real multi-process path (NCCL) is exercised by scripts/multi_gpu_check.py under torchrun."""
import ctypes

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _single(case, steps, sparse, strict):
    lb = case.make_solver(sparse=sparse, strict=strict)
    lb.run(steps)
    return lb.F.to_numpy(), lb.rho.to_numpy(), lb.v.to_numpy(), lb.get_max_v()


def _configure(slab, case):
    for face, kind, val in case.bc:
        getattr(slab, (cases.FACE_SETTERS_RHO if kind == "rho" else cases.FACE_SETTERS_VEL)[face])(val)
    if case.force is not None:
        slab.set_force(case.force)
    if case.niu is not None:
        slab.set_viscosity(case.niu)


@pytest.mark.parametrize("sparse", [False, True])
@pytest.mark.parametrize("overlap", [False, True])
@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("transport", ["native", "torch"])
def test_single_slab_ring_equals_plain_solver(cuda, sparse, overlap, strict, transport):
    """world = 1: the slab's ghost planes are fed by its own opposite faces (periodic ring);
    x-face BCs live on the first / last owned plane."""
    from taichi_lbm3d_b200.multi_gpu import SlabSolver
    for case in (cases.case_mixed_bc((12, 10, 9)), cases.case_periodic_force((11, 7, 13))):
        case.perturb = 0.0
        steps = 9
        F, rho, v, mv = _single(case, steps, sparse, strict)
        ss = SlabSolver(*case.shape, sparse_storage=sparse, strict=strict, overlap=overlap, transport=transport)
        ss.set_solid(case.solid)
        _configure(ss, case)
        ss.init_simulation()
        ss.run(4)
        ss.local_field("rho")            # mid-run read must not disturb the pipeline
        ss.run(steps - 4)
        fl = case.solid == 0
        assert np.array_equal(ss.local_field("F")[fl], F[fl])
        assert np.array_equal(ss.local_field("rho")[fl], rho[fl])
        assert np.array_equal(ss.local_field("v")[fl], v[fl])
        assert abs(ss.get_max_v() - mv) < 1e-7


@pytest.mark.parametrize("sparse", [False, True])
@pytest.mark.parametrize("world", [2, 3])
def test_emulated_ranks_equal_plain_solver(cuda, sparse, world):
    """P halo_x contexts on one GPU exchanging through the same pack/unpack entry points the
    NCCL path uses, stepped in lockstep with the overlapped plane schedule."""
    import torch
    from taichi_lbm3d_b200.multi_gpu import SlabPartition, _CudaBackend, _LocalSlab
    case = cases.case_mixed_bc((13, 8, 9))
    case.perturb = 0.0
    steps = 7
    F, rho, v, _ = _single(case, steps, sparse, False)
    parts = [SlabPartition(case.shape[0], world, r) for r in range(world)]
    slabs = []
    for p in parts:
        s = _LocalSlab(p, case.shape[1], case.shape[2], sparse_storage=sparse)
        s.solid.from_numpy(p.local_solid(case.solid))
        _configure(s, case)
        s.init_simulation()
        slabs.append(s)
    backs = [_CudaBackend(s) for s in slabs]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def exchange(which):
        packed = [(b.pack(0, which).clone(), b.pack(1, which).clone()) for b in backs]
        for r, b in enumerate(backs):
            b.unpack(0, packed[parts[r].left][1], which)     # left ghost <- left rank's right face
            b.unpack(1, packed[parts[r].right][0], which)    # right ghost <- right rank's left face

    for s in slabs:
        s._ck(s._lib.lbm_step_begin(s._ctx, st), "begin")
    exchange(0)
    for _ in range(steps - 1):
        for s, p in zip(slabs, parts):
            lib, ctx = s._lib, s._ctx
            s._ck(lib.lbm_step_planes(ctx, 1, 2, st), "planes")
            if p.own > 1:
                s._ck(lib.lbm_step_planes(ctx, p.own, p.own + 1, st), "planes")
        exchange(1)
        for s, p in zip(slabs, parts):
            if p.own > 2:
                s._ck(s._lib.lbm_step_planes(s._ctx, 2, p.own, st), "planes")
            s._ck(s._lib.lbm_step_flip(s._ctx), "flip")
    Fs = np.concatenate([p.owned(s.F.to_numpy()) for s, p in zip(slabs, parts)], axis=0)
    rs = np.concatenate([p.owned(s.rho.to_numpy()) for s, p in zip(slabs, parts)], axis=0)
    vs = np.concatenate([p.owned(s.v.to_numpy()) for s, p in zip(slabs, parts)], axis=0)
    fl = case.solid == 0
    assert np.array_equal(Fs[fl], F[fl]) and np.array_equal(rs[fl], rho[fl]) and np.array_equal(vs[fl], v[fl])
