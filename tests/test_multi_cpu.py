"""CPU tests of the multi-GPU host logic: partition arithmetic, and the halo exchange driven
over gloo with world_size 2 and 3 (one process per rank) against the single-domain oracle."""
import os
import socket

import numpy as np
import pytest

from tests import cases


def test_partition_arithmetic():
    from taichi_lbm3d_b200.multi_gpu import SlabPartition
    for gnx, world in ((10, 1), (10, 3), (256, 8), (7, 7)):
        parts = [SlabPartition(gnx, world, r) for r in range(world)]
        assert sum(p.own for p in parts) == gnx
        assert parts[0].x0 == 0 and parts[-1].x1 == gnx
        for r, p in enumerate(parts):
            assert p.local_nx == p.own + 2
            assert p.left == (r - 1) % world and p.right == (r + 1) % world
            assert p.x1 == parts[(r + 1) % world].x0 or r == world - 1
            planes = p.local_planes()
            assert planes[0] == (p.x0 - 1) % gnx and planes[-1] == p.x1 % gnx and planes[1] == p.x0
        assert parts[0].x_face_mask & 1 and parts[-1].x_face_mask & 2
        if world > 2:
            assert parts[1].x_face_mask == 0
    g = np.arange(10 * 2 * 2).reshape(10, 2, 2)
    p = SlabPartition(10, 3, 0)
    loc = p.local_solid(g)
    assert np.array_equal(loc[0], g[9]) and np.array_equal(loc[1], g[0]) and np.array_equal(loc[-1], g[4])
    assert np.array_equal(p.owned(loc), g[0:4])
    with pytest.raises(ValueError):
        SlabPartition(2, 3, 0)


def test_crossing_populations():
    from taichi_lbm3d_b200.constants import CROSS_LEFT, CROSS_RIGHT, E
    assert CROSS_RIGHT == [1, 7, 9, 11, 13] and CROSS_LEFT == [2, 8, 10, 12, 14]
    assert all(E[s, 0] == 1 for s in CROSS_RIGHT) and all(E[s, 0] == -1 for s in CROSS_LEFT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, steps, out_dir):
    import torch.distributed as dist
    from taichi_lbm3d_b200.multi_gpu import HaloExchanger, SlabPartition
    from tests.slab_oracle import OracleSlab
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        case = _case()
        part = SlabPartition(case.shape[0], world, rank)
        slab = OracleSlab(part, case)
        halo = HaloExchanger(part, slab, dist)
        # same schedule as SlabSolver.run without overlap
        slab.begin()
        halo.exchange(0)
        for _ in range(steps - 1):
            slab.stream_bc_macro()
            slab.collide()
            halo.exchange(0)
        slab.stream_bc_macro()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), F=slab.owned("F"), rho=slab.owned("rho"),
                 v=slab.owned("v"), x0=part.x0)
    finally:
        dist.destroy_process_group()


def _case():
    solid = cases.random_porous((11, 6, 7), 0.3, 17)
    return cases.Case("slabs", solid, bc=[(2, "rho", 1.0), (3, "rho", 0.99), (5, "vel", [0.0, 0.01, 0.0])],
                      force=[2e-5, 0.0, -1e-5], perturb=1e-3)


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_over_gloo_matches_single_domain(world, tmp_path):
    import torch.multiprocessing as mp
    from oracle.ref_single_phase import RefSinglePhase
    steps = 6
    port = _free_port()
    mp.spawn(_worker, args=(world, port, steps, str(tmp_path)), nprocs=world, join=True)
    case = _case()
    ref = case.make_oracle(RefSinglePhase)
    for _ in range(steps):
        ref.step()
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    F = np.concatenate([p["F"] for p in parts], axis=0)
    rho = np.concatenate([p["rho"] for p in parts], axis=0)
    v = np.concatenate([p["v"] for p in parts], axis=0)
    fl = case.solid == 0
    # streaming is pure copies and collision is node-local: the decomposition is bit-exact
    assert np.array_equal(F[fl], ref.F[fl])
    assert np.array_equal(rho[fl], ref.rho[fl])
    assert np.array_equal(v[fl], ref.v[fl])


# ---- two-phase: the staged exchange (two exchanges per step) over gloo --------------------------
class _TagBackend:
    """stand-in for one rank's two-phase slab: packs tensors that say who sent what, and keeps
    what it is asked to unpack"""

    def __init__(self, rank):
        import torch
        self.torch, self.rank, self.got = torch, rank, {}
        self.n = [15 * 6, 6]              # stage 0: 15 floats per face node, stage 1: 1

    def pack(self, stage, side):
        return self.torch.full((self.n[stage],), float(self.rank * 100 + stage * 10 + side))

    def unpack(self, stage, side, tensor):
        assert tensor.numel() == self.n[stage]
        vals = set(tensor.tolist())
        assert len(vals) == 1
        self.got[(stage, side)] = vals.pop()

    def recv_buffer(self, stage, side):
        return self.torch.empty(self.n[stage])


def _worker_staged(rank, world, port, out_dir):
    import json
    import torch.distributed as dist
    from taichi_lbm3d_b200.multi_gpu import SlabPartition, StagedHaloExchanger
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        part = SlabPartition(12, world, rank)
        b = _TagBackend(rank)
        halo = StagedHaloExchanger(part, b, dist)
        for stage in (0, 1, 0, 1):        # the schedule of TwoPhaseSlabSolver.run
            halo.exchange(stage)
        with open(os.path.join(out_dir, "staged%d.json" % rank), "w") as fh:
            json.dump({"%d,%d" % k: v for k, v in b.got.items()}, fh)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_staged_exchange_routes_planes_to_the_right_ghosts(world, tmp_path):
    """left ghost <- what the LEFT rank packed for its right side (side 1), right ghost <- what the
    RIGHT rank packed for its left side (side 0), for both stages, also when left == right
    (two ranks) and in the ring of one"""
    import json
    import torch.multiprocessing as mp
    from taichi_lbm3d_b200.multi_gpu import SlabPartition, StagedHaloExchanger
    if world == 1:
        b = _TagBackend(0)
        halo = StagedHaloExchanger(SlabPartition(12, 1, 0), b, None)
        halo.exchange(0)
        halo.exchange(1)
        got = [{"%d,%d" % k: v for k, v in b.got.items()}]
    else:
        mp.spawn(_worker_staged, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
        got = [json.load(open(os.path.join(str(tmp_path), "staged%d.json" % r))) for r in range(world)]
    for r in range(world):
        left, right = (r - 1) % world, (r + 1) % world
        for stage in (0, 1):
            assert got[r]["%d,0" % stage] == left * 100 + stage * 10 + 1
            assert got[r]["%d,1" % stage] == right * 100 + stage * 10 + 0


def _case2p():
    from tests import cases2p
    shape = (11, 6, 7)
    solid = (np.random.default_rng(19).random(shape) < 0.25).astype(np.int8)
    psi = np.ones(shape, np.float32)
    psi[:, :2] = -1.0
    # constant psi on the y-left face, pressure faces in z, x periodic (crosses the cuts)
    return cases2p.Case2P("slabs2p", solid, psi, flow_bc=[(4, 1, 1.0), (5, 1, 0.995)], psi_bc=[(2, -1.0)],
                          force=(3e-5, -1e-5, 0.0))


def _worker_2p(rank, world, port, steps, out_dir):
    import torch.distributed as dist
    from taichi_lbm3d_b200.multi_gpu import SlabPartition, StagedHaloExchanger
    from tests.slab_oracle import OracleSlab2P
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        case = _case2p()
        part = SlabPartition(case.shape[0], world, rank)
        slab = OracleSlab2P(part, case)
        halo = StagedHaloExchanger(part, slab, dist)
        # the schedule of TwoPhaseSlabSolver.run / lbm2p_run_slab
        slab.stage(0)
        halo.exchange(0)
        for _ in range(steps - 1):
            slab.stage(1)
            halo.exchange(1)
            slab.stage(2)
            halo.exchange(0)
        slab.stage(1)                     # close the last step (what the getters do on demand)
        np.savez(os.path.join(out_dir, "tp_rank%d.npz" % rank),
                 **{n: slab.owned(n) for n in ("F", "rho", "v", "psi", "rho_r", "rho_b")})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_phase_staged_exchange_over_gloo_matches_single_domain(world, tmp_path):
    """the two-exchange schedule of the two-phase slabs (psi after the colour stage; populations and
    colour data after the collision), driven by the PRODUCT's partition and exchanger over gloo with
    the NumPy oracle as the stepper: bit-identical to the single-domain oracle"""
    import torch.multiprocessing as mp
    from oracle.ref_two_phase import RefTwoPhase
    steps = 4
    mp.spawn(_worker_2p, args=(world, _free_port(), steps, str(tmp_path)), nprocs=world, join=True)
    case = _case2p()
    ref = case.make_oracle(RefTwoPhase)
    for _ in range(steps):
        ref.step()
    fl = case.solid == 0
    parts = [np.load(os.path.join(str(tmp_path), "tp_rank%d.npz" % r)) for r in range(world)]
    for n in ("F", "rho", "v", "psi", "rho_r", "rho_b"):
        got = np.concatenate([p[n] for p in parts], axis=0)
        assert np.array_equal(got[fl], getattr(ref, n)[fl]), n
