#!/usr/bin/env python
"""Multi-process parity check of the two-phase x-slabs (run under torchrun on N GPUs): the
TwoPhaseSlabSolver over NCCL must be bit-identical (verification arithmetic) to the single-GPU
solver on the same problem; also times the production arithmetic.  Rank 0 prints."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase  # noqa: E402
from taichi_lbm3d_b200.geometry import sphere_pack  # noqa: E402
from taichi_lbm3d_b200.multi_gpu import TwoPhaseSlabSolver  # noqa: E402

FIELDS = ("rho", "v", "psi", "rho_r", "rho_b")


def configure(lb):
    lb.niu_l, lb.niu_g, lb.CapA, lb.psi_solid = 0.05, 0.2, 0.005, 0.7


def main():
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = dist.get_world_size()
    ok = True
    solid = sphere_pack(96, 64, 64, 0.7, 3.0, 6.0, seed=3, periodic=False)
    psi = np.ones(solid.shape, np.float32)
    psi[:20] = -1.0
    for transport in ("native", "torch"):
        ss = TwoPhaseSlabSolver(*solid.shape, strict=True, transport=transport)
        ss.set_fields(solid, psi)
        configure(ss.local)
        ss.init_simulation()
        ss.run(30)
        got = {n: ss.gather_field(n) for n in FIELDS}
        if rank == 0:
            ref = LB3D_Solver_Two_Phase(*solid.shape, strict=True)
            ref.solid.from_numpy(solid)
            ref.psi.from_numpy(psi)
            configure(ref)
            ref.init_simulation()
            ref.run(30)
            fl = solid == 0
            same = all(np.array_equal(got[n][fl], getattr(ref, n).to_numpy()[fl]) for n in FIELDS)
            ok &= same
            print("multi_gpu_check_2p world=%d porous 96x64x64 transport=%-6s bit_identical=%s" % (world, transport, same),
                  flush=True)
        dist.barrier()
    # throughput of the production arithmetic, weak scaling: 128 planes of 256^2 per GPU, droplet row
    n = 256
    gnx = 128 * world
    x, y, z = np.meshgrid(np.arange(gnx), np.arange(n), np.arange(n), indexing='ij')
    r = np.sqrt(((x % 128) - 64) ** 2 + (y - n / 2) ** 2 + (z - n / 2) ** 2)
    psi = np.where(r < 40, -1.0, 1.0).astype(np.float32)
    ss = TwoPhaseSlabSolver(gnx, n, n)
    ss.set_fields(np.zeros((gnx, n, n), np.int8), psi)
    configure(ss.local)
    ss.local.bc_psi_x_left = 0
    ss.init_simulation()
    ss.run(10)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ss.run(50)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item()) / 50
        print("multi_gpu_check_2p world=%d droplets %dx%dx%d: %.4f ms/step, %.0f MLUPS"
              % (world, gnx, n, n, ms, gnx * n * n / (ms * 1e-3) / 1e6), flush=True)
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
