#!/usr/bin/env python
"""Probe of the overlapped two-phase slab schedule (LBM3D_2P_OVERLAP=1) under torchrun:
argv = planes_per_rank ny steps.  Prints ms/step on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_lbm3d_b200.multi_gpu import TwoPhaseSlabSolver  # noqa: E402

planes, n, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
world, rank = dist.get_world_size(), dist.get_rank()
gnx = planes * world
x, y, z = np.meshgrid(np.arange(gnx), np.arange(n), np.arange(n), indexing='ij')
r = np.sqrt(((x % planes) - planes / 2) ** 2 + (y - n / 2) ** 2 + (z - n / 2) ** 2)
psi = np.where(r < planes / 3, -1.0, 1.0).astype(np.float32)
ss = TwoPhaseSlabSolver(gnx, n, n)
ss.set_fields(np.zeros((gnx, n, n), np.int8), psi)
ss.local.niu_l, ss.local.niu_g, ss.local.CapA, ss.local.psi_solid = 0.05, 0.2, 0.005, 0.7
ss.local.bc_psi_x_left = 0
ss.init_simulation()
ss.run(4)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ss.run(steps)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    ms = e0.elapsed_time(e1) / steps
    print("probe world=%d %dx%dx%d overlap=%s sync_every=%s: %.4f ms/step %.0f MLUPS" % (
        world, gnx, n, n, os.environ.get("LBM3D_2P_OVERLAP", "0"), os.environ.get("LBM3D_2P_SYNC_EVERY", "0"),
        ms, gnx * n * n / (ms * 1e-3) / 1e6), flush=True)
dist.destroy_process_group()
