#!/usr/bin/env python
"""Event timeline of the overlapped slab step (run under torchrun on N GPUs):

    LBM3D_TIMELINE=gpurun_out/timeline python -m torch.distributed.run ... scripts/timeline_slab.py [planes_per_gpu]

Every rank writes <file>.rank<r>: when, on its two streams, the interior kernel, the boundary
kernel and the pack / ncclSend+Recv / unpack of the halo exchange began and ended during the last
four steps (CUDA events, microseconds).  The headline cavity slabs of bench.py."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

env = bench.Env()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ny = nz = int(sys.argv[2]) if len(sys.argv) > 2 else 256
lb = bench.make_cavity_solver(env, n * env.world, ny, nz)
tl = os.environ.pop("LBM3D_TIMELINE", None)
lb.run(50)                      # warm-up without recording
env.barrier()
if tl:
    os.environ["LBM3D_TIMELINE"] = tl
lb.run(20)
env.barrier()
if env.rank == 0:
    print("timeline written to %s.rank* (peer_memory=%s)" % (tl, getattr(lb, "peer_memory", False)))
bench.release(lb)
if env.dist:
    dist.destroy_process_group()
