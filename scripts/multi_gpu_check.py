#!/usr/bin/env python
"""Multi-process parity check (run under torchrun on N GPUs): the SlabSolver over NCCL must be
bit-identical to the single-GPU solver on the same problem.  Prints one line per rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase  # noqa: E402
from taichi_lbm3d_b200.geometry import cavity, sphere_pack  # noqa: E402
from taichi_lbm3d_b200.multi_gpu import SlabSolver  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = dist.get_world_size()
    ok = True
    for name, solid, setup, sparse in (
            ("cavity96 dense", cavity(96, 64, 64), lambda s: s.set_bc_vel_x1([0.0, 0.0, 0.1]), False),
            ("porous96 sparse, pressure x", sphere_pack(96, 64, 64, 0.7, 3.0, 6.0, seed=3, periodic=False),
             lambda s: (s.set_bc_rho_x0(1.0), s.set_bc_rho_x1(0.99), s.set_force([1e-6, 0, 0])), True)):
        # the peer-memory case twice: the second solver gets the first one's buffers back from the library's
        # cache and exports them to its neighbours again
        for overlap, transport in ((False, "native"), (True, "native"), (True, "native"), (True, "torch")):
            ss = SlabSolver(*solid.shape, sparse_storage=sparse, overlap=overlap, transport=transport)
            ss.set_solid(solid)
            setup(ss)
            ss.init_simulation()
            ss.run(50)
            rho = ss.gather_field("rho")
            v = ss.gather_field("v")
            mv = ss.get_max_v()
            if rank == 0:
                ref = LB3D_Solver_Single_Phase(*solid.shape, sparse_storage=sparse)
                ref.solid.from_numpy(solid)
                setup(ref)
                ref.init_simulation()
                ref.run(50)
                fl = solid == 0
                same = np.array_equal(rho[fl], ref.rho.to_numpy()[fl]) and np.array_equal(v[fl], ref.v.to_numpy()[fl])
                ok &= same
                print("multi_gpu_check world=%d %-28s overlap=%d transport=%-6s peer_memory=%d bit_identical=%s max_v=%.6g (ref %.6g)"
                      % (world, name, overlap, transport, ss.peer_memory, same, mv, ref.get_max_v()), flush=True)
            ss.close()
            dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
