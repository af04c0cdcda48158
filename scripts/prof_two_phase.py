#!/usr/bin/env python
"""A few steps of a two-phase case for ncu: droplet256 | porous384 | cfg4 | cfg4s (sparse)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase  # noqa: E402
from taichi_lbm3d_b200.geometry import ftb131_standin, sphere_pack  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "droplet256"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.cuda.set_device(0)
sparse = False
if case == "droplet256":
    n = 256
    x, y, z = np.meshgrid(*[np.arange(n, dtype=np.float32)] * 3, indexing='ij', sparse=True)
    r2 = (x - n / 2) ** 2 + (y - n / 2) ** 2 + (z - n / 2) ** 2
    solid = np.zeros((n, n, n), np.int8)
    psi = np.where(r2 < (n / 4) ** 2, -1.0, 1.0).astype(np.float32)
elif case == "porous384":
    n = 384
    solid = sphere_pack(n, n, n, 0.80, 6.0, 12.0, seed=n, periodic=True)
    psi = np.ones(solid.shape, np.float32)
    psi[:n // 4] = -1.0
    sparse = True
else:
    solid = ftb131_standin()
    psi = np.ones(solid.shape, np.float32)
    psi[:13] = -1.0
    sparse = case == "cfg4s"
lb = LB3D_Solver_Two_Phase(*solid.shape, sparse_storage=sparse)
lb.solid.from_numpy(solid)
lb.psi.from_numpy(psi)
lb.niu_l, lb.niu_g, lb.CapA, lb.psi_solid = 0.05, 0.2, 0.005, 0.7
lb.init_simulation()
lb.run(steps)
lb.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lb.run(steps)
e1.record()
torch.cuda.synchronize()
print(case, "ms/step", e0.elapsed_time(e1) / steps)
