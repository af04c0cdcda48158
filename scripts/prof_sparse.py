#!/usr/bin/env python
"""A few steps of BASELINE config 3 (512^3 sphere pack, sparse storage) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
from taichi_lbm3d_b200.geometry import sphere_pack
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
solid = sphere_pack(n, n, n, 0.80, 8.0 * n / 512, 16.0 * n / 512, seed=n, periodic=True)
lb = LB3D_Solver_Single_Phase(n, n, n, sparse_storage=True, in_place=len(sys.argv) > 2 and sys.argv[2] == "aa")
lb.solid.from_numpy(solid)
lb.set_force([1e-6, 0.0, 0.0])
lb.init_simulation()
lb.run(12)
lb.synchronize()
