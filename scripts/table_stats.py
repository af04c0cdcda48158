import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
from taichi_lbm3d_b200.geometry import sphere_pack
os.environ["LBM3D_DEBUG"]="1"
for n,r0 in ((256,4.0),(512,8.0)):
    solid = sphere_pack(n,n,n,0.8,r0,2*r0,seed=n,periodic=True)
    lb = LB3D_Solver_Single_Phase(n,n,n,sparse_storage=True)
    lb.solid.from_numpy(solid); lb.set_force([1e-6,0,0]); lb.init_simulation()
    lb.run(10); lb.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); lb.run(50); e1.record(); torch.cuda.synchronize()
    print(n, e0.elapsed_time(e1)/50, "ms/step", flush=True)
    del lb
