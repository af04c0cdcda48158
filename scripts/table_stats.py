import os, sys
sys.path.insert(0, os.getcwd())
import torch
from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
from taichi_lbm3d_b200.geometry import sphere_pack
n=512
solid = sphere_pack(n,n,n,0.8,8.0,16.0,seed=n,periodic=True)
lb = LB3D_Solver_Single_Phase(n,n,n,sparse_storage=True)
lb.solid.from_numpy(solid); lb.set_force([1e-6,0,0]); lb.init_simulation()
lb.run(10); lb.synchronize()
best=1e9
for rep in range(3):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); lb.run(50); e1.record(); torch.cuda.synchronize()
    best=min(best,e0.elapsed_time(e1)/50)
print(os.environ.get("LBM3D_LIB","default")[-12:], os.environ.get("LBM3D_SPARSE_STREAM","1"), n, best, "ms/step", lb.get_max_v(), flush=True)
