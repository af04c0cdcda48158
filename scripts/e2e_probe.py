import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
torch.cuda.set_device(0)
g = bench.cavity_planes(256,256,256,range(256))
for it in range(3):
    t0=time.perf_counter()
    lb = LB3D_Solver_Single_Phase(256,256,256)
    t1=time.perf_counter()
    lb.solid.from_numpy(g)
    t2=time.perf_counter()
    lb.set_bc_vel_x1([0,0,0.1])
    lb.init_simulation()
    t3=time.perf_counter()
    lb.run(5); lb.synchronize()
    t4=time.perf_counter()
    print("ctor %.4f from_numpy %.4f init %.4f run %.4f"%(t1-t0,t2-t1,t3-t2,t4-t3))
    del lb
    torch.cuda.empty_cache()
