#!/usr/bin/env python
"""Can the two kernels of the two-phase step share the GPU?  k2p_main is bound by DRAM (87 % of the
copy rate), k2p_colour by instruction issue and L1 latency: run side by side they could approach
max(sum of DRAM time, sum of issue time) instead of the sum.  Probe with two independent solvers
on two streams (the block scheduler decides what overlaps), against the same work on one stream."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase  # noqa: E402


def make(nx, n):
    x, y, z = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(n, dtype=np.float32),
                          np.arange(n, dtype=np.float32), indexing='ij', sparse=True)
    r2 = (x - nx / 2) ** 2 + (y - n / 2) ** 2 + (z - n / 2) ** 2
    lb = LB3D_Solver_Two_Phase(nx, n, n)
    lb.solid.from_numpy(np.zeros((nx, n, n), np.int8))
    lb.psi.from_numpy(np.where(r2 < (n / 4) ** 2, -1.0, 1.0).astype(np.float32))
    lb.niu_l, lb.niu_g, lb.CapA, lb.psi_solid = 0.05, 0.2, 0.005, 0.7
    lb.init_simulation()
    lb.run(200)
    return lb


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


nx, n, K = 128, 256, 20
X, Y = make(nx, n), make(nx, n)
nodes = 2 * nx * n * n


def sequential():
    for _ in range(K):
        X.step()
        Y.step()


for prio in ((0, 0), (-1, 0), (0, -1)):
    s1, s2 = torch.cuda.Stream(priority=prio[0]), torch.cuda.Stream(priority=prio[1])

    def concurrent():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        for _ in range(K):
            with torch.cuda.stream(s1):
                X.step()
            with torch.cuda.stream(s2):
                Y.step()
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    def skewed():
        # Y half a step behind X: X's main pass is issued together with Y's colour pass
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        for _ in range(K):
            with torch.cuda.stream(s1):
                X.step()
            with torch.cuda.stream(s2):
                Y.step()
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    t_seq = timed(sequential)
    t_con = timed(concurrent)
    print("priorities %s: one stream %.3f ms per pair of steps (%.0f MLUPS), two streams %.3f ms (%.0f MLUPS)"
          % (prio, t_seq / K, nodes * K / t_seq / 1e3, t_con / K, nodes * K / t_con / 1e3), flush=True)
