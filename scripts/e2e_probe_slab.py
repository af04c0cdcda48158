#!/usr/bin/env python
"""set-up cost of a slab solver under torchrun, call by call (rank 0 prints); diagnostic for bench.py's e2e at N > 1"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402
from taichi_lbm3d_b200 import _lib  # noqa: E402


class Timed:
    def __init__(self, lib):
        self._lib, self.t = lib, collections.OrderedDict()

    def __getattr__(self, name):
        fn = getattr(self._lib, name)

        def call(*a):
            t0 = time.perf_counter()
            r = fn(*a)
            self.t[name] = self.t.get(name, 0.0) + time.perf_counter() - t0
            return r
        return call


env = bench.Env()
n = 256
real = _lib.load()
from taichi_lbm3d_b200.multi_gpu import SlabPartition  # noqa: E402
part = SlabPartition(n * env.world, env.world, env.rank)
pinned = torch.from_numpy(bench.cavity_planes(n * env.world, n, n, part.local_planes())).pin_memory()
lb = bench.make_cavity_solver(env, n * env.world, n, n)
lb.run(30)                                   # like the bench: a device-timed solver (peer memory) comes first
torch.cuda.synchronize()
bench.release(lb)
for rep in range(5):
    timed = Timed(real)
    _lib._lib = timed
    env.barrier()
    t0 = time.perf_counter()
    lb = bench.make_cavity_solver(env, n * env.world, n, n, pinned=pinned.numpy())
    t1 = time.perf_counter()
    for _ in range(20):
        lb.step()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    bench.release(lb)
    t3 = time.perf_counter()
    _lib._lib = real
    if env.rank == 0:
        print("rep %d init %.4f steps %.4f close %.4f | %s" % (
            rep, t1 - t0, t2 - t1, t3 - t2,
            " ".join("%s=%.4f" % (k.replace("lbm_", ""), v) for k, v in timed.t.items() if v > 5e-4)), flush=True)
if env.dist:
    torch.distributed.destroy_process_group()
