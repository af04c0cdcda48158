#!/usr/bin/env python
"""where does the set-up part of bench.py's end-to-end job go?  Times every C-ABI call of
init_simulation() / close() for a few repetitions of the 256^3 cavity (diagnostic)."""
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
pinned = torch.from_numpy(bench.cavity_planes(n, n, n, range(n))).pin_memory()
for rep in range(6):
    timed = Timed(real)
    _lib._lib = timed
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lb = bench.make_cavity_solver(env, n, n, n, pinned=pinned.numpy())
    t1 = time.perf_counter()
    for _ in range(20):
        lb.step()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    lb.close()
    t3 = time.perf_counter()
    _lib._lib = real
    print("rep %d init %.4f steps %.4f close %.4f | %s" % (
        rep, t1 - t0, t2 - t1, t3 - t2,
        " ".join("%s=%.4f" % (k.replace("lbm_", ""), v) for k, v in timed.t.items() if v > 2e-4)), flush=True)

import cProfile  # noqa: E402
import pstats  # noqa: E402
pr = cProfile.Profile()
pr.enable()
lb = bench.make_cavity_solver(env, n, n, n, pinned=pinned.numpy())
pr.disable()
lb.close()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
