#!/usr/bin/env python
"""Throughput of the two-phase colour-gradient step (BASELINE config 4 shape at 131^3 and a
dense 256^3 droplet box).  Reports MLUPS and the fraction of the HBM roofline at the declared
176 algorithmic bytes per fluid-node update."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase  # noqa: E402
from taichi_lbm3d_b200.geometry import ftb131_standin  # noqa: E402

B2P = 176.0


def run(name, solid, psi, steps=200, warmup=20, sparse=False):
    lb = LB3D_Solver_Two_Phase(*solid.shape, sparse_storage=sparse)
    lb.solid.from_numpy(solid)
    lb.psi.from_numpy(psi)
    lb.niu_l, lb.niu_g, lb.CapA, lb.psi_solid = 0.05, 0.2, 0.005, 0.7
    lb.init_simulation()
    lb.run(warmup)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lb.launch_count
    e0.record()
    lb.run(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nfl = int((solid == 0).sum())
    mlups = nfl * steps / (ms * 1e-3) / 1e6
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
    ach = B2P * nfl * steps / (ms * 1e-3) / 1e9
    return {"workload": name, "fluid_nodes": nfl, "mlups": mlups, "ms_per_step": ms / steps,
            "launches_per_step": (lb.launch_count - l0) / steps,
            "roofline": {"bytes_per_update": B2P, "achieved_GBps": ach, "peak_GBps": peak, "frac": ach / peak},
            "psi_range": [float(lb.psi.to_numpy()[solid == 0].min()), float(lb.psi.to_numpy()[solid == 0].max())]}


def main():
    torch.cuda.set_device(0)
    out = []
    solid = ftb131_standin()
    psi = np.ones(solid.shape, np.float32)
    psi[:13] = -1.0
    out.append(run("drainage 131^3 sphere-pack stand-in (config 4 parameters), dense storage", solid, psi))
    out.append(run("drainage 131^3 sphere-pack stand-in (config 4 parameters), sparse storage", solid, psi, sparse=True))
    from taichi_lbm3d_b200.geometry import sphere_pack
    n = 384
    solid = sphere_pack(n, n, n, 0.80, 6.0, 12.0, seed=n, periodic=True)
    psi = np.ones(solid.shape, np.float32)
    psi[:n // 4] = -1.0
    out.append(run("drainage %d^3 periodic sphere pack (porosity 0.2), sparse storage" % n, solid, psi,
                   steps=100, warmup=10, sparse=True))
    n = 256
    x, y, z = np.meshgrid(*[np.arange(n)] * 3, indexing='ij')
    r = np.sqrt((x - n / 2) ** 2 + (y - n / 2) ** 2 + (z - n / 2) ** 2)
    out.append(run("droplet in a periodic 256^3 box, dense", np.zeros((n, n, n), np.int8),
                   np.where(r < n / 4, -1.0, 1.0).astype(np.float32), steps=100, warmup=10))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
