#!/usr/bin/env python
"""BASELINE config 5 on ONE GPU: the 1024^3 lid-driven cavity, dense storage stepped in place
(one population buffer, 82 GB).  Prints MLUPS and the fraction of the 152-byte roofline; also
the 256^3 cavity in place vs two buffers."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase  # noqa: E402
from taichi_lbm3d_b200.geometry import cavity  # noqa: E402


def run(n, in_place, steps, warm):
    solid = cavity(n, n, n)
    nfl = int((solid == 0).sum(dtype=np.int64))
    lb = LB3D_Solver_Single_Phase(n, n, n, in_place=in_place)
    lb.solid.from_numpy(solid)
    del solid
    lb.set_bc_vel_x1([0.0, 0.0, 0.1])
    lb.init_simulation()
    lb.run(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lb.run(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    mlups = nfl / (ms * 1e-3) / 1e6
    out = {"workload": "lid-driven cavity %d^3, dense storage, %s" % (n, "in place (AA)" if in_place else "two buffers"),
           "fluid_nodes": nfl, "ms_per_step": ms, "mlups": mlups, "frac_of_measured_roofline": mlups * 152e6 / 1e9 / 6554.6,
           "max_v": lb.get_max_v(), "gpu_mem_GB": torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9}
    print(json.dumps(out), flush=True)
    del lb
    torch.cuda.empty_cache()


if __name__ == "__main__":
    torch.cuda.set_device(0)
    run(256, False, 100, 10)
    run(256, True, 100, 10)
    if "--big" in sys.argv:
        run(1024, True, 20, 4)
