#!/usr/bin/env python
"""A/B timing of sparse-kernel build variants in ONE process (geometry generated once).

    python scripts/sparse_ab.py SIZE name[:lib_tag][:ENV=VAL,...] ...

Each variant loads lib/liblbm3d_b200.<lib_tag>.so (default build when the tag is empty) into
the same process (ctypes handles are independent), steps the periodic sphere pack of
bench.py --workload porous --size SIZE, and prints MLUPS / fraction of the 152-byte roofline.
Variants are also cross-checked: rho/v after the timed steps must agree to 1e-6.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from taichi_lbm3d_b200 import _lib, build
    from taichi_lbm3d_b200.geometry import sphere_pack
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    n = int(sys.argv[1])
    cache = "/tmp/lbm3d_geo_porous_%d.npy" % n
    if os.path.exists(cache):
        solid = np.load(cache)
    else:
        r0 = max(3.0, 8.0 * n / 512.0)
        solid = sphere_pack(n, n, n, 0.80, r0, 2 * r0, seed=n, periodic=True)
        np.save(cache, solid)
    nfl = int((solid == 0).sum())
    steps, warm = 60, 10
    peak = 6554.6
    ref = None
    for spec in sys.argv[2:]:
        parts = spec.split(":")
        name = parts[0]
        tag = parts[1] if len(parts) > 1 else ""
        envs = dict(kv.split("=") for kv in parts[2].split(",")) if len(parts) > 2 and parts[2] else {}
        saved = {k: os.environ.get(k) for k in envs}
        os.environ.update(envs)
        if tag:
            os.environ["LBM3D_LIB"] = os.path.join(build.LIBDIR, "liblbm3d_b200.%s.so" % tag)
        else:
            os.environ.pop("LBM3D_LIB", None)
        _lib._lib = None
        lb = LB3D_Solver_Single_Phase(n, n, n, sparse_storage=True)
        lb.solid.from_numpy(solid)
        lb.set_force([1e-6, 0.0, 0.0])
        lb.init_simulation()
        lb.run(warm)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lb.run(steps)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
        mlups = nfl / (best * 1e-3) / 1e6
        v = lb.v.to_numpy()
        fl = solid == 0
        chk = ""
        if ref is None:
            ref = v[fl].copy()
        else:
            chk = " dv_vs_first %.2e" % float(np.abs(v[fl] - ref).max() / np.abs(ref).max())
        print("%-28s ms/step %.4f  MLUPS %.0f  frac %.4f%s" % (name, best, mlups, mlups * 152e6 / 1e9 / peak, chk),
              flush=True)
        del lb
        torch.cuda.empty_cache()
        for k, val in saved.items():
            if val is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = val


if __name__ == "__main__":
    main()
