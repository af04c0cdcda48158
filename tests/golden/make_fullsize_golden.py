#!/usr/bin/env python
"""Full-size oracle fixtures for BASELINE configs 2, 3 and 4 at the north_star's horizon.

    OMP_NUM_THREADS=6 python tests/golden/make_fullsize_golden.py cfg4 [cfg2] [cfg3] [cfg3k]

The C oracle (oracle/ref_*.c, the build whose results are bit-identical to the NumPy restatement
and to the reference source run through tests/taichi_shim) needs minutes to an hour of host time
at these sizes, so it runs HERE, offline, and the GPU tests compare against what it left behind:

  * a sample of fluid nodes (every fluid node of a few lattice planes, capped) with ALL fields of
    the oracle at the final step -- verification arithmetic must reproduce them bit for bit,
    production arithmetic within the tolerance stored next to them;
  * float64 sums of every field over ALL fluid nodes (a checksum of the whole lattice);
  * per field, the oracle's own fp32 round-off at that horizon: |oracle32 - oracle64| (max over the
    fluid nodes, and its median / 90th / 99th percentile over the sample), the yardstick of SURVEY
    8c for quantities that are differences of O(1) numbers (v in creeping flow) or that interface
    dynamics amplify (two-phase fields: at 1000 steps of config 4 the fp32 oracle is 1e-2 away from
    its own fp64 form at the worst node, 1e-5 at the median).

Nothing here reads /root/reference; geometry comes from the seeded generators of
taichi_lbm3d_b200.geometry, so the GPU test rebuilds the identical case.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.cref import RefSinglePhaseC, RefTwoPhaseC  # noqa: E402
from tests import cases, cases2p  # noqa: E402

PCTS = (50.0, 90.0, 99.0)
SP_FIELDS = ("F", "rho", "v")
TP_FIELDS = ("F", "rho", "v", "psi", "rho_r", "rho_b")


def fullsize_case(name):
    """(case, steps, fields, planes to sample, oracle class) -- shared with the GPU tests"""
    from taichi_lbm3d_b200.geometry import ftb131_standin, sphere_pack
    if name == "cfg2":        # 256^3 lid-driven cavity, dense (Single_phase/example_cavity.py scaled up)
        return cases.case_cavity(256), 1000, SP_FIELDS, (3, 128, 254), RefSinglePhaseC
    if name == "cfg3":        # 512^3 periodic sphere pack, porosity 0.20, fx = 1e-6, all faces periodic
        solid = sphere_pack(512, 512, 512, 0.80, 8.0, 16.0, seed=512, periodic=True)
        return cases.Case("cfg3", solid, force=[1e-6, 0.0, 0.0]), 100, SP_FIELDS, (0, 200, 511), RefSinglePhaseC
    if name == "cfg3k":       # the same medium at 256^3 over the north_star's full horizon of 1000 steps
        solid = sphere_pack(256, 256, 256, 0.80, 8.0, 16.0, seed=256, periodic=True)
        return cases.Case("cfg3k", solid, force=[1e-6, 0.0, 0.0]), 1000, SP_FIELDS, (0, 100, 255), RefSinglePhaseC
    if name == "cfg4":        # 131^3 drainage, README parameters, psi = -1 entering from x0
        solid = ftb131_standin()
        psi = np.ones(solid.shape, np.float32)
        psi[:13] = -1.0
        case = cases2p.Case2P("cfg4", solid, psi, niu_l=0.05, niu_g=0.2, CapA=0.005, psi_solid=0.7)
        return case, 1000, TP_FIELDS, (1, 14, 40, 90, 129), RefTwoPhaseC
    raise ValueError(name)


def sample_index(solid, planes, cap=24000):
    """flat indices of the fluid nodes of the given x planes (evenly thinned to at most `cap`)"""
    nx, ny, nz = solid.shape
    idx = []
    for x in planes:
        flat = np.flatnonzero(solid[x].reshape(-1) == 0).astype(np.int64) + np.int64(x) * ny * nz
        idx.append(flat)
    idx = np.concatenate(idx)
    if idx.size > cap:
        idx = idx[np.linspace(0, idx.size - 1, cap).astype(np.int64)]
    return idx


def take(field, idx):
    a = np.asarray(field)
    n = int(np.prod(a.shape[:3]))
    return np.ascontiguousarray(a.reshape((n,) + a.shape[3:])[idx])


def main(names):
    for name in names:
        case, steps, fields, planes, cls = fullsize_case(name)
        fl = case.solid == 0
        idx = sample_index(case.solid, planes)
        out = {"steps": np.int64(steps), "index": idx, "shape": np.array(case.shape, np.int64),
               "n_fluid": np.int64(fl.sum())}
        t0 = time.time()
        o = case.make_oracle(cls)
        o.run(steps)
        print("%s: fp32 oracle, %d steps, %.0f s" % (name, steps, time.time() - t0), flush=True)
        for f in fields:
            a = getattr(o, f)
            out[f] = take(a, idx)
            out["sum_" + f] = np.float64(a[fl].astype(np.float64).sum())
            out["max_" + f] = np.float64(np.abs(a[fl]).max())
        keep = {f: getattr(o, f)[fl].astype(np.float64) for f in fields if f != "F"}
        keepF = take(o.F, idx).astype(np.float64)
        keep_s = {f: take(getattr(o, f), idx).astype(np.float64) for f in fields}
        del o
        t0 = time.time()
        o64 = case.make_oracle(cls, dtype=np.float64)
        o64.run(steps)
        print("%s: fp64 oracle, %.0f s" % (name, time.time() - t0), flush=True)
        for f in fields:
            if f == "F":      # populations: round-off measured on the sample (the full array is 19 x larger)
                d = np.abs(keepF - take(o64.F, idx)).max()
            else:
                d = np.abs(keep[f] - getattr(o64, f)[fl]).max()
            out["roundoff_" + f] = np.float64(d)
            # how that round-off is distributed over the sample: median, 90th and 99th percentile
            ds = np.abs(keep_s[f] - take(getattr(o64, f), idx))
            out["roundoff_pct_" + f] = np.percentile(ds.reshape(ds.shape[0], -1).max(axis=1), PCTS)
            print("   %-6s max %.6g  fp32 round-off %.3g (relative %.3g)"
                  % (f, out["max_" + f], d, d / out["max_" + f]), flush=True)
        del o64
        path = os.path.join(HERE, "full_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("wrote %s (%.1f MB)" % (path, os.path.getsize(path) / 1e6), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["cfg4", "cfg2", "cfg3"])
