"""Geometry input: the reference's text format plus binary formats, and the generators the
benchmark configurations need.

Reference: ``init_geo`` (Single_phase/LBM_3D_SinglePhase_Solver.py:173-177) reads an ASCII
file of 0/1 with x fastest (README.md:17-23); ``flow_domain_geo_generation_2D.py:18-27``
writes the cavity file.  All functions return int8 arrays of shape (nx, ny, nz), C order,
1 = solid, which is what ``solid.to_numpy()`` exposes in the reference.
"""
import os

import numpy as np


def load_geometry(filename, nx, ny, nz):
    """init_geo :173-177.  ``.npy`` (any integer/bool array of shape (nx,ny,nz)) and ``.raw``
    (uint8, x fastest like the text format) are accepted besides the reference's text."""
    ext = os.path.splitext(filename)[1].lower()
    if ext == ".npy":
        a = np.load(filename)
        if a.shape != (nx, ny, nz):
            raise ValueError("%s has shape %s, expected %s" % (filename, a.shape, (nx, ny, nz)))
        return (a > 0).astype(np.int8)
    if ext == ".raw":
        in_dat = np.fromfile(filename, dtype=np.uint8)
    else:
        in_dat = _fast_loadtxt(filename)
    if in_dat.size != nx * ny * nz:
        raise ValueError("%s holds %d values, expected %d" % (filename, in_dat.size, nx * ny * nz))
    in_dat = (in_dat > 0).astype(np.int8)                          # :175
    return np.ascontiguousarray(np.reshape(in_dat, (nx, ny, nz), order='F'))   # :176


def _fast_loadtxt(filename):
    """np.loadtxt(filename) for the 0/1 format, without the per-token Python overhead."""
    with open(filename, "rb") as fh:
        raw = fh.read()
    toks = np.frombuffer(raw, dtype=np.uint8)
    digit = (toks == 48) | (toks == 49)
    if toks.size and np.all(digit | (toks == 10) | (toks == 13) | (toks == 32)):
        # fast only when every token is ONE character ("10" or "01" is one number for np.loadtxt):
        # no two digits may be adjacent
        if not np.any(digit[1:] & digit[:-1]):
            return (toks[digit] - 48).astype(np.float64)
    return np.loadtxt(filename).reshape(-1)        # general numbers: the reference's own call


def save_geometry_text(filename, solid):
    """flow_domain_geo_generation_2D.py:25-28: Fortran-order flatten, one value per line."""
    out = np.asarray(solid).reshape(-1, order='F')
    np.savetxt(filename, out.T, fmt='%d')


def load_grey_scale(filename, nx, ny, nz):
    """Solid fraction per node as the grey-scale script reads it (Grey_Scale/
    lbm_solver_3d_Macro_Sukop.py:142-145): text, one value per line, Fortran order; for
    ``solver.ns.from_numpy``."""
    in_dat = np.loadtxt(filename)
    if in_dat.size != nx * ny * nz:
        raise ValueError("%s holds %d values, expected %d" % (filename, in_dat.size, nx * ny * nz))
    return np.ascontiguousarray(np.reshape(in_dat, (nx, ny, nz), order='F').astype(np.float32))


def grey_channel(nx=60, ny=50, nz=5, layer=19, fraction=0.2):
    """The case of the reference's Grey_Scale/BC.dat: a channel between solid walls at y = 0 and
    y = ny-1 with a grey layer of the given solid fraction on the lower wall (y = 1 .. layer)."""
    ns = np.zeros((nx, ny, nz), np.float32)
    ns[:, 0, :] = 1.0
    ns[:, ny - 1, :] = 1.0
    ns[:, 1:1 + layer, :] = fraction
    return ns


def cavity(nx, ny, nz):
    """flow_domain_geo_generation_2D.py:18-23: walls on x=0, y=0, y=-1, z=0, z=-1 (the lid is
    the open x=nx-1 face, driven by set_bc_vel_x1)."""
    g = np.zeros((nx, ny, nz), np.int8)
    g[0, :, :] = 1
    g[:, 0, :] = 1
    g[:, -1, :] = 1
    g[:, :, 0] = 1
    g[:, :, -1] = 1
    return g


def sphere_pack(nx, ny, nz, solid_fraction=0.80, r_min=8.0, r_max=16.0, seed=512, periodic=True,
                batch=64):
    """Seeded overlapping-sphere pack: spheres with radii
    U[r_min, r_max] are added until the solid fraction reaches ``solid_fraction``."""
    rng = np.random.default_rng(seed)
    g = np.zeros((nx, ny, nz), bool)
    n = np.array([nx, ny, nz])
    target = solid_fraction * g.size
    count = 0                       # solid cells so far, kept incrementally (checked once per batch)
    while count < target:
        for _ in range(batch):
            c = rng.random(3) * n
            r = r_min + (r_max - r_min) * rng.random()
            lo = np.floor(c - r).astype(int)
            hi = np.ceil(c + r).astype(int) + 1
            axes = []
            for d in range(3):
                idx = np.arange(lo[d], hi[d])
                dist = idx - c[d]
                if periodic:
                    idx = idx % n[d]
                else:
                    keep = (idx >= 0) & (idx < n[d])
                    idx, dist = idx[keep], dist[keep]
                axes.append((idx, dist))
            (ix, dx), (iy, dy), (iz, dz) = axes
            mask = (dx[:, None, None] ** 2 + dy[None, :, None] ** 2 + dz[None, None, :] ** 2) <= r * r
            if periodic and (len(set(ix.tolist())) < ix.size or len(set(iy.tolist())) < iy.size
                             or len(set(iz.tolist())) < iz.size):
                # sphere wider than the box: wrapped indices repeat, count the slow way
                g[np.ix_(ix, iy, iz)] |= mask
                count = int(g.sum())
                continue
            sel = np.ix_(ix, iy, iz)
            sub = g[sel]
            count += int(np.count_nonzero(mask & ~sub))
            g[sel] = sub | mask
    return g.astype(np.int8)


def ftb131_standin():
    """Synthetic stand-in for the missing img_ftb131.txt: 131^3, seed 131,
    radii U[4,9], solid fraction >= 0.80, non-periodic in x."""
    return sphere_pack(131, 131, 131, 0.80, 4.0, 9.0, seed=131, periodic=False)


def ftb131_geometry():
    """(solid, where it came from): the micro-CT image of BASELINE configs 1 and 4 when the user
    has it -- the file named by ``LBM3D_FTB131``, else ``./img_ftb131.txt`` (the name the
    reference's example scripts open, Single_phase/example_porous_medium.py:13) -- and otherwise
    the seeded stand-in.  A path given through the environment that does not load is
    an error, not a silent fall-back."""
    named = os.environ.get("LBM3D_FTB131")
    for path in ([named] if named else []) + ["./img_ftb131.txt"]:
        if os.path.exists(path):
            return load_geometry(path, 131, 131, 131), "ftb131 image %s" % os.path.abspath(path)
        if path == named:
            raise FileNotFoundError("LBM3D_FTB131=%s does not exist" % named)
    return ftb131_standin(), "sphere-pack stand-in for the missing ftb131 files"
