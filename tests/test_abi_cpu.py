"""CPU tests of the boundary: the library loads, exports every symbol include/lbm3d.h
declares, fails loudly without a GPU, and the host-side mirror behaves like the reference
class (setters, defaults, geometry loader, VTK writer)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions(name):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbm2?p?_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from taichi_lbm3d_b200 import _lib
    lib = _lib.load()
    declared = _header_functions("lbm3d.h")
    assert len(declared) >= 30
    for fn in declared:
        assert hasattr(lib, fn), "missing export %s" % fn
        assert fn in _lib.SIGNATURES, "ctypes binding lacks %s" % fn
    assert lib.lbm_abi_version() == 1
    declared2 = _header_functions("lbm3d_2phase.h")
    assert len(declared2) >= 20
    for fn in declared2:
        assert hasattr(lib, fn), "missing export %s" % fn
        assert fn in _lib.SIGNATURES_2P, "ctypes binding lacks %s" % fn


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from taichi_lbm3d_b200 import _lib, LB3D_Solver_Single_Phase
    lib = _lib.load()
    cfg = _lib.LbmConfig(nx=4, ny=4, nz=4, sparse=0, strict=0, halo_x=0, device=0, x_face_mask=0)
    ctx = ctypes.c_void_p()
    assert lib.lbm_create(ctypes.byref(cfg), ctypes.byref(ctx)) == -2       # LBM_ERR_CUDA
    assert b"no CPU fallback" in lib.lbm_last_error(None)
    lb = LB3D_Solver_Single_Phase(4, 4, 4)
    with pytest.raises(_lib.LbmError):
        lb.init_simulation()
    with pytest.raises(_lib.LbmError):
        lb.step()
    from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase
    cfg2 = _lib.Lbm2pConfig(nx=4, ny=4, nz=4, strict=0, device=0, reserved=0)
    assert lib.lbm2p_create(ctypes.byref(cfg2), ctypes.byref(ctx)) == -2
    with pytest.raises(_lib.LbmError):
        LB3D_Solver_Two_Phase(4, 4, 4).init_simulation()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "taichi_lbm3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "from oracle" not in txt and "import oracle" not in txt, f
                assert "libref_" not in txt and "oracle/_build" not in txt, f


def test_class_surface_matches_reference():
    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
    lb = LB3D_Solver_Single_Phase(nx=5, ny=6, nz=7, sparse_storage=True)
    # defaults of the reference ctor (:13-28)
    assert (lb.nx, lb.ny, lb.nz) == (5, 6, 7) and lb.niu == 0.16667 and (lb.fx, lb.fy, lb.fz) == (0.0, 0.0, 0.0)
    assert lb.bc_x_left == 0 and lb.rho_bcxl == 1.0 and lb.sparse_storage is True
    for name in ("init_geo", "init_simulation", "step", "get_max_v", "export_VTK", "set_viscosity", "set_force",
                 "set_bc_vel_x0", "set_bc_vel_x1", "set_bc_vel_y0", "set_bc_vel_y1", "set_bc_vel_z0",
                 "set_bc_vel_z1", "set_bc_rho_x0", "set_bc_rho_x1", "set_bc_rho_y0", "set_bc_rho_y1",
                 "set_bc_rho_z0", "set_bc_rho_z1"):
        assert callable(getattr(lb, name))
    lb.set_bc_vel_x1([0.0, 0.0, 0.1])
    assert lb.bc_x_right == 2 and lb.vz_bcxr == 0.1                      # :405-407
    lb.set_bc_rho_y0(0.98)
    assert lb.bc_y_left == 1 and lb.rho_bcyl == 0.98                     # :437-439
    lb.set_force([1e-6, 0, 0])
    lb.set_viscosity(0.1)
    assert lb.fx == 1e-6 and lb.niu == 0.1
    geo = np.zeros((5, 6, 7))
    geo[:, :, 0] = 3.0
    lb.solid.from_numpy(geo)                                             # example_poiseuille_flow.py:24
    assert lb.solid.to_numpy().dtype == np.int8 and lb.solid.to_numpy().max() == 1
    for f in ("rho", "v", "f", "F", "solid"):
        assert hasattr(getattr(lb, f), "to_numpy") and hasattr(getattr(lb, f), "from_numpy")
    assert np.array_equal(lb.x, np.linspace(0, 5, 5))                    # :112


def test_relaxation_rates_rounding():
    from taichi_lbm3d_b200 import relaxation_rates
    from oracle.ref_single_phase import relaxation_rates as oracle_rates
    for niu in (0.16667, 0.1667, 0.1, 0.05):
        for mode in ("class", "textbook"):
            assert np.array_equal(relaxation_rates(niu, mode), oracle_rates(niu, mode).astype(np.float32))


def test_geometry_loader_matches_reference_semantics(tmp_path):
    from taichi_lbm3d_b200 import geometry
    rng = np.random.default_rng(0)
    g = (rng.random((6, 5, 4)) < 0.4).astype(np.int8)
    txt = tmp_path / "geo.dat"
    geometry.save_geometry_text(str(txt), g)
    # the reference's own loader lines (:174-176)
    in_dat = np.loadtxt(str(txt))
    in_dat[in_dat > 0] = 1
    want = np.reshape(in_dat, (6, 5, 4), order='F')
    assert np.array_equal(geometry.load_geometry(str(txt), 6, 5, 4), want)
    assert np.array_equal(want, g)
    np.save(str(tmp_path / "geo.npy"), g)
    assert np.array_equal(geometry.load_geometry(str(tmp_path / "geo.npy"), 6, 5, 4), g)
    g.reshape(-1, order='F').astype(np.uint8).tofile(str(tmp_path / "geo.raw"))
    assert np.array_equal(geometry.load_geometry(str(tmp_path / "geo.raw"), 6, 5, 4), g)
    with pytest.raises(ValueError):
        geometry.load_geometry(str(txt), 6, 5, 5)
    with pytest.raises(FileNotFoundError):
        geometry.load_geometry(str(tmp_path / "missing.dat"), 6, 5, 4)


def test_generators():
    from taichi_lbm3d_b200 import geometry
    c = geometry.cavity(8, 9, 10)
    assert c[0].all() and c[:, 0].all() and c[:, -1].all() and c[:, :, 0].all() and c[:, :, -1].all()
    assert int((c == 0).sum()) == 7 * 7 * 8
    a = geometry.sphere_pack(24, 24, 24, 0.7, 2.0, 4.0, seed=1)
    b = geometry.sphere_pack(24, 24, 24, 0.7, 2.0, 4.0, seed=1)
    assert np.array_equal(a, b) and 0.7 <= a.mean() < 0.9


def test_grey_scale_geometry(tmp_path):
    """grey_channel() is the reference's Grey_Scale/BC.dat (walls at y = 0, 49; solid fraction 0.2 on
    y = 1..19; checked against the file where the reference is mounted), and load_grey_scale reads
    the script's text format (Fortran order, :142-145)"""
    import os
    from taichi_lbm3d_b200 import geometry
    ns = geometry.grey_channel()
    assert ns.shape == (60, 50, 5) and ns.dtype == np.float32
    assert np.all(ns[:, 0] == 1) and np.all(ns[:, 49] == 1) and np.all(ns[:, 1:20] == np.float32(0.2)) and not ns[:, 20:49].any()
    path = str(tmp_path / "BC.dat")
    np.savetxt(path, ns.reshape(-1, order="F"), fmt="%g")
    assert np.array_equal(geometry.load_grey_scale(path, 60, 50, 5), ns)
    ref = "/root/reference/Grey_Scale/BC.dat"
    if os.path.exists(ref):
        assert np.array_equal(geometry.load_grey_scale(ref, 60, 50, 5), ns)


def test_vtr_roundtrip(tmp_path):
    from taichi_lbm3d_b200 import vtk
    rng = np.random.default_rng(1)
    nx, ny, nz = 4, 5, 6
    solid = (rng.random((nx, ny, nz)) < 0.5).astype(np.int8)
    rho = rng.random((nx, ny, nz)).astype(np.float32)
    v = rng.random((nx, ny, nz, 3)).astype(np.float32)
    x, y, z = np.linspace(0, nx, nx), np.linspace(0, ny, ny), np.linspace(0, nz, nz)
    fn = vtk.grid_to_vtr(str(tmp_path / "LB_SingelPhase_7"), x, y, z,
                         {"Solid": solid, "rho": rho, "velocity": (v[..., 0].copy(), v[..., 1].copy(), v[..., 2].copy())})
    assert fn.endswith("LB_SingelPhase_7.vtr")
    head = open(fn, "rb").read(400).decode("ascii", "replace")
    assert 'type="RectilinearGrid"' in head and 'WholeExtent="0 3 0 4 0 5"' in head
    x2, y2, z2, data = vtk.read_vtr(fn)
    assert np.array_equal(x2, x) and np.array_equal(z2, z)
    assert np.array_equal(data["Solid"], solid) and np.array_equal(data["rho"], rho)
    for k in range(3):
        assert np.array_equal(data["velocity"][k], v[..., k])


def test_two_phase_class_keeps_script_globals():
    """attribute names and defaults of 2phase/lbm_solver_3d_2phase.py:16-39"""
    from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase
    lb = LB3D_Solver_Two_Phase(6, 5, 4)
    assert (lb.fx, lb.fy, lb.fz) == (5.0e-5, -2e-5, 0.0) and lb.niu_l == 0.1 and lb.niu_g == 0.1
    assert lb.psi_solid == 0.7 and lb.CapA == 0.005 and lb.rho_bcxr == 0.995
    assert (lb.bc_psi_x_left, lb.psi_x_left) == (1, -1.0) and lb.bc_psi_z_right == 0
    lb.set_bc_psi(3, 1.0)
    assert lb.bc_psi_y_right == 1 and lb.psi_y_right == 1.0
    lb.set_bc_rho(1, 0.99)
    assert lb.bc_x_right == 1 and lb.rho_bcxr == 0.99
    psi = np.where(np.arange(6)[:, None, None] < 2, -1.0, 1.0) * np.ones((6, 5, 4))
    lb.psi.from_numpy(psi)
    assert np.array_equal(lb.psi.to_numpy(), psi.astype(np.float32))


def test_two_phase_init_geo(tmp_path):
    from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase, geometry
    rng = np.random.default_rng(2)
    g = (rng.random((5, 4, 3)) < 0.3).astype(np.int8)
    ph = np.where(rng.random((5, 4, 3)) < 0.5, -1.0, 1.0)
    geometry.save_geometry_text(str(tmp_path / "g.dat"), g)
    np.savetxt(str(tmp_path / "p.dat"), ph.reshape(-1, order='F'))
    lb = LB3D_Solver_Two_Phase(5, 4, 3)
    s, p = lb.init_geo(str(tmp_path / "g.dat"), str(tmp_path / "p.dat"))     # script :194-202
    assert np.array_equal(s, g) and np.array_equal(p, ph.astype(np.float32))


def test_create_rejects_bad_and_oversized_lattices():
    """argument errors are reported before any device work (no GPU needed)"""
    import ctypes
    from taichi_lbm3d_b200 import _lib
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    for nx, ny, nz, halo in ((0, 4, 4, 0), (4, -1, 4, 0), (2048, 2048, 2048, 0), (2, 4, 4, 1)):
        cfg = _lib.LbmConfig(nx=nx, ny=ny, nz=nz, sparse=0, strict=0, halo_x=halo, device=0, x_face_mask=0)
        assert lib.lbm_create(ctypes.byref(cfg), ctypes.byref(ctx)) == -1        # LBM_ERR_INVALID
        assert lib.lbm_last_error(None)
    assert lib.lbm_create(None, ctypes.byref(ctx)) == -1
    # NULL contexts are refused, not dereferenced
    assert lib.lbm_step(None, 1, None) == -1 and lib.lbm_init(None) == -1 and lib.lbm_launch_count(None) == -1
    assert lib.lbm2p_step(None, 1, None) == -1


def test_headers_are_plain_c_and_a_c_host_links(tmp_path):
    """include/*.h compile as C99 (no C++-isms, no torch types) and a plain C program links against
    liblbm3d_b200.so; without a GPU its lbm_create fails with the library's own message"""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    from taichi_lbm3d_b200 import _lib
    _lib.load()
    libdir = os.path.dirname(_lib.library_path())
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include "lbm3d.h"
#include "lbm3d_2phase.h"
int main(void) {
    lbm_config cfg = {4, 4, 4, 0, 0, 0, 0, 0};
    lbm_ctx *ctx = NULL;
    int rc = lbm_create(&cfg, &ctx);
    printf("abi %d rc %d msg %s\n", lbm_abi_version(), rc, rc ? lbm_last_error(NULL) : "ok");
    if (ctx) lbm_destroy(ctx);
    lbm2p_config cfg2 = {4, 4, 4, 0, 0, LBM2P_SPARSE};
    lbm2p_ctx *c2 = NULL;
    rc = lbm2p_create(&cfg2, &c2);
    if (c2) lbm2p_destroy(c2);
    return 0;
}
''')
    exe = tmp_path / "host"
    subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-llbm3d_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, check=True).stdout.decode()
    assert out.startswith("abi 1 rc ")
    import torch
    if not torch.cuda.is_available():
        assert "no CPU fallback" in out
