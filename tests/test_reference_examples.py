"""The reference's own case scripts (Single_phase/example_cavity.py, example_poiseuille_flow.py,
example_porous_medium.py) against the drop-in class: SURVEY 8(b) end to end.

CPU (this container, where /root/reference exists): each script is executed VERBATIM -- only a
no-op `taichi` module is put in its way and its loop is capped -- (a) against a recorder, whose
call sequence must equal the committed trace, and (b) against the product class with the device
entry points stubbed out, which proves constructor signature, setters, `init_geo` on the
reference's own geometry file and `solid.from_numpy` on the script's float64 array.

GPU (no /root/reference there): the committed call sequence is replayed on the product class and
the state is compared with the oracle driven the same way; `export_VTK` is called on the live
solver and the file it writes is decoded independently of the product's reader.
"""
import json
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from tests import refscripts

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LBM3D_REFERENCE", "/root/reference")
with open(os.path.join(HERE, "golden", "ref_example_traces.json")) as _fh:
    TRACES = json.load(_fh)
have_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "Single_phase")),
                              reason="the reference tree is not present")


def _product_module():
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    try:
        sys.modules.pop("LBM_3D_SinglePhase_Solver", None)
        import LBM_3D_SinglePhase_Solver as m          # examples/: the module name the scripts import
    finally:
        sys.path.remove(os.path.join(ROOT, "examples"))
        sys.modules.pop("LBM_3D_SinglePhase_Solver", None)
    return m


@have_ref
def test_committed_traces_are_what_the_reference_scripts_do():
    from tests.golden import make_example_traces
    assert json.loads(json.dumps(make_example_traces.traces())) == TRACES


@have_ref
@pytest.mark.parametrize("rel", ["Single_phase/example_cavity.py", "Single_phase/example_poiseuille_flow.py"])
def test_reference_script_verbatim_up_to_the_device(rel, tmp_path, monkeypatch):
    """the unmodified script text runs against the product class (device calls stubbed: no GPU here)"""
    import types
    prod = _product_module()
    seen = {"steps": 0, "vtk": []}

    class Dry(prod.LB3D_Solver_Single_Phase):
        def init_simulation(self):
            seen["geometry"] = self.solid.to_numpy()
            seen["bc"] = [self._bc_tuple(f) for f in range(6)]
            seen["force"], seen["niu"] = (self.fx, self.fy, self.fz), self.niu

        def step(self):
            seen["steps"] += 1

        def get_max_v(self):
            return 0.0

        def export_VTK(self, n):
            seen["vtk"].append(n)

    mod = types.ModuleType("LBM_3D_SinglePhase_Solver")
    mod.LB3D_Solver_Single_Phase = Dry
    path = os.path.join(REF, rel)
    monkeypatch.chdir(os.path.dirname(path))          # the cavity script reads ./geo_cavity.dat
    with open(path) as fh:
        refscripts.run_script(fh.read(), mod, TRACES[rel]["max_iter"], path)
    assert seen["steps"] == TRACES[rel]["max_iter"] and seen["vtk"] == [0]
    if "cavity" in rel:
        from taichi_lbm3d_b200.geometry import cavity
        assert np.array_equal(seen["geometry"], cavity(50, 50, 50))
        assert seen["bc"][1] == (2, 1.0, [0.0, 0.0, 0.1])
    else:
        g = np.zeros((5, 20, 16), np.int8)
        g[:, :, 0] = 1
        g[:, :, -1] = 1
        assert np.array_equal(seen["geometry"], g)
        assert seen["force"] == (0.0, 0.0001, 0.0) and seen["niu"] == 0.1667


def _decode_vtr(fname):
    """independent of taichi_lbm3d_b200.vtk.read_vtr: XML header through ElementTree, appended
    blocks by their offsets, point order x fastest (VTK XML RectilinearGrid)"""
    raw = open(fname, "rb").read()
    cut = raw.index(b"<AppendedData")
    root = ET.fromstring(raw[:cut] + b"</VTKFile>")
    assert root.tag == "VTKFile" and root.get("type") == "RectilinearGrid" and root.get("byte_order") == "LittleEndian"
    assert root.get("header_type") == "UInt64"
    grid = root.find("RectilinearGrid")
    ext = [int(t) for t in grid.get("WholeExtent").split()]
    n = (ext[1] + 1, ext[3] + 1, ext[5] + 1)
    base = raw.index(b"_", cut) + 1
    dt = {"Int8": "i1", "Float32": "<f4", "Float64": "<f8"}
    out = {}
    for da in grid.find("Piece").iter("DataArray"):
        off = base + int(da.get("offset"))
        nbytes = struct.unpack("<Q", raw[off:off + 8])[0]
        a = np.frombuffer(raw[off + 8:off + 8 + nbytes], dtype=dt[da.get("type")])
        nc = int(da.get("NumberOfComponents"))
        if da.get("Name").endswith("_coordinates"):
            out[da.get("Name")] = a
        else:
            assert a.size == n[0] * n[1] * n[2] * nc
            a = a.reshape((n[2], n[1], n[0]) + ((nc,) if nc > 1 else ()))      # z slowest, x fastest
            out[da.get("Name")] = np.transpose(a, (2, 1, 0) + ((3,) if nc > 1 else ()))
    return n, out


def _drive(o, calls):
    """the setters of a recorded call sequence on an oracle"""
    for name, arg in calls:
        if name == "set_bc_vel_x1":
            o.set_bc_vel(1, arg[0])
        elif name == "set_force":
            o.set_force(arg[0])
        elif name == "set_viscosity":
            o.set_viscosity(arg[0])


@pytest.mark.gpu
@pytest.mark.parametrize("rel", ["Single_phase/example_cavity.py", "Single_phase/example_poiseuille_flow.py"])
def test_replayed_reference_script_matches_the_oracle(cuda, rel, tmp_path, monkeypatch):
    from oracle.cref import RefSinglePhaseC
    from taichi_lbm3d_b200 import geometry
    from tests.cases import TOL, rel_linf
    monkeypatch.chdir(tmp_path)
    if "cavity" in rel:                                # the file the reference ships next to the script
        geometry.save_geometry_text("./geo_cavity.dat", geometry.cavity(50, 50, 50))
    prod = _product_module()
    calls = TRACES[rel]["calls"]
    lb = refscripts.replay(calls, prod.LB3D_Solver_Single_Phase)
    steps = sum(arg for name, arg in calls if name == "step")
    assert steps == TRACES[rel]["max_iter"]
    # the oracle, driven by the same sequence
    solid = lb.solid.to_numpy()
    o = RefSinglePhaseC(*solid.shape)
    o.set_solid(solid)
    _drive(o, calls)
    o.init_simulation()
    o.run(steps)
    fl = solid == 0
    assert rel_linf(lb.F.to_numpy()[fl], o.F[fl]) <= TOL
    assert rel_linf(lb.rho.to_numpy()[fl], o.rho[fl]) <= TOL
    # v is a difference of O(0.1) populations: where it is small (Poiseuille, 6e-4) the bar is the
    # oracle's own fp32 round-off, measured against its fp64 form (tests/cases.py: v_abs_tolerance)
    o64 = RefSinglePhaseC(*solid.shape, dtype=np.float64)
    o64.set_solid(solid)
    _drive(o64, calls)
    o64.init_simulation()
    o64.run(steps)
    from tests.cases import v_abs_tolerance
    assert np.abs(lb.v.to_numpy()[fl].astype(np.float64) - o.v[fl]).max() <= v_abs_tolerance(o, o64)
    # export_VTK(0) ran after the first step (the scripts export at iter 0): file name and content
    fname = str(tmp_path / "LB_SingelPhase_0.vtr")     # [sic], reference :464
    assert os.path.exists(fname)
    o1 = RefSinglePhaseC(*solid.shape)
    o1.set_solid(solid)
    _drive(o1, calls)
    o1.init_simulation()
    o1.run(1)
    n, data = _decode_vtr(fname)
    assert n == solid.shape and set(data) >= {"Solid", "rho", "velocity", "x_coordinates"}
    assert np.array_equal(data["Solid"], solid)
    assert np.array_equal(data["x_coordinates"], np.linspace(0, n[0], n[0]))
    assert rel_linf(data["rho"][fl], o1.rho[fl]) <= TOL
    assert np.abs(data["velocity"][fl] - o1.v[fl]).max() <= TOL * np.abs(o1.v[fl]).max() + 3e-7
    # and a fresh export of the live GPU solver at the end of the run
    lb.export_VTK(steps)
    n2, d2 = _decode_vtr(str(tmp_path / ("LB_SingelPhase_%d.vtr" % steps)))
    assert np.array_equal(d2["rho"], lb.rho.to_numpy()) and np.array_equal(d2["velocity"], lb.v.to_numpy())
