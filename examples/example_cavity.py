"""Lid-driven cavity, the case of the reference's Single_phase/example_cavity.py: 50^3 box from
geo_cavity.dat, lid velocity (0, 0, 0.1) on the x-right face, 2000 steps, VTK every 1000.

A reference case script itself runs against this package after deleting its two Taichi lines
(`import taichi as ti`, `ti.init(...)`): examples/LBM_3D_SinglePhase_Solver.py re-exports the
class under the module name those scripts import."""
import os

import LBM_3D_SinglePhase_Solver as lb3dsp
from _progress import Progress

GEOMETRY = "./geo_cavity.dat"
if not os.path.exists(GEOMETRY):          # the reference ships this file; generate it when absent
    from taichi_lbm3d_b200 import geometry
    geometry.save_geometry_text(GEOMETRY, geometry.cavity(50, 50, 50))

solver = lb3dsp.LB3D_Solver_Single_Phase(nx=50, ny=50, nz=50, sparse_storage=False)
solver.init_geo(GEOMETRY)
solver.set_bc_vel_x1([0.0, 0.0, 0.1])
solver.init_simulation()

progress = Progress()
for step in range(2001):
    solver.step()
    if step % 500 == 0:
        progress.report(step, max_v=solver.get_max_v())
    if step % 1000 == 0:
        solver.export_VTK(step)           # ./LB_SingelPhase_<step>.vtr
