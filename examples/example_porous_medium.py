"""Pressure-driven flow through a porous sample, the case of the reference's
Single_phase/example_porous_medium.py (BASELINE config 1): 131^3, rho = 1.0 on the x-left face
and 0.99 on the x-right face, default viscosity, 50 000 steps.  The micro-CT image
img_ftb131.txt is not redistributed with the reference mount (.MISSING_LARGE_BLOBS); when it is
absent a seeded sphere-pack stand-in of the same size and porosity is written in its place."""
import os

import LBM_3D_SinglePhase_Solver as lb3dsp
from _progress import Progress

IMAGE = os.environ.get("LBM3D_FTB131", "./img_ftb131.txt")
if not os.path.exists(IMAGE):
    if "LBM3D_FTB131" in os.environ:
        raise FileNotFoundError("LBM3D_FTB131=%s does not exist" % IMAGE)
    from taichi_lbm3d_b200 import geometry
    geometry.save_geometry_text(IMAGE, geometry.ftb131_standin())

# dense storage as in the reference script (:12); sparse_storage=True is the faster choice at 20 % porosity
solver = lb3dsp.LB3D_Solver_Single_Phase(nx=131, ny=131, nz=131, sparse_storage=False)
solver.init_geo(IMAGE)
solver.set_bc_rho_x0(1.0)
solver.set_bc_rho_x1(0.99)
solver.init_simulation()

progress = Progress()
for step in range(50001):
    solver.step()
    if step % 500 == 0:
        progress.report(step, max_v=solver.get_max_v())
    if step % 2000 == 0:
        solver.export_VTK(step)
