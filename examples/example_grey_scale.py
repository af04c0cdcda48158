"""Flow over a grey (partially permeable) layer, the case of the reference's
Grey_Scale/lbm_solver_3d_Macro_Sukop.py: 60 x 50 x 5 channel from BC.dat -- solid walls at y = 0 and
y = 49, solid fraction 0.2 on y = 1..19 -- all faces periodic, body force (1e-6, 0, 0), niu = 0.1,
5000 steps, VTK every 500.  The script's physics is the single-phase class with tau = 3 niu + 1/2,
the un-scaled force term and the partial bounce-back streaming switched on by assigning `ns`."""
import os

import numpy as np

import LBM_3D_SinglePhase_Solver as lb3dsp
from _progress import Progress
from taichi_lbm3d_b200 import geometry

nx, ny, nz = 60, 50, 5
GEOMETRY = "./BC.dat"
if os.path.exists(GEOMETRY):              # the reference ships this file; generate the same case when absent
    ns = geometry.load_grey_scale(GEOMETRY, nx, ny, nz)
else:
    ns = geometry.grey_channel(nx, ny, nz, layer=19, fraction=0.2)

solver = lb3dsp.LB3D_Solver_Single_Phase(nx=nx, ny=ny, nz=nz, tau_mode="textbook", guo_mode="unscaled")
solver.ns.from_numpy(ns)                  # also makes the nodes with int(ns) >= 1 solid, as the script does
solver.set_force([1.0e-6, 0.0, 0.0])
solver.set_viscosity(0.1)
solver.init_simulation()

progress = Progress()
for step in range(5001):
    solver.step()
    if step % 100 == 0:
        v = solver.v.to_numpy()
        progress.report(step, max_v=solver.get_max_v(), u_grey=float(np.abs(v[:, 10, :, 0]).mean()),
                        u_open=float(np.abs(v[:, 35, :, 0]).mean()))
    if step % 500 == 0:
        solver.export_VTK(step)
