# Drainage through the ftb131 sandstone with the colour-gradient two-phase solver: the case the
# reference script 2phase/lbm_solver_3d_2phase.py (and its _sparse twin) hard-codes -- 131^3,
# niu_l = 0.05, niu_g = 0.2, CapA = 0.005, psi_solid = 0.7, psi = -1 fixed on the x-left face,
# force (5e-5, -2e-5, 0), the loop of its lines 626-659.  The script's module-level globals are
# attributes of the class here.  img_ftb131.txt / phase_ftb131.dat are not redistributed with
# the reference mount (.MISSING_LARGE_BLOBS); when absent, a seeded sphere-pack stand-in and a
# phase field with the non-wetting phase in the first 13 planes are used.
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase, geometry  # noqa: E402
from _progress import Progress  # noqa: E402

nx = ny = nz = 131
lb = LB3D_Solver_Two_Phase(nx, ny, nz, sparse_storage=True)     # what ..._2phase_sparse.py does

if os.path.exists('./img_ftb131.txt') and os.path.exists('./phase_ftb131.dat'):
    lb.init_geo('./img_ftb131.txt', './phase_ftb131.dat')
else:
    solid = geometry.ftb131_standin()
    phase = np.ones((nx, ny, nz), np.float32)
    phase[:13] = -1.0
    lb.solid.from_numpy(solid)
    lb.psi.from_numpy(phase)

lb.niu_l, lb.niu_g = 0.05, 0.2
lb.CapA = 0.005
lb.psi_solid = 0.7
lb.fx, lb.fy, lb.fz = 5.0e-5, -2e-5, 0.0
lb.bc_psi_x_left, lb.psi_x_left = 1, -1.0
lb.init_simulation()

progress = Progress()
for step in range(80001):
    lb.step()
    if step % 500 == 0:
        fluid = lb.solid.to_numpy() == 0
        progress.report(step, max_v=lb.get_max_v(), non_wetting_saturation=float((lb.psi.to_numpy()[fluid] < 0).mean()))
    if step % 10000 == 0:
        lb.export_VTK(step)               # ./structured<step>.vtr: Solid, rho, phase, velocity
