# Drainage through the ftb131 sandstone with the colour-gradient two-phase solver: the case the
# reference script 2phase/lbm_solver_3d_2phase.py (and its _sparse twin) hard-codes -- 131^3,
# niu_l = 0.05, niu_g = 0.2, CapA = 0.005, psi_solid = 0.7, psi = -1 fixed on the x-left face,
# force (5e-5, -2e-5, 0), the loop of its lines 626-659.  The script's module-level globals are
# attributes of the class here.  img_ftb131.txt / phase_ftb131.dat are not redistributed with
# the reference mount (.MISSING_LARGE_BLOBS); when absent, a seeded sphere-pack stand-in and a
# phase field with the non-wetting phase in the first 13 planes are used.
import os
import time

import numpy as np

from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase, geometry

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

time_init = time_now = time.time()
for iter in range(80000 + 1):
    lb.step()

    if (iter % 500 == 0):
        time_pre, time_now = time_now, time.time()
        diff_time, elap_time = int(time_now - time_pre), int(time_now - time_init)
        m_diff, s_diff = divmod(diff_time, 60)
        h_diff, m_diff = divmod(m_diff, 60)
        m_elap, s_elap = divmod(elap_time, 60)
        h_elap, m_elap = divmod(m_elap, 60)
        print('----------Time between two outputs is %dh %dm %ds; elapsed time is %dh %dm %ds----------------------'
              % (h_diff, m_diff, s_diff, h_elap, m_elap, s_elap))
        print('The %dth iteration, max |v| = %g, non-wetting saturation = %.4f\n\n '
              % (iter, lb.get_max_v(), float((lb.psi.to_numpy()[lb.solid.to_numpy() == 0] < 0).mean())))

        if (iter % 10000 == 0):
            lb.export_VTK(iter)             # ./structured<iter>.vtr: Solid, rho, phase, velocity
