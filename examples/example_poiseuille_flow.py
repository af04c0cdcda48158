"""Plane Poiseuille flow, the case of the reference's Single_phase/example_poiseuille_flow.py:
5 x 20 x 16 lattice, walls on z = 0 and z = 15, body force 1e-4 along y, niu = 0.1667.  The only
case of the reference with a closed-form answer: with the class's constants the steady momentum
profile is the parabola g/(2 nu) (z - 1/2)(29/2 - z) with g = f/9 and nu = niu/9
(tests/test_oracle.py::test_poiseuille_known_answer)."""
import numpy as np

import LBM_3D_SinglePhase_Solver as lb3dsp
from _progress import Progress

STEPS = 20000                             # the reference runs 150 000
NX, NY, NZ = 5, 20, 16

walls = np.zeros((NX, NY, NZ))
walls[:, :, [0, NZ - 1]] = 1

solver = lb3dsp.LB3D_Solver_Single_Phase(nx=NX, ny=NY, nz=NZ)
solver.solid.from_numpy(walls)
solver.set_force([0.0, 1.0e-4, 0.0])
solver.set_viscosity(0.1667)
solver.init_simulation()

progress = Progress()
for step in range(STEPS + 1):
    solver.step()
    if step % 2000 == 0:
        progress.report(step, max_v=solver.get_max_v())
    if step % 10000 == 0:
        solver.export_VTK(step)

z = np.arange(1, NZ - 1)
profile = solver.v.to_numpy()[NX // 2, NY // 2, 1:NZ - 1, 1]
analytic = (1.0e-4 / 9.0) / (2 * 0.1667 / 9.0) * (z - 0.5) * (NZ - 1.5 - z)
print("centre-line v_y %.6g, parabola %.6g" % (profile[len(z) // 2], analytic[len(z) // 2]))
