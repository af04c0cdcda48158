"""Generates tests/golden/*.npz.

The reference ships no golden vectors and Taichi cannot run here, so these fixtures are
produced by the NumPy oracle (oracle/ref_single_phase.py, fp32, literal evaluation order);
they pin the oracle (and, through the parity tests, the CUDA path) against regressions and
travel to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py

geo_cavity_50.npz is different: it is the reference's own fixture
Single_phase/geo_cavity.dat (50^3), read with the reference's loader semantics
(init_geo :173-177) and stored bit-packed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_single_phase import RefSinglePhase  # noqa: E402
from oracle.ref_two_phase import RefTwoPhase  # noqa: E402
from tests import cases, cases2p  # noqa: E402


def main():
    for make, steps in ((cases.case_mixed_bc, 25), (cases.case_all_faces, 10), (cases.case_periodic_force, 10),
                        (cases.case_force_field, 10), (cases.case_other_copy, 10)):
        case = make()
        o = case.make_oracle(RefSinglePhase)
        F0, rho0, v0 = o.F.copy(), o.rho.copy(), o.v.copy()
        for _ in range(steps):
            o.step()
        np.savez_compressed(os.path.join(HERE, "sp_%s.npz" % case.name), solid=case.solid, steps=steps,
                            F0=F0, rho0=rho0, v0=v0, F=o.F, rho=o.rho, v=o.v)
        print(case.name, steps, float(np.abs(o.v).max()))
    # two-phase colour-gradient path (oracle/ref_two_phase.py, fp32, literal evaluation order)
    for make, steps in ((cases2p.case_drainage, 12), (cases2p.case_bcs, 12), (cases2p.case_periodic_bubble, 12)):
        case = make()
        o = case.make_oracle(RefTwoPhase)
        for _ in range(steps):
            o.step()
        np.savez_compressed(os.path.join(HERE, "tp_%s.npz" % case.name), solid=case.solid, psi0=case.psi, steps=steps,
                            F=o.F, rho=o.rho, v=o.v, psi=o.psi, rho_r=o.rho_r, rho_b=o.rho_b)
        print("two-phase", case.name, steps, float(np.abs(o.v).max()))
    ref_geo = "/root/reference/Single_phase/geo_cavity.dat"
    if os.path.exists(ref_geo):
        d = np.loadtxt(ref_geo)
        d[d > 0] = 1
        g = np.reshape(d, (50, 50, 50), order='F').astype(np.uint8)
        np.savez_compressed(os.path.join(HERE, "geo_cavity_50.npz"), packed=np.packbits(g.reshape(-1)),
                            shape=np.array([50, 50, 50]))
        print("geo_cavity", int(g.sum()))


if __name__ == "__main__":
    main()
