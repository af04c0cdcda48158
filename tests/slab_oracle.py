"""CPU stand-in for one rank's slab, built on the NumPy oracle (test infrastructure).

It exposes the backend interface of taichi_lbm3d_b200.multi_gpu.HaloExchanger (pack / unpack /
recv_buffer) so the gloo tests drive the PRODUCT's partition and exchange code with a CPU
stepper.  x faces must be periodic here (y/z face BCs and forces are fine); x-face BCs across
slabs are covered on the GPU against the single-domain solver.
"""
import numpy as np
import torch

from oracle.ref_single_phase import RefSinglePhase, pull_stream
from taichi_lbm3d_b200.constants import CROSS_LEFT, CROSS_RIGHT


class OracleSlab:
    def __init__(self, part, case):
        assert all(face >= 2 for face, _, _ in case.bc), "x faces must be periodic in the CPU slab test"
        self.part = part
        lnx = part.local_nx
        o = RefSinglePhase(lnx, case.shape[1], case.shape[2], tau_mode=case.tau_mode)
        o.set_solid(part.local_solid(case.solid))
        for face, kind, val in case.bc:
            (o.set_bc_rho if kind == "rho" else o.set_bc_vel)(face, val)
        if case.force is not None:
            o.set_force(case.force)
        if case.niu is not None:
            o.set_viscosity(case.niu)
        o.init_simulation()
        if case.perturb:
            Fg = case.start_F()
            fl = o.solid == 0
            o.F[fl] = np.take(Fg, part.local_planes(), axis=0)[fl]
            o.streaming3()
        self.o = o

    # ---- backend interface of HaloExchanger --------------------------------------------------
    def pack(self, side, which=0):
        o = self.o
        plane = 1 if side == 0 else o.nx - 2
        dirs = CROSS_LEFT if side == 0 else CROSS_RIGHT
        return torch.from_numpy(np.ascontiguousarray(o.f[plane][..., dirs]))

    def unpack(self, side, tensor, which=0):
        o = self.o
        plane = 0 if side == 0 else o.nx - 1
        dirs = CROSS_RIGHT if side == 0 else CROSS_LEFT
        o.f[plane][..., dirs] = tensor.numpy().reshape(o.ny, o.nz, 5)

    def recv_buffer(self, side):
        o = self.o
        return torch.empty((o.ny, o.nz, 5), dtype=torch.float32)

    # ---- stepping in the fused order of the CUDA pipeline ------------------------------------------
    def begin(self):
        self.o.colission()

    def stream_bc_macro(self):
        o = self.o
        o.F[...] = pull_stream(o.f, o.solid)
        o.Boundary_condition()
        o.streaming3()

    def collide(self):
        self.o.colission()

    def owned(self, name):
        return self.part.owned(getattr(self.o, name))
