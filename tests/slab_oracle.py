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


# ---------------------------------------------------------------------------------------------
# two-phase: CPU stand-in for one rank's slab behind StagedHaloExchanger
# ---------------------------------------------------------------------------------------------
class _SlabTwoPhase:
    """oracle/ref_two_phase.RefTwoPhase whose colission() stops before the colour accumulation
    (:365-372), so that the neighbours' g_r, g_b can arrive in between"""

    def __new__(cls, *args, **kw):
        from oracle.ref_two_phase import RefTwoPhase

        class Split(RefTwoPhase):
            def _accumulate_colour(self):
                pass

            def accumulate_now(self):
                RefTwoPhase._accumulate_colour(self)

        return Split(*args, **kw)


class OracleSlab2P:
    """One rank's slab of the two-phase solver on the CPU: the NumPy oracle on the local lattice
    (owned planes + one ghost plane either side; np.roll wraps the two ghost planes onto each other,
    which never reaches an owned node), stepped in the three stages of the CUDA schedule
    (include/lbm3d_2phase.h) and exchanging what those stages need: stage 0 the post-collision f
    and the recoloured g_r, g_b of the boundary plane (the CUDA path sends 5 populations and the
    24-byte colour record instead), stage 1 psi.  x faces must be periodic here."""

    def __init__(self, part, case):
        assert all(face >= 2 for face, _, _ in case.flow_bc) and all(face >= 2 for face, _ in case.psi_bc), \
            "x faces must be periodic in the CPU slab test"
        self.part = part
        o = _SlabTwoPhase(part.local_nx, case.shape[1], case.shape[2])
        o.set_solid(part.local_solid(case.solid))
        o.set_psi(np.ascontiguousarray(np.take(case.psi, part.local_planes(), axis=0)))
        o.fx, o.fy, o.fz = case.force
        o.niu_l, o.niu_g, o.CapA, o.psi_solid = case.niu_l, case.niu_g, case.CapA, case.psi_solid
        o.bc_type = [0] * 6
        for face, t, rho in case.flow_bc:
            o.bc_type[face] = t
            o.bc_rho[face] = rho
        o.bc_psi_type = [0] * 6
        for face, val in case.psi_bc:
            o.bc_psi_type[face] = 1
            o.bc_psi_val[face] = val
        o.init_simulation()
        self.o = o

    # ---- backend interface of StagedHaloExchanger -------------------------------------------------
    def _fields(self, stage):
        o = self.o
        return (o.f, o.g_r, o.g_b) if stage == 0 else (o.psi[..., None],)

    def pack(self, stage, side):
        plane = 1 if side == 0 else self.o.nx - 2
        return torch.from_numpy(np.concatenate([a[plane].reshape(-1) for a in self._fields(stage)]).astype(np.float32))

    def unpack(self, stage, side, tensor):
        plane = 0 if side == 0 else self.o.nx - 1
        flat, pos = tensor.numpy(), 0
        for a in self._fields(stage):
            n = a[plane].size
            a[plane] = flat[pos:pos + n].reshape(a[plane].shape)
            pos += n

    def recv_buffer(self, stage, side):
        return torch.empty(sum(a[1].size for a in self._fields(stage)), dtype=torch.float32)

    # ---- the three stages --------------------------------------------------------------------------
    def stage(self, k):
        o = self.o
        if k in (0, 2):                   # collision of the next step: f*, g_r, g_b (needs ghost psi)
            o.colission()
        else:                             # colour accumulation, stream, BCs, macro, psi (needs ghost f*, g)
            o.accumulate_now()
            o.streaming1()
            o.Boundary_condition()
            o.streaming3()
            o.Boundary_condition_psi()

    def owned(self, name):
        return self.part.owned(getattr(self.o, name))
