"""Two-phase case definitions shared by the oracle and the CUDA solver."""
import numpy as np

_FACE = ("x_left", "x_right", "y_left", "y_right", "z_left", "z_right")
_SFX = ("xl", "xr", "yl", "yr", "zl", "zr")


class Case2P:
    def __init__(self, name, solid, psi, flow_bc=(), psi_bc=((0, -1.0),), force=(5e-5, -2e-5, 0.0),
                 niu_l=0.05, niu_g=0.2, CapA=0.005, psi_solid=0.7):
        self.name = name
        self.solid = np.ascontiguousarray(solid, np.int8)
        self.psi = np.ascontiguousarray(psi, np.float32)
        self.shape = self.solid.shape
        self.flow_bc = list(flow_bc)      # (face, type, rho)
        self.psi_bc = list(psi_bc)        # (face, value): constant-psi faces; all others periodic
        self.force, self.niu_l, self.niu_g, self.CapA, self.psi_solid = force, niu_l, niu_g, CapA, psi_solid

    def make_oracle(self, cls, **kw):
        o = cls(*self.shape, **kw)
        o.set_solid(self.solid)
        o.set_psi(self.psi)
        o.fx, o.fy, o.fz = self.force
        o.niu_l, o.niu_g, o.CapA, o.psi_solid = self.niu_l, self.niu_g, self.CapA, self.psi_solid
        o.bc_type = [0] * 6
        for face, t, rho in self.flow_bc:
            o.bc_type[face] = t
            o.bc_rho[face] = rho
        o.bc_psi_type = [0] * 6
        for face, val in self.psi_bc:
            o.bc_psi_type[face] = 1
            o.bc_psi_val[face] = val
        o.init_simulation()
        return o

    def make_solver(self, strict=False, sparse=False):
        from taichi_lbm3d_b200 import LB3D_Solver_Two_Phase
        lb = LB3D_Solver_Two_Phase(*self.shape, strict=strict, sparse_storage=sparse)
        lb.solid.from_numpy(self.solid)
        lb.psi.from_numpy(self.psi)
        lb.set_force(self.force)
        lb.niu_l, lb.niu_g, lb.CapA, lb.psi_solid = self.niu_l, self.niu_g, self.CapA, self.psi_solid
        for face in range(6):
            setattr(lb, "bc_" + _FACE[face], 0)
            setattr(lb, "bc_psi_" + _FACE[face], 0)
        for face, t, rho in self.flow_bc:
            setattr(lb, "bc_" + _FACE[face], t)
            setattr(lb, "rho_bc" + _SFX[face], rho)
        for face, val in self.psi_bc:
            lb.set_bc_psi(face, val)
        lb.init_simulation()
        return lb


def _slab_psi(shape, cut):
    psi = np.ones(shape, np.float32)
    psi[:cut] = -1.0
    return psi


def case_drainage(shape=(14, 10, 9), seed=3):
    """config-4 shape in small: porous block, psi=-1 slab entering from x0 (constant-psi face)"""
    solid = (np.random.default_rng(seed).random(shape) < 0.3).astype(np.int8)
    return Case2P("drainage", solid, _slab_psi(shape, shape[0] // 3))


def case_bcs(shape=(12, 9, 10), seed=8):
    """pressure faces in x, the script's velocity form on z1, constant psi on x0 and y1"""
    solid = (np.random.default_rng(seed).random(shape) < 0.25).astype(np.int8)
    return Case2P("bcs", solid, _slab_psi(shape, 4), flow_bc=[(0, 1, 1.0), (1, 1, 0.995), (5, 2, 1.0)],
                  psi_bc=[(0, -1.0), (3, 1.0)], force=(0.0, 0.0, 0.0), niu_l=0.1, niu_g=0.1)


def case_periodic_bubble(shape=(16, 16, 16)):
    """fully periodic box, no solid: a droplet of phase -1 (tests surface tension + recolouring)"""
    x, y, z = np.meshgrid(*[np.arange(n) for n in shape], indexing='ij')
    r = np.sqrt((x - 7.5) ** 2 + (y - 7.5) ** 2 + (z - 7.5) ** 2)
    psi = np.where(r < 4.5, -1.0, 1.0).astype(np.float32)
    return Case2P("bubble", np.zeros(shape, np.int8), psi, psi_bc=(), force=(1e-5, 0.0, 0.0))
