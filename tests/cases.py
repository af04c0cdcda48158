"""Shared case definitions: each case configures the oracle and the CUDA solver identically
(same geometry, parameters, boundary conditions, start state)."""
import numpy as np

FACE_SETTERS_RHO = ["set_bc_rho_x0", "set_bc_rho_x1", "set_bc_rho_y0", "set_bc_rho_y1",
                    "set_bc_rho_z0", "set_bc_rho_z1"]
FACE_SETTERS_VEL = ["set_bc_vel_x0", "set_bc_vel_x1", "set_bc_vel_y0", "set_bc_vel_y1",
                    "set_bc_vel_z0", "set_bc_vel_z1"]


class Case:
    def __init__(self, name, solid, bc=(), force=None, niu=None, perturb=0.0, tau_mode="class",
                 force_field=None, guo_mode="class"):
        self.name = name
        self.guo_mode = guo_mode
        self.force_field = force_field   # (nx,ny,nz,3): per-node force (cal_local_force override)
        self.solid = np.ascontiguousarray(solid, dtype=np.int8)
        self.shape = self.solid.shape
        self.bc = list(bc)              # (face, "rho", value) | (face, "vel", [vx,vy,vz])
        self.force = force
        self.niu = niu
        self.perturb = perturb
        self.tau_mode = tau_mode

    # ---- oracle -------------------------------------------------------------------------
    def make_oracle(self, cls, **kw):
        o = cls(*self.shape, tau_mode=self.tau_mode, guo_mode=self.guo_mode, **kw)
        o.set_solid(self.solid)
        for face, kind, val in self.bc:
            (o.set_bc_rho if kind == "rho" else o.set_bc_vel)(face, val)
        if self.force is not None:
            o.set_force(self.force)
        if self.niu is not None:
            o.set_viscosity(self.niu)
        if self.force_field is not None:
            o.set_force_field(self.force_field)
        o.init_simulation()
        if self.perturb:
            fl = o.solid == 0
            o.F[fl] = self.start_F(o.F.dtype)[fl]
            o.streaming3()                       # rho, v consistent with the perturbed F
        return o

    def start_F(self, dtype=np.float32):
        from taichi_lbm3d_b200.constants import W
        noise = np.random.default_rng(7).standard_normal(self.shape + (19,))
        F = (W[None, None, None, :].astype(np.float32)
             * (1 + np.float32(self.perturb) * noise.astype(np.float32))).astype(np.float32)
        return F.astype(dtype)

    # ---- CUDA solver -----------------------------------------------------------------------
    def make_solver(self, sparse=False, strict=False):
        """sparse: False (dense), True (compact list, two buffers), "aa" (compact list, in place),
        "daa" (dense, in place)"""
        from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase
        lb = LB3D_Solver_Single_Phase(*self.shape, sparse_storage=sparse in (True, "aa"), strict=strict,
                                      tau_mode=self.tau_mode, in_place=sparse in ("aa", "daa"),
                                      guo_mode=self.guo_mode)
        lb.solid.from_numpy(self.solid)
        for face, kind, val in self.bc:
            getattr(lb, (FACE_SETTERS_RHO if kind == "rho" else FACE_SETTERS_VEL)[face])(val)
        if self.force is not None:
            lb.set_force(self.force)
        if self.niu is not None:
            lb.set_viscosity(self.niu)
        if self.force_field is not None:
            lb.set_force_field(self.force_field)
        lb.init_simulation()
        return lb

    def apply_start(self, lb, oracle):
        """Give the solver the oracle's (perturbed) start state, as a user would via from_numpy."""
        if self.perturb:
            lb.F.from_numpy(oracle.F)
            lb.rho.from_numpy(oracle.rho)
            lb.v.from_numpy(oracle.v)


def random_porous(shape=(12, 10, 9), frac=0.35, seed=3):
    return (np.random.default_rng(seed).random(shape) < frac).astype(np.int8)


def case_mixed_bc(shape=(12, 10, 9)):
    """every BC type at once, force, perturbed start, overlapping faces on edges"""
    return Case("mixed_bc", random_porous(shape), bc=[(0, "rho", 1.0), (1, "rho", 0.99),
                (4, "vel", [0.01, 0.02, 0.0]), (3, "rho", 1.01)], force=[1e-4, -2e-5, 3e-5],
                perturb=1e-3)


def case_periodic_force(shape=(11, 7, 13)):
    return Case("periodic_force", random_porous(shape, 0.3, 11), force=[1e-5, 2e-5, -1e-5], perturb=1e-3)


def case_force_field(shape=(11, 9, 13)):
    """buoyancy-like per-node force (what Phase_change/...Solute_Solver.py:185-190 returns from
    its cal_local_force override): uniform part + a smooth field + noise, pressure faces in z"""
    rng = np.random.default_rng(17)
    x, y, z = np.meshgrid(*[np.arange(n) for n in shape], indexing='ij')
    ff = np.zeros(shape + (3,), np.float32)
    ff[..., 0] = 1e-5 + 2e-6 * np.sin(2 * np.pi * y / shape[1])
    ff[..., 1] = -3e-6 * np.cos(2 * np.pi * z / shape[2])
    ff[..., 2] = 4e-6 * rng.standard_normal(shape)
    return Case("force_field", random_porous(shape, 0.3, 23), bc=[(4, "rho", 1.0), (5, "rho", 0.995)],
                force_field=ff, perturb=1e-3)


def case_other_copy(shape=(10, 9, 12)):
    """the physics of the solver's other copy (Phase_change/LBM_3D_SinglePhase_Solver.py:126,235):
    tau = 3 niu + 0.5 and the un-scaled Guo term"""
    return Case("other_copy", random_porous(shape, 0.25, 31), force=[2e-5, -1e-5, 5e-6], niu=0.1,
                perturb=1e-3, tau_mode="textbook", guo_mode="unscaled")


def case_all_faces():
    s = random_porous((9, 8, 10), 0.2, 21)
    return Case("all_faces", s, bc=[(0, "vel", [0.02, 0.0, 0.0]), (1, "rho", 0.98), (2, "rho", 1.0),
                (3, "vel", [0.0, -0.01, 0.01]), (4, "rho", 1.02), (5, "vel", [0.0, 0.0, 0.03])],
                perturb=5e-4)


def case_cavity(n=50):
    from taichi_lbm3d_b200.geometry import cavity
    return Case("cavity%d" % n, cavity(n, n, n), bc=[(1, "vel", [0.0, 0.0, 0.1])])


def case_poiseuille():
    g = np.zeros((5, 20, 16), np.int8)
    g[:, :, 0] = 1
    g[:, :, -1] = 1
    return Case("poiseuille", g, force=[0.0, 1e-4, 0.0], niu=0.1667)


def case_porous(n=48, seed=131):
    from taichi_lbm3d_b200.geometry import sphere_pack
    g = sphere_pack(n, n, n, 0.70, 3.0, 6.0, seed=seed, periodic=False)
    return Case("porous%d" % n, g, bc=[(0, "rho", 1.0), (1, "rho", 0.99)])


def rel_linf(a, b):
    """relative L-infinity distance: max|a-b| / max|b| (the parity metric of BASELINE.json)."""
    d = float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
    s = float(np.abs(np.asarray(b, np.float64)).max())
    return d / s if s > 0 else d


TOL = 1e-5      # relative L-infinity bar of BASELINE.json (populations, rho, v; fp32, 1000 steps)


def v_abs_tolerance(o32, o64):
    """Absolute tolerance on v for the production (re-ordered) arithmetic.

    v = sum_s e_s F_s / rho is a difference of O(0.1) populations, so in fp32 it carries an
    ABSOLUTE round-off of ~1e-7 whatever its magnitude.  Where max|v| is large (cavity lid
    0.1) the bar is the plain relative 1e-5.  In creeping porous-media flow (max|v| ~ 1e-3)
    two fp32 evaluations that differ only in summation order -- e.g. the oracle built with
    and without -ffast-math, or Taichi's own fast_math reassociation -- already differ by
    more than 1e-5 * max|v|, so there the bar is the oracle's own fp32 round-off, measured
    against its fp64 form ("accept if the distance is no larger than the fp32
    oracle's distance to the fp64 oracle"), with a factor 2.  Verification mode
    (strict=True) is held to bit-identity instead and needs none of this.
    """
    fl = o32.solid == 0
    scale = float(np.abs(o32.v[fl]).max())
    roundoff = float(np.abs(o32.v[fl].astype(np.float64) - o64.v[fl]).max())
    return max(TOL * scale, 2.0 * roundoff)
