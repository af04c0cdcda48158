"""CPU restatement (oracle) of the reference single-phase D3Q19 MRT time step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
legs may import it, and only as the checker.  The product path
(``taichi_lbm3d_b200``) never imports this module and has no CPU fallback.

PARITY PIN.  The reference (yjhp1016/taichi_LBM3D) ships no tests, no golden vectors
and no known-answer fixtures, and its runtime (Taichi) is not installable in this
image.  The pin is the reference's OWN SOURCE: ``Single_phase/LBM_3D_SinglePhase_Solver.py``
is imported unmodified and executed through a pure-Python stand-in for the Taichi
constructs it uses (``tests/taichi_shim``; ``tests/golden/make_reference_fixtures.py``
writes ``tests/golden/ref_sp_*.npz``), and this restatement reproduces those outputs
BIT FOR BIT (``tests/test_reference_pin.py``: pressure, velocity and periodic faces, body
force, non-default viscosity).  What the stand-in cannot know is how real Taichi's LLVM
backend reassociates or contracts under its default ``fast_math=True``; against real
Taichi output a round-off tolerance would remain.  The restatement follows the reference
statement by statement and keeps its four-pass structure (collide -> push-stream -> face
BCs -> macro) and every quirk of the code (tau = niu/3 + 0.5, the /3 and /9 in the Guo
term, equilibrium-overwrite boundary conditions, m3/m5/m7 = u rather than rho*u).
Further pins in tests/: rest state, mass conservation, push == pull, analytic Poiseuille
with the effective force f/9, and the bit-identity between this NumPy form, the
pure-Python loop form below and the C form in ``ref_single_phase.c``.

Evaluation order.  Taichi's default ``fast_math=True`` leaves the summation order
of ``M @ F`` and ``.sum()`` unspecified; here every reduction runs in ascending
index order, one rounding per operation, in ``dtype`` (float32 to mirror ti.f32,
float64 as the round-off yardstick).

All line numbers cite ``/root/reference/Single_phase/LBM_3D_SinglePhase_Solver.py``.
"""
import numpy as np

# --- constants: :64-110 (M, inv_M, LR) and :183-197 (e, w) -------------------------
M_INT = np.array([
    [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1],
    [-1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1],
    [1, -2, -2, -2, -2, -2, -2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1],
    [0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0],
    [0, -2, 2, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0],
    [0, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1],
    [0, 0, 0, -2, 2, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1],
    [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1],
    [0, 0, 0, 0, 0, -2, 2, 0, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1],
    [0, 2, 2, -1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2],
    [0, -2, -2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2],
    [0, 0, 0, 1, 1, -1, -1, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0],
    [0, 0, 0, -1, -1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1],
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, -1, 1, -1, 1, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 0, 0, -1, 1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1],
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, -1, 1, -1, 1, 1, -1]], dtype=np.int64)

# :85
LR = np.array([0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15, 18, 17])

# :183-187
E = np.array([[0, 0, 0],
              [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
              [1, 1, 0], [-1, -1, 0], [1, -1, 0], [-1, 1, 0],
              [1, 0, 1], [-1, 0, -1], [1, 0, -1], [-1, 0, 1],
              [0, 1, 1], [0, -1, -1], [0, 1, -1], [0, -1, 1]], dtype=np.int64)

# :195-197 (Python float64 values, rounded to the field dtype on store)
W64 = np.array([1.0 / 3.0] + [1.0 / 18.0] * 6 + [1.0 / 36.0] * 12)


def inv_M64():
    """:83  inv_M_np = np.linalg.inv(M_np)  (float64; cast to f32 by the field store :110)."""
    return np.linalg.inv(M_INT.astype(np.float64))


def relaxation_rates(niu, tau_mode="class"):
    """:126-131.  tau_mode="class" is the class solver (niu/3+0.5); "textbook" is the
    commented-out line :126 that every other copy of the solver uses (3*niu+0.5)."""
    if tau_mode == "class":
        tau_f = niu / 3.0 + 0.5
    elif tau_mode == "textbook":
        tau_f = 3.0 * niu + 0.5
    else:
        raise ValueError(tau_mode)
    s_v = 1.0 / tau_f
    s_other = 8.0 * (2.0 - s_v) / (8.0 - s_v)
    return np.array([0, s_v, s_v, 0, s_other, 0, s_other, 0, s_other, s_v, s_v, s_v,
                     s_v, s_v, s_v, s_v, s_other, s_other, s_other], dtype=np.float64)


class RefSinglePhase:
    """Structure-faithful restatement of class LB3D_Solver_Single_Phase (:9-481).

    Field layout mirrors what the Taichi fields expose through ``to_numpy``:
    ``f, F`` -> (nx, ny, nz, 19); ``rho`` -> (nx, ny, nz); ``v`` -> (nx, ny, nz, 3);
    ``solid`` -> (nx, ny, nz) int8, C order.  Dense storage semantics (:31-35): solid
    nodes keep f = F = w, rho = 1, v = 0.
    """

    def __init__(self, nx, ny, nz, dtype=np.float32, tau_mode="class", guo_mode="class", vel_bc_mode="class"):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.dtype = np.dtype(dtype)
        self.tau_mode = tau_mode
        self.guo_mode = guo_mode            # "unscaled": Phase_change/LBM_3D_SinglePhase_Solver.py:235
        self.vel_bc_mode = vel_bc_mode      # "script": Single_phase/lbm_solver_3d.py:253,268
        # :17-18
        self.fx, self.fy, self.fz = 0.0e-6, 0.0, 0.0
        self.force_field = None             # array form of an overridden cal_local_force
        self.ns = None                      # solid fraction per node: Grey_Scale/lbm_solver_3d_Macro_Sukop.py:42
        self.niu = 0.16667
        # :23-28   [type, rho, vx, vy, vz] per face, order x0,x1,y0,y1,z0,z1
        self.bc_type = [0] * 6
        self.bc_rho = [1.0] * 6
        self.bc_vel = [[0.0, 0.0, 0.0] for _ in range(6)]
        dt = self.dtype
        self.f = np.zeros((nx, ny, nz, 19), dt)
        self.F = np.zeros((nx, ny, nz, 19), dt)
        self.rho = np.zeros((nx, ny, nz), dt)
        self.v = np.zeros((nx, ny, nz, 3), dt)
        self.solid = np.zeros((nx, ny, nz), np.int8)
        self.M = M_INT.astype(dt)
        self.inv_M = inv_M64().astype(dt)           # :110
        self.w = W64.astype(dt)                     # :195-197
        self.e_f = E.astype(dt)                     # :189-193

    # ---- API mirrors -------------------------------------------------------------
    def init_geo(self, filename):
        """:173-177"""
        in_dat = np.loadtxt(filename)
        in_dat[in_dat > 0] = 1
        in_dat = np.reshape(in_dat, (self.nx, self.ny, self.nz), order='F')
        self.solid[...] = in_dat.astype(np.int8)

    def set_solid(self, arr):
        self.solid[...] = (np.asarray(arr) > 0).astype(np.int8)

    def set_grey_scale(self, ns):
        """Grey_Scale/lbm_solver_3d_Macro_Sukop.py:338-343: the solid fraction of every node; nodes
        with int(ns) >= 1 are solid.  Switches the streaming to that script's partial bounce-back."""
        a = np.asarray(ns)
        self.ns = np.ascontiguousarray(a.astype(self.dtype))
        self.solid[...] = (a.astype(int) > 0).astype(np.int8)

    def set_bc_vel(self, face, vel):      # :405-427
        self.bc_type[face] = 2
        self.bc_vel[face] = [float(vel[0]), float(vel[1]), float(vel[2])]

    def set_bc_rho(self, face, rho):      # :429-451
        self.bc_type[face] = 1
        self.bc_rho[face] = float(rho)

    def set_viscosity(self, niu):         # :454
        self.niu = niu

    def set_force(self, force):           # :457
        self.fx, self.fy, self.fz = force[0], force[1], force[2]

    def set_force_field(self, force):
        """what a subclass overriding cal_local_force(i,j,k) (:217-220) returns, as an array
        (nx,ny,nz,3); None = the uniform force"""
        self.force_field = None if force is None else np.ascontiguousarray(np.asarray(force).astype(self.dtype))

    def cal_local_force(self, fl):
        """:217-220 for every fluid node: (n,3) or the uniform (1,3)"""
        if getattr(self, "force_field", None) is not None:
            return self.force_field[fl]
        return self.ext_f[None, :]

    def init_simulation(self):
        """:118-149 then init() :160-170 (dense branch: every node)."""
        dt = self.dtype
        self.S = relaxation_rates(self.niu, self.tau_mode).astype(dt)   # :127-131
        self.ext_f = np.array([self.fx, self.fy, self.fz]).astype(dt)   # :134-136
        self.force_flag = int(abs(self.fx) > 0 or abs(self.fy) > 0 or abs(self.fz) > 0
                              or getattr(self, "force_field", None) is not None)
        self.rho[...] = 1.0
        self.v[...] = 0.0
        for s in range(19):
            val = self._feq(s, dt.type(1.0), np.zeros(3, dt))
            self.f[..., s] = val
            self.F[..., s] = val

    # ---- helpers -----------------------------------------------------------------
    def _feq(self, k, rho_local, u):
        """:152-158.  u has a trailing axis of 3; e[k].dot(u) in ascending component order."""
        dt = self.dtype.type
        e = self.e_f[k]
        eu = e[0] * u[..., 0] + e[1] * u[..., 1] + e[2] * u[..., 2]
        uv = u[..., 0] * u[..., 0] + u[..., 1] * u[..., 1] + u[..., 2] * u[..., 2]
        return self.w[k] * rho_local * (dt(1.0) + dt(3.0) * eu + dt(4.5) * eu * eu - dt(1.5) * uv)

    def _matvec(self, A, x):
        """A @ x per node, x: (n,19); ascending-l accumulation, zeros skipped (exact)."""
        out = np.zeros_like(x)
        for s in range(19):
            acc = np.zeros(x.shape[0], self.dtype)
            for l in range(19):
                if A[s, l] != 0:
                    acc = acc + A[s, l] * x[:, l]
            out[:, s] = acc
        return out

    def _meq(self, rho, u):
        """:209-215"""
        dt = self.dtype.type
        out = np.zeros((rho.shape[0], 19), self.dtype)
        ux, uy, uz = u[:, 0], u[:, 1], u[:, 2]
        out[:, 0] = rho
        out[:, 3] = ux
        out[:, 5] = uy
        out[:, 7] = uz
        out[:, 1] = ux * ux + uy * uy + uz * uz
        out[:, 9] = dt(2) * ux * ux - uy * uy - uz * uz
        out[:, 11] = uy * uy - uz * uz
        out[:, 13] = ux * uy
        out[:, 14] = uy * uz
        out[:, 15] = ux * uz
        return out

    # ---- the four passes -----------------------------------------------------------
    def colission(self):
        """:222-241"""
        dt = self.dtype.type
        fl = self.solid == 0
        Fn = self.F[fl]                      # (n,19)
        rho = self.rho[fl]
        v = self.v[fl]
        m = self._matvec(self.M, Fn)                           # :226
        meq = self._meq(rho, v)                                # :227
        m = m - self.S[None, :] * (m - meq)                    # :228
        if self.force_flag == 1:                               # :230-238
            f = self.cal_local_force(fl)                       # :231
            for s in range(19):
                f_guo = np.zeros(Fn.shape[0], self.dtype)
                for l in range(19):
                    if self.M[s, l] == 0:
                        continue            # term multiplied by M[s,l]==0 contributes exactly 0
                    e = self.e_f[l]
                    emv_f = (e[0] - v[:, 0]) * f[:, 0] + (e[1] - v[:, 1]) * f[:, 1] + (e[2] - v[:, 2]) * f[:, 2]
                    ev = e[0] * v[:, 0] + e[1] * v[:, 1] + e[2] * v[:, 2]
                    ef = e[0] * f[:, 0] + e[1] * f[:, 1] + e[2] * f[:, 2]
                    term = emv_f + (ev * ef) if self.guo_mode == "unscaled" else emv_f / dt(3.0) + (ev * ef) / dt(9.0)
                    f_guo = f_guo + self.w[l] * term * self.M[s, l]
                m[:, s] = m[:, s] + (dt(1) - dt(0.5) * self.S[s]) * f_guo
        self.f[fl] = self._matvec(self.inv_M, m)               # :240-241

    def streaming1(self):
        """:259-268  push with periodic wrap (:247-257) and half-way bounce-back."""
        fluid = self.solid == 0
        for s in range(19):
            ex, ey, ez = (int(c) for c in E[s])
            # value arriving at ip = i + e_s (wrapped) is f[i][s]
            arriving = np.roll(self.f[..., s], (ex, ey, ez), axis=(0, 1, 2))
            src_fluid = np.roll(fluid, (ex, ey, ez), axis=(0, 1, 2))
            take = src_fluid & fluid                         # i fluid and ip fluid, viewed at ip
            self.F[..., s][take] = arriving[take]
            # i fluid, ip solid: F[i][LR[s]] = f[i][s]
            nb_solid = np.roll(~fluid, (-ex, -ey, -ez), axis=(0, 1, 2))   # solid[ip] viewed at i
            bounce = fluid & nb_solid
            self.F[..., LR[s]][bounce] = self.f[..., s][bounce]

    def streaming_grey(self):
        """Grey_Scale/lbm_solver_3d_Macro_Sukop.py:233-247 (streaming0 + streaming1): every fluid node
        blends its post-collision population s with the opposite one of the node it moves to,
            f2[i][s] = f[i][s] + ns[i] * (f[i + e_s][LR[s]] - f[i][s]),
        and pushes f2[i][s] to i + e_s whatever that node is -- no bounce-back.  Solid nodes are
        never collided (f = w for ever) and never push, so F[i][s] of a fluid node whose source
        i - e_s is solid keeps its previous value."""
        fluid = self.solid == 0
        for s in range(19):
            ex, ey, ez = (int(c) for c in E[s])
            fs = self.f[..., s]
            opp = np.roll(self.f[..., LR[s]], (-ex, -ey, -ez), axis=(0, 1, 2))    # f[ip][LR[s]] viewed at i
            f2 = fs + self.ns * (opp - fs)
            arriving = np.roll(f2, (ex, ey, ez), axis=(0, 1, 2))
            src_fluid = np.roll(fluid, (ex, ey, ez), axis=(0, 1, 2))
            self.F[..., s][src_fluid] = arriving[src_fluid]

    def Boundary_condition(self):
        """:272-370  faces in order x0,x1,y0,y1,z0,z1; later faces overwrite earlier."""
        dt = self.dtype.type
        n = (self.nx, self.ny, self.nz)
        for face in range(6):
            t = self.bc_type[face]
            if t == 0:
                continue
            axis, side = face // 2, face % 2
            idx = [slice(None)] * 3
            idx_in = [slice(None)] * 3
            idx[axis] = 0 if side == 0 else n[axis] - 1
            idx_in[axis] = 1 if side == 0 else n[axis] - 2
            idx, idx_in = tuple(idx), tuple(idx_in)
            fl = self.solid[idx] == 0
            if t == 1:                                        # :274-281 etc.
                v_face = self.v[idx]
                v_in = self.v[idx_in]
                use_in = self.solid[idx_in] > 0
                u = np.where(use_in[..., None], v_in, v_face)
                for s in range(19):
                    val = self._feq(s, dt(self.bc_rho[face]), u)
                    Fs = self.F[idx + (s,)]
                    Fs[fl] = val[fl]
            elif self.vel_bc_mode == "script":
                # Single_phase/lbm_solver_3d.py:253: F[s] = feq(LR[s],1,u) - F[LR[s]] + feq(s,1,u), in
                # place for s = 0..18 -- for LR[s] < s the F[LR[s]] read is the value just written
                u = np.array(self.bc_vel[face]).astype(self.dtype)
                for s in range(19):
                    Fs = self.F[idx + (s,)]
                    Fo = self.F[idx + (int(LR[s]),)]
                    val = self._feq(int(LR[s]), dt(1.0), u) - Fo + self._feq(s, dt(1.0), u)
                    Fs[fl] = val[fl]
            else:                                             # :283-288 etc.
                u = np.array(self.bc_vel[face]).astype(self.dtype)
                for s in range(19):
                    val = self._feq(s, dt(1.0), u)
                    Fs = self.F[idx + (s,)]
                    Fs[fl] = val

    def streaming3(self):
        """:372-392"""
        dt = self.dtype.type
        fl = self.solid == 0
        self.f[fl] = self.F[fl]
        Fn = self.F[fl]
        rho = np.zeros(Fn.shape[0], self.dtype)
        for s in range(19):
            rho = rho + Fn[:, s]
        v = np.zeros((Fn.shape[0], 3), self.dtype)
        for s in range(19):
            for c in range(3):
                if E[s, c] != 0:
                    v[:, c] = v[:, c] + self.e_f[s, c] * Fn[:, s]
        v = v / rho[:, None]
        fvec = self.cal_local_force(fl)                        # :385
        v = v + (fvec / dt(2)) / rho[:, None]
        self.rho[fl] = rho
        self.v[fl] = v
        self.rho[~fl] = 1.0
        self.v[~fl] = 0.0

    def step(self):
        """:477-481"""
        self.colission()
        if self.ns is not None:
            self.streaming_grey()
        else:
            self.streaming1()
        self.Boundary_condition()
        self.streaming3()

    def get_max_v(self):
        """:394-402"""
        nrm = np.sqrt(self.v[..., 0] * self.v[..., 0] + self.v[..., 1] * self.v[..., 1]
                      + self.v[..., 2] * self.v[..., 2])
        return float(max(nrm.max(), -1e10))

    # ---- slow literal forms, for tiny cases only ----------------------------------------
    def streaming1_loops(self):
        """:259-268 as literal per-node Python loops (the push exactly as written)."""
        nx, ny, nz = self.nx, self.ny, self.nz
        for i in range(nx):
            for j in range(ny):
                for k in range(nz):
                    if self.solid[i, j, k] != 0:
                        continue
                    for s in range(19):
                        ip = [i + int(E[s, 0]), j + int(E[s, 1]), k + int(E[s, 2])]
                        if ip[0] < 0: ip[0] = nx - 1
                        if ip[0] > nx - 1: ip[0] = 0
                        if ip[1] < 0: ip[1] = ny - 1
                        if ip[1] > ny - 1: ip[1] = 0
                        if ip[2] < 0: ip[2] = nz - 1
                        if ip[2] > nz - 1: ip[2] = 0
                        if self.solid[ip[0], ip[1], ip[2]] == 0:
                            self.F[ip[0], ip[1], ip[2], s] = self.f[i, j, k, s]
                        else:
                            self.F[i, j, k, LR[s]] = self.f[i, j, k, s]


def pull_stream(f, solid):
    """Pull restatement of :259-268 used to pin push == pull (SURVEY 8a'.2):
    F[i][s] = f[i - e_s][s] if that node is fluid else f[i][LR[s]] (fluid i only)."""
    fluid = solid == 0
    F = np.array(f, copy=True)
    for s in range(19):
        ex, ey, ez = (int(c) for c in E[s])
        from_nb = np.roll(f[..., s], (ex, ey, ez), axis=(0, 1, 2))
        nb_fluid = np.roll(fluid, (ex, ey, ez), axis=(0, 1, 2))
        F[..., s] = np.where(fluid & nb_fluid, from_nb, np.where(fluid, f[..., LR[s]], f[..., s]))
    return F
