"""CPU restatement (oracle) of the reference two-phase colour-gradient step.

TEST INFRASTRUCTURE ONLY (see oracle/ref_single_phase.py: only tests/, smoke() and bench.py's
CPU legs may use anything under oracle/).  PARITY PIN: the reference has no tests or golden
vectors and Taichi cannot run in this image; the kernels of the reference script itself are
executed through tests/taichi_shim with only its hand-edited parameter lines replaced
(tests/golden/make_reference_fixtures.py -> tests/golden/ref_tp_*.npz).  The script accumulates
rho_r / rho_b with float atomics whose order is open (:365-372): with them executed in node
order (accumulate_colour_in_push_order, what the shim's sequential run does) this restatement
reproduces the script BIT FOR BIT; in its own deterministic order (_accumulate_colour, the
order of the CUDA path) it differs by that summation order only (tests/test_reference_pin.py).

Follows ``2phase/lbm_solver_3d_2phase.py`` (the dense script; ``..._sparse.py`` differs only in
allocation) statement by statement; line numbers below cite that file.  The script is a
module of globals; here they are attributes of ``RefTwoPhase``.  Quirks kept as written:

  * relaxation from wl = 1/(3 niu_l + 1/2), wg = 1/(3 niu_g + 1/2) (:103-104) with the quadratic
    blend for |psi| <= 0.1 (:278-299);
  * the Guo force term is un-scaled and applied every step, all 19 moments (:242-247, :331);
  * psi = rho_r - rho_b/(rho_r+rho_b) -- operator precedence as written (:605);
  * velocity boundary faces use ``bc_vel_*[None]`` fields that static_init never writes
    (:223-228 assigns kernel-local names), i.e. u_bc = 0, and update F[s] in place for
    s = 0..18 so later directions see already-updated opposites (:500-504 ...);
  * Compute_C wraps on periodic psi faces and CLAMPS on constant-psi faces (:390-428) and
    vanishes next to a solid where |rho_r - rho_b| > 0.9 (:271-273).

One thing cannot be restated literally: colission pushes g_r, g_b into the neighbours'
accumulators with ``+=`` inside a parallel loop (:365-372) -- float atomics whose order is
not defined.  The oracle fixes the order as: at the destination, ascending direction index s
(the pull form; the CUDA kernel uses the same order).  ``push_colour_loops`` is the literal
sequential push, for checking that the two agree to round-off on tiny cases.
"""
import numpy as np

from .ref_single_phase import E, LR, M_INT, W64, inv_M64


class RefTwoPhase:
    def __init__(self, nx, ny, nz, dtype=np.float32):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.dtype = np.dtype(dtype)
        # :18-39 script defaults
        self.fx, self.fy, self.fz = 5.0e-5, -2e-5, 0.0
        self.niu_l, self.niu_g = 0.1, 0.1
        self.psi_solid = 0.7
        self.CapA = 0.005
        self.bc_type = [0] * 6
        self.bc_rho = [1.0, 0.995, 1.0, 1.0, 1.0, 1.0]
        self.bc_psi_type = [1, 0, 0, 0, 0, 0]
        self.bc_psi_val = [-1.0, 1.0, 1.0, 1.0, 1.0, 1.0]
        dt = self.dtype
        shp = (nx, ny, nz)
        self.f = np.zeros(shp + (19,), dt)
        self.F = np.zeros(shp + (19,), dt)
        self.rho = np.zeros(shp, dt)
        self.v = np.zeros(shp + (3,), dt)
        self.psi = np.zeros(shp, dt)
        self.rho_r = np.zeros(shp, dt)
        self.rho_b = np.zeros(shp, dt)
        self.rhor = np.zeros(shp, dt)
        self.rhob = np.zeros(shp, dt)
        self.solid = np.zeros(shp, np.int8)
        self.M = M_INT.astype(dt)
        self.inv_M = inv_M64().astype(dt)       # :140, :145
        self.w = W64.astype(dt)                 # :234-236
        self.e_f = E.astype(dt)

    # ---- setup ---------------------------------------------------------------------------
    def set_solid(self, arr):
        self.solid[...] = (np.asarray(arr) > 0).astype(np.int8)

    def set_psi(self, arr):
        self.psi[...] = np.asarray(arr).astype(self.dtype)

    def init_geo(self, filename, filename2):
        """:194-202"""
        in_dat = np.loadtxt(filename)
        in_dat[in_dat > 0] = 1
        self.solid[...] = np.reshape(in_dat, (self.nx, self.ny, self.nz), order='F').astype(np.int8)
        ph = np.loadtxt(filename2)
        self.psi[...] = np.reshape(ph, (self.nx, self.ny, self.nz), order='F').astype(self.dtype)

    def init_simulation(self):
        """derived constants :100-108 (module level, Python floats), static_init :205-228, init :173-186"""
        dt = self.dtype.type
        wl = 1.0 / (self.niu_l / (1.0 / 3.0) + 0.5)
        wg = 1.0 / (self.niu_g / (1.0 / 3.0) + 0.5)
        lg0 = 2 * wl * wg / (wl + wg)
        l1 = 2 * (wl - lg0) * 10
        l2 = -l1 / 0.2
        g1 = 2 * (lg0 - wg) * 10
        g2 = g1 / 0.2
        self.wl, self.wg, self.lg0, self.l1, self.l2, self.g1, self.g2 = (dt(x) for x in (wl, wg, lg0, l1, l2, g1, g2))
        self.ext_f = np.array([self.fx, self.fy, self.fz]).astype(self.dtype)     # :147
        fl = self.solid == 0
        self.rho[fl] = 1.0
        self.v[fl] = 0.0
        self.rho_r[fl] = (self.psi[fl] + dt(1.0)) / dt(2.0)
        self.rho_b[fl] = dt(1.0) - self.rho_r[fl]
        self.rhor[fl] = 0.0
        self.rhob[fl] = 0.0
        for s in range(19):
            self.f[..., s][fl] = self.w[s]
            self.F[..., s][fl] = self.w[s]

    # ---- helpers ---------------------------------------------------------------------------
    def _feq(self, k, rho_local, u):
        """:161-170"""
        dt = self.dtype.type
        e = self.e_f[k]
        eu = e[0] * u[..., 0] + e[1] * u[..., 1] + e[2] * u[..., 2]
        uv = u[..., 0] * u[..., 0] + u[..., 1] * u[..., 1] + u[..., 2] * u[..., 2]
        return self.w[k] * rho_local * (dt(1.0) + dt(3.0) * eu + dt(4.5) * eu * eu - dt(1.5) * uv)

    def _psi_shift(self, arr, s):
        """value of arr at periodic_index_for_psi(i + e_s) (:390-428): wrap on periodic psi faces,
        clamp on constant-psi faces."""
        out = arr
        for axis in range(3):
            d = int(E[s, axis])
            if d == 0:
                continue
            n = arr.shape[axis]
            rolled = np.roll(out, -d, axis=axis)          # rolled[i] = out[i + d] with wrap
            idx = [slice(None)] * 3
            if d > 0 and self.bc_psi_type[2 * axis + 1] != 0:     # i+d > n-1 -> clamp to n-1
                idx[axis] = n - 1
                rolled = rolled.copy()
                rolled[tuple(idx)] = out[tuple(idx)]
            if d < 0 and self.bc_psi_type[2 * axis] != 0:         # i+d < 0 -> clamp to 0
                idx[axis] = 0
                rolled = rolled.copy()
                rolled[tuple(idx)] = out[tuple(idx)]
            out = rolled
        return out

    def Compute_C(self):
        """:259-275 for every node at once; returns C (nx,ny,nz,3)"""
        dt = self.dtype.type
        C = np.zeros(self.solid.shape + (3,), self.dtype)
        ind_S = np.zeros(self.solid.shape, bool)
        for s in range(19):
            sol = self._psi_shift(self.solid, s) != 0
            val = np.where(sol, dt(self.psi_solid), self._psi_shift(self.psi, s))
            ind_S |= sol
            for c in range(3):
                # C += 3.0*w[s]*e_f[s]*psi[ip]   (component-wise, left to right)
                C[..., c] = C[..., c] + dt(3.0) * self.w[s] * self.e_f[s, c] * val
        kill = (np.abs(self.rho_r - self.rho_b) > dt(0.9)) & ind_S
        C[kill] = 0
        return C

    def Compute_S_local(self, psi):
        """:278-299 -> (sv, sother)"""
        dt = self.dtype.type
        sv_pos = np.where(psi > dt(0.1), self.wl, self.lg0 + self.l1 * psi + self.l2 * psi * psi)
        sv_neg = np.where(psi < dt(-0.1), self.wg, self.lg0 + self.g1 * psi + self.g2 * psi * psi)
        sv = np.where(psi > 0, sv_pos, sv_neg).astype(self.dtype)
        sother = dt(8.0) * (dt(2.0) - sv) / (dt(8.0) - sv)
        return sv, sother

    def _matvec(self, A, x):
        out = np.zeros_like(x)
        for s in range(19):
            acc = np.zeros(x.shape[0], self.dtype)
            for l in range(19):
                if A[s, l] != 0:
                    acc = acc + A[s, l] * x[:, l]
            out[:, s] = acc
        return out

    # ---- passes --------------------------------------------------------------------------------
    def colission(self):
        """:302-372.  Leaves the post-collision f and the recoloured g_r, g_b (self.g_r, self.g_b,
        fluid nodes) and accumulates them into rhor / rhob."""
        dt = self.dtype.type
        fl = self.solid == 0
        Call = self.Compute_C()
        C = Call[fl]
        v = self.v[fl]
        rho = self.rho[fl]
        cc = np.sqrt(C[:, 0] * C[:, 0] + C[:, 1] * C[:, 1] + C[:, 2] * C[:, 2])       # C.norm()
        pos = cc > 0
        safe = np.where(pos, cc, dt(1.0))
        normal = np.where(pos[:, None], C / safe[:, None], dt(0.0)).astype(self.dtype)
        m = self._matvec(self.M, self.F[fl])                                  # multiply_M :231-238
        ux, uy, uz = v[:, 0], v[:, 1], v[:, 2]
        meq = np.zeros_like(m)                                                # meq_vec :250-256
        meq[:, 0] = rho
        meq[:, 3], meq[:, 5], meq[:, 7] = ux, uy, uz
        meq[:, 1] = ux * ux + uy * uy + uz * uz
        meq[:, 9] = dt(2) * ux * ux - uy * uy - uz * uz
        meq[:, 11] = uy * uy - uz * uz
        meq[:, 13], meq[:, 14], meq[:, 15] = ux * uy, uy * uz, ux * uz
        nx_, ny_, nz_ = normal[:, 0], normal[:, 1], normal[:, 2]
        CapA = dt(self.CapA)
        meq[:, 1] = meq[:, 1] + CapA * cc                                      # :316-321
        meq[:, 9] = meq[:, 9] + dt(0.5) * CapA * cc * (dt(2) * nx_ * nx_ - ny_ * ny_ - nz_ * nz_)
        meq[:, 11] = meq[:, 11] + dt(0.5) * CapA * cc * (ny_ * ny_ - nz_ * nz_)
        meq[:, 13] = meq[:, 13] + dt(0.5) * CapA * cc * (nx_ * ny_)
        meq[:, 14] = meq[:, 14] + dt(0.5) * CapA * cc * (ny_ * nz_)
        meq[:, 15] = meq[:, 15] + dt(0.5) * CapA * cc * (nx_ * nz_)
        sv, so = self.Compute_S_local(self.psi[fl])                            # :323
        zero = np.zeros_like(sv)
        S = [zero, sv, sv, zero, so, zero, so, zero, so, sv, sv, sv, sv, sv, sv, sv, so, so, so]
        fo = self.ext_f
        for s in range(19):                                                    # :329-331
            m[:, s] = m[:, s] - S[s] * (m[:, s] - meq[:, s])
            guo = np.zeros(m.shape[0], self.dtype)                             # GuoF :241-247
            for l in range(19):
                if self.M[s, l] == 0:
                    continue
                e = self.e_f[l]
                emu_f = (e[0] - ux) * fo[0] + (e[1] - uy) * fo[1] + (e[2] - uz) * fo[2]
                eu = e[0] * ux + e[1] * uy + e[2] * uz
                ef = e[0] * fo[0] + e[1] * fo[1] + e[2] * fo[2]
                guo = guo + self.w[l] * (emu_f + (eu * ef)) * self.M[s, l]
            m[:, s] = m[:, s] + (dt(1) - dt(0.5) * S[s]) * guo
        self.f[fl] = self._matvec(self.inv_M, m)                               # :340-343
        rr, rb = self.rho_r[fl], self.rho_b[fl]
        g_r = np.stack([self._feq(s, rr, v) for s in range(19)], axis=1)       # :345-346
        g_b = np.stack([self._feq(s, rb, v) for s in range(19)], axis=1)
        for kk in (1, 3, 5, 7, 9, 11, 13, 15, 17):                             # :351-363
            e = self.e_f[kk]
            ef = e[0] * C[:, 0] + e[1] * C[:, 1] + e[2] * C[:, 2]
            cospsi = np.where(g_r[:, kk] < g_r[:, kk + 1], g_r[:, kk], g_r[:, kk + 1])
            cospsi = np.where(cospsi < g_b[:, kk], cospsi, g_b[:, kk])
            cospsi = np.where(cospsi < g_b[:, kk + 1], cospsi, g_b[:, kk + 1])
            cospsi = cospsi * (ef / safe)
            cospsi = np.where(pos, cospsi, dt(0.0)).astype(self.dtype)         # only if cc > 0
            g_r[:, kk] = g_r[:, kk] + cospsi
            g_r[:, kk + 1] = g_r[:, kk + 1] - cospsi
            g_b[:, kk] = g_b[:, kk] - cospsi
            g_b[:, kk + 1] = g_b[:, kk + 1] + cospsi
        self.g_r = np.zeros(self.solid.shape + (19,), self.dtype)
        self.g_b = np.zeros(self.solid.shape + (19,), self.dtype)
        self.g_r[fl] = g_r
        self.g_b[fl] = g_b
        self.C = Call
        self._accumulate_colour()

    def _accumulate_colour(self):
        """:365-372 in pull form, ascending s at the destination (see module docstring)"""
        fluid = self.solid == 0
        for name, g in (("rhor", self.g_r), ("rhob", self.g_b)):
            acc = getattr(self, name)
            for s in range(19):
                ex, ey, ez = (int(c) for c in E[s])
                from_nb = np.roll(g[..., s], (ex, ey, ez), axis=(0, 1, 2))
                nb_fluid = np.roll(fluid, (ex, ey, ez), axis=(0, 1, 2))
                contrib = np.where(nb_fluid, from_nb, g[..., LR[s]])
                acc[fluid] = acc[fluid] + contrib[fluid]

    def accumulate_colour_in_push_order(self):
        """:365-372 with the float atomics executed one after the other in node order (what a
        sequential run of the script does): drop-in for _accumulate_colour; tiny cases only"""
        rr, rb = self.push_colour_loops()
        fl = self.solid == 0
        self.rhor[fl] = self.rhor[fl] + rr[fl]
        self.rhob[fl] = self.rhob[fl] + rb[fl]

    def push_colour_loops(self):
        """:365-372 literally (sequential node order); tiny cases only.  Returns (rhor, rhob)."""
        nx, ny, nz = self.nx, self.ny, self.nz
        rhor = np.zeros_like(self.rhor)
        rhob = np.zeros_like(self.rhob)
        for i in range(nx):
            for j in range(ny):
                for k in range(nz):
                    if self.solid[i, j, k] != 0:
                        continue
                    for s in range(19):
                        ip = ((i + int(E[s, 0])) % nx, (j + int(E[s, 1])) % ny, (k + int(E[s, 2])) % nz)
                        tgt = ip if self.solid[ip] == 0 else (i, j, k)
                        rhor[tgt] += self.g_r[i, j, k, s]
                        rhob[tgt] += self.g_b[i, j, k, s]
        return rhor, rhob

    def streaming1(self):
        """:431-442 (same push as the single-phase class)"""
        fluid = self.solid == 0
        for s in range(19):
            ex, ey, ez = (int(c) for c in E[s])
            arriving = np.roll(self.f[..., s], (ex, ey, ez), axis=(0, 1, 2))
            src_fluid = np.roll(fluid, (ex, ey, ez), axis=(0, 1, 2))
            take = src_fluid & fluid
            self.F[..., s][take] = arriving[take]
            nb_solid = np.roll(~fluid, (-ex, -ey, -ez), axis=(0, 1, 2))
            bounce = fluid & nb_solid
            self.F[..., LR[s]][bounce] = self.f[..., s][bounce]

    def Boundary_condition(self):
        """:491-583"""
        dt = self.dtype.type
        n = (self.nx, self.ny, self.nz)
        zero_u = np.zeros(3, self.dtype)         # bc_vel_* fields are never written
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
            if t == 1:
                u = np.where((self.solid[idx_in] > 0)[..., None], self.v[idx_in], self.v[idx])
                for s in range(19):
                    val = self._feq(s, dt(self.bc_rho[face]), u)
                    Fs = self.F[idx + (s,)]
                    Fs[fl] = val[fl]
            else:
                for s in range(19):              # in place, ascending s (:500-504)
                    Fs = self.F[idx + (s,)]
                    Fo = self.F[idx + (int(LR[s]),)]
                    val = self._feq(int(LR[s]), dt(1.0), zero_u) - Fo + self._feq(s, dt(1.0), zero_u)
                    Fs[fl] = val[fl]

    def streaming3(self):
        """:587-605"""
        fl = self.solid == 0
        self.rho_r[fl] = self.rhor[fl]
        self.rho_b[fl] = self.rhob[fl]
        self.rhor[fl] = 0
        self.rhob[fl] = 0
        self.f[fl] = self.F[fl]
        Fn = self.F[fl]
        rho = np.zeros(Fn.shape[0], self.dtype)
        v = np.zeros((Fn.shape[0], 3), self.dtype)
        for s in range(19):                      # interleaved accumulation as written
            rho = rho + Fn[:, s]
            for c in range(3):
                if E[s, c] != 0:
                    v[:, c] = v[:, c] + self.e_f[s, c] * Fn[:, s]
        v = v / rho[:, None]
        v = v + (self.ext_f[None, :] / self.dtype.type(2)) / rho[:, None]
        self.rho[fl] = rho
        self.v[fl] = v
        rr, rb = self.rho_r[fl], self.rho_b[fl]
        self.psi[fl] = rr - rb / (rr + rb)       # precedence as written (:605)

    def Boundary_condition_psi(self):
        """:445-486"""
        dt = self.dtype.type
        n = (self.nx, self.ny, self.nz)
        for face in range(6):
            if self.bc_psi_type[face] != 1:
                continue
            axis, side = face // 2, face % 2
            idx = [slice(None)] * 3
            idx[axis] = 0 if side == 0 else n[axis] - 1
            idx = tuple(idx)
            fl = self.solid[idx] == 0
            val = dt(self.bc_psi_val[face])
            self.psi[idx][fl] = val
            self.rho_r[idx][fl] = (val + dt(1.0)) / dt(2.0)
            self.rho_b[idx][fl] = dt(1.0) - (val + dt(1.0)) / dt(2.0)

    def step(self):
        """main loop body :626-632"""
        self.colission()
        self.streaming1()
        self.Boundary_condition()
        self.streaming3()
        self.Boundary_condition_psi()
