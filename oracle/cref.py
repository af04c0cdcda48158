"""ctypes front end of the C restatement (oracle/ref_single_phase.c, ref_two_phase.c).

TEST INFRASTRUCTURE ONLY (see oracle/ref_single_phase.py).  ``RefSinglePhaseC`` has the
same surface as ``ref_single_phase.RefSinglePhase`` but runs the C code, which is fast
enough for 131^3 x 1000-step parity runs and serves as the CPU timing baseline.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import ref_single_phase as _np_ref

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.environ.get("LBM3D_ORACLE_BUILD") or os.path.join(_HERE, "_build")      # override: a scratch build


def build(force=False):
    """Compile the C oracle (gcc, OpenMP) into oracle/_build/."""
    targets = [os.path.join(_BUILD, n) for n in ("libref_strict.so", "libref_fast.so")]
    if force or not all(os.path.exists(t) for t in targets):
        subprocess.run(["make", "-C", _HERE, "OUT=" + _BUILD] + (["-B"] if force else []), check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return targets


_libs = {}


def load(kind="strict"):
    if kind not in _libs:
        path = os.path.join(_BUILD, "libref_%s.so" % kind)
        if not os.path.exists(path):
            build()
        _libs[kind] = ctypes.CDLL(path)
    return _libs[kind]


def _params_struct(ctype):
    class P(ctypes.Structure):
        _fields_ = [("nx", ctypes.c_int), ("ny", ctypes.c_int), ("nz", ctypes.c_int),
                    ("force_flag", ctypes.c_int), ("bc_type", ctypes.c_int * 6),
                    ("S", ctype * 19), ("invM", ctype * 361), ("w", ctype * 19),
                    ("force", ctype * 3), ("bc_rho", ctype * 6), ("bc_vel", (ctype * 3) * 6),
                    ("force_field", ctypes.c_void_p), ("guo_unscaled", ctypes.c_int),
                    ("vel_bc_script", ctypes.c_int), ("ns", ctypes.c_void_p)]
    return P


_P32 = _params_struct(ctypes.c_float)
_P64 = _params_struct(ctypes.c_double)


class RefSinglePhaseC(_np_ref.RefSinglePhase):
    """Same state and setters as the NumPy oracle; the four passes run in C."""

    def __init__(self, nx, ny, nz, dtype=np.float32, tau_mode="class", kind="strict", guo_mode="class",
                 vel_bc_mode="class"):
        super().__init__(nx, ny, nz, dtype=dtype, tau_mode=tau_mode, guo_mode=guo_mode, vel_bc_mode=vel_bc_mode)
        self._lib = load(kind)
        self._suf = "f32" if self.dtype == np.float32 else "f64"
        self._ct = ctypes.c_float if self.dtype == np.float32 else ctypes.c_double
        P = _P32 if self.dtype == np.float32 else _P64
        sz = getattr(self._lib, "ref_sp_sizeof_params_" + self._suf)
        sz.restype = ctypes.c_size_t
        assert sz() == ctypes.sizeof(P), "ctypes mirror of ref_params out of date"
        self._P = P

    def _fn(self, name):
        fn = getattr(self._lib, "%s_%s" % (name, self._suf))
        fn.restype = None
        return fn

    def _ptr(self, a):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(ctypes.c_void_p)

    def init_simulation(self):
        super().init_simulation()
        p = self._P()
        p.nx, p.ny, p.nz = self.nx, self.ny, self.nz
        p.force_flag = self.force_flag
        for i in range(6):
            p.bc_type[i] = self.bc_type[i]
            p.bc_rho[i] = self.dtype.type(self.bc_rho[i])
            for c in range(3):
                p.bc_vel[i][c] = self.dtype.type(self.bc_vel[i][c])
        for s in range(19):
            p.S[s] = self.S[s]
            p.w[s] = self.w[s]
        flat = self.inv_M.reshape(-1)
        for i in range(361):
            p.invM[i] = flat[i]
        for c in range(3):
            p.force[c] = self.ext_f[c]
        ff = getattr(self, "force_field", None)
        p.force_field = None if ff is None else ff.ctypes.data
        p.guo_unscaled = 1 if self.guo_mode == "unscaled" else 0
        p.vel_bc_script = 1 if self.vel_bc_mode == "script" else 0
        p.ns = None if self.ns is None else self.ns.ctypes.data
        self._p = p

    def set_force_field(self, force):
        super().set_force_field(force)
        if getattr(self, "_p", None) is not None:
            self._p.force_field = None if self.force_field is None else self.force_field.ctypes.data
            self._p.force_flag = 1 if self.force_field is not None else self.force_flag

    def colission(self):
        self._fn("ref_sp_colission")(ctypes.byref(self._p), self._ptr(self.solid), self._ptr(self.F),
                                     self._ptr(self.rho), self._ptr(self.v), self._ptr(self.f))

    def streaming1(self):
        self._fn("ref_sp_streaming1")(ctypes.byref(self._p), self._ptr(self.solid), self._ptr(self.f),
                                      self._ptr(self.F))

    def streaming_grey(self):
        self._fn("ref_sp_streaming_grey")(ctypes.byref(self._p), self._ptr(self.solid), self._ptr(self.f),
                                          self._ptr(self.F))

    def Boundary_condition(self):
        self._fn("ref_sp_boundary_condition")(ctypes.byref(self._p), self._ptr(self.solid),
                                              self._ptr(self.v), self._ptr(self.F))

    def streaming3(self):
        self._fn("ref_sp_streaming3")(ctypes.byref(self._p), self._ptr(self.solid), self._ptr(self.F),
                                      self._ptr(self.f), self._ptr(self.rho), self._ptr(self.v))

    def run(self, nsteps):
        self._fn("ref_sp_step")(ctypes.byref(self._p), self._ptr(self.solid), self._ptr(self.f),
                                self._ptr(self.F), self._ptr(self.rho), self._ptr(self.v),
                                ctypes.c_int(int(nsteps)))


# ---- two-phase ---------------------------------------------------------------------------------
from . import ref_two_phase as _np_ref2  # noqa: E402


def _params2_struct(ctype):
    class P(ctypes.Structure):
        _fields_ = [("nx", ctypes.c_int), ("ny", ctypes.c_int), ("nz", ctypes.c_int),
                    ("bc_type", ctypes.c_int * 6), ("bc_psi_type", ctypes.c_int * 6),
                    ("invM", ctype * 361), ("w", ctype * 19), ("force", ctype * 3),
                    ("bc_rho", ctype * 6), ("bc_psi_val", ctype * 6), ("psi_solid", ctype), ("CapA", ctype),
                    ("wl", ctype), ("wg", ctype), ("lg0", ctype), ("l1", ctype), ("l2", ctype),
                    ("g1", ctype), ("g2", ctype)]
    return P


_P2_32 = _params2_struct(ctypes.c_float)
_P2_64 = _params2_struct(ctypes.c_double)


class RefTwoPhaseC(_np_ref2.RefTwoPhase):
    """Same state and setters as the NumPy two-phase oracle; the passes run in C."""

    def __init__(self, nx, ny, nz, dtype=np.float32, kind="strict"):
        super().__init__(nx, ny, nz, dtype=dtype)
        self._lib = load(kind)
        self._suf = "f32" if self.dtype == np.float32 else "f64"
        P = _P2_32 if self.dtype == np.float32 else _P2_64
        sz = getattr(self._lib, "ref2p_sizeof_params_" + self._suf)
        sz.restype = ctypes.c_size_t
        assert sz() == ctypes.sizeof(P), "ctypes mirror of ref2p_params out of date"
        self._P = P
        self.g_r = np.zeros((nx, ny, nz, 19), self.dtype)
        self.g_b = np.zeros((nx, ny, nz, 19), self.dtype)

    def _fn(self, name):
        fn = getattr(self._lib, "%s_%s" % (name, self._suf))
        fn.restype = None
        return fn

    @staticmethod
    def _ptr(a):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(ctypes.c_void_p)

    def init_simulation(self):
        super().init_simulation()
        t = self.dtype.type
        p = self._P()
        p.nx, p.ny, p.nz = self.nx, self.ny, self.nz
        for i in range(6):
            p.bc_type[i] = self.bc_type[i]
            p.bc_psi_type[i] = self.bc_psi_type[i]
            p.bc_rho[i] = t(self.bc_rho[i])
            p.bc_psi_val[i] = t(self.bc_psi_val[i])
        flat = self.inv_M.reshape(-1)
        for i in range(361):
            p.invM[i] = flat[i]
        for s in range(19):
            p.w[s] = self.w[s]
        for c in range(3):
            p.force[c] = self.ext_f[c]
        p.psi_solid, p.CapA = t(self.psi_solid), t(self.CapA)
        for n in ("wl", "wg", "lg0", "l1", "l2", "g1", "g2"):
            setattr(p, n, getattr(self, n))
        self._p = p

    def colission(self):
        P = self._ptr
        self._fn("ref2p_collide")(ctypes.byref(self._p), P(self.solid), P(self.F), P(self.rho), P(self.v),
                                  P(self.psi), P(self.rho_r), P(self.rho_b), P(self.f), P(self.g_r), P(self.g_b))
        self._fn("ref2p_accumulate")(ctypes.byref(self._p), P(self.solid), P(self.g_r), P(self.g_b),
                                     P(self.rhor), P(self.rhob))

    def streaming1(self):
        self._fn("ref2p_streaming1")(ctypes.byref(self._p), self._ptr(self.solid), self._ptr(self.f), self._ptr(self.F))

    def Boundary_condition(self):
        self._fn("ref2p_boundary_condition")(ctypes.byref(self._p), self._ptr(self.solid), self._ptr(self.v),
                                             self._ptr(self.F))

    def streaming3(self):
        P = self._ptr
        self._fn("ref2p_streaming3")(ctypes.byref(self._p), P(self.solid), P(self.F), P(self.f), P(self.rho),
                                     P(self.v), P(self.psi), P(self.rho_r), P(self.rho_b), P(self.rhor), P(self.rhob))

    def Boundary_condition_psi(self):
        P = self._ptr
        self._fn("ref2p_boundary_condition_psi")(ctypes.byref(self._p), P(self.solid), P(self.psi), P(self.rho_r),
                                                 P(self.rho_b))

    def run(self, nsteps):
        P = self._ptr
        self._fn("ref2p_step")(ctypes.byref(self._p), P(self.solid), P(self.f), P(self.F), P(self.rho), P(self.v),
                               P(self.psi), P(self.rho_r), P(self.rho_b), P(self.rhor), P(self.rhob), P(self.g_r),
                               P(self.g_b), ctypes.c_int(int(nsteps)))
