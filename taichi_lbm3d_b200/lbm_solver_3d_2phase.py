"""Two-phase colour-gradient D3Q19 MRT solver on B200, as a class.

The reference ships this solver only as a flat script of module-level globals and kernels
(``2phase/lbm_solver_3d_2phase.py``; README to-do: "wrap functions into class").  This class
keeps the script's names -- its globals are attributes with the same defaults (:16-39), its
``init_geo(filename, filename2)`` (:194) and the kernel sequence of its main loop (:626-632)
become ``init_simulation()`` / ``step()`` -- and runs the CUDA kernels behind
``include/lbm3d_2phase.h``.

    lb = LB3D_Solver_Two_Phase(131, 131, 131)
    lb.init_geo('./img_ftb131.txt', './phase_ftb131.dat')
    lb.niu_l, lb.niu_g, lb.CapA, lb.psi_solid = 0.05, 0.2, 0.005, 0.7
    lb.init_simulation()
    for it in range(80001):
        lb.step()
        if it % 10000 == 0:
            lb.export_VTK(it)          # ./structured<it>.vtr: Solid, rho, phase, velocity (:647-659)
"""
import ctypes

import numpy as np

from . import _lib
from . import vtk as _vtk

_FACE = ("x_left", "x_right", "y_left", "y_right", "z_left", "z_right")
_SFX = ("xl", "xr", "yl", "yr", "zl", "zr")


class _Field2:
    def __init__(self, solver, name, shape, dtype=np.float32):
        self._solver, self._name, self.shape, self.dtype = solver, name, shape, np.dtype(dtype)

    def to_numpy(self):
        return self._solver._get(self._name)

    def from_numpy(self, arr):
        self._solver._set(self._name, arr)


class LB3D_Solver_Two_Phase:
    def __init__(self, nx, ny, nz, strict=False, device=None, sparse_storage=False):
        self.nx, self.ny, self.nz = nx, ny, nz
        # compact fluid-node list instead of the full lattice: what the reference's second script,
        # 2phase/lbm_solver_3d_2phase_sparse.py, does with a pointer SNode tree
        self.sparse_storage = bool(sparse_storage)
        # script globals, same names and defaults (2phase/lbm_solver_3d_2phase.py:18-39)
        self.fx, self.fy, self.fz = 5.0e-5, -2e-5, 0.0
        self.niu_l = 0.1
        self.niu_g = 0.1
        self.psi_solid = 0.7
        self.CapA = 0.005
        self.bc_x_left, self.rho_bcxl, self.vx_bcxl, self.vy_bcxl, self.vz_bcxl = 0, 1.0, 0.0e-5, 0.0, 0.0
        self.bc_x_right, self.rho_bcxr, self.vx_bcxr, self.vy_bcxr, self.vz_bcxr = 0, 0.995, 0.0, 0.0, 0.0
        self.bc_y_left, self.rho_bcyl, self.vx_bcyl, self.vy_bcyl, self.vz_bcyl = 0, 1.0, 0.0, 0.0, 0.0
        self.bc_y_right, self.rho_bcyr, self.vx_bcyr, self.vy_bcyr, self.vz_bcyr = 0, 1.0, 0.0, 0.0, 0.0
        self.bc_z_left, self.rho_bczl, self.vx_bczl, self.vy_bczl, self.vz_bczl = 0, 1.0, 0.0, 0.0, 0.0
        self.bc_z_right, self.rho_bczr, self.vx_bczr, self.vy_bczr, self.vz_bczr = 0, 1.0, 0.0, 0.0, 0.0
        self.bc_psi_x_left, self.psi_x_left = 1, -1.0
        self.bc_psi_x_right, self.psi_x_right = 0, 1.0
        self.bc_psi_y_left, self.psi_y_left = 0, 1.0
        self.bc_psi_y_right, self.psi_y_right = 0, 1.0
        self.bc_psi_z_left, self.psi_z_left = 0, 1.0
        self.bc_psi_z_right, self.psi_z_right = 0, 1.0
        self.strict = bool(strict)
        self.device = device
        shp = (nx, ny, nz)
        self._solid_host = np.zeros(shp, np.int8)
        self._psi_host = np.zeros(shp, np.float32)
        self._ctx = None
        self._lib = None
        self.solid = _Field2(self, "solid", shp, np.int8)
        self.psi = _Field2(self, "psi", shp)
        self.rho = _Field2(self, "rho", shp)
        self.rho_r = _Field2(self, "rho_r", shp)
        self.rho_b = _Field2(self, "rho_b", shp)
        self.v = _Field2(self, "v", shp + (3,))
        self.F = _Field2(self, "F", shp + (19,))
        self.f = _Field2(self, "f", shp + (19,))
        self.x = np.linspace(0, nx, nx)      # :155-157
        self.y = np.linspace(0, ny, ny)
        self.z = np.linspace(0, nz, nz)

    # ---- convenience setters (the script edits globals by hand) ---------------------------------
    def set_force(self, force):
        self.fx, self.fy, self.fz = force[0], force[1], force[2]

    def set_viscosity(self, niu_l, niu_g):
        self.niu_l, self.niu_g = niu_l, niu_g

    def set_bc_rho(self, face, rho):
        setattr(self, "bc_" + _FACE[face], 1)
        setattr(self, "rho_bc" + _SFX[face], rho)

    def set_bc_psi(self, face, psi):
        setattr(self, "bc_psi_" + _FACE[face], 1)
        setattr(self, "psi_" + _FACE[face], psi)

    # ---- init_geo :194-202 ---------------------------------------------------------------------
    def init_geo(self, filename, filename2):
        from .geometry import load_geometry
        self._solid_host = load_geometry(filename, self.nx, self.ny, self.nz)
        ph = np.loadtxt(filename2) if not str(filename2).endswith(".npy") else np.load(filename2).reshape(-1, order='F')
        self._psi_host = np.ascontiguousarray(
            np.reshape(ph, (self.nx, self.ny, self.nz), order='F').astype(np.float32))
        return self._solid_host, self._psi_host

    def _config_flags(self):
        """lbm2p_config.reserved: storage flag of a whole lattice (x-slabs override this, multi_gpu.py)"""
        return 8 if self.sparse_storage else 0           # LBM2P_SPARSE

    # ---- static_init + init :205-228, :173-186 -----------------------------------------------------
    def init_simulation(self):
        import torch
        if not torch.cuda.is_available():
            raise _lib.LbmError("taichi_lbm3d_b200 needs a CUDA device (no CPU fallback)")
        lib = self._lib = _lib.load()
        if self._ctx is not None:
            lib.lbm2p_destroy(self._ctx)
            self._ctx = None
        dev = torch.cuda.current_device() if self.device is None else torch.device(self.device).index or 0
        cfg = _lib.Lbm2pConfig(nx=self.nx, ny=self.ny, nz=self.nz, strict=int(self.strict), device=int(dev),
                               reserved=int(self._config_flags()))
        ctx = ctypes.c_void_p()
        st = lib.lbm2p_create(ctypes.byref(cfg), ctypes.byref(ctx))
        if st < 0:
            raise _lib.LbmError("lbm2p_create failed (%d): %s" % (st, lib.lbm2p_last_error(None).decode()))
        self._ctx = ctx
        f32 = lambda x: float(np.float32(x))  # noqa: E731
        solid = np.ascontiguousarray(self._solid_host, np.int8)
        psi = np.ascontiguousarray(self._psi_host, np.float32)
        self._ck(lib.lbm2p_set_geometry(ctx, solid.ctypes.data_as(ctypes.c_void_p)), "lbm2p_set_geometry")
        self._ck(lib.lbm2p_set_phase(ctx, psi.ctypes.data_as(ctypes.c_void_p)), "lbm2p_set_phase")
        self._ck(lib.lbm2p_set_fluid(ctx, float(self.niu_l), float(self.niu_g), float(self.psi_solid),
                                     float(self.CapA)), "lbm2p_set_fluid")
        fc = (ctypes.c_float * 3)(f32(self.fx), f32(self.fy), f32(self.fz))
        self._ck(lib.lbm2p_set_force(ctx, fc), "lbm2p_set_force")
        zero = (ctypes.c_float * 3)(0.0, 0.0, 0.0)     # bc_vel_* fields are never written (:223-228)
        for face in range(6):
            t = int(getattr(self, "bc_" + _FACE[face]))
            rho = f32(getattr(self, "rho_bc" + _SFX[face]))
            self._ck(lib.lbm2p_set_bc(ctx, face, t, ctypes.c_float(rho), zero), "lbm2p_set_bc")
            pt = int(getattr(self, "bc_psi_" + _FACE[face]))
            pv = f32(getattr(self, "psi_" + _FACE[face]))
            self._ck(lib.lbm2p_set_psi_bc(ctx, face, pt, ctypes.c_float(pv)), "lbm2p_set_psi_bc")
        if self.strict:
            from .constants import M_np
            inv = np.ascontiguousarray(np.linalg.inv(M_np).astype(np.float32))      # :140, :145
            self._ck(lib.lbm2p_set_inverse_matrix(ctx, inv.ctypes.data_as(_lib._FP)), "lbm2p_set_inverse_matrix")
        self._ck(lib.lbm2p_init(ctx), "lbm2p_init")

    # ---- main loop body :626-632 ----------------------------------------------------------------------
    def _stream(self):
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step(self):
        self._ck(self._lib.lbm2p_step(self._need(), 1, self._stream()), "lbm2p_step")

    def run(self, nsteps):
        self._ck(self._lib.lbm2p_step(self._need(), int(nsteps), self._stream()), "lbm2p_step")

    def synchronize(self):
        self._ck(self._lib.lbm2p_synchronize(self._need()), "lbm2p_synchronize")

    def get_max_v(self):
        out = ctypes.c_float()
        self._ck(self._lib.lbm2p_get_max_v(self._need(), ctypes.byref(out)), "lbm2p_get_max_v")
        return out.value

    @property
    def launch_count(self):
        return int(self._lib.lbm2p_launch_count(self._need()))

    def set_state(self, F, rho, v, psi, rho_r, rho_b):
        """replace the whole state (restart / perturbed start)"""
        arrs = [np.ascontiguousarray(np.asarray(a, np.float32)) for a in (F, rho, v, psi, rho_r, rho_b)]
        self._psi_host = arrs[3].copy()
        self._ck(self._lib.lbm2p_set_state(self._need(), *[a.ctypes.data_as(ctypes.c_void_p) for a in arrs]),
                 "lbm2p_set_state")

    def export_VTK(self, n):
        """the script's gridToVTK call :647-659"""
        v = self.v.to_numpy()
        return _vtk.grid_to_vtr("./structured" + str(n), self.x, self.y, self.z,
                                {"Solid": np.ascontiguousarray(self.solid.to_numpy()),
                                 "rho": np.ascontiguousarray(self.rho.to_numpy()),
                                 "phase": np.ascontiguousarray(self.psi.to_numpy()),
                                 "velocity": (np.ascontiguousarray(v[..., 0]), np.ascontiguousarray(v[..., 1]),
                                              np.ascontiguousarray(v[..., 2]))})

    # ---- plumbing -----------------------------------------------------------------------------------
    def _ck(self, status, what):
        return _lib.check2(self._lib, self._ctx, status, what)

    def _need(self):
        if self._ctx is None:
            raise _lib.LbmError("init_simulation() has not been called")
        return self._ctx

    def _get(self, name):
        if name == "solid":
            return self._solid_host.copy()
        if self._ctx is None:
            if name == "psi":
                return self._psi_host.copy()
            raise _lib.LbmError("field %s is not available before init_simulation()" % name)
        out = np.empty(getattr(self, name).shape, np.float32)
        fn = getattr(self._lib, "lbm2p_get_" + ("F" if name == "f" else name))
        self._ck(fn(self._ctx, out.ctypes.data_as(ctypes.c_void_p)), "lbm2p_get_" + name)
        return out

    def _set(self, name, arr):
        if self._ctx is not None:
            if name not in ("solid", "psi"):
                raise _lib.LbmError("after init_simulation() use set_state() to replace fields")
            # new geometry / initial phase field (the script's init_geo): the running context is
            # stale, the next init_simulation() rebuilds it
            self._lib.lbm2p_destroy(self._ctx)
            self._ctx = None
        a = np.asarray(arr)
        if a.shape != (self.nx, self.ny, self.nz):
            raise ValueError("%s must have shape %s" % (name, (self.nx, self.ny, self.nz)))
        if name == "solid":
            self._solid_host = (a > 0).view(np.int8)
        elif name == "psi":
            self._psi_host = np.ascontiguousarray(a.astype(np.float32))
        else:
            raise _lib.LbmError("only solid and psi can be assigned before init_simulation()")

    def close(self):
        """release the device state now (addition); init_simulation() sets the solver up again"""
        if self._ctx is not None and self._lib is not None:
            self._lib.lbm2p_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            if self._ctx is not None and self._lib is not None:
                self._lib.lbm2p_destroy(self._ctx)
                self._ctx = None
        except Exception:  # noqa: BLE001
            pass
