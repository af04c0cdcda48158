"""Drop-in mirror of the reference's ``Single_phase/LBM_3D_SinglePhase_Solver.py``.

Same module name, class name, constructor, setters, ``init_geo``, ``init_simulation``,
``step``, ``get_max_v`` and ``export_VTK`` as the reference class
(``LB3D_Solver_Single_Phase``, reference file :9-481), with the Taichi kernels replaced by
the fused sm_100a CUDA step behind the C ABI of ``include/lbm3d.h``.  The case scripts of
the reference run after dropping their two Taichi lines (``import taichi`` / ``ti.init``).

Additions that the reference does not have: ``run(n)`` (n steps, one launch each, no
Python in between), ``in_place`` (sparse storage on ONE population buffer, AA pattern),
``set_force_field`` (per-node force: the array form of overriding ``cal_local_force``),
``guo_mode`` (the un-scaled force term of the solver's other copies), ``tau_mode`` (the textbook relaxation time the other copies of the
solver use), ``strict`` (oracle-order arithmetic for verification), ``to_torch`` on fields.

There is no CPU path: constructing the solver is cheap, ``init_simulation`` needs a GPU.
"""
import ctypes

import numpy as np

from . import _lib
from . import vtk as _vtk

_FACES = ("x_left", "x_right", "y_left", "y_right", "z_left", "z_right")
_SUFFIX = ("xl", "xr", "yl", "yr", "zl", "zr")


def relaxation_rates(niu, tau_mode="class"):
    """S_dig of init_simulation (reference :126-131), evaluated in Python floats (fp64) and
    rounded once to fp32 exactly as the Taichi field store does."""
    if tau_mode == "class":
        tau_f = niu / 3.0 + 0.5          # :127
    elif tau_mode == "textbook":
        tau_f = 3.0 * niu + 0.5          # :126 (commented out in the class)
    else:
        raise ValueError("tau_mode must be 'class' or 'textbook'")
    s_v = 1.0 / tau_f
    s_other = 8.0 * (2.0 - s_v) / (8.0 - s_v)
    return np.array([0, s_v, s_v, 0, s_other, 0, s_other, 0, s_other, s_v, s_v, s_v, s_v,
                     s_v, s_v, s_v, s_other, s_other, s_other], dtype=np.float64).astype(np.float32)


class _Field:
    """Stand-in for a Taichi field: ``to_numpy()`` / ``from_numpy(arr)`` with the shapes the
    reference exposes ((nx,ny,nz), (nx,ny,nz,3), (nx,ny,nz,19); C order)."""

    def __init__(self, solver, name, shape, dtype):
        self._solver, self._name = solver, name
        self.shape, self.dtype = shape, np.dtype(dtype)

    def to_numpy(self, out=None):
        """Fresh array like the reference's to_numpy(); `out` (addition) receives the copy
        instead, e.g. a pinned host buffer (torch.empty(..., pin_memory=True).numpy())."""
        return self._solver._get_field(self._name, out)

    def from_numpy(self, arr):
        self._solver._set_field(self._name, arr)

    def to_torch(self, device=None):
        import torch
        t = torch.from_numpy(self.to_numpy())
        return t.to(device) if device is not None else t

    def __getitem__(self, idx):
        return self.to_numpy()[idx]


class LB3D_Solver_Single_Phase:
    def __init__(self, nx, ny, nz, sparse_storage=False, strict=False, tau_mode="class", device=None,
                 in_place=None, guo_mode="class", vel_bc_mode="class"):
        # reference :13-28
        self.enable_projection = True
        self.sparse_storage = sparse_storage
        # step IN PLACE on one population buffer (AA pattern): half the memory; sparse storage
        # 82 % instead of 87 % of the HBM roofline on B200; default from LBM3D_AA
        self.in_place = in_place
        self.nx, self.ny, self.nz = nx, ny, nz
        self.fx, self.fy, self.fz = 0.0e-6, 0.0, 0.0
        self.niu = 0.16667
        self.bc_x_left, self.rho_bcxl, self.vx_bcxl, self.vy_bcxl, self.vz_bcxl = 0, 1.0, 0.0e-5, 0.0, 0.0
        self.bc_x_right, self.rho_bcxr, self.vx_bcxr, self.vy_bcxr, self.vz_bcxr = 0, 1.0, 0.0, 0.0, 0.0
        self.bc_y_left, self.rho_bcyl, self.vx_bcyl, self.vy_bcyl, self.vz_bcyl = 0, 1.0, 0.0, 0.0, 0.0
        self.bc_y_right, self.rho_bcyr, self.vx_bcyr, self.vy_bcyr, self.vz_bcyr = 0, 1.0, 0.0, 0.0, 0.0
        self.bc_z_left, self.rho_bczl, self.vx_bczl, self.vy_bczl, self.vz_bczl = 0, 1.0, 0.0, 0.0, 0.0
        self.bc_z_right, self.rho_bczr, self.vx_bczr, self.vy_bczr, self.vz_bczr = 0, 1.0, 0.0, 0.0, 0.0
        self.strict = bool(strict)
        self.tau_mode = tau_mode
        # "class": Guo term as the class writes it (:236, divided by 3 and 9); "unscaled": as in
        # Phase_change/LBM_3D_SinglePhase_Solver.py:235 (with tau_mode="textbook" that copy's physics)
        if guo_mode not in ("class", "unscaled"):
            raise ValueError("guo_mode must be 'class' or 'unscaled'")
        self.guo_mode = guo_mode
        # "class": a fixed-velocity face overwrites its nodes with feq(1, u) (:283-288); "script": the
        # form of the solver's script copies (Single_phase/lbm_solver_3d.py:253), in place for s = 0..18:
        # F[s] = feq(LR[s], 1, u) - F[LR[s]] + feq(s, 1, u)
        if vel_bc_mode not in ("class", "script"):
            raise ValueError("vel_bc_mode must be 'class' or 'script'")
        self.vel_bc_mode = vel_bc_mode
        self.device = device
        self._solid = None           # host copy of the geometry, all-fluid until assigned (see _solid_host)
        self._solid_dev = None       # or: the assigned array as it was copied to the device (see _set_field)
        self._force_field = None
        self._ns_host = None
        self._ctx = None
        self._lib = None
        self.solid = _Field(self, "solid", (nx, ny, nz), np.int8)
        # solid fraction per node of the grey-scale solver (Grey_Scale/lbm_solver_3d_Macro_Sukop.py:42):
        # ns.from_numpy(a) switches the streaming to that script's partial bounce-back (:233-247) and, as
        # the script does (:339), makes the nodes with int(a) >= 1 solid; with tau_mode="textbook",
        # guo_mode="unscaled" the time step is the script's.  Dense two-buffer storage only.
        self.ns = _Field(self, "ns", (nx, ny, nz), np.float32)
        self.rho = _Field(self, "rho", (nx, ny, nz), np.float32)
        self.v = _Field(self, "v", (nx, ny, nz, 3), np.float32)
        self.F = _Field(self, "F", (nx, ny, nz, 19), np.float32)
        self.f = _Field(self, "f", (nx, ny, nz, 19), np.float32)
        self.LR = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15, 18, 17]   # :85
        # :112-114
        self.x = np.linspace(0, nx, nx)
        self.y = np.linspace(0, ny, ny)
        self.z = np.linspace(0, nz, nz)

    @property
    def _solid_host(self):
        # made on first need: a solver whose geometry went straight to the device (or is never read
        # back) does not touch 16 MB of host memory per 256^3 for it
        if self._solid is None:
            if self._solid_dev is not None:
                self._solid = (self._solid_dev > 0).cpu().numpy().view(np.int8)    # init_geo :175
            else:
                self._solid = np.zeros((self.nx, self.ny, self.nz), np.int8)
        return self._solid

    @_solid_host.setter
    def _solid_host(self, arr):
        self._solid = arr
        self._solid_dev = None

    # ---- setters, reference :405-458 ----------------------------------------------------
    def set_bc_vel_x1(self, vel):
        self.bc_x_right = 2
        self.vx_bcxr = vel[0]; self.vy_bcxr = vel[1]; self.vz_bcxr = vel[2]

    def set_bc_vel_x0(self, vel):
        self.bc_x_left = 2
        self.vx_bcxl = vel[0]; self.vy_bcxl = vel[1]; self.vz_bcxl = vel[2]

    def set_bc_vel_y1(self, vel):
        self.bc_y_right = 2
        self.vx_bcyr = vel[0]; self.vy_bcyr = vel[1]; self.vz_bcyr = vel[2]

    def set_bc_vel_y0(self, vel):
        self.bc_y_left = 2
        self.vx_bcyl = vel[0]; self.vy_bcyl = vel[1]; self.vz_bcyl = vel[2]

    def set_bc_vel_z1(self, vel):
        self.bc_z_right = 2
        self.vx_bczr = vel[0]; self.vy_bczr = vel[1]; self.vz_bczr = vel[2]

    def set_bc_vel_z0(self, vel):
        self.bc_z_left = 2
        self.vx_bczl = vel[0]; self.vy_bczl = vel[1]; self.vz_bczl = vel[2]

    def set_bc_rho_x0(self, rho):
        self.bc_x_left = 1
        self.rho_bcxl = rho

    def set_bc_rho_x1(self, rho):
        self.bc_x_right = 1
        self.rho_bcxr = rho

    def set_bc_rho_y0(self, rho):
        self.bc_y_left = 1
        self.rho_bcyl = rho

    def set_bc_rho_y1(self, rho):
        self.bc_y_right = 1
        self.rho_bcyr = rho

    def set_bc_rho_z0(self, rho):
        self.bc_z_left = 1
        self.rho_bczl = rho

    def set_bc_rho_z1(self, rho):
        self.bc_z_right = 1
        self.rho_bczr = rho

    def set_viscosity(self, niu):
        self.niu = niu

    def set_force(self, force):
        self.fx = force[0]; self.fy = force[1]; self.fz = force[2]

    def set_force_field(self, force):
        """Per-node force, shape (nx, ny, nz, 3): the array form of the reference's override
        point ``cal_local_force(i, j, k)`` (:217-220; the solute solver overrides it to add
        buoyancy).  May be replaced between steps; ``None`` returns to ``set_force``."""
        if force is not None:
            force = np.ascontiguousarray(np.asarray(force, dtype=np.float32))
            if force.shape != (self.nx, self.ny, self.nz, 3):
                raise ValueError("force field must have shape %s" % ((self.nx, self.ny, self.nz, 3),))
        self._force_field = force
        if self._ctx is not None:
            self._apply_force_field()

    def _apply_force_field(self):
        ff = self._force_field
        ptr = ctypes.c_void_p(None) if ff is None else ff.ctypes.data_as(ctypes.c_void_p)
        self._ck(self._lib.lbm_set_force_field(self._ctx, ptr), "lbm_set_force_field")

    # ---- geometry, reference :173-177 ---------------------------------------------------------
    def init_geo(self, filename):
        from .geometry import load_geometry
        self._set_field("solid", load_geometry(filename, self.nx, self.ny, self.nz))

    # ---- init_simulation, reference :118-149 --------------------------------------------------
    def _bc_tuple(self, face):
        sfx = _SUFFIX[face]
        return (getattr(self, "bc_" + _FACES[face]), getattr(self, "rho_bc" + sfx),
                [getattr(self, "vx_bc" + sfx), getattr(self, "vy_bc" + sfx), getattr(self, "vz_bc" + sfx)])

    def _config(self):
        import torch
        if not torch.cuda.is_available():
            raise _lib.LbmError("taichi_lbm3d_b200 needs a CUDA device (no CPU fallback)")
        dev = torch.cuda.current_device() if self.device is None else torch.device(self.device).index or 0
        import os
        in_place = os.environ.get("LBM3D_AA", "0") == "1" if self.in_place is None else bool(self.in_place)
        mode = (3 if in_place else 0) if not self.sparse_storage else (2 if in_place else 1)
        return _lib.LbmConfig(nx=self.nx, ny=self.ny, nz=self.nz, sparse=mode,
                              strict=int(self.strict), halo_x=0, device=int(dev), x_face_mask=0)

    def init_simulation(self):
        lib = self._lib = _lib.load()
        if self._ctx is not None:
            lib.lbm_destroy(self._ctx)
            self._ctx = None
        cfg = self._config()
        ctx = ctypes.c_void_p()
        st = lib.lbm_create(ctypes.byref(cfg), ctypes.byref(ctx))
        if st < 0:
            raise _lib.LbmError("lbm_create failed (%d): %s" % (st, lib.lbm_last_error(None).decode()))
        self._ctx = ctx
        self.bc_vel_x_left = [self.vx_bcxl, self.vy_bcxl, self.vz_bcxl]      # :119-124
        self.bc_vel_x_right = [self.vx_bcxr, self.vy_bcxr, self.vz_bcxr]
        self.bc_vel_y_left = [self.vx_bcyl, self.vy_bcyl, self.vz_bcyl]
        self.bc_vel_y_right = [self.vx_bcyr, self.vy_bcyr, self.vz_bcyr]
        self.bc_vel_z_left = [self.vx_bczl, self.vy_bczl, self.vz_bczl]
        self.bc_vel_z_right = [self.vx_bczr, self.vy_bczr, self.vz_bczr]
        self.tau_f = self.niu / 3.0 + 0.5 if self.tau_mode == "class" else 3.0 * self.niu + 0.5
        self.s_v = 1.0 / self.tau_f
        self.s_other = 8.0 * (2.0 - self.s_v) / (8.0 - self.s_v)
        S = relaxation_rates(self.niu, self.tau_mode)
        self.force_flag = 1 if (abs(self.fx) > 0 or abs(self.fy) > 0 or abs(self.fz) > 0) else 0   # :137-140
        if self._solid_dev is not None and self._solid_dev.device.index == cfg.device:
            # the array went to the device when it was assigned; lbm_set_geometry turns > 0 into 1 there
            self._ck(lib.lbm_set_geometry(ctx, ctypes.c_void_p(self._solid_dev.data_ptr())), "lbm_set_geometry")
        else:
            solid = np.ascontiguousarray(self._solid_host, dtype=np.int8)
            self._ck(lib.lbm_set_geometry(ctx, solid.ctypes.data_as(ctypes.c_void_p)), "lbm_set_geometry")
        for face in range(6):
            t, rho, vel = self._bc_tuple(face)
            velc = (ctypes.c_float * 3)(*[float(np.float32(c)) for c in vel])
            self._ck(lib.lbm_set_bc(ctx, face, int(t), ctypes.c_float(float(np.float32(rho))), velc), "lbm_set_bc")
        fc = (ctypes.c_float * 3)(float(np.float32(self.fx)), float(np.float32(self.fy)), float(np.float32(self.fz)))
        self._ck(lib.lbm_set_force(ctx, fc), "lbm_set_force")
        self._ck(lib.lbm_set_guo_form(ctx, 1 if self.guo_mode == "unscaled" else 0), "lbm_set_guo_form")
        self._ck(lib.lbm_set_vel_bc_form(ctx, 1 if self.vel_bc_mode == "script" else 0), "lbm_set_vel_bc_form")
        if self._ns_host is not None:
            self._ck(lib.lbm_set_grey_scale(ctx, self._ns_host.ctypes.data_as(ctypes.c_void_p)), "lbm_set_grey_scale")
        self._ck(lib.lbm_set_relaxation(ctx, S.ctypes.data_as(_lib._FP)), "lbm_set_relaxation")
        if self.strict:
            from .constants import M_np
            inv = np.ascontiguousarray(np.linalg.inv(M_np).astype(np.float32))   # :83, :110
            self._ck(lib.lbm_set_inverse_matrix(ctx, inv.ctypes.data_as(_lib._FP)), "lbm_set_inverse_matrix")
        self._ck(lib.lbm_init(ctx), "lbm_init")
        if self._force_field is not None:
            self._apply_force_field()

    # ---- time stepping, reference :477-481 -----------------------------------------------------
    def _stream(self):
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step(self):
        self._ck(self._lib.lbm_step(self._require_ctx(), 1, self._stream()), "lbm_step")

    def run(self, nsteps):
        """nsteps reference steps back to back (addition; the reference loops in Python)."""
        self._ck(self._lib.lbm_step(self._require_ctx(), int(nsteps), self._stream()), "lbm_step")

    def synchronize(self):
        self._ck(self._lib.lbm_synchronize(self._require_ctx()), "lbm_synchronize")

    def get_max_v(self):
        out = ctypes.c_float()
        self._ck(self._lib.lbm_get_max_v(self._require_ctx(), ctypes.byref(out)), "lbm_get_max_v")
        return out.value

    @property
    def launch_count(self):
        return int(self._lib.lbm_launch_count(self._require_ctx()))

    def sample(self, index, fields=("F", "rho", "v")):
        """F, rho, v at the nodes with linear index ``i*ny*nz + j*nz + k`` (addition: a probe that
        does not copy whole lattices to the host); returns a dict of arrays [n,19], [n], [n,3]"""
        idx = np.ascontiguousarray(np.asarray(index, dtype=np.int64).reshape(-1))
        out = {"F": np.empty((idx.size, 19), np.float32) if "F" in fields else None,
               "rho": np.empty(idx.size, np.float32) if "rho" in fields else None,
               "v": np.empty((idx.size, 3), np.float32) if "v" in fields else None}
        ptr = lambda a: ctypes.c_void_p(None) if a is None else a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        self._ck(self._lib.lbm_get_nodes(self._require_ctx(), idx.size, idx.ctypes.data_as(ctypes.c_void_p),
                                         ptr(out["F"]), ptr(out["rho"]), ptr(out["v"])), "lbm_get_nodes")
        return {k: v for k, v in out.items() if v is not None}

    # ---- output, reference :462-475 ----------------------------------------------------------
    def export_VTK(self, n):
        v = self.v.to_numpy()
        _vtk.grid_to_vtr("./LB_SingelPhase_" + str(n), self.x, self.y, self.z,
                         {"Solid": np.ascontiguousarray(self.solid.to_numpy()),
                          "rho": np.ascontiguousarray(self.rho.to_numpy()),
                          "velocity": (np.ascontiguousarray(v[0:self.nx, 0:self.ny, 0:self.nz, 0]),
                                       np.ascontiguousarray(v[0:self.nx, 0:self.ny, 0:self.nz, 1]),
                                       np.ascontiguousarray(v[0:self.nx, 0:self.ny, 0:self.nz, 2]))})

    # ---- checkpoint / restart (the reference cannot resume: its VTK dumps lack populations) ----
    @staticmethod
    def _ckpt_path(path):
        path = str(path)
        return path if path.endswith(".npz") else path + ".npz"      # what np.savez appends

    def _case_settings(self):
        """everything besides the state that decides how the run continues"""
        bcs = []
        for face in range(6):
            t, rho, vel = self._bc_tuple(face)
            bcs.append([float(t), float(rho)] + [float(c) for c in vel])
        return {"niu": float(self.niu), "force": [float(self.fx), float(self.fy), float(self.fz)],
                "tau_mode": self.tau_mode, "guo_mode": self.guo_mode, "vel_bc_mode": self.vel_bc_mode, "bc": bcs}

    def save_checkpoint(self, path):
        """state of the run (F, rho, v), geometry, per-node force and the case settings (viscosity,
        force, BCs, tau/guo mode) as a compressed .npz (``.npz`` is appended if missing)"""
        import json
        extra = {} if self._force_field is None else {"force_field": self._force_field}
        if self._ns_host is not None:
            extra["ns"] = self._ns_host
        np.savez_compressed(self._ckpt_path(path), nx=self.nx, ny=self.ny, nz=self.nz, solid=self._solid_host,
                            F=self.F.to_numpy(), rho=self.rho.to_numpy(), v=self.v.to_numpy(),
                            settings=np.array(json.dumps(self._case_settings())), **extra)

    def load_checkpoint(self, path, strict_settings=True):
        """restore F, rho, v (and the per-node force) saved by save_checkpoint, after
        init_simulation() on the same geometry.  The case settings stored with the state must equal
        this solver's (a restart with other setters would silently continue a different case);
        ``strict_settings=False`` only warns."""
        import json
        d = np.load(self._ckpt_path(path))
        if (int(d["nx"]), int(d["ny"]), int(d["nz"])) != (self.nx, self.ny, self.nz) or \
                not np.array_equal(d["solid"], self._solid_host):
            raise ValueError("checkpoint was written for a different lattice")
        if ("ns" in d.files) != (self._ns_host is not None) or \
                ("ns" in d.files and not np.array_equal(d["ns"], self._ns_host)):
            raise ValueError("checkpoint was written for a different grey-scale lattice (ns differs)")
        if "settings" in d.files:
            saved, mine = json.loads(str(d["settings"])), self._case_settings()
            if saved != mine:
                msg = "checkpoint settings differ from this solver's: saved %s, now %s" % (saved, mine)
                if strict_settings:
                    raise ValueError(msg)
                import warnings
                warnings.warn(msg)
        self.F.from_numpy(d["F"])
        self.rho.from_numpy(d["rho"])
        self.v.from_numpy(d["v"])
        if "force_field" in d.files:
            self.set_force_field(d["force_field"])

    # ---- sparse-storage tables (bit-exact compaction checks) ---------------------------------
    def num_fluid(self):
        n = ctypes.c_int64()
        self._ck(self._lib.lbm_get_num_fluid(self._require_ctx(), ctypes.byref(n)), "lbm_get_num_fluid")
        return n.value

    def fluid_index(self):
        out = np.empty(self.num_fluid(), np.int64)
        self._ck(self._lib.lbm_get_fluid_index(self._require_ctx(), out.ctypes.data_as(ctypes.c_void_p)),
                 "lbm_get_fluid_index")
        return out

    def neighbor_table(self):
        out = np.empty((18, self.num_fluid()), np.int32)
        self._ck(self._lib.lbm_get_neighbor_table(self._require_ctx(), out.ctypes.data_as(ctypes.c_void_p)),
                 "lbm_get_neighbor_table")
        return out

    def link_flags(self):
        n = self.num_fluid() if self.sparse_storage else self.nx * self.ny * self.nz
        out = np.empty(n, np.uint32)
        self._ck(self._lib.lbm_get_link_flags(self._require_ctx(), out.ctypes.data_as(ctypes.c_void_p)),
                 "lbm_get_link_flags")
        return out if self.sparse_storage else out.reshape(self.nx, self.ny, self.nz)

    # ---- plumbing ------------------------------------------------------------------------------
    def _ck(self, status, what):
        return _lib.check(self._lib, self._ctx, status, what)

    def _require_ctx(self):
        if self._ctx is None:
            raise _lib.LbmError("init_simulation() has not been called")
        return self._ctx

    def _get_field(self, name, out=None):
        if name == "solid":
            if out is not None:
                out[...] = self._solid_host
                return out
            return self._solid_host.copy()
        if name == "ns":
            if self._ns_host is None:
                raise _lib.LbmError("ns has not been assigned (the lattice is not a grey-scale one)")
            if out is not None:
                out[...] = self._ns_host
                return out
            return self._ns_host.copy()
        if self._ctx is None:
            raise _lib.LbmError("field %s is not available before init_simulation()" % name)
        shape = {"rho": (self.nx, self.ny, self.nz), "v": (self.nx, self.ny, self.nz, 3),
                 "F": (self.nx, self.ny, self.nz, 19), "f": (self.nx, self.ny, self.nz, 19)}[name]
        if out is None:
            out = np.empty(shape, np.float32)
        elif out.shape != shape or out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous float32 array of shape %s" % (shape,))
        fn = {"rho": self._lib.lbm_get_rho, "v": self._lib.lbm_get_v, "F": self._lib.lbm_get_F,
              "f": self._lib.lbm_get_F}[name]       # after streaming3 f == F (:379)
        self._ck(fn(self._ctx, out.ctypes.data_as(ctypes.c_void_p)), "lbm_get_" + name)
        return out

    def _device_for_geometry(self, a):
        """the CUDA device a geometry array can be copied to as it is, or None"""
        if a.dtype not in (np.int8, np.uint8, np.bool_) or not a.flags["C_CONTIGUOUS"] or not a.flags["WRITEABLE"]:
            return None
        if a.dtype == np.uint8 and a.size and int(a.max()) > 127:      # would read as negative = fluid
            return None
        try:
            import torch
            if not torch.cuda.is_available():
                return None
            idx = torch.cuda.current_device() if self.device is None else (torch.device(self.device).index or 0)
            return torch.device("cuda", idx)
        except Exception:  # noqa: BLE001
            return None

    def _upload(self, name, arr):
        fn = {"rho": self._lib.lbm_set_rho, "v": self._lib.lbm_set_v, "F": self._lib.lbm_set_F}[name]
        self._ck(fn(self._ctx, arr.ctypes.data_as(ctypes.c_void_p)), "lbm_set_" + name)

    def _set_field(self, name, arr):
        if name == "solid":
            a = np.asarray(arr)
            if a.shape != (self.nx, self.ny, self.nz):
                raise ValueError("solid must have shape %s" % ((self.nx, self.ny, self.nz),))
            # the reference allows init_geo(A); init_simulation(); ...; init_geo(B); init_simulation():
            # a new geometry makes the running context stale, the next init_simulation() rebuilds it
            if self._ctx is not None:
                self._lib.lbm_destroy(self._ctx)
                self._ctx = None
            # Like a Taichi field, the solver keeps a COPY taken now.  A byte array on a machine with
            # a GPU is copied straight to the device (one DMA; from pinned memory at PCIe speed) and
            # turned into 0 / 1 there; the host copy is made only if somebody asks for it.  Anything
            # else is reduced to 0 / 1 bytes on the host first (init_geo :175).
            dev = self._device_for_geometry(a)
            if dev is not None:
                import torch
                src = a.view(np.int8) if a.dtype != np.int8 else a
                self._solid = None
                self._solid_dev = torch.from_numpy(src).to(dev)
            else:
                self._solid_host = (a > 0).view(np.int8)    # bool viewed as 0/1 bytes: one pass
            return
        if name == "ns":
            a = np.asarray(arr)
            if a.shape != (self.nx, self.ny, self.nz):
                raise ValueError("ns must have shape %s" % ((self.nx, self.ny, self.nz),))
            if self._ctx is not None:
                self._lib.lbm_destroy(self._ctx)
                self._ctx = None
            self._ns_host = np.ascontiguousarray(a, dtype=np.float32)
            self._solid_host = (a.astype(int) > 0).view(np.int8)     # solid_np = ns_np.astype(int), :339
            return
        if name == "f":
            return      # scratch in the reference: colission overwrites it before any read (:240)
        shape = getattr(self, name).shape
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float32))
        if a.shape != shape:
            raise ValueError("%s must have shape %s" % (name, shape))
        if self._ctx is None:
            # the reference's init() (:160-170) resets rho, v, f, F, so values assigned before
            # init_simulation() never survive; refuse instead of silently dropping them
            raise _lib.LbmError("assign %s after init_simulation(), which resets it" % name)
        self._upload(name, a)

    def close(self):
        """release the device state now (addition; the reference leaves it to Taichi's runtime).  The
        large buffers stay cached in the library for the next solver of the same size
        (``lbm_pool_trim``); the object can be set up again with init_simulation()."""
        if self._ctx is not None and self._lib is not None:
            self._lib.lbm_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            if self._ctx is not None and self._lib is not None:
                self._lib.lbm_destroy(self._ctx)
                self._ctx = None
        except Exception:  # noqa: BLE001
            pass
