"""ctypes binding of liblbm3d_b200.so (the C ABI declared in include/lbm3d.h).

There is no CPU fallback: if the CUDA library cannot be loaded, importing this module
raises, and every call that needs a GPU returns the library's own error when none is
present.
"""
import ctypes
import os

from . import build as _build

_c = ctypes

LBM_OK = 0


class LbmConfig(ctypes.Structure):
    """lbm_config (include/lbm3d.h)."""
    _fields_ = [("nx", _c.c_int32), ("ny", _c.c_int32), ("nz", _c.c_int32),
                ("sparse", _c.c_int32), ("strict", _c.c_int32), ("halo_x", _c.c_int32),
                ("device", _c.c_int32), ("x_face_mask", _c.c_int32)]


class Lbm2pConfig(ctypes.Structure):
    """lbm2p_config (include/lbm3d_2phase.h)."""
    _fields_ = [("nx", _c.c_int32), ("ny", _c.c_int32), ("nz", _c.c_int32),
                ("strict", _c.c_int32), ("device", _c.c_int32), ("reserved", _c.c_int32)]


class LbmError(RuntimeError):
    pass


_VP = _c.c_void_p
_I = _c.c_int
_I64 = _c.c_int64
_FP = _c.POINTER(_c.c_float)

# name -> (restype, argtypes); every symbol include/lbm3d.h declares
SIGNATURES = {
    "lbm_abi_version": (_I, []),
    "lbm_create": (_I, [_c.POINTER(LbmConfig), _c.POINTER(_VP)]),
    "lbm_destroy": (_I, [_VP]),
    "lbm_last_error": (_c.c_char_p, [_VP]),
    "lbm_set_geometry": (_I, [_VP, _VP]),
    "lbm_set_bc": (_I, [_VP, _I, _I, _c.c_float, _FP]),
    "lbm_set_force": (_I, [_VP, _FP]),
    "lbm_set_force_field": (_I, [_VP, _VP]),
    "lbm_set_guo_form": (_I, [_VP, _I]),
    "lbm_set_vel_bc_form": (_I, [_VP, _I]),
    "lbm_set_grey_scale": (_I, [_VP, _VP]),
    "lbm_pool_trim": (ctypes.c_longlong, []),
    "lbm_set_viscosity": (_I, [_VP, _c.c_double, _I]),
    "lbm_set_relaxation": (_I, [_VP, _FP]),
    "lbm_set_inverse_matrix": (_I, [_VP, _FP]),
    "lbm_init": (_I, [_VP]),
    "lbm_step": (_I, [_VP, _I, _VP]),
    "lbm_launch_count": (_I64, [_VP]),
    "lbm_synchronize": (_I, [_VP]),
    "lbm_get_rho": (_I, [_VP, _VP]),
    "lbm_get_v": (_I, [_VP, _VP]),
    "lbm_get_F": (_I, [_VP, _VP]),
    "lbm_get_solid": (_I, [_VP, _VP]),
    "lbm_set_rho": (_I, [_VP, _VP]),
    "lbm_set_v": (_I, [_VP, _VP]),
    "lbm_set_F": (_I, [_VP, _VP]),
    "lbm_get_max_v": (_I, [_VP, _FP]),
    "lbm_get_nodes": (_I, [_VP, _I64, _VP, _VP, _VP, _VP]),
    "lbm_get_num_fluid": (_I, [_VP, _c.POINTER(_I64)]),
    "lbm_get_fluid_index": (_I, [_VP, _VP]),
    "lbm_get_neighbor_table": (_I, [_VP, _VP]),
    "lbm_get_link_flags": (_I, [_VP, _VP]),
    "lbm_halo_count": (_I64, [_VP, _I]),
    "lbm_halo_pack": (_I, [_VP, _I, _I, _VP, _VP]),
    "lbm_halo_unpack": (_I, [_VP, _I, _I, _VP, _VP]),
    "lbm_p2p_export": (_I, [_VP, _VP]),
    "lbm_p2p_connect": (_I, [_VP, _VP, _VP]),
    "lbm_p2p_enable": (_I, [_VP, _I]),
    "lbm_p2p_disconnect": (_I, [_VP]),
    "lbm_comm_ready": (_I, [_I, _I]),
    "lbm_comm_unique_id": (_I, [_VP]),
    "lbm_comm_init": (_I, [_VP, _VP, _I, _I]),
    "lbm_run_slab": (_I, [_VP, _I, _I, _VP]),
    "lbm_step_begin": (_I, [_VP, _VP]),
    "lbm_step_planes": (_I, [_VP, _I, _I, _VP]),
    "lbm_step_flip": (_I, [_VP]),
    "lbm_get_device_ptr": (_I, [_VP, _I, _c.POINTER(_VP), _c.POINTER(_c.c_size_t)]),
    "lbm_get_layout": (_I, [_VP, _c.POINTER(_I64)]),
}

# every symbol include/lbm3d_2phase.h declares
SIGNATURES_2P = {
    "lbm2p_create": (_I, [_c.POINTER(Lbm2pConfig), _c.POINTER(_VP)]),
    "lbm2p_destroy": (_I, [_VP]),
    "lbm2p_last_error": (_c.c_char_p, [_VP]),
    "lbm2p_set_geometry": (_I, [_VP, _VP]),
    "lbm2p_set_phase": (_I, [_VP, _VP]),
    "lbm2p_set_fluid": (_I, [_VP, _c.c_double, _c.c_double, _c.c_double, _c.c_double]),
    "lbm2p_set_force": (_I, [_VP, _FP]),
    "lbm2p_set_bc": (_I, [_VP, _I, _I, _c.c_float, _FP]),
    "lbm2p_set_psi_bc": (_I, [_VP, _I, _I, _c.c_float]),
    "lbm2p_set_inverse_matrix": (_I, [_VP, _FP]),
    "lbm2p_init": (_I, [_VP]),
    "lbm2p_step": (_I, [_VP, _I, _VP]),
    "lbm2p_launch_count": (_I64, [_VP]),
    "lbm2p_synchronize": (_I, [_VP]),
    "lbm2p_get_rho": (_I, [_VP, _VP]),
    "lbm2p_get_v": (_I, [_VP, _VP]),
    "lbm2p_get_F": (_I, [_VP, _VP]),
    "lbm2p_get_psi": (_I, [_VP, _VP]),
    "lbm2p_get_rho_r": (_I, [_VP, _VP]),
    "lbm2p_get_rho_b": (_I, [_VP, _VP]),
    "lbm2p_get_solid": (_I, [_VP, _VP]),
    "lbm2p_set_state": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "lbm2p_get_max_v": (_I, [_VP, _FP]),
    "lbm2p_halo_floats": (_I64, [_VP, _I]),
    "lbm2p_halo_count": (_I64, [_VP, _I]),
    "lbm2p_halo_pack": (_I, [_VP, _I, _I, _VP, _VP]),
    "lbm2p_halo_unpack": (_I, [_VP, _I, _I, _VP, _VP]),
    "lbm2p_slab_stage": (_I, [_VP, _I, _VP]),
    "lbm2p_comm_unique_id": (_I, [_VP]),
    "lbm2p_comm_init": (_I, [_VP, _VP, _I, _I]),
    "lbm2p_run_slab": (_I, [_VP, _I, _VP]),
}

_lib = None


def library_path():
    return _build.LIB


def load():
    """Load the CUDA library, (re)building it first when it is missing or older than its sources
    (csrc/, include/*.h) and nvcc exists.  LBM3D_LIB selects a differently tuned build of the same
    sources (python -m taichi_lbm3d_b200.build --tag NAME --flags "-D..."), used as it is."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LBM3D_LIB")
    if path:
        if not os.path.exists(path):
            raise ImportError("taichi_lbm3d_b200: LBM3D_LIB=%s does not exist" % path)
    else:
        path = _build.LIB
        try:
            # every rank of a multi-GPU job comes through here: one builder at a time
            with _build.build_lock():
                if _build.needs_build() and (_build.have_nvcc() or not os.path.exists(path)):
                    _build.build()
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(path):
                raise ImportError(
                    "taichi_lbm3d_b200: %s is missing and could not be built (%s). Run "
                    "`python -m taichi_lbm3d_b200.build`; there is no CPU fallback." % (path, e))
            raise ImportError("taichi_lbm3d_b200: %s is older than its sources and the rebuild failed: %s"
                              % (path, e))
    lib = ctypes.CDLL(path)
    for table in (SIGNATURES, SIGNATURES_2P):
        for name, (res, args) in table.items():
            fn = getattr(lib, name)        # AttributeError = ABI mismatch, fail loudly
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib


def check2(lib, ctx, status, what):
    if status < 0:
        msg = lib.lbm2p_last_error(ctx)
        raise LbmError("%s failed (%d): %s" % (what, status, msg.decode() if msg else "?"))
    return status


def check(lib, ctx, status, what):
    if status < 0:
        msg = lib.lbm_last_error(ctx)
        raise LbmError("%s failed (%d): %s" % (what, status, msg.decode() if msg else "?"))
    return status
