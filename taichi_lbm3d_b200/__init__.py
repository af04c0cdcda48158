"""taichi_lbm3d_b200 -- B200-native (sm_100a) D3Q19 MRT lattice-Boltzmann time step behind
the Python API of yjhp1016/taichi_LBM3D.

    from taichi_lbm3d_b200 import LB3D_Solver_Single_Phase

Host code is Python (PyTorch for device selection, streams and torch.distributed); all
lattice work runs in hand-written CUDA kernels reached through the C ABI in
``include/lbm3d.h``.  No Taichi, no Triton, no CPU fallback.
"""
from .LBM_3D_SinglePhase_Solver import LB3D_Solver_Single_Phase, relaxation_rates  # noqa: F401
from .lbm_solver_3d_2phase import LB3D_Solver_Two_Phase  # noqa: F401

__version__ = "0.1.0"
