"""Lets the reference's case scripts keep their line
    import LBM_3D_SinglePhase_Solver as lb3dsp
unchanged: this module re-exports the B200-native class under the reference's module name."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taichi_lbm3d_b200.LBM_3D_SinglePhase_Solver import *  # noqa: F401,F403,E402
from taichi_lbm3d_b200.LBM_3D_SinglePhase_Solver import LB3D_Solver_Single_Phase  # noqa: F401,E402
