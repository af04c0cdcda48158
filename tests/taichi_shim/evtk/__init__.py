"""`evtk` is the old name of pyevtk (Grey_Scale/lbm_solver_3d_Macro_Sukop.py:3 imports it): same stand-in."""
