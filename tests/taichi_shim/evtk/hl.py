from pyevtk.hl import *  # noqa: F401,F403
from pyevtk.hl import gridToVTK  # noqa: F401
