"""stand-in for pyevtk (absent from the image); the reference imports gridToVTK at module level"""
