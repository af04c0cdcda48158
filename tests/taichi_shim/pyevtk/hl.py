def gridToVTK(*args, **kwargs):  # noqa: N802
    raise NotImplementedError("pyevtk is not installed; VTK export is not part of the fixtures")
