"""Binary ``.vtr`` (VTK XML RectilinearGrid) writer/reader.

Replaces ``pyevtk.hl.gridToVTK`` as the reference calls it in ``export_VTK``
(Single_phase/LBM_3D_SinglePhase_Solver.py:462-475): 1-D coordinate arrays, point data
given as C-ordered (nx,ny,nz) arrays, vectors as 3-tuples of such arrays.  Same on-disk
layout as pyevtk: appended raw data, UInt64 block headers, little endian, x fastest.
"""
import struct

import numpy as np

_VTK_TYPE = {np.dtype(np.int8): "Int8", np.dtype(np.uint8): "UInt8", np.dtype(np.int16): "Int16",
             np.dtype(np.int32): "Int32", np.dtype(np.int64): "Int64", np.dtype(np.float32): "Float32",
             np.dtype(np.float64): "Float64"}


def grid_to_vtr(path, x, y, z, pointData):
    """Write ``path + '.vtr'``; returns the file name (as gridToVTK does)."""
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    nx, ny, nz = x.size, y.size, z.size
    blocks, arrays = [], []
    offset = 0

    def add(name, ncomp, payload):
        nonlocal offset
        arrays.append('<DataArray Name="%s" NumberOfComponents="%d" type="%s" format="appended" offset="%d"/>'
                      % (name, ncomp, _VTK_TYPE[payload.dtype], offset))
        blocks.append(payload)
        offset += 8 + payload.nbytes

    point_xml = []
    scalars = vectors = None
    for name, data in pointData.items():
        if isinstance(data, tuple):
            comps = [np.asarray(c) for c in data]
            for c in comps:
                if c.shape != (nx, ny, nz):
                    raise ValueError("%s: component shape %s != %s" % (name, c.shape, (nx, ny, nz)))
            inter = np.empty((nx * ny * nz, 3), comps[0].dtype)
            for k in range(3):
                inter[:, k] = comps[k].ravel(order='F')
            before = len(arrays)
            add(name, 3, np.ascontiguousarray(inter).reshape(-1))
            point_xml.append(arrays[before])
            vectors = vectors or name
        else:
            a = np.asarray(data)
            if a.shape != (nx, ny, nz):
                raise ValueError("%s: shape %s != %s" % (name, a.shape, (nx, ny, nz)))
            before = len(arrays)
            add(name, 1, np.ascontiguousarray(a.ravel(order='F')))
            point_xml.append(arrays[before])
            scalars = scalars or name
    coord_xml = []
    for name, c in (("x_coordinates", x), ("y_coordinates", y), ("z_coordinates", z)):
        before = len(arrays)
        add(name, 1, c)
        coord_xml.append(arrays[before])

    ext = "0 %d 0 %d 0 %d" % (nx - 1, ny - 1, nz - 1)
    attr = ""
    if scalars:
        attr += ' scalars="%s"' % scalars
    if vectors:
        attr += ' vectors="%s"' % vectors
    head = ['<?xml version="1.0"?>',
            '<VTKFile type="RectilinearGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">',
            '<RectilinearGrid WholeExtent="%s">' % ext,
            '<Piece Extent="%s">' % ext,
            '<PointData%s>' % attr] + point_xml + ['</PointData>', '<CellData>', '</CellData>',
            '<Coordinates>'] + coord_xml + ['</Coordinates>', '</Piece>', '</RectilinearGrid>',
            '<AppendedData encoding="raw">', '_']
    fname = path + ".vtr"
    with open(fname, "wb") as fh:
        fh.write("\n".join(head).encode("ascii"))
        for b in blocks:
            fh.write(struct.pack("<Q", b.nbytes))
            fh.write(b.astype(b.dtype.newbyteorder("<"), copy=False).tobytes())
        fh.write(b"\n</AppendedData>\n</VTKFile>\n")
    return fname


def read_vtr(fname):
    """Parse a file written by :func:`grid_to_vtr` back into
    ``(x, y, z, {name: array | (ax, ay, az)})`` with (nx,ny,nz) C-ordered arrays."""
    import re
    with open(fname, "rb") as fh:
        raw = fh.read()
    marker = raw.index(b'<AppendedData encoding="raw">')
    start = raw.index(b"_", marker) + 1
    header = raw[:marker].decode("ascii")
    ext = [int(t) for t in re.search(r'WholeExtent="([^"]+)"', header).group(1).split()]
    nx, ny, nz = ext[1] + 1, ext[3] + 1, ext[5] + 1
    inv = {v: k for k, v in _VTK_TYPE.items()}
    out, coords = {}, {}
    for m in re.finditer(r'<DataArray Name="([^"]+)" NumberOfComponents="(\d+)" type="(\w+)" '
                         r'format="appended" offset="(\d+)"/>', header):
        name, ncomp, vt, off = m.group(1), int(m.group(2)), m.group(3), int(m.group(4))
        p = start + off
        nbytes = struct.unpack("<Q", raw[p:p + 8])[0]
        a = np.frombuffer(raw, dtype=inv[vt].newbyteorder("<"), count=nbytes // inv[vt].itemsize, offset=p + 8)
        if name.endswith("_coordinates"):
            coords[name[0]] = a.copy()
        elif ncomp == 3:
            a = a.reshape(-1, 3)
            out[name] = tuple(a[:, k].reshape((nx, ny, nz), order='F').copy() for k in range(3))
        else:
            out[name] = a.reshape((nx, ny, nz), order='F').copy()
    return coords["x"], coords["y"], coords["z"], out
