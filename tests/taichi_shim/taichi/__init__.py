"""A minimal pure-Python stand-in for the parts of Taichi the reference solvers use.

TEST INFRASTRUCTURE ONLY.  Taichi cannot be installed in this image (no network), so the
reference's own source (/root/reference/Single_phase/LBM_3D_SinglePhase_Solver.py, imported
UNMODIFIED) is executed through this shim on tiny lattices to produce the fixtures
tests/golden/ref_*.npz (tests/golden/make_reference_fixtures.py); the oracle and the CUDA path
are then checked against what the reference's code computes.  Nothing under taichi_lbm3d_b200/
imports this.

Semantics implemented (and only these):
  * ti.field / ti.Vector.field / ti.Matrix.field with shape=(...): NumPy arrays; element access
    returns a scalar or a writable VIEW, so ``F[ip][s] = x`` and ``f[i,j,k] += v`` write through;
  * @ti.kernel / @ti.func / @ti.data_oriented / ti.static: plain Python (kernels run
    sequentially: the reference's kernels are race-free except cal_max_v's atomic_max, which is
    order-independent);
  * struct-for over a field and ti.grouped / ti.ndrange: all indices in C order;
  * arithmetic in float32 throughout (default_fp = f32): integer operands are converted when
    they meet floats, Python float literals are weak scalars (NumPy >= 2 promotion);
  * Vector.dot, Matrix @ Vector: sums in ascending index order -- Taichi unrolls them the same
    way, but its LLVM backend may reassociate / contract under fast_math, so a comparison with
    real Taichi output would still need a round-off tolerance.
  * ti.root.pointer(ijk, n).dense(ijk, b).place(fields): fields of extent n*b per axis that share
    one activity mask per pointer cell; a WRITE (also through an element view) activates the
    block, reads of inactive cells give 0, struct-fors visit the cells of active blocks only --
    the semantics the reference's sparse_storage=True mode relies on (:36-44, :164, :225).
    Reads outside a dense field's extent give 0 (Taichi leaves them undefined; the reference
    guards every such read with i<nx, :225, :262, :376).
"""
import itertools

import numpy as np

f32 = np.float32
f64 = np.float64
i32 = np.int32
i8 = np.int8
ijk = "ijk"
ijkl = "ijkl"
i = "i"
cpu = "cpu"
gpu = "gpu"


def init(*args, **kwargs):
    return None


def data_oriented(cls):
    return cls


def kernel(fn):
    return fn


def func(fn):
    return fn


def static(x):
    return x


def _coerce(a, b):
    """Taichi's implicit cast: an integer operand meeting a float becomes f32."""
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f" or b.dtype.kind == "f":
        return a.astype(np.float32, copy=False), b.astype(np.float32, copy=False)
    return a, b


class Vec(np.ndarray):
    """ti.Vector value / view of one field element (a view of an SNode-placed field activates
    its block when it is written through)."""
    _touch = None

    def __setitem__(self, key, value):
        if self._touch is not None:
            self._touch()
        np.ndarray.__setitem__(self, key, value)

    def __iadd__(self, o):
        if self._touch is not None:
            self._touch()
        return np.ndarray.__iadd__(self, np.asarray(o).astype(self.dtype, copy=False))

    def __isub__(self, o):
        if self._touch is not None:
            self._touch()
        return np.ndarray.__isub__(self, np.asarray(o).astype(self.dtype, copy=False))

    def __itruediv__(self, o):
        if self._touch is not None:
            self._touch()
        return np.ndarray.__itruediv__(self, np.asarray(o).astype(self.dtype, copy=False))

    def dot(self, other):
        a, b = _coerce(self, other)
        acc = a[0] * b[0]
        for k in range(1, a.shape[0]):
            acc = acc + a[k] * b[k]
        return acc

    def norm(self):
        return np.sqrt(self.dot(self))

    def norm_sqr(self):
        return self.dot(self)

    def sum(self):          # noqa: A003  ascending order
        acc = self[0]
        for k in range(1, self.shape[0]):
            acc = acc + self[k]
        return acc

    @property
    def x(self):
        return self[0]

    @property
    def y(self):
        return self[1]

    @property
    def z(self):
        return self[2]

    def _binary(self, other, op):
        a, b = _coerce(np.asarray(self), other)
        return op(a, b).view(Vec)

    def __add__(self, o):
        return self._binary(o, np.add)

    def __radd__(self, o):
        return self._binary(o, np.add)

    def __sub__(self, o):
        return self._binary(o, np.subtract)

    def __rsub__(self, o):
        a, b = _coerce(o, np.asarray(self))
        return np.subtract(a, b).view(Vec)

    def __mul__(self, o):
        return self._binary(o, np.multiply)

    def __rmul__(self, o):
        return self._binary(o, np.multiply)

    def __truediv__(self, o):
        a, b = _coerce(np.asarray(self, dtype=np.float32), o)
        return np.divide(a, b).view(Vec)


class Mat(np.ndarray):
    def __matmul__(self, v):
        a, b = _coerce(np.asarray(self), np.asarray(v))
        out = np.empty(a.shape[0], a.dtype)
        for r in range(a.shape[0]):
            acc = a[r, 0] * b[0]
            for k in range(1, a.shape[1]):
                acc = acc + a[r, k] * b[k]
            out[r] = acc
        return out.view(Vec)


def _is_float_list(vals):
    return any(isinstance(v, (float, np.floating)) for v in np.asarray(vals, dtype=object).reshape(-1))


def Vector(vals):  # noqa: N802
    arr = np.asarray(vals)
    if arr.dtype.kind == "f" or _is_float_list(vals):
        return np.array(vals, dtype=np.float32).view(Vec)
    return np.array(vals, dtype=np.int32).view(Vec)


def Matrix(vals):  # noqa: N802
    arr = np.asarray(vals)
    if arr.dtype.kind == "f":
        return arr.astype(np.float32).view(Mat)
    return arr.astype(np.int32).view(Mat)


class _Field:
    def __init__(self, dtype, shape, elem=()):
        self.elem = tuple(elem)
        self.dtype = dtype
        self.block = None             # SNode-placed: cells per pointer cell, activity mask (shared)
        self.active = None
        self.shape = None
        self.data = None
        if shape is not None:
            self._allocate((shape,) if isinstance(shape, int) else tuple(shape))

    def _allocate(self, shape):
        self.shape = tuple(shape)
        self.data = np.zeros(self.shape + self.elem, self.dtype)

    @staticmethod
    def _idx(key):
        if key is None:
            return ()
        if isinstance(key, np.ndarray):
            return tuple(int(k) for k in key)
        if isinstance(key, tuple):          # F[ip, s]: a vector index followed by scalars
            out = []
            for k in key:
                out.extend(int(q) for q in k) if isinstance(k, np.ndarray) and k.ndim == 1 else out.append(int(k))
            return tuple(out)
        return (int(key),)

    def _inside(self, idx):
        return all(0 <= i < n for i, n in zip(idx, self.shape))

    def _activate(self, idx):
        if self.active is not None:
            self.active[tuple(i // b for i, b in zip(idx, self.block))] = True

    def __getitem__(self, key):
        if self.shape == () and not self.elem:
            return self.data              # 0-d view: ti.atomic_max(fld[None], x) can write through it
        idx = self._idx(key)
        if not self._inside(idx[:len(self.shape)]):
            return np.zeros(self.elem, self.dtype).view(Vec) if self.elem else self.dtype(0)
        v = self.data[idx]
        if len(self.elem) == 1:
            v = v.view(Vec)
            if self.active is not None:
                v._touch = lambda: self._activate(idx)
            return v
        if len(self.elem) == 2:
            return v.view(Mat)
        return v                      # scalar (np.float32 / np.int8 ...)

    def __setitem__(self, key, value):
        idx = self._idx(key)
        self._activate(idx[:len(self.shape)])
        self.data[idx] = np.asarray(value).astype(self.dtype, copy=False)

    def __iter__(self):               # struct-for: every index (of the active blocks), C order
        if self.active is None:
            return itertools.product(*[range(n) for n in self.shape])
        return (idx for idx in itertools.product(*[range(n) for n in self.shape])
                if self.active[tuple(i // b for i, b in zip(idx, self.block))])

    def from_numpy(self, arr):
        self.data[...] = np.asarray(arr).astype(self.dtype).reshape(self.data.shape)

    def to_numpy(self):
        return self.data.copy()

    def fill(self, val):
        self.data[...] = val


def field(dtype, shape=None):
    return _Field(dtype, shape)


def _vector_field(n, dtype, shape=None):
    return _Field(dtype, shape, (n,))


def _matrix_field(n, m, dtype, shape=None):
    return _Field(dtype, shape, (n, m))


Vector.field = _vector_field
Matrix.field = _matrix_field


def grouped(fld):
    for idx in fld:
        yield np.array(idx, dtype=np.int32).view(Vec)


def ndrange(*ranges):
    its = [range(r[0], r[1]) if isinstance(r, (tuple, list)) else range(r) for r in ranges]
    return itertools.product(*its)


def atomic_max(target, value):
    """target is the 0-d view a scalar field returns for fld[None]"""
    if value > target:
        target[...] = value


class _SNode:
    """pointer / dense cells along ijk(l): just enough of the SNode tree for
    ti.root.pointer(axes, n).dense(axes, b).place(...).  The activity mask belongs to the
    POINTER node: every field placed below it (through any of its dense children) shares it."""

    def __init__(self, cells=(), block=None, active=None):
        self.cells, self.block, self.active = tuple(cells), block, active

    def pointer(self, axes, dims):
        if self.cells:
            raise NotImplementedError("taichi shim: one pointer level only")
        return _SNode(dims, None, np.zeros(tuple(dims), bool))

    def dense(self, axes, dims):
        if not self.cells or self.block is not None:
            raise NotImplementedError("taichi shim: dense directly under one pointer level only")
        return _SNode(self.cells, tuple(dims), self.active)

    def place(self, *fields):
        if self.block is None:
            raise NotImplementedError("taichi shim: place under pointer().dense() only")
        for f in fields:
            f._allocate(tuple(c * b for c, b in zip(self.cells, self.block)))
            f.block, f.active = self.block, self.active


root = _SNode()
