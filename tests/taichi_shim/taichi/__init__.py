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
The pointer / dense SNode tree (sparse_storage=True) is not modelled.
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
    """ti.Vector value / view of one field element."""

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
        if shape is None:
            raise NotImplementedError("taichi shim: SNode-placed fields (sparse_storage=True) are not modelled")
        if isinstance(shape, int):
            shape = (shape,)
        self.shape = tuple(shape)
        self.elem = tuple(elem)
        self.data = np.zeros(self.shape + self.elem, dtype)
        self.dtype = dtype

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

    def __getitem__(self, key):
        if self.shape == () and not self.elem:
            return self.data              # 0-d view: ti.atomic_max(fld[None], x) can write through it
        v = self.data[self._idx(key)]
        if len(self.elem) == 1:
            return v.view(Vec)
        if len(self.elem) == 2:
            return v.view(Mat)
        return v                      # scalar (np.float32 / np.int8 ...)

    def __setitem__(self, key, value):
        self.data[self._idx(key)] = np.asarray(value).astype(self.dtype, copy=False)

    def __iter__(self):               # struct-for: every index, C order
        return itertools.product(*[range(n) for n in self.shape])

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


class _Root:
    def pointer(self, *a, **k):
        raise NotImplementedError("taichi shim: the pointer SNode tree (sparse_storage=True) is not modelled")


root = _Root()
