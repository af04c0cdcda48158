"""Unit tests of tests/taichi_shim (the pure-Python stand-in the reference pin runs on): the
semantics the reference relies on must hold there, or the pin means nothing."""
import os
import sys

import numpy as np
import pytest

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "taichi_shim")


@pytest.fixture()
def ti():
    saved = {m: sys.modules.pop(m, None) for m in ("taichi",)}
    sys.path.insert(0, SHIM)
    try:
        import taichi
        yield taichi
    finally:
        sys.path.remove(SHIM)
        sys.modules.pop("taichi", None)
        for m, mod in saved.items():
            if mod is not None:
                sys.modules[m] = mod


def test_f32_arithmetic_and_casts(ti):
    e = ti.Vector([1, -1, 0])                      # i32, like self.e[k]
    u = ti.Vector([0.1, 0.2, 0.3])                 # f32
    assert e.dtype == np.int32 and u.dtype == np.float32
    d = e.dot(u)
    assert isinstance(d, np.float32)               # int operand cast to f32, no float64 promotion
    assert d == np.float32(np.float32(0.1) * np.float32(1) + np.float32(0.2) * np.float32(-1)) + np.float32(0.3) * 0
    w = np.float32(1.0 / 18.0)
    assert (w * 1.0 * (1.0 + 3.0 * d + 4.5 * d * d - 1.5 * u.dot(u))).dtype == np.float32   # feq :152-158
    assert ((e - u) / 3.0).dtype == np.float32 and (u / np.float32(2)).dtype == np.float32
    assert abs(float(u.norm()) - float(np.sqrt(np.float32(0.14)))) < 1e-7


def test_matrix_vector_product_is_the_ascending_sum(ti):
    rng = np.random.default_rng(0)
    A = ti.Matrix(rng.standard_normal((19, 19)))
    b = ti.Vector(list(rng.standard_normal(19)))
    got = A @ b
    want = np.empty(19, np.float32)
    for r in range(19):
        acc = np.float32(0)
        for k in range(19):
            acc = np.float32(acc + np.float32(A[r, k] * b[k]))
        want[r] = acc
    assert got.dtype == np.float32 and np.array_equal(np.asarray(got), want)


def test_field_views_write_through(ti):
    F = ti.Vector.field(19, ti.f32, shape=(2, 3, 4))
    F[1, 2, 3][5] = 0.25                                   # self.F[ip][s] = ...   (:266)
    assert F.to_numpy()[1, 2, 3, 5] == np.float32(0.25)
    F[ti.Vector([0, 1, 2])] = ti.Vector([1.0] * 19)        # self.f[i] = self.F[i]   (:379)
    assert np.all(F.to_numpy()[0, 1, 2] == 1.0)
    v = ti.Vector.field(3, ti.f32, shape=(2, 2, 2))
    v[0, 0, 0] += ti.Vector([1.0, 2.0, 3.0])               # self.v[i] += e_f[s]*f[i][s]   (:383)
    v[0, 0, 0] /= np.float32(2)
    assert v.to_numpy()[0, 0, 0].tolist() == [0.5, 1.0, 1.5]
    rho = ti.field(ti.f32, shape=(2, 2, 2))
    rho[1, 1, 1] += 3.0
    assert rho[1, 1, 1] == np.float32(3.0) and isinstance(rho[1, 1, 1], np.float32)
    m = ti.field(ti.f32, shape=())
    m[None] = -1e10
    ti.atomic_max(m[None], np.float32(0.5))                # cal_max_v :402
    assert float(m[None]) == 0.5
    G = ti.field(ti.f32, shape=(2, 2, 2, 19))              # the two-phase script's 4-D scalar fields
    G[ti.Vector([1, 0, 1]), 7] = 2.0
    assert G.to_numpy()[1, 0, 1, 7] == 2.0
    assert rho[5, 0, 0] == 0 and np.all(np.asarray(v[0, 9, 0]) == 0)      # reads outside the extent: 0


def test_struct_for_grouped_and_ndrange(ti):
    rho = ti.field(ti.f32, shape=(2, 3, 2))
    assert list(rho)[:3] == [(0, 0, 0), (0, 0, 1), (0, 1, 0)] and len(list(rho)) == 12
    idx = list(ti.grouped(rho))
    assert idx[4].tolist() == [0, 2, 0] and (idx[4] + ti.Vector([1, 0, 1])).tolist() == [1, 2, 1] and idx[4].y == 2
    assert list(ti.ndrange((0, 2), (1, 3))) == [(0, 1), (0, 2), (1, 1), (1, 2)]
    assert list(ti.static(range(3))) == [0, 1, 2]


def test_pointer_dense_snode_activation(ti):
    rho, f = ti.field(ti.f32), ti.Vector.field(19, ti.f32)
    cell = ti.root.pointer(ti.ijk, (2, 2, 2))
    cell.dense(ti.ijk, (3, 3, 3)).place(rho, f)            # :43-44
    assert rho.shape == (6, 6, 6) and f.to_numpy().shape == (6, 6, 6, 19)
    assert list(rho) == []                                 # nothing active yet
    assert rho[4, 4, 4] == 0                               # reading does not activate
    assert list(rho) == []
    f[4, 4, 4][3] = 1.0                                    # a write through a view activates the block ...
    cells = list(rho)                                      # ... for every field under the same pointer
    assert len(cells) == 27 and cells[0] == (3, 3, 3) and cells[-1] == (5, 5, 5)
    rho[0, 1, 2] = 2.0
    assert len(list(f)) == 54
    other = ti.field(ti.f32)
    ti.root.pointer(ti.ijk, (2, 2, 2)).dense(ti.ijk, (3, 3, 3)).place(other)
    assert list(other) == []                               # a different pointer node: its own activity
