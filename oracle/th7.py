"""A minimal Torch7 tensor model -- TEST INFRASTRUCTURE, NOT PRODUCT.

Just enough of the `torch.Tensor` surface to let oracle/lua_literal.py restate the reference's
criterion files STATEMENT BY STATEMENT instead of as closed forms: 1-based inclusive index tables that return views
sharing storage, in-place methods that return self so they chain, `r:add(a, v, b)` that overwrites, reductions that
keep the reduced dimension, operator overloads that allocate, Byte masks from `torch.ge / le` that `:cuda()` turns
into floats.  Backed by numpy views; `dtype()` selects the arithmetic type (float64 to compare with the closed
forms at 1e-12, float32 to follow the reference's rounding).

Deliberately NOT modelled: `resizeAs` of a size-mismatched view (SURVEY Q9 -- first-order SmoothnessCriterion has its
own storage-level model, oracle/b2f_oracle.py:_TorchStorageTensor); any such call raises.
"""
from __future__ import annotations

import contextlib

import numpy as np

_DT = [np.float64]
ALL = slice(None)


@contextlib.contextmanager
def dtype(dt):
    _DT.append(dt)
    try:
        yield
    finally:
        _DT.pop()


def _dt():
    return _DT[-1]


def _num(v):
    return _dt()(v)


class Tensor:
    """A view of a numpy array with Torch7 method semantics.  1-based `size(d)`, `t[r1, r2, ...]` with ranges
    `ALL`, `(i,)` (= {i}: a range of one, keeps the dimension) or `(i, j)` (inclusive)."""

    def __init__(self, a):
        self.a = a

    # ---- shape --------------------------------------------------------------------
    def size(self, d=None):
        return tuple(self.a.shape) if d is None else self.a.shape[d - 1]

    def nElement(self):
        return int(self.a.size)

    def clone(self):
        return Tensor(self.a.copy())

    def new(self):
        return Tensor(np.empty((0,), self.a.dtype))

    def resize(self, *size):
        if len(size) == 1 and isinstance(size[0], tuple):
            size = size[0]
        if tuple(self.a.shape) != tuple(size):
            if self.a.base is not None:
                raise NotImplementedError("resize of a view (Torch7 would re-lay out shared storage: SURVEY Q9)")
            self.a = np.full(size, np.nan, self.a.dtype)   # uninitialised: NaN so that a read shows up
        return self

    def resizeAs(self, other):
        return self.resize(*other.a.shape)

    def zero(self):
        self.a[...] = 0
        return self

    def fill(self, v):
        self.a[...] = v
        return self

    def copy(self, src):
        assert src.a.size == self.a.size
        self.a[...] = src.a.reshape(self.a.shape)
        return self

    def cuda(self):
        # Byte mask -> float tensor (CudaTensor); float tensors are returned as they are
        if self.a.dtype == np.uint8:
            return Tensor(self.a.astype(_dt()))
        return self

    def double(self):
        return Tensor(self.a.astype(np.float64))

    # ---- index tables ---------------------------------------------------------------
    @staticmethod
    def _key(key):
        if not isinstance(key, tuple):
            key = (key,)
        out = []
        for k in key:
            if isinstance(k, slice):
                out.append(k)
            elif isinstance(k, (int, np.integer)):     # a number selects (the dimension is dropped)
                out.append(int(k) - 1)
            else:
                lo, hi = (k[0], k[0]) if len(k) == 1 else k
                out.append(slice(lo - 1, hi))
        return tuple(out)

    def __getitem__(self, key):
        return Tensor(self.a[self._key(key)])

    def __setitem__(self, key, value):
        dst = self.a[self._key(key)]
        if isinstance(value, Tensor):
            assert value.a.size == dst.size
            dst[...] = value.a.reshape(dst.shape)
        else:
            dst[...] = value

    def transpose(self, d1, d2):
        return Tensor(np.swapaxes(self.a, d1 - 1, d2 - 1))

    def repeatTensor(self, *reps):
        a = self.a
        if a.ndim < len(reps):
            a = a.reshape((1,) * (len(reps) - a.ndim) + a.shape)
        return Tensor(np.tile(a, reps))

    # ---- in-place arithmetic (all return self) -----------------------------------------
    def _put(self, value, ref_shape):
        if tuple(self.a.shape) != tuple(ref_shape):
            if self.a.size != int(np.prod(ref_shape)):
                raise NotImplementedError("result resize of a size-mismatched tensor (SURVEY Q9)")
            value = value.reshape(self.a.shape)
        self.a[...] = value
        return self

    def add(self, *args):
        if len(args) == 1:                       # r:add(tensor) / r:add(number): r += x
            x = args[0]
            self.a[...] = self.a + (x.a.reshape(self.a.shape) if isinstance(x, Tensor) else _num(x))
            return self
        if len(args) == 2 and not isinstance(args[0], Tensor):   # r:add(v, tensor): r += v * b
            v, b = args
            self.a[...] = self.a + _num(v) * b.a.reshape(self.a.shape)
            return self
        if len(args) == 2:                       # r:add(a, b): r = a + b (overwrite)
            a, b = args
            return self._put(a.a + (b.a if isinstance(b, Tensor) else _num(b)), a.a.shape)
        a, v, b = args                           # r:add(a, v, b): r = a + v * b (overwrite)
        return self._put(a.a + _num(v) * b.a, a.a.shape)

    def cmul(self, b):
        self.a[...] = self.a * b.a.reshape(self.a.shape) if self.a.dtype != np.uint8 else (self.a & b.a)
        return self

    def cdiv(self, b):
        self.a[...] = self.a / b.a.reshape(self.a.shape)
        return self

    def mul(self, k):
        self.a[...] = self.a * _num(k)
        return self

    def div(self, k):
        self.a[...] = self.a / _num(k)
        return self

    def pow(self, e):
        self.a[...] = np.power(self.a, _num(e))
        return self

    def sqrt(self):
        self.a[...] = np.sqrt(self.a)
        return self

    def abs(self):
        self.a[...] = np.abs(self.a)
        return self

    def sum(self, dim=None):
        if dim is None:
            return float(self.a.sum(dtype=np.float64))
        return sum_(self, dim)

    def view(self, *size):
        return Tensor(self.a.reshape(size))

    # ---- operator overloads (each allocates and rounds separately) ----------------------------
    def __add__(self, o):
        return Tensor(self.a + (o.a if isinstance(o, Tensor) else _num(o)))

    __radd__ = __add__

    def __sub__(self, o):
        return Tensor(self.a - (o.a if isinstance(o, Tensor) else _num(o)))

    def __rsub__(self, o):
        return Tensor(_num(o) - self.a)

    def __mul__(self, o):
        assert not isinstance(o, Tensor), "tensor * tensor is a matrix product in Torch7"
        return Tensor(self.a * _num(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        assert not isinstance(o, Tensor)
        return Tensor(self.a / _num(o))

    def __neg__(self):
        return Tensor(-self.a)


# ---- the `torch.` functions the criterion files call (all allocate) ---------------------------

def tensor(array):
    """A Tensor holding a copy of `array` in the working dtype."""
    return Tensor(np.array(array, dtype=_dt()))


def Tensor_(*size):
    """torch.Tensor(sizes): uninitialised (NaN, so that a read before :zero() shows up)."""
    if len(size) == 1 and isinstance(size[0], tuple):
        size = size[0]
    return Tensor(np.full(size, np.nan, _dt()))


def range_(lo, hi):
    return Tensor(np.arange(lo, hi + 1, dtype=_dt()))


def add(a, *rest):
    if len(rest) == 1:
        b = rest[0]
        return Tensor(a.a + (b.a if isinstance(b, Tensor) else _num(b)))
    v, b = rest
    return Tensor(a.a + _num(v) * b.a)


def cmul(a, b):
    return Tensor(a.a * b.a)


def cdiv(a, b):
    return Tensor(a.a / b.a)


def mul(a, k):
    return Tensor(a.a * _num(k))


def div(a, k):
    return Tensor(a.a / _num(k))


def pow_(a, e):
    return Tensor(np.power(a.a, _num(e)))


def abs_(a):
    return Tensor(np.abs(a.a))


def log(a):
    return Tensor(np.log(a.a))


def exp(a):
    return Tensor(np.exp(a.a))


def sum_(a, dim):
    return Tensor(a.a.sum(axis=dim - 1, keepdims=True).astype(a.a.dtype))


def mean(a, dim):
    return Tensor(a.a.mean(axis=dim - 1, keepdims=True).astype(a.a.dtype))


def ge(a, v):
    return Tensor((a.a >= _num(v)).astype(np.uint8))


def le(a, v):
    return Tensor((a.a <= _num(v)).astype(np.uint8))


def repeatTensor(a, *reps):
    return a.repeatTensor(*reps)


def expandAs(a, ref):
    return Tensor(np.broadcast_to(a.a, ref.a.shape))


class LuaTable:
    """A 1-based Lua array."""

    def __init__(self, items=()):
        self.items = list(items)

    def __getitem__(self, k):
        return self.items[k - 1]

    def __setitem__(self, k, v):
        if k == len(self.items) + 1:
            self.items.append(v)
        else:
            self.items[k - 1] = v

    def __len__(self):
        return len(self.items)

    def insert(self, v):
        self.items.append(v)
