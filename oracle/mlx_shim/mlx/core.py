"""NumPy restatement of the `mlx.core` primitives the reference calls (TEST INFRASTRUCTURE ONLY).

Why this exists: the reference (jishnuvenugopal/pocket-tts-mlx) delegates all arithmetic to the
third-party package `mlx` (`pyproject.toml:27`, pinned only as `mlx>=0.20.0`), which is not
installable in this image.  This module restates the *published* semantics of exactly the
`mx.*` entry points the reference's hot path uses (SURVEY.md section 2.3 lists the call sites), so
that `oracle/gen_golden.py` can execute the reference's own, unmodified Python from
`/root/reference` and record golden vectors.  It is never imported by the product package.

Semantics restated (MLX docs):
  * default floating dtype is float32; float64 results are narrowed to float32 (MLX has no
    float64 arithmetic on its default device);
  * `conv_transpose1d(x[N,L,C_in], w[C_out,K,C_in/groups])` is the gradient of `conv1d`
    (scatter form: y[n, l*stride + j, o] += x[n,l,c] * w[o,j,c]);
  * `var(ddof=d)` divides by N-d; `softmax` is the max-subtracted exponential normalisation;
  * `slice_update(a, u, start, axes)` returns a copy of `a` with the block at `start` replaced.
"""

from __future__ import annotations

import numpy as np

nan = float("nan")
inf = float("inf")


class Dtype:
    """Stand-in for mlx.core.Dtype (has `.size`, compares by value)."""

    def __init__(self, np_dtype):
        self._np = np.dtype(np_dtype)
        self.size = self._np.itemsize
        self.name = self._np.name

    def __eq__(self, other):
        return _np_dtype(other) == self._np if other is not None else False

    def __hash__(self):
        return hash(self._np)

    def __repr__(self):
        return f"mlx_shim.{self._np.name}"


float32 = Dtype(np.float32)
float16 = Dtype(np.float16)
int64 = Dtype(np.int64)
int32 = Dtype(np.int32)
uint32 = Dtype(np.uint32)
bool_ = Dtype(np.bool_)


def _np_dtype(d):
    if d is None:
        return None
    if isinstance(d, Dtype):
        return d._np
    return np.dtype(d)


def _narrow(a: np.ndarray) -> np.ndarray:
    if a.dtype == np.float64:
        return a.astype(np.float32)
    return a


class array(np.ndarray):
    """ndarray view subclass: float64 never escapes, `.dtype` is a shim Dtype."""

    def __new__(cls, value, dtype=None):
        base = np.asarray(value)
        if base.dtype == np.float64:
            base = base.astype(np.float32)
        if dtype is not None:
            base = base.astype(_np_dtype(dtype))
        return base.view(cls)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        raw = tuple(np.asarray(i) if isinstance(i, array) else i for i in inputs)
        if out is not None:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, array) else o for o in out)
        res = getattr(ufunc, method)(*raw, **kwargs)
        if isinstance(res, tuple):
            return tuple(_wrap(r) for r in res)
        return _wrap(res)

    @property
    def dtype(self):  # python-level only; numpy internals keep the C dtype
        return Dtype(np.asarray(self).dtype)

    def astype(self, dtype, *a, **k):
        return _wrap(np.asarray(self).astype(_np_dtype(dtype)))

    def item(self):
        return np.asarray(self).item()

    def tolist(self):
        return np.asarray(self).tolist()

    def __getitem__(self, idx):
        if isinstance(idx, array):
            idx = np.asarray(idx)
        elif isinstance(idx, tuple):
            idx = tuple(np.asarray(i) if isinstance(i, array) else i for i in idx)
        return _wrap(np.asarray(self)[idx])


def _wrap(x):
    if isinstance(x, np.ndarray):
        return _narrow(np.asarray(x)).view(array)
    if isinstance(x, np.generic):
        return _narrow(np.asarray(x)).view(array)
    return x


def _raw(x):
    if isinstance(x, array):
        return np.asarray(x)
    if isinstance(x, (list, tuple)):
        return [_raw(v) for v in x]
    return x


# ---- creation ---------------------------------------------------------------------------

def zeros(shape, dtype=float32):
    return _wrap(np.zeros(shape, dtype=_np_dtype(dtype)))


def ones(shape, dtype=float32):
    return _wrap(np.ones(shape, dtype=_np_dtype(dtype)))


def full(shape, vals, dtype=None):
    if dtype is None:
        v = np.asarray(_raw(vals))
        dt = np.float32 if v.dtype.kind == "f" else v.dtype
    else:
        dt = _np_dtype(dtype)
    return _wrap(np.full(shape, _raw(vals), dtype=dt))


def zeros_like(a):
    return _wrap(np.zeros_like(_raw(a)))


def arange(*args, dtype=None):
    out = np.arange(*args)
    if dtype is not None:
        out = out.astype(_np_dtype(dtype))
    elif out.dtype.kind == "i":
        out = out.astype(np.int32)
    return _wrap(out)


def linspace(start, stop, num=50, dtype=float32):
    return _wrap(np.linspace(start, stop, num).astype(_np_dtype(dtype)))


def tril(x, k=0):
    return _wrap(np.tril(_raw(x), k))


# ---- shape ------------------------------------------------------------------------------

def concatenate(arrays, axis=0):
    return _wrap(np.concatenate([_raw(a) for a in arrays], axis=axis))


def stack(arrays, axis=0):
    return _wrap(np.stack([_raw(a) for a in arrays], axis=axis))


def split(a, indices_or_sections, axis=0):
    return [_wrap(p) for p in np.split(_raw(a), indices_or_sections, axis=axis)]


def transpose(a, axes=None):
    return _wrap(np.transpose(_raw(a), axes))


def broadcast_to(a, shape):
    return _wrap(np.broadcast_to(_raw(a), shape))


def pad(a, pad_width, mode="constant", constant_values=0):
    return _wrap(np.pad(_raw(a), pad_width, mode=mode, constant_values=constant_values))


def slice_update(a, update, start_indices, axes):
    out = np.array(_raw(a), copy=True)
    upd = _raw(update)
    starts = [int(s) for s in np.asarray(_raw(start_indices)).tolist()]
    index = [slice(None)] * out.ndim
    for ax, st in zip(axes, starts):
        index[ax] = slice(st, st + upd.shape[ax])
    out[tuple(index)] = upd
    return _wrap(out)


# ---- elementwise / reductions -------------------------------------------------------------

def where(c, a, b):
    return _wrap(np.where(_raw(c), _raw(a), _raw(b)))


def isnan(a):
    return _wrap(np.isnan(_raw(a)))


def exp(a):
    return _wrap(np.exp(_raw(a)))


def cos(a):
    return _wrap(np.cos(_raw(a)))


def sin(a):
    return _wrap(np.sin(_raw(a)))


def sqrt(a):
    return _wrap(np.sqrt(_raw(a)))


def rsqrt(a):
    r = np.asarray(_raw(a))
    return _wrap((1.0 / np.sqrt(r)).astype(r.dtype if r.dtype.kind == "f" else np.float32))


def sigmoid(a):
    r = np.asarray(_raw(a))
    return _wrap(1.0 / (1.0 + np.exp(-r)))


def erf(a):
    from scipy.special import erf as _erf

    r = np.asarray(_raw(a))
    return _wrap(_erf(r).astype(r.dtype))


def clip(a, lo, hi):
    return _wrap(np.clip(_raw(a), lo, hi))


def mean(a, axis=None, keepdims=False):
    return _wrap(np.mean(_raw(a), axis=axis, keepdims=keepdims))


def var(a, axis=None, keepdims=False, ddof=0):
    return _wrap(np.var(_raw(a), axis=axis, keepdims=keepdims, ddof=ddof))


def softmax(a, axis=-1):
    r = np.asarray(_raw(a))
    m = np.max(r, axis=axis, keepdims=True)
    e = np.exp(r - m)
    return _wrap(e / np.sum(e, axis=axis, keepdims=True))


def matmul(a, b):
    return _wrap(np.matmul(_raw(a), _raw(b)))


def eval(*args):  # MLX is lazy; NumPy is not
    return None


# ---- convolution ------------------------------------------------------------------------

def conv1d(x, w, stride=1, padding=0, dilation=1, groups=1):
    """x [N,L,C_in], w [C_out,K,C_in/groups] -> [N,L_out,C_out] (cross-correlation)."""
    x = np.asarray(_raw(x))
    w = np.asarray(_raw(w))
    if padding:
        x = np.pad(x, [(0, 0), (padding, padding), (0, 0)])
    n, length, c_in = x.shape
    c_out, k, c_in_g = w.shape
    span = (k - 1) * dilation + 1
    l_out = (length - span) // stride + 1
    og = c_out // groups
    y = np.zeros((n, l_out, c_out), dtype=np.result_type(x.dtype, w.dtype))
    for g in range(groups):
        xs = x[:, :, g * c_in_g:(g + 1) * c_in_g]
        ws = w[g * og:(g + 1) * og]
        for j in range(k):
            tap = xs[:, j * dilation: j * dilation + (l_out - 1) * stride + 1: stride, :]
            y[:, :, g * og:(g + 1) * og] += tap @ ws[:, j, :].T
    return _wrap(y)


def conv_transpose1d(x, w, stride=1, padding=0, dilation=1, output_padding=0, groups=1):
    """x [N,L,C_in], w [C_out,K,C_in/groups] -> [N,(L-1)*stride+(K-1)*dilation+1,C_out]."""
    x = np.asarray(_raw(x))
    w = np.asarray(_raw(w))
    n, length, c_in = x.shape
    c_out, k, c_in_g = w.shape
    og = c_out // groups
    l_full = (length - 1) * stride + (k - 1) * dilation + 1 + output_padding
    y = np.zeros((n, l_full, c_out), dtype=np.result_type(x.dtype, w.dtype))
    for g in range(groups):
        xs = x[:, :, g * c_in_g:(g + 1) * c_in_g]
        ws = w[g * og:(g + 1) * og]
        for j in range(k):
            contrib = xs @ ws[:, j, :].T  # [N, L, og]
            y[:, j * dilation: j * dilation + (length - 1) * stride + 1: stride,
              g * og:(g + 1) * og] += contrib
    if padding:
        y = y[:, padding: l_full - padding]
    return _wrap(y)


# ---- random -----------------------------------------------------------------------------

class _Random:
    """`mx.random.normal` backed by a recordable NumPy generator.

    gen_golden.py seeds it and reads `draws` back so that the very same noise can be
    injected into the oracle and the CUDA path."""

    def __init__(self):
        self.rng = np.random.Generator(np.random.PCG64(0))
        self.draws = []

    def seed(self, s):
        self.rng = np.random.Generator(np.random.PCG64(s))
        self.draws = []

    def normal(self, shape=(), dtype=float32, loc=0.0, scale=1.0, key=None):
        shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        z = self.rng.standard_normal(shape).astype(np.float32)
        self.draws.append(z.copy())
        return _wrap((z * np.float32(scale) + np.float32(loc)).astype(_np_dtype(dtype)))


random = _Random()
