"""Stand-in `tensorflow` module backed by torch-CPU.  TEST INFRASTRUCTURE ONLY.

TensorFlow is not installed in this image, so the reference's four hot-path
files (src/tf_smpl/batch_smpl.py, batch_lbs.py, projection.py, src/ops.py)
cannot run as-is.  This module implements exactly the ~45 `tf.*` symbols those
files touch so that they execute UNCHANGED from
/root/reference, with torch autograd standing in for `tf.GradientTape`.

Only `oracle/run_reference.py` and `oracle/make_golden.py` put this directory
on sys.path; nothing in the product imports it.

`set_float(torch.float64)` switches what `tf.float32` means, so the same
reference code can be run as an fp64 oracle.
"""
import contextlib

import numpy as np
import torch

_FLOAT = torch.float32


def set_float(dt):
    global _FLOAT, float32
    _FLOAT = dt
    float32 = dt


float32 = _FLOAT
int32 = torch.int32
int64 = torch.int64


class _Shape(tuple):
    def as_list(self):
        return list(self)


class Tensor(torch.Tensor):
    """torch.Tensor whose .shape has TF's .as_list() (batch_lbs.py:23,47)."""

    @property
    def shape(self):
        return _Shape(super().shape)

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).numpy()


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        y = x if isinstance(x, Tensor) else x.as_subclass(Tensor)
        if dtype is not None and y.dtype != dtype:
            y = y.to(dtype)
        return y
    a = np.asarray(x)
    if dtype is None:
        if a.dtype.kind == "f":
            dtype = _FLOAT
        elif a.dtype.kind in "iu":
            dtype = torch.int64
        elif a.dtype.kind == "b":
            dtype = torch.bool
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).as_subclass(Tensor)


def convert_to_tensor(x, dtype=None):
    return _t(x, dtype)


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield


def constant(value, dtype=None, name=None):
    return _t(value, dtype)


def Variable(value, name=None, dtype=None, trainable=True):
    if isinstance(value, np.matrix):
        value = np.asarray(value)
    return _t(value, dtype)


def reshape(x, shape, name=None):
    shape = [int(s) for s in shape] if not isinstance(shape, torch.Tensor) else [int(s) for s in shape.tolist()]
    return _t(x).reshape(shape)


def shape(x):
    return _Shape(_t(x).shape)


def range(start, limit=None, delta=1):
    if limit is None:
        start, limit = 0, start
    return torch.arange(start, limit, delta).as_subclass(Tensor)


def stack(values, axis=0, name=None):
    return torch.stack([_t(v) for v in values], dim=axis).as_subclass(Tensor)


def concat(values, axis, name=None):
    return torch.cat([_t(v) for v in values], dim=axis).as_subclass(Tensor)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def tile(x, multiples):
    return _t(x).repeat(*[int(m) for m in multiples])


def eye(n, dtype=None):
    return torch.eye(n, dtype=dtype or _FLOAT).as_subclass(Tensor)


def ones(shape, dtype=None):
    return torch.ones([int(s) for s in shape], dtype=dtype or _FLOAT).as_subclass(Tensor)


def zeros(shape, dtype=None):
    return torch.zeros([int(s) for s in shape], dtype=dtype or _FLOAT).as_subclass(Tensor)


def ones_like(x):
    return torch.ones_like(_t(x))


def pad(x, paddings):
    flat = []
    for lo, hi in reversed(list(paddings)):
        flat += [int(lo), int(hi)]
    return torch.nn.functional.pad(_t(x), flat)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def norm(x, ord="euclidean", axis=None, name=None):
    x = _t(x)
    if ord == "euclidean" or ord == 2:
        if axis is None:
            return torch.sqrt(torch.sum(x * x))
        return torch.sqrt(torch.sum(x * x, dim=axis))
    if ord == 1:
        return torch.sum(torch.abs(x), dim=axis)
    raise NotImplementedError(ord)


def cos(x):
    return torch.cos(_t(x))


def sin(x):
    return torch.sin(_t(x))


def abs(x):
    return torch.abs(_t(x))


def square(x):
    x = _t(x)
    return x * x


def scatter_nd(indices, updates, shape):
    # TF semantics: zeros(shape) with updates ACCUMULATED at indices.
    indices = _t(indices)
    assert indices.shape[-1] == 1 and len(shape) == 1
    out = torch.zeros([int(shape[0])], dtype=updates.dtype).as_subclass(Tensor)
    return out.index_add(0, indices.reshape(-1).long(), _t(updates).reshape(-1))


def multiply(a, b):
    return _t(a) * _t(b, _t(a).dtype if not isinstance(b, torch.Tensor) else None)


def add(a, b):
    return _t(a) + b


def scalar_mul(s, x):
    return s * _t(x)


def reduce_sum(x, axis=None):
    return torch.sum(_t(x)) if axis is None else torch.sum(_t(x), dim=axis)


def reduce_mean(x, axis=None):
    return torch.mean(_t(x)) if axis is None else torch.mean(_t(x), dim=axis)


def argmin(x, axis=0):
    # torch.argmin returns the FIRST minimal index, like TF.
    return torch.argmin(_t(x), dim=axis)


def gather(params, indices, axis=0):
    return torch.index_select(_t(params), axis, _t(indices).long().reshape(-1))


def gather_nd(params, indices):
    indices = _t(indices).long()
    assert indices.shape[-1] == 1
    return _t(params)[indices[:, 0]]


def where(cond):
    return torch.nonzero(_t(cond)).as_subclass(Tensor)  # row-major, like tf.where


def equal(a, b):
    return _t(a) == b


def greater(a, b):
    return _t(a) > b


def cast(x, dtype):
    return _t(x).to(dtype)


def print(*a, **k):
    import builtins

    builtins.print(*[v.item() if isinstance(v, torch.Tensor) and v.numel() == 1 else v for v in a])


def tensordot(a, b, axes):
    return torch.tensordot(_t(a), _t(b), dims=axes).as_subclass(Tensor)


def transpose(x, perm=None):
    x = _t(x)
    if perm is None:
        perm = list(reversed(range(x.dim())))
    return x.permute(*perm)


class _Linalg:
    @staticmethod
    def diag_part(x):
        return torch.diagonal(_t(x), dim1=-2, dim2=-1)


linalg = _Linalg()


class _Math:
    @staticmethod
    def divide(a, b):
        return _t(a) / b

    @staticmethod
    def add(a, b):
        return _t(a) + b


math = _Math()


class _Losses:
    @staticmethod
    def absolute_difference(labels, predictions, weights=1.0):
        """tf.compat.v1.losses.absolute_difference with the default reduction
        SUM_BY_NONZERO_WEIGHTS: sum(|pred-labels|*w) / #{w != 0 after broadcast
        to the loss shape}, 0 if that count is 0 (div_no_nan)."""
        labels, predictions = _t(labels), _t(predictions)
        losses = torch.abs(predictions - labels)
        w = _t(weights).to(losses.dtype)
        weighted = losses * w
        present = (w != 0).to(losses.dtype) * torch.ones_like(losses)
        n = present.sum()
        total = weighted.sum()
        if float(n) == 0.0:
            return total * 0.0
        return total / n


class _V1:
    losses = _Losses()


class _Compat:
    v1 = _V1()


compat = _Compat()


# ---- the two symbols hpe_b200/tf_adapter.py needs ----------------
def _set_shape(self, shape):
    """tf.Tensor.set_shape: a static-shape assertion; here it only checks."""
    want = [int(s) if s is not None else None for s in list(shape)]
    have = list(self.shape)
    assert len(want) == len(have) and all(w is None or w == h for w, h in zip(want, have)), (want, have)


Tensor.set_shape = _set_shape


def numpy_function(func, inp, Tout, name=None):
    """tf.numpy_function: call `func` on numpy copies of the inputs; outputs come back as tensors of
    dtypes Tout (outside the autograd graph, as in TensorFlow)."""
    args = [x.detach().as_subclass(torch.Tensor).numpy() if isinstance(x, torch.Tensor) else np.asarray(x) for x in inp]
    out = func(*args)
    single = not isinstance(Tout, (list, tuple))
    outs = [out] if single else list(out)
    touts = [Tout] if single else list(Tout)
    res = [_t(np.asarray(o), dt) for o, dt in zip(outs, touts)]
    return res[0] if single else res


def custom_gradient(f):
    """tf.custom_gradient: f(*args) -> (outputs, grad_fn); grad_fn(*upstream) -> gradients w.r.t. args.
    Implemented as a torch.autograd.Function; outputs that receive no gradient get None upstream is
    NOT TensorFlow's behaviour (it passes zeros), so zeros are materialised."""

    def wrapped(*args):
        targs = [_t(a) for a in args]
        holder = {}

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *xs):
                outs, grad_fn = f(*[x.as_subclass(Tensor) for x in xs])
                holder["grad"] = grad_fn
                holder["single"] = not isinstance(outs, (list, tuple))
                outs = [outs] if holder["single"] else list(outs)
                holder["like"] = [o.detach() for o in outs]
                return tuple(o.detach().as_subclass(torch.Tensor) for o in outs)

            @staticmethod
            def backward(ctx, *ups):
                ups = [(_t(u) if u is not None else _t(torch.zeros_like(l))) for u, l in zip(ups, holder["like"])]
                g = holder["grad"](*ups)
                g = list(g) if isinstance(g, (list, tuple)) else [g]
                return tuple(None if x is None else x.as_subclass(torch.Tensor) for x in g)

        outs = _Fn.apply(*[a.as_subclass(torch.Tensor) for a in targs])
        outs = [o.as_subclass(Tensor) for o in outs]
        return outs[0] if holder["single"] else tuple(outs)

    return wrapped
