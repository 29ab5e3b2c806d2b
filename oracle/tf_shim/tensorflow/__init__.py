"""NumPy stand-in for the TensorFlow primitives the magenta/ddsp hot path uses.

TEST INFRASTRUCTURE ONLY.  TensorFlow cannot be installed in this image (no
network, Python 3.12), so the UNMODIFIED reference sources under
/root/reference/ddsp are imported on top of this package instead
(oracle/ref_on_shim.py): every line of ddsp/core.py, synths.py, processors.py,
dags.py, losses.py, spectral_ops.py runs as written - op order, shapes, index
arithmetic, error checks - and only the ~70 TF primitives below are ours.  Each
keeps TF's eager semantics where they are observable: float32 arithmetic (Python
scalars and float64 ndarrays are cast to the tensor's dtype, never the reverse),
sequential float32 cumsum, float32/complex64 FFTs, `tf.signal` framing /
overlap-add / periodic Hann, and the legacy (`tf.compat.v1`, no half-pixel)
image resize kernels.  What cannot be reproduced is the last ulp of Eigen's
elementary functions (sin, exp, pow) and the association order of its
reductions; numpy's are used.

Nothing under ddsp_b200/ imports this; only tests/, oracle/ and the golden
fixture generator do.
"""
import builtins
import math as _math
import types as _types

import numpy as np
import scipy.fft as _sfft

__version__ = '2.11.0-numpy-shim'
newaxis = None


# ----------------------------------------------------------------------------
# dtypes
# ----------------------------------------------------------------------------
class DType:
  def __init__(self, name, np_dtype):
    self.name = name
    self.as_numpy_dtype = np_dtype

  def __repr__(self):
    return 'tf.' + self.name

  def __eq__(self, other):
    return isinstance(other, DType) and other.name == self.name

  def __hash__(self):
    return hash(self.name)

  @property
  def is_complex(self):
    return self.name.startswith('complex')


float16 = DType('float16', np.float16)
float32 = DType('float32', np.float32)
float64 = DType('float64', np.float64)
int32 = DType('int32', np.int32)
int64 = DType('int64', np.int64)
bool = DType('bool', np.bool_)   # pylint: disable=redefined-builtin
complex64 = DType('complex64', np.complex64)
complex128 = DType('complex128', np.complex128)
_DTYPES = {np.dtype(d.as_numpy_dtype): d for d in
           (float16, float32, float64, int32, int64, bool, complex64, complex128)}


# "Wide" mode: every float32 the reference asks for becomes float64 (and
# complex64 -> complex128), so the UNMODIFIED reference code is evaluated in
# double precision - the arbiter for the 1e-4 parity gate (the reference's own
# float32 phase accumulation drifts by 1e-2 .. 3e-1).
_WIDE = [False]


def set_wide(flag):
  _WIDE[0] = builtins.bool(flag)


def is_wide():
  return _WIDE[0]


def _F32():
  return np.dtype(np.float64 if _WIDE[0] else np.float32)


def _np_dtype(dtype):
  if dtype is None:
    return None
  nd = np.dtype(dtype.as_numpy_dtype) if isinstance(dtype, DType) else np.dtype(dtype)
  if _WIDE[0]:
    if nd == np.float32:
      return np.dtype(np.float64)
    if nd == np.complex64:
      return np.dtype(np.complex128)
  return nd


class TensorShape(tuple):
  def as_list(self):
    return list(self)

  @property
  def rank(self):
    return len(self)

  ndims = rank


# ----------------------------------------------------------------------------
# Tensor
# ----------------------------------------------------------------------------
class Tensor:
  """Eager tensor: an immutable view of a numpy array with TF's operator
  semantics (the other operand is converted to THIS tensor's dtype)."""
  __slots__ = ('_a',)
  __array_priority__ = 100

  def __init__(self, a):
    self._a = a

  # -- interop --
  def numpy(self):
    return self._a

  def __array__(self, dtype=None, copy=None):
    return self._a if dtype is None else self._a.astype(dtype)

  @property
  def shape(self):
    return TensorShape(self._a.shape)

  @property
  def dtype(self):
    return _DTYPES[self._a.dtype]

  @property
  def ndim(self):
    return self._a.ndim

  def get_shape(self):
    return self.shape

  def __len__(self):
    return self._a.shape[0]

  def __iter__(self):
    return (Tensor(x) for x in self._a)

  def __repr__(self):
    return 'tf.Tensor(%r, shape=%s, dtype=%s)' % (self._a, self._a.shape,
                                                   self.dtype.name)

  def __float__(self):
    return builtins.float(self._a)

  def __int__(self):
    return builtins.int(self._a)

  def __bool__(self):
    return builtins.bool(self._a)

  __hash__ = object.__hash__

  def __getitem__(self, idx):
    def conv(i):
      return i._a if isinstance(i, Tensor) else i
    idx = tuple(conv(i) for i in idx) if isinstance(idx, tuple) else conv(idx)
    return Tensor(np.asarray(self._a[idx]))

  # -- arithmetic --
  def _other(self, o):
    if isinstance(o, Tensor):
      if o._a.dtype != self._a.dtype:
        raise TypeError('tf_shim: dtype mismatch %s vs %s (TensorFlow would '
                        'raise InvalidArgumentError)' % (self._a.dtype, o._a.dtype))
      return o._a
    return np.asarray(o, dtype=self._a.dtype)

  def __add__(self, o): return Tensor(self._a + self._other(o))
  def __radd__(self, o): return Tensor(self._other(o) + self._a)
  def __sub__(self, o): return Tensor(self._a - self._other(o))
  def __rsub__(self, o): return Tensor(self._other(o) - self._a)
  def __mul__(self, o): return Tensor(self._a * self._other(o))
  def __rmul__(self, o): return Tensor(self._other(o) * self._a)
  def __truediv__(self, o): return Tensor(self._a / self._other(o))
  def __rtruediv__(self, o): return Tensor(self._other(o) / self._a)
  def __floordiv__(self, o): return Tensor(np.floor_divide(self._a, self._other(o)))
  def __mod__(self, o): return Tensor(np.mod(self._a, self._other(o)))
  def __pow__(self, o): return Tensor(np.power(self._a, self._other(o)))
  def __rpow__(self, o): return Tensor(np.power(self._other(o), self._a))
  def __neg__(self): return Tensor(-self._a)
  def __abs__(self): return Tensor(np.abs(self._a))
  def __eq__(self, o): return Tensor(np.asarray(self._a == self._other(o)))
  def __ne__(self, o): return Tensor(np.asarray(self._a != self._other(o)))
  def __lt__(self, o): return Tensor(np.asarray(self._a < self._other(o)))
  def __le__(self, o): return Tensor(np.asarray(self._a <= self._other(o)))
  def __gt__(self, o): return Tensor(np.asarray(self._a > self._other(o)))
  def __ge__(self, o): return Tensor(np.asarray(self._a >= self._other(o)))
  def __and__(self, o): return Tensor(np.logical_and(self._a, self._other(o)))
  def __or__(self, o): return Tensor(np.logical_or(self._a, self._other(o)))
  def __invert__(self): return Tensor(np.logical_not(self._a))


class Variable(Tensor):
  __slots__ = ()


def _arr(x, dtype=None):
  """ndarray of anything tensor-like.  Python floats default to float32 and Python
  ints to int32, as in tf.convert_to_tensor."""
  if isinstance(x, Tensor):
    a = x._a
    return a if dtype is None else a.astype(_np_dtype(dtype), copy=False)
  if dtype is not None:
    return np.asarray(x, dtype=_np_dtype(dtype))
  if isinstance(x, np.ndarray) or isinstance(x, np.generic):
    a = np.asarray(x)
    if _WIDE[0] and a.dtype == np.float32:
      a = a.astype(np.float64)        # wide mode: no float32 survives
    return a
  a = np.asarray(x)
  if a.dtype == np.float64:
    a = a.astype(_F32())
  elif a.dtype == np.int64:
    a = a.astype(np.int32)
  return a


def _like(x, *refs):
  """x as an ndarray whose dtype follows the first Tensor / ndarray among refs
  (how TF's binary ops convert Python scalars)."""
  if isinstance(x, Tensor):
    return x._a
  for r in refs:
    if isinstance(r, Tensor):
      return np.asarray(x, dtype=r._a.dtype)
  for r in refs:
    if isinstance(r, np.ndarray):
      return np.asarray(x, dtype=r.dtype)
  return _arr(x)


def _t(a):
  return Tensor(np.asarray(a))


# ----------------------------------------------------------------------------
# construction / casting
# ----------------------------------------------------------------------------
def convert_to_tensor(value, dtype=None, dtype_hint=None, name=None):
  del name, dtype_hint
  return _t(_arr(value, dtype))


constant = convert_to_tensor


def cast(x, dtype, name=None):
  del name
  a = _arr(x)
  nd = _np_dtype(dtype)
  if np.iscomplexobj(a) and not np.issubdtype(nd, np.complexfloating):
    a = a.real
  return _t(a.astype(nd, copy=False))


def identity(x, name=None):
  return _t(_arr(x))


def stop_gradient(x, name=None):
  return _t(_arr(x))


def zeros(shape, dtype=float32, name=None):
  return _t(np.zeros(_shape_arg(shape), dtype=_np_dtype(dtype)))


def ones(shape, dtype=float32, name=None):
  return _t(np.ones(_shape_arg(shape), dtype=_np_dtype(dtype)))


def zeros_like(x, dtype=None, name=None):
  return _t(np.zeros_like(_arr(x), dtype=_np_dtype(dtype)))


def ones_like(x, dtype=None, name=None):
  return _t(np.ones_like(_arr(x), dtype=_np_dtype(dtype)))


def fill(dims, value, name=None):
  return _t(np.full(_shape_arg(dims), _arr(value)))


def eye(num_rows, num_columns=None, dtype=float32, name=None):
  return _t(np.eye(num_rows, num_columns, dtype=_np_dtype(dtype)))


def _shape_arg(shape):
  if isinstance(shape, Tensor):
    shape = shape._a
  if np.ndim(shape) == 0:
    return (builtins.int(shape),)
  return tuple(builtins.int(s) for s in shape)


def range(start, limit=None, delta=1, dtype=None, name=None):  # pylint: disable=redefined-builtin
  if limit is None:
    start, limit = 0, start
  args = [builtins.float(_arr(v)) if not isinstance(v, builtins.int) else v
          for v in (start, limit, delta)]
  if dtype is None:
    dtype = int32 if all(isinstance(v, builtins.int) for v in args) else float32
  nd = _np_dtype(dtype)
  n = builtins.max(0, builtins.int(_math.ceil((args[1] - args[0]) / args[2])))
  # RangeOp: value starts at `start` and is incremented by `delta` in the output
  # dtype; for the integer-valued ranges used here this is exact.
  return _t((np.arange(n, dtype=nd) * np.asarray(args[2], nd) + np.asarray(args[0], nd)).astype(nd))


def linspace(start, stop, num, name=None, axis=0):
  """math_ops.linspace_nd: start + delta * [0..num-2], then `stop` appended."""
  del axis
  num = builtins.int(num)
  s = _arr(start).astype(_F32()) if not isinstance(start, Tensor) else start._a
  e = np.asarray(_arr(stop), dtype=s.dtype)
  if num == 1:
    return _t(np.reshape(s, (1,)))
  delta = (e - s) / np.asarray(num - 1, s.dtype)
  body = s + delta * np.arange(num - 1, dtype=s.dtype)
  return _t(np.concatenate([body, np.reshape(e, (1,))]).astype(s.dtype))


# ----------------------------------------------------------------------------
# shape ops
# ----------------------------------------------------------------------------
def shape(x, out_type=int32, name=None):
  return _t(np.asarray(_arr(x).shape, dtype=_np_dtype(out_type)))


def size(x, name=None):
  return _t(np.asarray(_arr(x).size, np.int32))


def rank(x, name=None):
  return _t(np.asarray(_arr(x).ndim, np.int32))


def reshape(x, shape, name=None):  # pylint: disable=redefined-outer-name
  return _t(np.reshape(_arr(x), _shape_arg(shape) if np.ndim(_arr(shape)) else (builtins.int(shape),)))


def transpose(x, perm=None, conjugate=False, name=None):
  a = np.transpose(_arr(x), perm)
  return _t(np.conj(a) if conjugate else a)


def squeeze(x, axis=None, name=None):
  if isinstance(axis, (list, tuple)):
    axis = tuple(axis)
  return _t(np.squeeze(_arr(x), axis=axis))


def expand_dims(x, axis, name=None):
  return _t(np.expand_dims(_arr(x), axis))


def concat(values, axis, name=None):
  arrs = [_arr(v) for v in values]
  dts = {a.dtype for a in arrs}
  if len(dts) != 1:
    raise TypeError('tf_shim.concat: mixed dtypes %s' % dts)
  return _t(np.concatenate(arrs, axis=axis))


def stack(values, axis=0, name=None):
  return _t(np.stack([_arr(v) for v in values], axis=axis))


def unstack(x, num=None, axis=0, name=None):
  a = _arr(x)
  return [_t(np.take(a, i, axis=axis)) for i in builtins.range(a.shape[axis])]


def tile(x, multiples, name=None):
  return _t(np.tile(_arr(x), _shape_arg(multiples)))


def broadcast_to(x, shape, name=None):  # pylint: disable=redefined-outer-name
  return _t(np.broadcast_to(_arr(x), _shape_arg(shape)).copy())


def pad(x, paddings, mode='CONSTANT', constant_values=0, name=None):
  a = _arr(x)
  p = [tuple(builtins.int(v) for v in row) for row in _arr(paddings).tolist()]
  mode = mode.upper()
  if mode == 'CONSTANT':
    return _t(np.pad(a, p, mode='constant',
                     constant_values=np.asarray(constant_values, a.dtype)))
  return _t(np.pad(a, p, mode={'REFLECT': 'reflect', 'SYMMETRIC': 'symmetric'}[mode]))


def slice(x, begin, size, name=None):  # pylint: disable=redefined-builtin,redefined-outer-name
  a = _arr(x)
  idx = tuple(builtins.slice(b, None if s == -1 else b + s)
              for b, s in zip(_arr(begin).tolist(), _arr(size).tolist()))
  return _t(a[idx])


def gather(params, indices, axis=0, batch_dims=0, name=None):
  a, i = _arr(params), _arr(indices)
  if batch_dims:
    return _t(np.take_along_axis(a, i, axis=axis))
  return _t(np.take(a, i, axis=axis))


def where(condition, x=None, y=None, name=None):
  c = _arr(condition)
  if x is None and y is None:
    return _t(np.argwhere(c).astype(np.int64))
  xa, ya = _like(x, y), _like(y, x)
  hard = lambda v: isinstance(v, (Tensor, np.ndarray, np.generic))
  if xa.dtype != ya.dtype and not (hard(x) and hard(y)):
    # Python scalars follow the other operand (dtype hint of convert_to_tensor);
    # two Python scalars: a float wins over an int.
    if hard(x) or (not hard(y) and xa.dtype.kind == 'f'):
      ya = ya.astype(xa.dtype)
    else:
      xa = xa.astype(ya.dtype)
  if xa.dtype != ya.dtype:
    raise TypeError('tf_shim.where: dtype mismatch %s vs %s' % (xa.dtype, ya.dtype))
  return _t(np.where(c, xa, ya))


def meshgrid(*args, indexing='xy', name=None):
  return [_t(m) for m in np.meshgrid(*[_arr(a) for a in args], indexing=indexing)]


def sort(values, axis=-1, direction='ASCENDING', name=None):
  a = np.sort(_arr(values), axis=axis)
  return _t(a if direction == 'ASCENDING' else np.flip(a, axis))


def argsort(values, axis=-1, direction='ASCENDING', stable=False, name=None):
  a = _arr(values)
  i = np.argsort(a if direction == 'ASCENDING' else -a, axis=axis, kind='stable')
  return _t(i.astype(np.int32))


def searchsorted(sorted_sequence, values, side='left', out_type=int32, name=None):
  s, v = _arr(sorted_sequence), _arr(values)
  out = np.empty(v.shape, _np_dtype(out_type))
  for idx in np.ndindex(s.shape[:-1]):
    out[idx] = np.searchsorted(s[idx], v[idx], side=side)
  return _t(out)


# ----------------------------------------------------------------------------
# elementwise math (float32 in, float32 out)
# ----------------------------------------------------------------------------
def _unary(fn):
  def op(x, name=None):
    del name
    return _t(fn(_arr(x)))
  return op


def _binary(fn):
  def op(x, y, name=None):
    del name
    return _t(fn(_like(x, y), _like(y, x)))
  return op


sin = _unary(np.sin)
cos = _unary(np.cos)
tan = _unary(np.tan)
exp = _unary(np.exp)
sqrt = _unary(np.sqrt)
square = _unary(np.square)
floor = _unary(np.floor)
sign = _unary(np.sign)
tanh = _unary(np.tanh)
negative = _unary(np.negative)
add = _binary(np.add)
subtract = _binary(np.subtract)
multiply = _binary(np.multiply)
divide = _binary(np.divide)
maximum = _binary(np.maximum)
minimum = _binary(np.minimum)
equal = _binary(np.equal)
not_equal = _binary(np.not_equal)
less = _binary(np.less)
less_equal = _binary(np.less_equal)
greater = _binary(np.greater)
greater_equal = _binary(np.greater_equal)
logical_and = _binary(np.logical_and)
logical_or = _binary(np.logical_or)
logical_not = _unary(np.logical_not)
pow = _binary(np.power)  # pylint: disable=redefined-builtin


def abs(x, name=None):  # pylint: disable=redefined-builtin
  return _t(np.abs(_arr(x)))       # complex64 -> float32 magnitude, as tf.abs


def round(x, name=None):  # pylint: disable=redefined-builtin
  return _t(np.rint(_arr(x)))       # round half to even, as tf.round


def clip_by_value(x, clip_value_min, clip_value_max, name=None):
  a = _arr(x)
  return _t(np.clip(a, np.asarray(_arr(clip_value_min), a.dtype),
                    np.asarray(_arr(clip_value_max), a.dtype)))


def complex(real, imag, name=None):  # pylint: disable=redefined-builtin
  r, i = _arr(real), _arr(imag)
  out = np.empty(np.broadcast(r, i).shape,
                 np.complex64 if r.dtype == np.float32 else np.complex128)
  out.real, out.imag = r, i
  return _t(out)


def _reduce(fn):
  def op(x, axis=None, keepdims=False, name=None):
    del name
    if isinstance(axis, list):
      axis = tuple(axis)
    a = _arr(x)
    return _t(np.asarray(fn(a, axis=axis, keepdims=keepdims)).astype(a.dtype, copy=False))
  return op


reduce_sum = _reduce(np.sum)
reduce_mean = _reduce(np.mean)
reduce_max = _reduce(np.max)
reduce_min = _reduce(np.min)
reduce_prod = _reduce(np.prod)


def reduce_any(x, axis=None, keepdims=False, name=None):
  return _t(np.any(_arr(x), axis=axis, keepdims=keepdims))


def reduce_all(x, axis=None, keepdims=False, name=None):
  return _t(np.all(_arr(x), axis=axis, keepdims=keepdims))


def argmax(x, axis=None, output_type=int64, name=None):
  return _t(np.argmax(_arr(x), axis=axis).astype(_np_dtype(output_type)))


def cumsum(x, axis=0, exclusive=False, reverse=False, name=None):
  """Sequential running sum in the input dtype (what Eigen's scan does on CPU;
  np.add.accumulate is sequential too)."""
  a = _arr(x)
  if reverse:
    a = np.flip(a, axis)
  out = np.add.accumulate(a, axis=axis, dtype=a.dtype)
  if exclusive:
    out = np.concatenate([np.zeros_like(np.take(out, [0], axis=axis)),
                          np.delete(out, -1, axis=axis)], axis=axis)
  if reverse:
    out = np.flip(out, axis)
  return _t(out)


def tensordot(a, b, axes, name=None):
  return _t(np.tensordot(_arr(a), _arr(b), axes))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
  a, b = _arr(a), _arr(b)
  if transpose_a:
    a = np.swapaxes(a, -1, -2)
  if transpose_b:
    b = np.swapaxes(b, -1, -2)
  return _t(np.matmul(a, b))


def executing_eagerly():
  return True


def function(func=None, **kwargs):
  if func is None:
    return lambda f: f
  return func


def name_scope(name):
  import contextlib
  return contextlib.nullcontext()


def random_normal_initializer(mean=0.0, stddev=0.05, seed=None):
  rng = np.random.default_rng(seed)
  return lambda shape, dtype=float32: _t(
      rng.normal(mean, stddev, _shape_arg(shape)).astype(_np_dtype(dtype)))


def constant_initializer(value=0):
  return lambda shape, dtype=float32: _t(
      np.full(_shape_arg(shape), value, dtype=_np_dtype(dtype)))


def _ns(name, **members):
  m = _types.ModuleType(__name__ + '.' + name)
  m.__dict__.update(members)
  import sys
  sys.modules[m.__name__] = m
  return m


# -- tf.math ------------------------------------------------------------------
def _log(x, name=None):
  return _t(np.log(_arr(x)))


def _floormod(x, y, name=None):
  return _t(np.mod(_like(x, y), _like(y, x)))


math = _ns(
    'math', log=_log, exp=exp, sin=sin, cos=cos, tan=tan, sqrt=sqrt, abs=abs,
    real=_unary(np.real), imag=_unary(np.imag), conj=_unary(np.conj),
    angle=_unary(np.angle), is_nan=_unary(np.isnan), is_inf=_unary(np.isinf),
    is_finite=_unary(np.isfinite), cumsum=cumsum, reduce_sum=reduce_sum,
    reduce_mean=reduce_mean, reduce_max=reduce_max, reduce_min=reduce_min,
    argmax=argmax, maximum=maximum, minimum=minimum, pow=pow, square=square,
    floor=floor, ceil=_unary(np.ceil), round=round, sign=sign, tanh=tanh,
    floormod=_floormod, mod=_floormod, add=add, subtract=subtract,
    multiply=multiply, divide=divide, equal=equal, less=less, greater=greater,
    less_equal=less_equal, greater_equal=greater_equal,
    log1p=_unary(np.log1p), expm1=_unary(np.expm1),
    sigmoid=lambda x, name=None: nn.sigmoid(x),
    softplus=lambda x, name=None: nn.softplus(x))
real = math.real
imag = math.imag


# -- tf.nn --------------------------------------------------------------------
def _sigmoid(x, name=None):
  a = _arr(x)
  one = np.asarray(1, a.dtype)
  return _t(one / (one + np.exp(-a)))


def _softplus(x, name=None):
  a = _arr(x)
  return _t(np.logaddexp(a, np.asarray(0, a.dtype)))


def _softmax(x, axis=-1, name=None):
  a = _arr(x)
  e = np.exp(a - np.max(a, axis=axis, keepdims=True))
  return _t(e / np.sum(e, axis=axis, keepdims=True))


def _moments(x, axes, keepdims=False, name=None):
  a = _arr(x)
  axes = tuple(axes) if isinstance(axes, (list, tuple)) else axes
  m = np.mean(a, axis=axes, keepdims=True)
  v = np.mean(np.square(a - m), axis=axes, keepdims=keepdims)
  return _t(m if keepdims else np.squeeze(m, axes)), _t(v)


nn = _ns('nn', sigmoid=_sigmoid, softplus=_softplus, softmax=_softmax,
         tanh=tanh, relu=lambda x, name=None: _t(np.maximum(_arr(x), 0)),
         moments=_moments)
sigmoid = _sigmoid


# -- tf.random ----------------------------------------------------------------
_RANDOM = {'rng': np.random.default_rng(0), 'uniform_queue': []}


def _random_uniform(shape, minval=0, maxval=None, dtype=float32, seed=None,  # pylint: disable=redefined-outer-name
                    name=None):
  """tf.random.uniform.  Tests that need the reference to see specific noise push
  arrays with `tf.random.inject_uniform(...)`; they are handed out first (shape
  checked), otherwise a NumPy generator is used."""
  shp = _shape_arg(shape)
  nd = _np_dtype(dtype)
  if _RANDOM['uniform_queue']:
    a = _RANDOM['uniform_queue'].pop(0)
    if tuple(a.shape) != shp:
      raise ValueError('tf_shim: injected uniform noise has shape %s, asked %s'
                       % (a.shape, shp))
    return _t(a.astype(nd, copy=False))
  maxval = 1 if maxval is None else maxval
  lo, hi = np.asarray(minval, nd), np.asarray(maxval, nd)
  u = _RANDOM['rng'].random(shp, dtype=np.float32 if nd == np.float32 else np.float64)
  return _t((u * (hi - lo) + lo).astype(nd))


def _random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):  # pylint: disable=redefined-outer-name
  rng = _RANDOM['rng'] if seed is None else np.random.default_rng(seed)
  return _t(rng.normal(mean, stddev, _shape_arg(shape)).astype(_np_dtype(dtype)))


def _set_seed(seed):
  _RANDOM['rng'] = np.random.default_rng(seed)


random = _ns('random', uniform=_random_uniform, normal=_random_normal,
             set_seed=_set_seed,
             inject_uniform=lambda a: _RANDOM['uniform_queue'].append(np.asarray(a)))


# -- tf.signal ----------------------------------------------------------------
def _raised_cosine_window(window_length, periodic, dtype, a, b):
  """window_ops._raised_cosine_window: everything in `dtype` (float32)."""
  nd = _np_dtype(dtype)
  window_length = builtins.int(_arr(window_length))
  if window_length == 1:
    return _t(np.ones([1], nd))
  even = 1 - window_length % 2
  n = np.asarray(window_length + builtins.int(periodic) * even - 1, nd)
  count = np.arange(window_length, dtype=nd)
  cos_arg = np.asarray(2 * np.pi, nd) * count / n
  return _t((np.asarray(a, nd) - np.asarray(b, nd) * np.cos(cos_arg)).astype(nd))


def _hann_window(window_length, periodic=True, dtype=float32, name=None):
  return _raised_cosine_window(window_length, periodic, dtype, 0.5, 0.5)


def _hamming_window(window_length, periodic=True, dtype=float32, name=None):
  return _raised_cosine_window(window_length, periodic, dtype, 0.54, 0.46)


def _frame(signal, frame_length, frame_step, pad_end=False, pad_value=0, axis=-1,
           name=None):
  """shape_ops.frame: frames of `frame_length` every `frame_step` along `axis`;
  pad_end -> ceil(N / step) frames, zero (pad_value) padded."""
  a = _arr(signal)
  axis = axis % a.ndim
  a = np.moveaxis(a, axis, -1)
  n = a.shape[-1]
  frame_length, frame_step = builtins.int(frame_length), builtins.int(frame_step)
  if pad_end:
    n_frames = -(-n // frame_step)
    need = (n_frames - 1) * frame_step + frame_length
    if need > n:
      pw = [(0, 0)] * (a.ndim - 1) + [(0, need - n)]
      a = np.pad(a, pw, mode='constant', constant_values=np.asarray(pad_value, a.dtype))
  else:
    n_frames = builtins.max(0, 1 + (n - frame_length) // frame_step)
  idx = (np.arange(n_frames)[:, None] * frame_step + np.arange(frame_length)[None, :])
  out = a[..., idx]                               # [..., frames, frame_length]
  out = np.moveaxis(out, (-2, -1), (axis, axis + 1)) if axis != a.ndim - 1 else out
  return _t(np.ascontiguousarray(out))


def _overlap_and_add(signal, frame_step, name=None):
  """reconstruction_ops.overlap_and_add: [..., frames, frame_length] ->
  [..., (frames - 1) * step + frame_length].  Frames are accumulated in frame
  order (TF sums the up-to-ceil(length/step) overlapping segments with one
  reduce_sum; for the 2-term sums of the Hann upsampler the order is immaterial,
  for the 4-term sums of fft_convolve it can differ in the last ulp)."""
  a = _arr(signal)
  frame_step = builtins.int(_arr(frame_step))
  frames, length = a.shape[-2], a.shape[-1]
  out = np.zeros(a.shape[:-2] + ((frames - 1) * frame_step + length,), a.dtype)
  if length % frame_step == 0:
    # vectorised: segment s of every frame lands at offset (frame + s) * step
    segs = length // frame_step
    a5 = a.reshape(a.shape[:-1] + (segs, frame_step))
    ov = out.reshape(out.shape[:-1] + (frames - 1 + segs, frame_step))
    for s in builtins.range(segs):
      ov[..., s:s + frames, :] += a5[..., s, :]
    return _t(out)
  for f in builtins.range(frames):
    out[..., f * frame_step:f * frame_step + length] += a[..., f, :]
  return _t(out)


def _fft_len(fft_length):
  if fft_length is None:
    return None
  v = _arr(fft_length)
  return builtins.int(v.reshape(-1)[0])


def _rfft(x, fft_length=None, name=None):
  """tf.signal.rfft: float32 -> complex64 (single-precision transform), input
  zero-padded or cropped to fft_length."""
  a = _arr(x)
  n = _fft_len(fft_length)
  return _t(_sfft.rfft(a, n=n, axis=-1).astype(
      np.complex64 if a.dtype == np.float32 else np.complex128, copy=False))


def _irfft(x, fft_length=None, name=None):
  a = _arr(x)
  n = _fft_len(fft_length)
  if n is None:
    n = 2 * (a.shape[-1] - 1)
  return _t(_sfft.irfft(a, n=n, axis=-1).astype(
      np.float32 if a.dtype == np.complex64 else np.float64, copy=False))


def _fftshift(x, axes=None, name=None):
  return _t(np.fft.fftshift(_arr(x), axes=axes))


def _enclosing_power_of_two(v):
  return builtins.int(2 ** _math.ceil(_math.log(v) / _math.log(2.0)))


def _stft(signals, frame_length, frame_step, fft_length=None,
          window_fn=_hann_window, pad_end=False, name=None):
  """spectral_ops.stft: frame (pad_end), periodic window in the signal dtype,
  rfft of fft_length (default: enclosing power of two)."""
  a = _arr(signals)
  if fft_length is None:
    fft_length = _enclosing_power_of_two(frame_length)
  framed = _frame(a, frame_length, frame_step, pad_end=pad_end)._a
  if window_fn is not None:
    framed = framed * window_fn(frame_length, dtype=_DTYPES[a.dtype])._a
  return _rfft(framed, [fft_length])


def _unavailable(name):
  def fn(*args, **kwargs):
    raise NotImplementedError('tf_shim: %s is outside the decoder path' % name)
  return fn


signal = _ns('signal', hann_window=_hann_window, hamming_window=_hamming_window,
             frame=_frame, overlap_and_add=_overlap_and_add, rfft=_rfft,
             irfft=_irfft, fftshift=_fftshift, stft=_stft,
             linear_to_mel_weight_matrix=_unavailable('linear_to_mel_weight_matrix'),
             mfccs_from_log_mel_spectrograms=_unavailable('mfccs_from_log_mel_spectrograms'))


# -- tf.nest ------------------------------------------------------------------
def _map_structure(fn, *structs):
  s0 = structs[0]
  if isinstance(s0, dict):
    return {k: _map_structure(fn, *[s[k] for s in structs]) for k in s0}
  if isinstance(s0, (list, tuple)):
    return type(s0)(_map_structure(fn, *xs) for xs in zip(*structs))
  return fn(*structs)


nest = _ns('nest', map_structure=_map_structure)


# -- tf.Module / tf.keras ------------------------------------------------------
class Module:
  def __init__(self, name=None):
    self._name = name or type(self).__name__.lower()

  @property
  def name(self):
    return self._name


class _Layer(Module):
  """tf.keras.layers.Layer as far as ddsp's Processor / DAGLayer / Loss use it:
  a named Module whose __call__ forwards to call()."""

  def __init__(self, name=None, trainable=True, dtype=None, autocast=True, **kwargs):
    super().__init__(name=name)
    self.trainable = trainable
    self.built = False

  def build(self, input_shape):
    self.built = True

  def call(self, *args, **kwargs):
    raise NotImplementedError

  def __call__(self, *args, **kwargs):
    if not self.built:
      self.build(None)
      self.built = True
    return self.call(*args, **kwargs)

  def add_weight(self, name=None, shape=None, dtype=float32, initializer=None,
                 trainable=True, **kwargs):
    if initializer is None:
      initializer = constant_initializer(0)
    return initializer(shape, dtype=dtype or float32)


class _KerasModel(_Layer):
  pass


class _LazyLayers(_types.ModuleType):
  Layer = _Layer

  def __getattr__(self, item):
    return _unavailable('tf.keras.layers.' + item)


_layers = _LazyLayers(__name__ + '.keras.layers')
keras = _ns('keras', layers=_layers, Model=_KerasModel,
            models=_ns('keras.models', load_model=_unavailable('load_model')))
losses = _ns('losses', cosine_distance=_unavailable('tf.losses.cosine_distance'))


# -- tf.test ------------------------------------------------------------------
import unittest as _unittest


class _TestCase(_unittest.TestCase):
  """tf.test.TestCase assertions used by the reference's tests."""

  @staticmethod
  def _np(x):
    if isinstance(x, Tensor):
      return x._a
    if isinstance(x, (list, tuple)):
      return np.asarray([_TestCase._np(v) for v in x])
    return np.asarray(x)

  def assertAllClose(self, a, b, rtol=1e-6, atol=1e-6, msg=None):
    if isinstance(a, dict):
      self.assertEqual(sorted(a), sorted(b))
      for k in a:
        self.assertAllClose(a[k], b[k], rtol, atol, msg)
      return
    np.testing.assert_allclose(self._np(a).astype(np.float64), self._np(b).astype(np.float64),
                               rtol=rtol, atol=atol, err_msg=msg or '')

  def assertAllEqual(self, a, b, msg=None):
    np.testing.assert_array_equal(self._np(a), self._np(b), err_msg=msg or '')

  def assertShapeEqual(self, np_array, tf_tensor, msg=None):
    self.assertEqual(tuple(np_array.shape), tuple(tf_tensor.shape))

  def evaluate(self, x):
    return _map_structure(lambda t: t.numpy() if isinstance(t, Tensor) else t, x)

  def get_temp_dir(self):
    import tempfile
    return tempfile.mkdtemp()


def _test_main(argv=None):
  _unittest.main(argv=argv)


test = _ns('test', TestCase=_TestCase, main=_test_main)


# -- tf.image / tf.compat.v1.image --------------------------------------------
class _ResizeMethod:
  BILINEAR = 0
  NEAREST_NEIGHBOR = 1
  BICUBIC = 2
  AREA = 3


def _resize_scale(in_size, out_size, align_corners):
  """image_resizer_state.h CalculateResizeScale: a float32."""
  if align_corners and out_size > 1:
    return _F32().type(in_size - 1) / _F32().type(out_size - 1)
  return _F32().type(in_size) / _F32().type(out_size)


_K_TABLE = 1 << 10


def _cubic_table():
  """resize_bicubic_op.cc InitCoeffsTable(a = -0.75), the legacy (non
  half-pixel) Keys kernel sampled at 1025 offsets, float32 entries."""
  a = -0.75
  # float32 in BOTH modes: the table is a constant of TensorFlow's kernel (`static
  # const float*`), not one of the reference's float32 tensors - and it is built once
  # per process, so its precision must not depend on which mode asked for it first
  tab = np.empty(((_K_TABLE + 1) * 2,), np.float32)
  for i in builtins.range(_K_TABLE + 1):
    x = i * 1.0 / _K_TABLE
    tab[2 * i] = ((a + 2) * x - (a + 3)) * x * x + 1
    x += 1.0
    tab[2 * i + 1] = ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
  return tab


_CUBIC = []


def _resize_axis(a, out_size, method, align_corners, axis):
  """One spatial axis of the legacy resize kernels (resize_bilinear_op.cc,
  resize_nearest_neighbor_op.cc, resize_bicubic_op.cc with half_pixel_centers =
  false, which is what tf.compat.v1.image.resize runs)."""
  in_size = a.shape[axis]
  if in_size == out_size and method != _ResizeMethod.BICUBIC:
    return a.astype(_F32().type) if method != _ResizeMethod.NEAREST_NEIGHBOR else a
  a = np.moveaxis(a, axis, 0)
  scale = _resize_scale(in_size, out_size, align_corners)
  src = (np.arange(out_size, dtype=_F32().type) * scale).astype(_F32().type)
  if method == _ResizeMethod.NEAREST_NEIGHBOR:
    # roundf (half away from zero), not rint: resize_nearest_neighbor_op.cc
    idx = np.minimum((np.floor(src + _F32().type(0.5)) if align_corners
                      else np.floor(src)).astype(np.int64), in_size - 1)
    out = a[idx]
  elif method == _ResizeMethod.BILINEAR:
    lo_f = np.floor(src)
    lo = lo_f.astype(np.int64)
    hi = np.minimum(np.ceil(src).astype(np.int64), in_size - 1)
    lerp = (src - lo_f).astype(_F32().type).reshape((-1,) + (1,) * (a.ndim - 1))
    top = a[lo].astype(_F32().type)
    bottom = a[hi].astype(_F32().type)
    out = (top + (bottom - top) * lerp).astype(_F32().type)
  elif method == _ResizeMethod.BICUBIC:
    if not _CUBIC:
      _CUBIC.append(_cubic_table())
    tab = _CUBIC[0]
    loc = np.floor(src).astype(np.int64)
    delta = (src - loc.astype(_F32().type)).astype(_F32().type)
    off = np.rint(delta * _F32().type(_K_TABLE)).astype(np.int64)   # lrintf
    w = [tab[off * 2 + 1], tab[off * 2], tab[(_K_TABLE - off) * 2],
         tab[(_K_TABLE - off) * 2 + 1]]
    bound = lambda v: np.clip(v, 0, in_size - 1)
    ix = [bound(loc - 1), bound(loc), bound(loc + 1), bound(loc + 2)]
    shp = (-1,) + (1,) * (a.ndim - 1)
    af = a.astype(_F32().type)
    # Interpolate(): ((w0 v0 + w1 v1) + w2 v2) + w3 v3 in float32
    out = (af[ix[0]] * w[0].reshape(shp) + af[ix[1]] * w[1].reshape(shp)
           + af[ix[2]] * w[2].reshape(shp) + af[ix[3]] * w[3].reshape(shp)).astype(_F32().type)
  else:
    raise NotImplementedError('tf_shim: resize method %r' % (method,))
  return np.moveaxis(out, 0, axis)


def _v1_resize(images, size, method=_ResizeMethod.BILINEAR, align_corners=False,
               preserve_aspect_ratio=False, name=None):
  """tf.compat.v1.image.resize on [batch, height, width, channels]."""
  a = _arr(images)
  if a.ndim != 4:
    raise ValueError("'images' must have either 3 or 4 dimensions.")
  h, w = [builtins.int(_arr(s)) for s in size]
  out = _resize_axis(a, h, method, align_corners, 1)
  out = _resize_axis(out, w, method, align_corners, 2)
  if method != _ResizeMethod.NEAREST_NEIGHBOR:
    out = out.astype(_F32().type, copy=False)
  return _t(out)


image = _ns('image', ResizeMethod=_ResizeMethod,
            resize=_unavailable('tf.image.resize (v2, half-pixel centres)'))


from tensorflow import compat  # noqa: E402,F401  (tf.compat.v1.image.resize, tf.compat.v2)
