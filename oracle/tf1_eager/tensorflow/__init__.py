# coding=utf-8
"""An EAGER stand-in for the ~90 TensorFlow-1.15 symbols that the reference's
``code/pred_models.py`` touches, backed by torch CPU tensors (fp64 by default).

THIS IS TEST INFRASTRUCTURE (part of the oracle), NOT PRODUCT CODE.  Its one purpose: let the
UNMODIFIED reference file ``/root/reference/code/pred_models.py`` be imported and its
``Model.__init__ -> build_forward / build_loss`` and ``Trainer.__init__`` be EXECUTED line by line
in this container, where TensorFlow 1.15 cannot be installed (Python 3.12, no wheel, no network).
The reference's own wiring (scopes, raw_rnn loop functions, beam bookkeeping, back-trace, gathers,
reshapes, loss) then runs as written; only the per-op semantics below are restated from the TF
1.15 sources (SURVEY.md section 8c sheet; each op cites what it stands for).

Execution model: eager.  ``tf.placeholder`` returns the value registered for it in the active
``feeding(...)`` context (a probe pass discovers which Model attribute each placeholder is, see
``oracle/tf1_eager/run_reference.py``); ``tf.cond`` evaluates the taken branch; ``tf.while_loop``,
``tf.nn.raw_rnn`` and ``tf.nn.dynamic_rnn`` are Python loops that follow TF's control-flow
protocol (python/ops/rnn.py); ``tf.gradients`` is torch autograd over the same eager values.
Every ``float``/``float32`` tensor is held in ``COMPUTE_DTYPE`` (float64 = the parity truth).
"""
from __future__ import annotations

import collections
import contextlib
import math
import re as _re

import numpy as np
import torch
import torch.nn.functional as F

COMPUTE_DTYPE = torch.float64
__version__ = "1.15.0-eager-standin"


# --------------------------------------------------------------------------------------------- #
# dtypes, shapes, tensors
# --------------------------------------------------------------------------------------------- #
class DType(object):
  def __init__(self, name):
    self.name = name

  @property
  def is_floating(self):
    return self.name.startswith("float")

  def __eq__(self, other):
    return self.name == (other.name if isinstance(other, DType) else _dtype(other).name)

  def __ne__(self, other):
    return not self == other

  def __hash__(self):
    return hash(self.name)

  def __repr__(self):
    return "tf." + self.name


float32 = DType("float32")
float64 = DType("float64")
int32 = DType("int32")
int64 = DType("int64")
bool_ = DType("bool")
_ALIASES = {"float": float32, "float32": float32, "float64": float64, "int32": int32,
            "int": int64, "int64": int64, "bool": bool_}


def _dtype(d):
  if isinstance(d, DType):
    return d
  if isinstance(d, str):
    return _ALIASES[d]
  if d is None:
    return float32
  raise TypeError(d)


def _torch_dtype(d):
  d = _dtype(d)
  return {"float32": COMPUTE_DTYPE, "float64": torch.float64, "int32": torch.int32,
          "int64": torch.int64, "bool": torch.bool}[d.name]


def _tf_dtype(t):
  if t.dtype.is_floating_point:
    return float32
  return {torch.int32: int32, torch.int64: int64, torch.bool: bool_}[t.dtype]


class Dimension(object):
  def __init__(self, value):
    self.value = None if value is None else int(value)

  def __int__(self):
    return self.value

  __index__ = __int__

  def __eq__(self, other):
    return self.value == (other.value if isinstance(other, Dimension) else other)

  def __hash__(self):
    return hash(self.value)

  def __mul__(self, other):
    return Dimension(self.value * int(other))

  __rmul__ = __mul__

  def __repr__(self):
    return "Dimension(%r)" % self.value


class TensorShape(object):
  def __init__(self, dims):
    self.dims = [Dimension(d) for d in dims]

  def as_list(self):
    return [d.value for d in self.dims]

  @property
  def ndims(self):
    return len(self.dims)

  def __len__(self):
    return len(self.dims)

  def __iter__(self):
    return iter(self.dims)

  def __getitem__(self, i):
    return self.dims[i]

  def __repr__(self):
    return "TensorShape(%r)" % self.as_list()


class _Op(object):
  def __init__(self, name):
    self.name = name


def _raw(x, like=None):
  """torch tensor of anything tensor-like."""
  if isinstance(x, Tensor):
    return x.t
  if isinstance(x, torch.Tensor):
    return x
  if isinstance(x, Dimension):
    x = x.value
  if isinstance(x, (bool, np.bool_)):
    return torch.tensor(bool(x))
  if isinstance(x, (int, np.integer)):
    if like is not None and like.dtype.is_floating_point:
      return torch.tensor(float(x), dtype=like.dtype)
    return torch.tensor(int(x), dtype=like.dtype if like is not None and like.dtype != torch.bool
                        else torch.int32)
  if isinstance(x, (float, np.floating)):
    return torch.tensor(float(x), dtype=COMPUTE_DTYPE)
  if isinstance(x, np.ndarray):
    t = torch.from_numpy(np.ascontiguousarray(x))
    return t.to(COMPUTE_DTYPE) if t.dtype.is_floating_point else t
  if isinstance(x, (list, tuple)):
    if any(isinstance(e, (Tensor, torch.Tensor)) for e in x):
      return torch.stack([_raw(e) for e in x])
    return _raw(np.asarray(x))
  raise TypeError(type(x))


def _int(x):
  """Python int of a static dimension / scalar int tensor (graph-time 'shape' values)."""
  if isinstance(x, Tensor):
    return int(x.t.item())
  if isinstance(x, torch.Tensor):
    return int(x.item())
  if isinstance(x, Dimension):
    return x.value
  return int(x)


def _shape(shape):
  if isinstance(shape, Tensor):
    return [int(v) for v in shape.t.tolist()]
  return [_int(s) for s in shape]


class Tensor(object):
  """Eager tensor with the slice of tf.Tensor's interface the reference uses."""
  __array_priority__ = 100

  def __init__(self, t, name=None):
    self.t = t
    self.name = (name or "Tensor") + ":0"
    self.op = _Op(name or "Tensor")

  # -- shape / dtype ---------------------------------------------------------------------- #
  def get_shape(self):
    return TensorShape(list(self.t.shape))

  @property
  def shape(self):
    return self.get_shape()

  @property
  def dtype(self):
    return _tf_dtype(self.t)

  def numpy(self):
    return self.t.detach().numpy()

  def __repr__(self):
    return "<tf1_eager.Tensor %s %s %s>" % (self.name, list(self.t.shape), self.dtype)

  # -- python protocol -------------------------------------------------------------------- #
  def __bool__(self):
    raise TypeError("a tf.Tensor is not a Python bool (graph-mode semantics kept on purpose)")

  def __len__(self):
    return self.t.shape[0]

  def __iter__(self):
    raise TypeError("tf.Tensor is not iterable in graph mode")

  def __getitem__(self, idx):
    if not isinstance(idx, tuple):
      idx = (idx,)
    idx = tuple(_int(i) if isinstance(i, (Tensor, Dimension)) else i for i in idx)
    return Tensor(self.t[idx])

  def _bin(self, other, fn, reverse=False):
    o = _raw(other, like=self.t)
    a, b = (o, self.t) if reverse else (self.t, o)
    if a.dtype != b.dtype and a.dtype.is_floating_point != b.dtype.is_floating_point:
      # python scalars adopt the tensor's dtype; tensors must agree like in TF
      if not isinstance(other, (int, float, np.integer, np.floating)):
        raise TypeError("dtype mismatch %s vs %s" % (a.dtype, b.dtype))
      o = o.to(self.t.dtype)
      a, b = (o, self.t) if reverse else (self.t, o)
    return Tensor(fn(a, b))

  def __add__(self, o): return self._bin(o, torch.add)
  def __radd__(self, o): return self._bin(o, torch.add, True)
  def __sub__(self, o): return self._bin(o, torch.sub)
  def __rsub__(self, o): return self._bin(o, torch.sub, True)
  def __mul__(self, o): return self._bin(o, torch.mul)
  def __rmul__(self, o): return self._bin(o, torch.mul, True)
  def __truediv__(self, o): return self._bin(o, torch.true_divide)
  def __rtruediv__(self, o): return self._bin(o, torch.true_divide, True)
  def __floordiv__(self, o): return self._bin(o, lambda a, b: torch.div(a, b, rounding_mode="floor"))
  def __mod__(self, o): return self._bin(o, torch.remainder)
  def __neg__(self): return Tensor(-self.t)
  def __pow__(self, o): return self._bin(o, torch.pow)
  def __ge__(self, o): return self._bin(o, torch.ge)
  def __gt__(self, o): return self._bin(o, torch.gt)
  def __le__(self, o): return self._bin(o, torch.le)
  def __lt__(self, o): return self._bin(o, torch.lt)
  __hash__ = object.__hash__

  def eval(self, session=None, feed_dict=None):
    return self.numpy()


def _wrap(t):
  return t if isinstance(t, Tensor) else Tensor(_raw(t))


# --------------------------------------------------------------------------------------------- #
# graph bookkeeping: variable scopes, variables, placeholders
# --------------------------------------------------------------------------------------------- #
AUTO_REUSE = "AUTO_REUSE"


class VariableScope(object):
  def __init__(self, name):
    self.name = name


class _State(object):
  def __init__(self):
    self.reset()

  def reset(self):
    self.scope = [""]            # stack of absolute variable-scope names
    self.variables = collections.OrderedDict()
    self.var_values = {}         # name -> numpy array supplied by the harness
    self.placeholders = []       # creation order
    self.feed = None             # callable(index, name, dtype, shape) -> array, or None = probe
    self.requires_grad = False
    self.rng = np.random.default_rng(0)
    # SimAug's random draws are injected by the harness (TF's streams cannot be reproduced):
    self.beta_sample = None      # value of tf.distributions.Beta(...).sample()
    self.uniform_hook = None     # callable(shape, minval, maxval, dtype) -> array for tf.random_uniform / tf.random.uniform


_S = _State()


class _StopBuild(Exception):
  """Raised by the first op after the placeholders in probe mode."""


def reset_default_graph():
  _S.reset()


@contextlib.contextmanager
def variable_scope(name_or_scope, reuse=None, default_name=None, **kw):
  """tf.variable_scope: a string nests under the current scope, a captured VariableScope object
  re-enters its absolute name (python/ops/variable_scope.py)."""
  if isinstance(name_or_scope, VariableScope):
    full = name_or_scope.name
  else:
    name = name_or_scope if name_or_scope is not None else default_name
    full = (_S.scope[-1] + "/" + name) if _S.scope[-1] else name
  _S.scope.append(full)
  try:
    yield VariableScope(full)
  finally:
    _S.scope.pop()


@contextlib.contextmanager
def name_scope(name, *a, **kw):
  yield name


@contextlib.contextmanager
def device(name):
  yield


def get_variable_scope():
  return VariableScope(_S.scope[-1])


class Variable(Tensor):
  def __init__(self, t, name, trainable):
    Tensor.__init__(self, t, name)
    self.trainable = trainable

  def assign(self, value):
    with torch.no_grad():
      self.t.copy_(_raw(value))
    return self


def constant_initializer(value=0.0, dtype=None):
  return lambda shape, dt: np.full(shape, value, dtype=np.float64)


def variance_scaling_initializer(scale=1.0, mode="fan_in", distribution="truncated_normal", **kw):
  """python/ops/init_ops.py VarianceScaling, fan_in / truncated normal (only used when the harness
  supplies no value, i.e. never in the parity runs)."""
  def init(shape, dt):
    fan_in = int(np.prod(shape[:-1])) if len(shape) > 1 else shape[0]
    std = math.sqrt(scale / max(1.0, fan_in)) / .87962566103423978
    return np.clip(_S.rng.standard_normal(shape), -2, 2) * std
  return init


def glorot_uniform_initializer():
  def init(shape, dt):
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    lim = math.sqrt(6.0 / (rf * (shape[-2] + shape[-1])))
    return _S.rng.uniform(-lim, lim, size=shape)
  return init


def truncated_normal(shape, mean=0.0, stddev=1.0, **kw):
  return Tensor(_raw(np.clip(_S.rng.standard_normal(_shape(shape)), -2, 2) * stddev + mean))


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
  """tf.get_variable under AUTO_REUSE semantics: full name = current variable scope + name.  The
  value comes from the harness' TF-named weight dict; a name the harness does not know is an
  error unless an initializer can stand in (global_step, or an explicit allow_init run)."""
  full = (_S.scope[-1] + "/" + name) if _S.scope[-1] else name
  if full in _S.variables:
    return _S.variables[full]
  shape = _shape(shape) if shape is not None else None
  td = _torch_dtype(dtype)
  if full in _S.var_values:
    val = np.asarray(_S.var_values[full])
    assert list(val.shape) == list(shape), (full, val.shape, shape)
  elif full == "global_step" or _S.var_values.get("__allow_init__", False):
    init = initializer or glorot_uniform_initializer()   # TF default for float variables
    val = np.asarray(init(shape, dtype))
  else:
    raise KeyError("reference created variable %r %s that the weight dict does not hold"
                   % (full, shape))
  t = torch.from_numpy(np.ascontiguousarray(val)).to(td).clone()
  if trainable and _S.requires_grad and t.dtype.is_floating_point:
    t.requires_grad_(True)
  v = Variable(t, full, trainable)
  _S.variables[full] = v
  return v


def trainable_variables():
  return [v for v in _S.variables.values() if v.trainable]


def global_variables():
  return list(_S.variables.values())


def placeholder(dtype, shape=None, name=None):
  """Eager placeholder: its value is looked up NOW from the active feed (by creation index)."""
  idx = len(_S.placeholders)
  if _S.feed is None:           # probe pass: hand back a token the harness can identify
    t = Tensor(torch.zeros(0), name or "Placeholder")
    t.placeholder_index = idx
    _S.placeholders.append(t)
    return t
  val = _S.feed(idx, name, dtype, shape)
  if isinstance(val, torch.Tensor):      # fed as is: tf.gradients(loss, <tensor computed from this placeholder>)
    t = val
  else:
    t = torch.from_numpy(np.ascontiguousarray(val)) if isinstance(val, np.ndarray) else torch.tensor(val)
    t = t.to(_torch_dtype(dtype))
  if shape is not None:
    want = [None if s is None else _int(s) for s in shape]
    assert len(want) == t.dim() and all(w is None or w == g for w, g in zip(want, t.shape)), \
        ("feed for placeholder #%d %r has shape %s, declared %s" % (idx, name, list(t.shape), want))
  out = Tensor(t, name or "Placeholder")
  out.placeholder_index = idx
  _S.placeholders.append(out)
  return out


@contextlib.contextmanager
def building(weights, feed, requires_grad=False, allow_init=False):
  """Harness context: fresh graph state, TF-named weights, a feed callable (None = probe)."""
  _S.reset()
  _S.var_values = dict(weights)
  if allow_init:
    _S.var_values["__allow_init__"] = True
  _S.feed = feed
  _S.requires_grad = requires_grad
  try:
    yield _S
  finally:
    pass


def _probe_guard():
  if _S.feed is None:
    raise _StopBuild()


# --------------------------------------------------------------------------------------------- #
# array ops
# --------------------------------------------------------------------------------------------- #
def constant(value, dtype=None, shape=None, name=None):
  if shape is not None:
    t = torch.full(_shape(shape), float(value) if not isinstance(value, int) or dtype else value)
    t = t.to(_torch_dtype(dtype) if dtype is not None else COMPUTE_DTYPE)
    return Tensor(t)
  t = _raw(value)
  if dtype is not None:
    t = t.to(_torch_dtype(dtype))
  return Tensor(t)


def identity(x, name=None):
  return Tensor(_raw(x))


def reshape(tensor, shape, name=None):
  return Tensor(_raw(tensor).reshape(_shape(shape)))


def expand_dims(x, axis=None, name=None):
  return Tensor(_raw(x).unsqueeze(axis))


def squeeze(x, axis=None, name=None):
  t = _raw(x)
  return Tensor(t.squeeze() if axis is None else t.squeeze(axis))


def tile(x, multiples, name=None):
  return Tensor(_raw(x).repeat(*_shape(multiples)))


def shape(x, name=None):
  return Tensor(torch.tensor(list(_raw(x).shape), dtype=torch.int32))


def transpose(x, perm=None, name=None):
  t = _raw(x)
  return Tensor(t.permute(*(perm if perm is not None else reversed(range(t.dim())))))


def reverse(x, axis, name=None):
  return Tensor(torch.flip(_raw(x), dims=list(axis)))


def concat(values, axis, name=None):
  return Tensor(torch.cat([_raw(v) for v in values], dim=axis))


def stack(values, axis=0, name=None):
  return Tensor(torch.stack([_raw(v) for v in values], dim=axis))


def split(value, num_or_size_splits, axis=0):
  return [Tensor(p) for p in torch.chunk(_raw(value), num_or_size_splits, dim=axis)]


def zeros(shape, dtype=float32, name=None):
  return Tensor(torch.zeros(_shape(shape), dtype=_torch_dtype(dtype)))


def ones(shape, dtype=float32, name=None):
  return Tensor(torch.ones(_shape(shape), dtype=_torch_dtype(dtype)))


def zeros_like(x):
  return Tensor(torch.zeros_like(_raw(x)))


def range(start, limit=None, delta=1, dtype=None, name=None):  # pylint: disable=redefined-builtin
  if limit is None:
    start, limit = 0, start
  return Tensor(torch.arange(_int(start), _int(limit), _int(delta),
                             dtype=_torch_dtype(dtype) if dtype else torch.int32))


def cast(x, dtype, name=None):
  return Tensor(_raw(x).to(_torch_dtype(dtype)))


def one_hot(indices, depth, dtype=float32, name=None, **kw):
  """tf.one_hot: float32 by default; out-of-range indices give an all-zero row."""
  idx = _raw(indices).long()
  d = _int(depth)
  ok = (idx >= 0) & (idx < d)
  out = F.one_hot(idx.clamp(0, d - 1), d) * ok.unsqueeze(-1)
  return Tensor(out.to(_torch_dtype(dtype)))


def gather(params, indices, axis=0, name=None):
  p, i = _raw(params), _raw(indices).long()
  assert axis == 0
  assert i.numel() == 0 or (int(i.min()) >= 0 and int(i.max()) < p.shape[0]), "gather index out of range"
  return Tensor(p[i])


def gather_nd(params, indices, name=None):
  p, i = _raw(params), _raw(indices).long()
  return Tensor(p[tuple(i[..., k] for k in builtins_range(i.shape[-1]))])


def where(condition, x=None, y=None, name=None):
  c = _raw(condition)
  if x is None:
    return Tensor(torch.nonzero(c))
  return Tensor(torch.where(c, _raw(x), _raw(y)))


def invert_permutation(x, name=None):
  p = _raw(x).long()
  inv = torch.empty_like(p)
  inv[p] = torch.arange(p.numel(), dtype=p.dtype)
  return Tensor(inv.to(torch.int32))


def map_fn(fn, elems, dtype=None, back_prop=True, **kw):
  e = _raw(elems)
  return Tensor(torch.stack([_raw(fn(Tensor(e[i]))) for i in builtins_range(e.shape[0])]))


import builtins as _builtins  # noqa: E402
builtins_range = _builtins.range


# --------------------------------------------------------------------------------------------- #
# math ops
# --------------------------------------------------------------------------------------------- #
def add(x, y, name=None): return _wrap(x) + y
def multiply(x, y, name=None): return _wrap(x) * y
def less(x, y, name=None): return _wrap(x) < y
def log(x, name=None): return Tensor(torch.log(_raw(x)))
def exp(x, name=None): return Tensor(torch.exp(_raw(x)))
def sqrt(x, name=None): return Tensor(torch.sqrt(_raw(x)))
def tanh(x, name=None): return Tensor(torch.tanh(_raw(x)))
def sigmoid(x, name=None): return Tensor(torch.sigmoid(_raw(x)))


def add_n(inputs, name=None):
  out = _raw(inputs[0])
  for v in inputs[1:]:
    out = out + _raw(v)
  return Tensor(out, name)


def _reduce(fn, x, axis, keepdims):
  t = _raw(x)
  if axis is None:
    return Tensor(fn(t))
  return Tensor(fn(t, dim=axis, keepdim=bool(keepdims)))


def reduce_sum(x, axis=None, keepdims=False, name=None):
  return _reduce(torch.sum, x, axis, keepdims)


def reduce_mean(x, axis=None, keepdims=False, name=None):
  return _reduce(torch.mean, x, axis, keepdims)


def reduce_max(x, axis=None, keepdims=False, name=None):
  t = _raw(x)
  return Tensor(t.max() if axis is None else t.max(dim=axis, keepdim=bool(keepdims)).values)


def reduce_all(x, axis=None, name=None):
  t = _raw(x)
  return Tensor(t.all() if axis is None else t.all(dim=axis))


def argmax(x, axis=None, output_type=int64, name=None):
  """tf.argmax: index of the FIRST maximum (torch.argmax documents the same)."""
  t = _raw(x)
  first = (t == t.max(dim=axis, keepdim=True).values).to(torch.int8).argmax(dim=axis)
  return Tensor(first.to(_torch_dtype(output_type)))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
  x, y = _raw(a), _raw(b)
  if transpose_a: x = x.transpose(-1, -2)
  if transpose_b: y = y.transpose(-1, -2)
  return Tensor(torch.matmul(x, y))


def clip_by_value(t, lo, hi, name=None):
  if isinstance(lo, (Tensor, torch.Tensor)) or isinstance(hi, (Tensor, torch.Tensor)):
    x = _raw(t)        # tensor bounds (SimAug/code/pred_models.py:404-408): min(max(t, lo), hi)
    lo_t = _raw(lo) if isinstance(lo, (Tensor, torch.Tensor)) else torch.as_tensor(lo, dtype=x.dtype)
    hi_t = _raw(hi) if isinstance(hi, (Tensor, torch.Tensor)) else torch.as_tensor(hi, dtype=x.dtype)
    return Tensor(torch.minimum(torch.maximum(x, lo_t.to(x.dtype)), hi_t.to(x.dtype)))
  return Tensor(torch.clamp(_raw(t), min=float(lo), max=float(hi)))


def group(*ops, **kw):
  return _GroupOp(ops)


class _GroupOp(object):
  def __init__(self, ops): self.ops = ops
  def run(self, *a, **kw):
    for o in self.ops: o.run()


# --------------------------------------------------------------------------------------------- #
# control flow
# --------------------------------------------------------------------------------------------- #
def cond(pred, true_fn=None, false_fn=None, name=None, **kw):
  """tf.cond, eagerly: the first tf op of build_forward, which is where a probe pass stops."""
  _probe_guard()
  p = _raw(pred)
  return true_fn() if bool(p.item()) else false_fn()


def while_loop(cond, body, loop_vars, back_prop=True, maximum_iterations=None, **kw):  # pylint: disable=redefined-outer-name
  """tf.while_loop, eagerly; `maximum_iterations` bounds the trip count (SimAug's PGD loop has cond = True, :146-155).
  A body that returns one tensor for a one-element loop_vars list is accepted, and the single result is returned
  unwrapped for a single loop variable - as TF does."""
  vars_ = list(loop_vars)
  it = 0
  def truth(v):
    return bool(v) if isinstance(v, (bool, np.bool_)) else bool(_raw(v).item())
  while truth(cond(*vars_)) and (maximum_iterations is None or it < int(maximum_iterations)):
    out = body(*vars_)
    vars_ = list(out) if isinstance(out, (list, tuple)) else [out]
    it += 1
  return vars_ if len(vars_) != 1 or isinstance(loop_vars, tuple) else vars_[0]


class TensorArray(object):
  """tf.TensorArray with write-once-per-index semantics relaxed to a dict; write() returns the
  array itself (the reference always rebinds the result)."""

  def __init__(self, dtype, size=0, dynamic_size=False, **kw):
    self.dtype = _dtype(dtype)
    self.items = {}
    self.size_ = _int(size)
    self.dynamic = dynamic_size

  def write(self, index, value):
    i = _int(index)
    assert self.dynamic or i < self.size_, "TensorArray write out of bounds"
    self.items[i] = _raw(value)
    return self

  def read(self, index):
    return Tensor(self.items[_int(index)])

  def unstack(self, value):
    t = _raw(value)
    for i in builtins_range(t.shape[0]):
      self.items[i] = t[i]
    return self

  def stack(self):
    n = max(self.items) + 1 if self.items else 0
    assert sorted(self.items) == list(builtins_range(n)), "TensorArray has holes"
    return Tensor(torch.stack([self.items[i] for i in builtins_range(n)]))

  def mark_used(self):
    pass

  def size(self):
    return Tensor(torch.tensor(max(self.size_, len(self.items)), dtype=torch.int32))


# --------------------------------------------------------------------------------------------- #
# nest
# --------------------------------------------------------------------------------------------- #
class _Nest(object):
  @staticmethod
  def map_structure(fn, *structs):
    s0 = structs[0]
    if isinstance(s0, tuple) and hasattr(s0, "_fields"):
      return type(s0)(*[_Nest.map_structure(fn, *[s[i] for s in structs])
                        for i in builtins_range(len(s0))])
    if isinstance(s0, (list, tuple)):
      return type(s0)(_Nest.map_structure(fn, *[s[i] for s in structs])
                      for i in builtins_range(len(s0)))
    return fn(*structs)

  @staticmethod
  def flatten(s):
    if isinstance(s, (list, tuple)):
      out = []
      for e in s:
        out.extend(_Nest.flatten(e))
      return out
    return [s]


nest = _Nest()


# --------------------------------------------------------------------------------------------- #
# nn
# --------------------------------------------------------------------------------------------- #
def _same_pad(in_size, k, stride, dilation=1):
  """SAME rule of tf.nn.conv2d (core/framework/common_shape_fns.cc GetWindowedOutputSize):
  out = ceil(in/stride); pad_total = max((out-1)*stride + (k-1)*dilation + 1 - in, 0);
  pad_before = pad_total // 2; the odd cell goes AFTER."""
  out = -(-in_size // stride)
  eff = (k - 1) * dilation + 1
  total = max((out - 1) * stride + eff - in_size, 0)
  return total // 2, total - total // 2


LSTMStateTuple = collections.namedtuple("LSTMStateTuple", ("c", "h"))


class _NN(object):
  tanh = staticmethod(tanh)
  sigmoid = staticmethod(sigmoid)

  @staticmethod
  def relu(x, name=None):
    return Tensor(torch.relu(_raw(x)))

  @staticmethod
  def leaky_relu(x, alpha=0.2, name=None):
    return Tensor(F.leaky_relu(_raw(x), alpha))

  @staticmethod
  def conv2d(input=None, filter=None, strides=None, padding=None, dilations=None,  # pylint: disable=redefined-builtin
             data_format="NHWC", name=None, filters=None):
    """tf.nn.conv2d, NHWC x HWIO, cross-correlation, explicit asymmetric SAME padding."""
    x, w = _raw(input), _raw(filter if filter is not None else filters)
    assert data_format == "NHWC" and padding in ("SAME", "VALID")
    sh, sw = (strides[1], strides[2]) if len(strides) == 4 else (strides[0], strides[-1])
    dh, dw = (1, 1) if dilations is None else ((dilations[1], dilations[2]) if not isinstance(dilations, int)
                                                else (dilations, dilations))
    x = x.permute(0, 3, 1, 2)
    if padding == "SAME":
      pt, pb = _same_pad(x.shape[2], w.shape[0], sh, dh)
      pl, pr = _same_pad(x.shape[3], w.shape[1], sw, dw)
      x = F.pad(x, (pl, pr, pt, pb))
    y = F.conv2d(x.contiguous(), w.permute(3, 2, 0, 1).contiguous(), stride=(sh, sw), dilation=(dh, dw))
    return Tensor(y.permute(0, 2, 3, 1))

  @staticmethod
  def bias_add(value, bias, data_format="NHWC", name=None):
    return Tensor(_raw(value) + _raw(bias))

  @staticmethod
  def embedding_lookup(params, ids, name=None):
    _probe_guard()        # SimAug/code/pred_models.py: the first op after the placeholders
    return gather(params, ids)

  @staticmethod
  def l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
    """python/ops/nn_impl.py: x * rsqrt(max(sum(x^2, axis, keepdims), epsilon))."""
    t = _raw(x)
    axis = axis if axis is not None else dim
    sq = (t * t).sum(dim=axis, keepdim=True)
    return Tensor(t * torch.rsqrt(torch.clamp(sq, min=epsilon)))

  @staticmethod
  def softmax(logits, axis=-1, name=None):
    return Tensor(torch.softmax(_raw(logits), dim=axis))

  @staticmethod
  def log_softmax(logits, axis=-1, name=None):
    return Tensor(torch.log_softmax(_raw(logits), dim=axis))

  @staticmethod
  def top_k(input, k=1, sorted=True, name=None):  # pylint: disable=redefined-builtin
    """tf.nn.top_k: descending values; among equal values the LOWER index comes first
    (core/kernels/topk_op.cc).  A stable descending sort gives exactly that order."""
    t = _raw(input)
    vals, idx = torch.sort(t, dim=-1, descending=True, stable=True)
    k = _int(k)
    return Tensor(vals[..., :k]), Tensor(idx[..., :k].to(torch.int32))

  @staticmethod
  def l2_loss(t, name=None):
    x = _raw(t)
    return Tensor((x * x).sum() / 2)

  @staticmethod
  def sparse_softmax_cross_entropy_with_logits(labels=None, logits=None, name=None):
    lg, lb = _raw(logits), _raw(labels).long()
    lp = torch.log_softmax(lg, dim=-1)
    return Tensor(-lp.gather(-1, lb.unsqueeze(-1)).squeeze(-1))

  @staticmethod
  def softmax_cross_entropy_with_logits(labels=None, logits=None, name=None):
    return Tensor(-(_raw(labels) * torch.log_softmax(_raw(logits), dim=-1)).sum(-1))

  @staticmethod
  def moments(x, axes, keep_dims=False, name=None):
    t = _raw(x)
    m = t.mean(dim=list(axes), keepdim=True)
    v = ((t - m) ** 2).mean(dim=list(axes), keepdim=True)
    if not keep_dims:
      m, v = m.squeeze(), v.squeeze()
    return Tensor(m), Tensor(v)

  @staticmethod
  def batch_normalization(x, mean, variance, offset, scale, variance_epsilon, name=None):
    inv = torch.rsqrt(_raw(variance) + variance_epsilon) * _raw(scale)
    return Tensor(_raw(x) * inv + (_raw(offset) - _raw(mean) * inv))

  # ---- recurrent drivers ---------------------------------------------------------------- #
  @staticmethod
  def dynamic_rnn(cell, inputs, sequence_length=None, initial_state=None, dtype=None,
                  time_major=False, scope=None, **kw):
    """python/ops/rnn.py dynamic_rnn: variables under variable_scope(scope or "rnn"); zero
    initial state; per step ``_rnn_step``: rows with time >= sequence_length emit zeros and copy
    their state through.  Returns (outputs [N,T,...], final_state)."""
    x = _raw(inputs)
    assert not time_major
    n, steps = x.shape[0], x.shape[1]
    with variable_scope(scope or "rnn"):
      state = initial_state if initial_state is not None else cell.zero_state(n, dtype)
      seq = _raw(sequence_length).long() if sequence_length is not None else None
      outs = []
      for t in builtins_range(steps):
        out, new_state = cell(Tensor(x[:, t]), state)
        if seq is not None:
          done = (t >= seq)
          sel = lambda new, old: Tensor(torch.where(
              done.reshape([-1] + [1] * (_raw(new).dim() - 1)), _raw(old), _raw(new)))
          out = sel(out, zeros_like(out))
          new_state = nest.map_structure(sel, new_state, state)
        outs.append(_raw(out))
        state = new_state
    return Tensor(torch.stack(outs, dim=1)), state

  @staticmethod
  def raw_rnn(cell, loop_fn, parallel_iterations=None, swap_memory=False, scope=None):
    """python/ops/rnn.py raw_rnn, protocol kept call for call:
      (finished, next_input, initial_state, emit_structure, loop_state) = loop_fn(0, None, None, None)
      while not all(finished):
        (output, cell_state) = cell(current_input, state)
        (next_finished, next_input, next_state, emit, loop_state') = loop_fn(time+1, output, cell_state, loop_state)
        loop_state = loop_state' unless it is None
        emit  = where(finished, zeros, emit)         # _copy_some_through, finished = flag BEFORE this step
        state = where(finished, state, next_state)
        emit_ta.write(time, emit); finished |= next_finished; time += 1
    emit_structure None => emit has the cell's output structure.  Returns (emit_ta, state, loop_state).
    """
    with variable_scope(scope or "rnn"):
      time = 0
      finished, next_input, state, emit_structure, loop_state = loop_fn(
          Tensor(torch.tensor(0, dtype=torch.int32)), None, None, None)
      assert emit_structure is None
      emit_ta = TensorArray(float32, size=0, dynamic_size=True)
      finished = _raw(finished)
      current_input = next_input
      while not bool(finished.all()):
        output, cell_state = cell(current_input, state)
        time += 1
        nf, next_input, next_state, emit, new_loop_state = loop_fn(
            Tensor(torch.tensor(time, dtype=torch.int32)), output, cell_state, loop_state)
        if new_loop_state is not None:
          loop_state = new_loop_state
        bc = lambda t: finished.reshape([-1] + [1] * (t.dim() - 1))
        emit_t = _raw(emit)
        emit_t = torch.where(bc(emit_t), torch.zeros_like(emit_t), emit_t)
        state = nest.map_structure(
            lambda cur, cand: Tensor(torch.where(bc(_raw(cand)), _raw(cur), _raw(cand))),
            state, next_state)
        emit_ta.write(time - 1, emit_t)
        finished = finished | _raw(nf)
        current_input = next_input
    return emit_ta, state, loop_state


class _DropoutWrapper(object):
  """tf.nn.rnn_cell.DropoutWrapper(cell, input_keep_prob): dropout on the INPUT only (2nd
  positional argument).  nn_ops.dropout with rate = 1 - keep: x * 1/keep * (u >= rate), u~U[0,1):
  the identity when keep == 1, which is what every published config feeds."""

  def __init__(self, cell, input_keep_prob=1.0, output_keep_prob=1.0, state_keep_prob=1.0, **kw):
    self.cell = cell
    self.keep = input_keep_prob
    assert output_keep_prob == 1.0 and state_keep_prob == 1.0

  def zero_state(self, batch_size, dtype):
    return self.cell.zero_state(batch_size, dtype)

  def __call__(self, inputs, state, scope=None):
    keep = float(_raw(self.keep).item())
    x = _raw(inputs)
    if keep < 1.0:
      u = torch.from_numpy(_S.rng.uniform(size=tuple(x.shape))).to(x.dtype)
      x = x * (1.0 / keep) * (u >= (1.0 - keep)).to(x.dtype)
    return self.cell(Tensor(x), state)


class _RnnCell(object):
  DropoutWrapper = _DropoutWrapper
  LSTMStateTuple = LSTMStateTuple


_NN.rnn_cell = _RnnCell()
nn = _NN()


class _ConvLSTMCell(object):
  """tf.contrib.rnn.ConvLSTMCell (contrib/rnn/python/ops/rnn_cell.py, TF 1.15):
    cell, hidden = state
    new_hidden = _conv([inputs, hidden], kernel_shape, 4*output_channels, use_bias)
        = conv_SAME(concat([inputs, hidden], -1), kernel[kh,kw,Cin+Ch,4Ch]) + biases[4Ch]
    input_gate, new_input, forget_gate, output_gate = split(new_hidden, 4, axis=-1)
    new_cell = sigmoid(forget_gate + forget_bias) * cell + sigmoid(input_gate) * tanh(new_input)
    output   = tanh(new_cell) * sigmoid(output_gate)
    return output, LSTMStateTuple(new_cell, output)
  Variables "kernel"/"biases" live in variable_scope(<scope at first call>/<name>) (Layer
  semantics: built once, reused on every later call)."""

  def __init__(self, conv_ndims, input_shape, output_channels, kernel_shape, use_bias=True,
               skip_connection=False, forget_bias=1.0, initializers=None, name="conv_lstm_cell"):
    assert conv_ndims == 2 and not skip_connection
    self.input_shape = list(input_shape)
    self.oc = output_channels
    self.kernel_shape = list(kernel_shape)
    self.use_bias = use_bias
    self.forget_bias = forget_bias
    self.name = name
    self._scope = None

  @property
  def output_size(self):
    return TensorShape(self.input_shape[:-1] + [self.oc])

  def zero_state(self, batch_size, dtype):
    shp = [_int(batch_size)] + self.input_shape[:-1] + [self.oc]
    z = lambda: Tensor(torch.zeros(shp, dtype=COMPUTE_DTYPE))
    return LSTMStateTuple(z(), z())

  def __call__(self, inputs, state, scope=None):
    if self._scope is None:
      self._scope = VariableScope((_S.scope[-1] + "/" + self.name) if _S.scope[-1] else self.name)
    cell, hidden = state
    x = torch.cat([_raw(inputs), _raw(hidden)], dim=-1)
    with variable_scope(self._scope):
      kernel = get_variable("kernel", self.kernel_shape + [x.shape[-1], 4 * self.oc], dtype=float32)
      res = _raw(nn.conv2d(Tensor(x), kernel, [1, 1, 1, 1], "SAME"))
      if self.use_bias:
        res = res + _raw(get_variable("biases", [4 * self.oc], dtype=float32,
                                      initializer=constant_initializer(0.0)))
    i, j, f, o = torch.chunk(res, 4, dim=-1)
    new_cell = torch.sigmoid(f + self.forget_bias) * _raw(cell)
    new_cell = new_cell + torch.sigmoid(i) * torch.tanh(j)
    output = torch.tanh(new_cell) * torch.sigmoid(o)
    return Tensor(output), LSTMStateTuple(Tensor(new_cell), Tensor(output))


class _Contrib(object):
  class rnn(object):
    ConvLSTMCell = _ConvLSTMCell


contrib = _Contrib()


# --------------------------------------------------------------------------------------------- #
# losses, gradients, optimizers (Trainer, code/pred_models.py:1636-1717)
# --------------------------------------------------------------------------------------------- #
class _Losses(object):
  class Reduction(object):
    MEAN = "weighted_mean"
    SUM = "weighted_sum"

  @staticmethod
  def huber_loss(labels, predictions, weights=1.0, delta=1.0, scope=None, reduction="weighted_mean",
                 **kw):
    """python/ops/losses/losses_impl.py: e=pred-labels; q=min(|e|,delta); 0.5 q^2 + delta(|e|-q);
    Reduction.MEAN with weights=1 = mean over all elements."""
    e = (_raw(predictions) - _raw(labels)).abs()
    q = torch.clamp(e, max=delta)
    l = 0.5 * q * q + delta * (e - q)
    assert reduction == "weighted_mean" and weights == 1.0
    return Tensor(l.mean())


losses = _Losses()


def gradients(ys, xs, **kw):
  """tf.gradients: d sum(ys) / d xs; `xs` may be one tensor (SimAug/code/pred_models.py:397) or a list."""
  y = _raw(ys)
  if isinstance(xs, (Tensor, torch.Tensor)):
    xs = [xs]
  gs = torch.autograd.grad(y.sum() if y.dim() else y, [_raw(x) for x in xs], allow_unused=True, retain_graph=True)
  return [None if g is None else Tensor(g) for g in gs]


class _ApplyOp(object):
  def __init__(self, fn):
    self.fn, self.done = fn, 0

  def run(self, *a, **kw):
    self.fn()
    self.done += 1


class _Optimizer(object):
  def __init__(self, learning_rate, **kw):
    self.lr = learning_rate
    self.kw = kw
    self.slots = {}

  def _lr(self):
    return float(_raw(self.lr).item()) if isinstance(self.lr, (Tensor, torch.Tensor)) else float(self.lr)

  def apply_gradients(self, grads_and_vars, global_step=None, name=None):
    gv = [(g, v) for g, v in grads_and_vars if g is not None]

    def run():
      lr = self._lr()
      with torch.no_grad():
        for g, v in gv:
          self._apply(_raw(g), v, lr)
        if global_step is not None:
          global_step.t += 1
    return _ApplyOp(run)


class _Adadelta(_Optimizer):
  """python/training/adadelta.py + core/kernels/training_ops.cc ApplyAdadelta (rho=.95, eps=1e-8):
    accum = rho*accum + (1-rho) g^2 ; update = sqrt(accum_update+eps) * rsqrt(accum+eps) * g
    accum_update = rho*accum_update + (1-rho) update^2 ; var -= lr * update"""

  def __init__(self, learning_rate=0.001, rho=0.95, epsilon=1e-8, **kw):
    _Optimizer.__init__(self, learning_rate)
    self.rho, self.eps = rho, epsilon

  def _apply(self, g, v, lr):
    a, au = self.slots.setdefault(v.op.name, (torch.zeros_like(v.t), torch.zeros_like(v.t)))
    a.mul_(self.rho).add_((1 - self.rho) * g * g)
    upd = torch.sqrt(au + self.eps) * torch.rsqrt(a + self.eps) * g
    au.mul_(self.rho).add_((1 - self.rho) * upd * upd)
    v.t.sub_(lr * upd)


class _Momentum(_Optimizer):
  def __init__(self, learning_rate, momentum, **kw):
    _Optimizer.__init__(self, learning_rate)
    self.m = momentum

  def _apply(self, g, v, lr):
    acc = self.slots.setdefault(v.op.name, torch.zeros_like(v.t))
    acc.mul_(self.m).add_(g)
    v.t.sub_(lr * acc)


class _Adam(_Optimizer):
  def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **kw):
    _Optimizer.__init__(self, learning_rate)
    self.b1, self.b2, self.eps, self.t = beta1, beta2, epsilon, 0

  def apply_gradients(self, grads_and_vars, global_step=None, name=None):
    op = _Optimizer.apply_gradients(self, grads_and_vars, global_step, name)
    inner = op.fn

    def run():
      self.t += 1
      inner()
    op.fn = run
    return op

  def _apply(self, g, v, lr):
    m, s = self.slots.setdefault(v.op.name, (torch.zeros_like(v.t), torch.zeros_like(v.t)))
    m.mul_(self.b1).add_((1 - self.b1) * g)
    s.mul_(self.b2).add_((1 - self.b2) * g * g)
    lr_t = lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
    v.t.sub_(lr_t * m / (torch.sqrt(s) + self.eps))


class _RMSProp(_Optimizer):
  def __init__(self, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10, **kw):
    _Optimizer.__init__(self, learning_rate)
    self.decay, self.mom, self.eps = decay, momentum, epsilon

  def _apply(self, g, v, lr):
    ms, mom = self.slots.setdefault(v.op.name, (torch.ones_like(v.t), torch.zeros_like(v.t)))
    ms.mul_(self.decay).add_((1 - self.decay) * g * g)
    mom.mul_(self.mom).add_(lr * g * torch.rsqrt(ms + self.eps))
    v.t.sub_(mom)


class _Train(object):
  AdadeltaOptimizer = _Adadelta
  MomentumOptimizer = _Momentum
  AdamOptimizer = _Adam
  RMSPropOptimizer = _RMSProp

  @staticmethod
  def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
    p = _raw(global_step).to(COMPUTE_DTYPE) / decay_steps
    if staircase:
      p = torch.floor(p)
    return Tensor(learning_rate * decay_rate ** p)

  @staticmethod
  def cosine_decay(learning_rate, global_step, decay_steps, alpha=0.0, name=None):
    g = torch.clamp(_raw(global_step).to(COMPUTE_DTYPE), max=decay_steps)
    c = 0.5 * (1 + torch.cos(math.pi * g / decay_steps))
    return Tensor(learning_rate * ((1 - alpha) * c + alpha))


train = _Train()


# --------------------------------------------------------------------------------------------- #
# the few extra symbols SimAug/code/pred_models.py touches
# --------------------------------------------------------------------------------------------- #
def sign(x, name=None): return Tensor(torch.sign(_raw(x)))
def stop_gradient(x, name=None): return Tensor(_raw(x).detach())
def greater(x, y, name=None): return _wrap(x) > y
def floormod(x, y, name=None): return Tensor(torch.remainder(_raw(x), _raw(y)))


def random_uniform(shape, minval=0, maxval=None, dtype=None, seed=None, name=None):
  """No hidden randomness: the harness supplies the draw (building(...) state `uniform_hook`)."""
  if _S.uniform_hook is None:
    raise RuntimeError("tf.random_uniform reached without a harness-supplied uniform_hook")
  shp = [int(_raw(d).item()) if isinstance(d, (Tensor, torch.Tensor)) else _int(d) for d in
         (shape if isinstance(shape, (list, tuple)) else _raw(shape).tolist())]
  v = _S.uniform_hook(tuple(shp), minval, maxval, dtype)
  t = torch.as_tensor(np.asarray(v))
  return Tensor(t.to(_torch_dtype(dtype)) if dtype is not None else t.to(COMPUTE_DTYPE))


class _Random(object):
  uniform = staticmethod(random_uniform)


random = _Random()


_pymath = math


class _MathNS(object):
  """tf.math; everything else falls through to Python's math module, which this file's own code uses."""
  @staticmethod
  def maximum(x, y, name=None):
    a = _raw(x)
    b = _raw(y) if isinstance(y, (Tensor, torch.Tensor)) else torch.as_tensor(y, dtype=a.dtype)
    return Tensor(torch.maximum(a, b.to(a.dtype)))

  def __getattr__(self, name):
    return getattr(_pymath, name)


math = _MathNS()


class _Beta(object):
  def __init__(self, concentration1, concentration0, **kw): pass

  def sample(self, *a, **kw):
    if _S.beta_sample is None:
      raise RuntimeError("tf.distributions.Beta.sample reached without a harness-supplied beta_sample")
    return Tensor(torch.tensor(float(_S.beta_sample), dtype=COMPUTE_DTYPE))


class _Distributions(object):
  Beta = _Beta


distributions = _Distributions()
_NN.softmax_cross_entropy_with_logits_v2 = staticmethod(_NN.softmax_cross_entropy_with_logits)


class Session(object):
  """Enough of tf.Session for harness code: fetch = value already computed eagerly."""

  def __init__(self, *a, **kw): pass
  def __enter__(self): return self
  def __exit__(self, *a): return False

  def run(self, fetches, feed_dict=None):
    if isinstance(fetches, (list, tuple)):
      return [self.run(f) for f in fetches]
    if fetches is None:
      return None
    if hasattr(fetches, "run") and not isinstance(fetches, Tensor):
      fetches.run()
      return None
    return fetches.numpy()
