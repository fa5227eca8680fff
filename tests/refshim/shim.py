"""See tests/refshim/__init__.py.  install(tree) registers the stand-in modules in sys.modules and puts
/root/reference/<tree> first on sys.path."""
import dataclasses
import functools
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("GDDIM_REFERENCE", "/root/reference")


# ---------------------------------------------------------------------------------------------------------------------
# arrays
# ---------------------------------------------------------------------------------------------------------------------
class JArr(np.ndarray):
  """numpy array with the jax `.at[idx].set/add/multiply/get` functional-update property."""
  __array_priority__ = 100.0

  @property
  def at(self):
    return _At(self)

  def block_until_ready(self):
    return self


class _At:
  def __init__(self, a):
    self.a = a

  def __getitem__(self, idx):
    return _AtIdx(self.a, idx)


class _AtIdx:
  def __init__(self, a, idx):
    self.a, self.idx = a, idx

  def _new(self):
    return np.array(self.a, copy=True).view(JArr)

  def set(self, v):
    out = self._new()
    out[self.idx] = v
    return out

  def add(self, v):
    out = self._new()
    out[self.idx] += v
    return out

  def multiply(self, v):
    out = self._new()
    out[self.idx] *= v
    return out

  def get(self):
    return wrap(np.asarray(self.a)[self.idx])


def wrap(x):
  if isinstance(x, np.ndarray) and not isinstance(x, JArr):
    return x.view(JArr)
  if isinstance(x, tuple):
    return tuple(wrap(v) for v in x)
  if isinstance(x, list):
    return [wrap(v) for v in x]
  return x


def _fix_kwargs(kw):
  if isinstance(kw.get("axis"), list):
    kw["axis"] = tuple(kw["axis"])
  return kw


def _np_wrapped(fn):
  @functools.wraps(fn)
  def f(*a, **kw):
    return wrap(fn(*a, **_fix_kwargs(kw)))
  return f


def _module(name, **attrs):
  m = types.ModuleType(name)
  m.__dict__.update(attrs)
  m.__path__ = []          # so that `import a.b.c` treats it as a package
  sys.modules[name] = m
  return m


class _Dummy:
  """Absorbs any attribute access / call (stubs for wandb, tensorflow, matplotlib ...)."""

  def __init__(self, name="dummy"):
    self._name = name

  def __getattr__(self, k):
    if k.startswith("__"):
      raise AttributeError(k)
    return _Dummy(f"{self._name}.{k}")

  def __call__(self, *a, **k):
    return _Dummy(self._name + "()")


def _stub_module(name):
  m = _module(name)
  def _getattr(k, _n=name):
    if k.startswith("__"):
      raise AttributeError(k)
    return _Dummy(f"{_n}.{k}")
  m.__getattr__ = _getattr
  return m


# ---------------------------------------------------------------------------------------------------------------------
# pytrees (dict / list / tuple / dataclass) -- enough for State and parameter dicts
# ---------------------------------------------------------------------------------------------------------------------
def tree_map(f, t, *rest):
  if isinstance(t, dict):
    return type(t)((k, tree_map(f, v, *[r[k] for r in rest])) for k, v in t.items())
  if isinstance(t, (list, tuple)):
    out = [tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(t)]
    return type(t)(out) if not hasattr(t, "_fields") else type(t)(*out)
  if dataclasses.is_dataclass(t) and not isinstance(t, type):
    return dataclasses.replace(t, **{fl.name: tree_map(f, getattr(t, fl.name), *[getattr(r, fl.name) for r in rest])
                                     for fl in dataclasses.fields(t)})
  if t is None:
    return None
  return f(t, *rest)


def tree_leaves(t):
  out = []
  tree_map(lambda x: out.append(x), t)
  return out


# ---------------------------------------------------------------------------------------------------------------------
# jax core transforms
# ---------------------------------------------------------------------------------------------------------------------
def jit(fn=None, static_argnums=None, **kw):
  if fn is None:
    return lambda f: f
  return fn


def _axis_size(args, in_axes):
  for a, ax in zip(args, in_axes):
    if ax is not None:
      return tree_leaves(a)[0].shape[ax]
  raise ValueError("vmap: no mapped argument")


def _take(a, ax, i):
  if ax is None:
    return a
  return tree_map(lambda x: wrap(np.take(np.asarray(x), i, axis=ax)) if np.ndim(x) > 0 else x, a)


def _stack(outs, out_axis=0):
  first = outs[0]
  if isinstance(first, tuple):
    return tuple(_stack([o[i] for o in outs], out_axis) for i in range(len(first)))
  if isinstance(first, list):
    return [_stack([o[i] for o in outs], out_axis) for i in range(len(first))]
  return wrap(np.stack([np.asarray(o) for o in outs], axis=out_axis))


def vmap(fn, in_axes=0, out_axes=0):
  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    assert len(axes) == len(args), (len(axes), len(args))
    n = _axis_size(args, axes)
    return _stack([fn(*[_take(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)], out_axes)
  return mapped


def pmap(fn, axis_name=None, static_broadcasted_argnums=(), **kw):
  static = (static_broadcasted_argnums,) if isinstance(static_broadcasted_argnums, int) else tuple(static_broadcasted_argnums)

  def mapped(*args):
    n = None
    for i, a in enumerate(args):
      if i not in static and a is not None:
        n = tree_leaves(a)[0].shape[0]
        break
    assert n == 1, "shim pmap: one local device"
    outs = []
    for d in range(n):
      outs.append(fn(*[a if (i in static or a is None) else tree_map(lambda x: wrap(np.asarray(x)[d]), a)
                       for i, a in enumerate(args)]))
    return _stack(outs)
  return mapped


def scan(f, init, xs, length=None):
  carry = init
  ys = []
  n = length if xs is None else tree_leaves(xs)[0].shape[0]
  for i in range(n):
    x = None if xs is None else tree_map(lambda a: wrap(np.asarray(a)[i]) if np.ndim(np.asarray(a)[i]) else np.asarray(a)[i], xs)
    carry, y = f(carry, x)
    ys.append(y)
  return carry, (_stack(ys) if ys and ys[0] is not None else None)


def fori_loop(lo, hi, body, init):
  val = init
  for i in range(int(lo), int(hi)):
    val = body(i, val)
  return val


def cond(pred, true_fn, false_fn, operand=None):
  return true_fn(operand) if bool(pred) else false_fn(operand)


# ---------------------------------------------------------------------------------------------------------------------
# jax.lax convolution / fft / slicing
# ---------------------------------------------------------------------------------------------------------------------
def _same_pads(size, k, s, dil=1):
  keff = (k - 1) * dil + 1
  out = -(-size // s)
  total = max((out - 1) * s + keff - size, 0)
  return total // 2, total - total // 2


def conv_general_dilated(lhs, rhs, window_strides, padding, lhs_dilation=None, rhs_dilation=None, dimension_numbers=None,
                         feature_group_count=1, **kw):
  import torch
  import torch.nn.functional as F
  assert lhs_dilation is None or tuple(lhs_dilation) == (1, 1)
  lspec, rspec, ospec = dimension_numbers
  assert rspec == "HWIO" and lspec == ospec and lspec in ("NHWC", "NCHW")
  x = torch.from_numpy(np.ascontiguousarray(np.asarray(lhs, dtype=np.float64)))
  w = torch.from_numpy(np.ascontiguousarray(np.asarray(rhs, dtype=np.float64))).permute(3, 2, 0, 1).contiguous()
  if lspec == "NHWC":
    x = x.permute(0, 3, 1, 2)
  dil = tuple(rhs_dilation) if rhs_dilation is not None else (1, 1)
  kh, kw_ = w.shape[2], w.shape[3]
  if isinstance(padding, str):
    if padding.upper() == "SAME":
      ph = _same_pads(x.shape[2], kh, window_strides[0], dil[0])
      pw = _same_pads(x.shape[3], kw_, window_strides[1], dil[1])
    elif padding.upper() == "VALID":
      ph = pw = (0, 0)
    else:
      raise ValueError(padding)
  else:
    ph, pw = tuple(padding[0]), tuple(padding[1])
  x = F.pad(x, (pw[0], pw[1], ph[0], ph[1]))
  y = F.conv2d(x, w, None, stride=tuple(window_strides), dilation=dil, groups=feature_group_count)
  if lspec == "NHWC":
    y = y.permute(0, 2, 3, 1)
  return wrap(np.ascontiguousarray(y.numpy()))


def conv_transpose(*a, **k):
  raise NotImplementedError("shim: jax.lax.conv_transpose (only reached by progressive='residual' output pyramids)")


class FftType:
  FFT, IFFT, RFFT, IRFFT = "FFT", "IFFT", "RFFT", "IRFFT"


def lax_fft(a, fft_type, fft_lengths):
  a = np.asarray(a)
  n = len(fft_lengths)
  axes = tuple(range(a.ndim - n, a.ndim))
  if fft_type == FftType.FFT:
    return wrap(np.fft.fftn(a, s=fft_lengths, axes=axes))
  if fft_type == FftType.IFFT:
    return wrap(np.fft.ifftn(a, s=fft_lengths, axes=axes))
  if fft_type == FftType.RFFT:
    return wrap(np.fft.rfftn(a, s=fft_lengths, axes=axes))
  if fft_type == FftType.IRFFT:
    return wrap(np.fft.irfftn(a, s=fft_lengths, axes=axes))
  raise ValueError(fft_type)


def slice_in_dim(x, start, limit, stride=1, axis=0):
  idx = [slice(None)] * np.ndim(x)
  idx[axis] = slice(start, limit, stride)
  return wrap(np.asarray(x)[tuple(idx)])


# ---------------------------------------------------------------------------------------------------------------------
# jax.random: recorded numpy stream (NOT threefry -- fixtures carry the drawn noise explicitly)
# ---------------------------------------------------------------------------------------------------------------------
_DRAWS = []


def draws():
  return _DRAWS


def clear_draws():
  del _DRAWS[:]


def PRNGKey(seed):
  return np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.uint32).view(JArr)


def _rng_of(key):
  return np.random.default_rng([int(v) for v in np.asarray(key).ravel()])


def split(key, num=2):
  r = _rng_of(key)
  return wrap(r.integers(0, 2**32, size=(num, 2), dtype=np.uint32))


def fold_in(key, data):
  return wrap(np.array([int(np.asarray(key).ravel()[1]), int(data) & 0xFFFFFFFF], dtype=np.uint32))


def normal(key, shape=(), dtype=np.float64):
  z = _rng_of(key).standard_normal(tuple(shape)).astype(np.float32).astype(np.float64)   # fp32-representable: fixtures store it losslessly
  _DRAWS.append(z)
  return wrap(z)


def uniform(key, shape=(), dtype=np.float64, minval=0.0, maxval=1.0):
  return wrap(_rng_of(key).uniform(minval, maxval, size=tuple(shape)))


def multivariate_normal(key, mean, cov, shape=None, dtype=np.float64, method="cholesky"):
  """jax.random.multivariate_normal: mean + factor @ N(0, I), factor = u*sqrt(s) (svd) | cholesky (jax/_src/random.py)."""
  mean, cov = np.asarray(mean, np.float64), np.asarray(cov, np.float64)
  if method == "svd":
    u, s, _ = np.linalg.svd(cov)
    # the sign of a singular vector is implementation-defined in jax (LAPACK on CPU, cuSolver on GPU); this stand-in
    # fixes it like the product and the oracle do: largest-magnitude entry of every column positive (2x2 only)
    assert cov.shape == (2, 2)
    for j in range(2):
      k = 0 if abs(u[0, j]) >= abs(u[1, j]) else 1
      if u[k, j] < 0:
        u[:, j] = -u[:, j]
    factor = u * np.sqrt(s[..., None, :])
  elif method == "eigh":
    w, v = np.linalg.eigh(cov)
    factor = v * np.sqrt(w[..., None, :])
  else:
    factor = np.linalg.cholesky(cov)
  z = normal(key, tuple(shape or ()) + mean.shape[-1:])
  return wrap(mean + np.einsum("...ij,...j->...i", factor, np.asarray(z)))


# ---------------------------------------------------------------------------------------------------------------------
# jax.nn
# ---------------------------------------------------------------------------------------------------------------------
def _sigmoid(x):
  x = np.asarray(x)
  return 1.0 / (1.0 + np.exp(-x))


def swish(x):
  return wrap(np.asarray(x) * _sigmoid(x))


def softmax(x, axis=-1):
  x = np.asarray(x)
  e = np.exp(x - x.max(axis=axis, keepdims=True))
  return wrap(e / e.sum(axis=axis, keepdims=True))


class _Init:
  """An initializer that remembers what it is (used by linen_collect to compare against oracle.collect_specs)."""

  def __init__(self, kind, scale=1.0):
    self.kind, self.scale = kind, scale

  def __call__(self, key, shape, dtype=np.float64):
    raise RuntimeError("shim initializers are descriptors; parameters are supplied by the caller")


def variance_scaling(scale, mode, distribution, **kw):
  if mode == "fan_avg" and distribution == "uniform":
    return _Init("vs", scale)
  return _Init(f"vs_{mode}_{distribution}", scale)      # not used by the score networks on the hot path


# ---------------------------------------------------------------------------------------------------------------------
# flax.linen
# ---------------------------------------------------------------------------------------------------------------------
_STACK = []
_CTX = {"params": None, "collect": None}


def compact(fn):
  return fn


class Module:
  _fields = ()

  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    fields = []
    for klass in reversed(cls.__mro__):
      for n in klass.__dict__.get("__annotations__", {}):
        if n not in fields and n not in ("name", "parent"):
          fields.append(n)
    cls._fields = tuple(fields)
    if "__call__" in cls.__dict__:
      orig = cls.__dict__["__call__"]

      @functools.wraps(orig)
      def call(self, *a, **k):
        _STACK.append(self)
        try:
          return orig(self, *a, **k)
        finally:
          _STACK.pop()
      cls.__call__ = call

  def __init__(self, *args, **kwargs):
    name = kwargs.pop("name", None)
    kwargs.pop("parent", None)
    assert len(args) <= len(self._fields), (type(self).__name__, args)
    vals = dict(zip(self._fields, args))
    for k, v in kwargs.items():
      assert k in self._fields and k not in vals, (type(self).__name__, k)
      vals[k] = v
    for f in self._fields:
      if f not in vals:
        if not hasattr(type(self), f):
          raise TypeError(f"{type(self).__name__}: missing field {f}")
        vals[f] = getattr(type(self), f)
      object.__setattr__(self, f, vals[f])
    self._counts = {}
    parent = _STACK[-1] if _STACK else None
    if parent is not None:
      if name is None:                       # flax: `<ClassName>_<k>`, numbered per class at construction time
        k = parent._counts.get(type(self).__name__, 0)
        parent._counts[type(self).__name__] = k + 1
        name = f"{type(self).__name__}_{k}"
      self._path = parent._path + (name,)
    else:
      self._path = ()
    self.name = name

  def param(self, name, init_fn, *init_args):
    shape = tuple(int(s) for s in init_args[0])
    full = "/".join(self._path + (name,))
    if _CTX["collect"] is not None:
      kind, scale = (init_fn.kind, init_fn.scale) if isinstance(init_fn, _Init) else ("?", 1.0)
      _CTX["collect"][full] = (shape, kind, float(scale))
      return wrap(np.zeros(shape))
    node = _CTX["params"]
    for p in self._path + (name,):
      node = node[p]
    a = np.asarray(node, dtype=np.float64)
    assert a.shape == shape, (full, a.shape, shape)
    return wrap(a)

  def apply(self, variables, *args, mutable=False, rngs=None, **kwargs):
    assert not _STACK, "shim: nested apply"
    _CTX["params"], _CTX["collect"] = variables["params"], None
    self._counts, self._path = {}, ()
    try:
      return self(*args, **kwargs)
    finally:
      _CTX["params"] = None


def linen_collect(module, *args, **kwargs):
  """Run `module` once with zero parameters and return {flax parameter path: (shape, init kind, init scale)} in creation
  order (what `model.init` would create)."""
  from collections import OrderedDict
  _CTX["collect"], _CTX["params"] = OrderedDict(), None
  module._counts, module._path = {}, ()
  try:
    module(*args, **kwargs)
    return _CTX["collect"]
  finally:
    _CTX["collect"] = None


class Dense(Module):
  features: int
  use_bias: bool = True
  kernel_init: object = None
  bias_init: object = None

  def __call__(self, x):
    w = self.param("kernel", self.kernel_init or _Init("lecun"), (np.shape(x)[-1], self.features))
    y = np.asarray(x) @ np.asarray(w)
    if self.use_bias:
      y = y + np.asarray(self.param("bias", self.bias_init or _Init("zeros"), (self.features,)))
    return wrap(y)


class Conv(Module):
  features: int
  kernel_size: tuple
  strides: tuple = None
  padding: object = "SAME"
  input_dilation: tuple = None
  kernel_dilation: tuple = None
  feature_group_count: int = 1
  use_bias: bool = True
  kernel_init: object = None
  bias_init: object = None

  def __call__(self, x):
    kh, kw = self.kernel_size
    w = self.param("kernel", self.kernel_init or _Init("lecun"), (kh, kw, np.shape(x)[-1] // self.feature_group_count, self.features))
    y = conv_general_dilated(x, w, self.strides or (1, 1), self.padding, rhs_dilation=self.kernel_dilation,
                             dimension_numbers=("NHWC", "HWIO", "NHWC"), feature_group_count=self.feature_group_count)
    if self.use_bias:
      y = y + np.asarray(self.param("bias", self.bias_init or _Init("zeros"), (self.features,)))
    return wrap(y)


class GroupNorm(Module):
  """flax.linen.GroupNorm (flax 0.3.x normalization.py): contiguous channel groups, statistics over (H, W, C/G) as
  E[x] and E[x^2] - E[x]^2, epsilon = 1e-6, per-channel scale and bias."""
  num_groups: int = 32
  group_size: int = None
  epsilon: float = 1e-6
  use_bias: bool = True
  use_scale: bool = True
  bias_init: object = None
  scale_init: object = None

  def __call__(self, x):
    x = np.asarray(x)
    C = x.shape[-1]
    G = self.num_groups if self.group_size is None else C // self.group_size
    assert C % G == 0
    xg = x.reshape(x.shape[:-1] + (G, C // G))
    red = tuple(range(1, x.ndim - 1)) + (x.ndim,)
    mean = xg.mean(axis=red, keepdims=True)
    mean2 = (xg * xg).mean(axis=red, keepdims=True)
    var = mean2 - mean * mean
    y = ((xg - mean) / np.sqrt(var + self.epsilon)).reshape(x.shape)
    if self.use_scale:
      y = y * np.asarray(self.param("scale", self.scale_init or _Init("ones"), (C,)))
    if self.use_bias:
      y = y + np.asarray(self.param("bias", self.bias_init or _Init("zeros"), (C,)))
    return wrap(y)


class Dropout(Module):
  rate: float

  def __call__(self, x, deterministic=False, rng=None):
    assert deterministic, "shim: dropout only in eval mode"
    return x


def avg_pool(x, window_shape, strides=None, padding="VALID"):
  import torch
  import torch.nn.functional as F
  assert tuple(window_shape) == (2, 2) and tuple(strides) == (2, 2)
  t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.float64))).permute(0, 3, 1, 2)
  return wrap(np.ascontiguousarray(F.avg_pool2d(t, 2, 2).permute(0, 2, 3, 1).numpy()))


def image_resize(x, shape, method):
  assert method == "nearest"
  x = np.asarray(x)
  fy, fx = shape[1] // x.shape[1], shape[2] // x.shape[2]
  return wrap(np.repeat(np.repeat(x, fy, axis=1), fx, axis=2))


# ---------------------------------------------------------------------------------------------------------------------
# module assembly
# ---------------------------------------------------------------------------------------------------------------------
class _Config:
  def __init__(self):
    self.values = {"jax_enable_x64": True}

  def read(self, k):
    return self.values[k]

  def update(self, k, v):
    self.values[k] = v


def _make_jnp(name):
  m = _module(name)
  special = dict(
      ndarray=np.ndarray, float32=np.float32, float64=np.float64, int32=np.int32, int64=np.int64, uint8=np.uint8,
      complex64=np.complex64, pi=np.pi, newaxis=None, inf=np.inf, e=np.e, bool_=np.bool_,
      asarray=lambda x, dtype=None: wrap(np.asarray(x, dtype=dtype)),
      array=lambda x, dtype=None, copy=True: wrap(np.array(x, dtype=dtype)),
      linalg=types.SimpleNamespace(inv=_np_wrapped(np.linalg.inv), cholesky=_np_wrapped(np.linalg.cholesky),
                                   norm=_np_wrapped(np.linalg.norm), svd=_np_wrapped(np.linalg.svd),
                                   eigh=_np_wrapped(np.linalg.eigh), det=_np_wrapped(np.linalg.det)),
      fft=types.SimpleNamespace(**{k: _np_wrapped(getattr(np.fft, k)) for k in ("fft", "ifft", "rfft", "irfft", "fftn", "ifftn")}),
  )
  m.__dict__.update(special)

  def _getattr(k):
    if k.startswith("__"):
      raise AttributeError(k)
    f = getattr(np, k)
    return _np_wrapped(f) if callable(f) and not isinstance(f, type) else f
  m.__getattr__ = _getattr
  return m


def _promote_dtypes_inexact(*args):
  return [wrap(np.asarray(a, dtype=np.float64)) for a in args]


def _struct_dataclass(cls):
  dc = dataclasses.dataclass(cls)
  dc.replace = lambda self, **kw: dataclasses.replace(self, **kw)
  return dc


def reset_modules():
  """Forget the reference's own top-level modules (both trees use the same names: sde_lib, sampling, utils, models ...)."""
  for k in list(sys.modules):
    top = k.split(".")[0]
    if top in ("sde_lib", "sampling", "deis", "utils", "models", "blur", "fft", "multistep", "losses", "datasets"):
      del sys.modules[k]
  for t in ("cld_jax", "blur_jax"):
    p = os.path.join(REF_ROOT, t)
    while p in sys.path:
      sys.path.remove(p)


def install(tree):
  """tree: 'cld_jax' | 'blur_jax'."""
  reset_modules()
  sys.dont_write_bytecode = True            # /root/reference is read-only; never try to write __pycache__ there
  jnp = _make_jnp("jax.numpy")
  lax = _module(
      "jax.lax", scan=scan, fori_loop=fori_loop, cond=cond, stop_gradient=lambda x: x, conv_general_dilated=conv_general_dilated,
      conv_transpose=conv_transpose, fft=lax_fft, slice_in_dim=slice_in_dim,
      rev=lambda x, dims: wrap(np.flip(np.asarray(x), axis=tuple(dims))),
      concatenate=lambda xs, dimension: wrap(np.concatenate([np.asarray(x) for x in xs], axis=dimension)),
      full=lambda shape, v, dtype=None: wrap(np.full(shape, v, dtype=dtype)),
      complex=lambda re, im: wrap(np.asarray(re) + 1j * np.asarray(im)))
  inits = _module("jax.nn.initializers", variance_scaling=variance_scaling, zeros=_Init("zeros"), ones=_Init("ones"),
                  normal=lambda stddev=1e-2: _Init("normal", stddev))
  jnn = _module("jax.nn", initializers=inits, softmax=softmax, swish=swish, silu=swish, sigmoid=lambda x: wrap(_sigmoid(x)),
                relu=lambda x: wrap(np.maximum(np.asarray(x), 0)), elu=lambda x: wrap(np.where(np.asarray(x) > 0, x, np.expm1(x))),
                leaky_relu=lambda x, negative_slope=0.01: wrap(np.where(np.asarray(x) >= 0, x, negative_slope * np.asarray(x))))
  rnd = _module("jax.random", PRNGKey=PRNGKey, split=split, fold_in=fold_in, normal=normal, uniform=uniform,
                multivariate_normal=multivariate_normal)
  image = _module("jax.image", resize=image_resize)
  ops = _module("jax.ops")
  xla_client = types.SimpleNamespace(FftType=FftType)
  jlib = _module("jax.lib", xla_client=xla_client)
  tree_util = _module("jax.tree_util", tree_map=tree_map, tree_leaves=tree_leaves)
  jax = _module("jax", numpy=jnp, lax=lax, nn=jnn, random=rnd, image=image, ops=ops, lib=jlib, tree_util=tree_util,
                jit=jit, vmap=vmap, pmap=pmap, config=_Config(), local_device_count=lambda: 1, device_count=lambda: 1,
                host_id=lambda: 0, process_index=lambda: 0, tree_map=tree_map, tree_multimap=tree_map,
                device_get=lambda x: x, devices=lambda: [0])
  src = _module("jax._src")
  src_np = _module("jax._src.numpy")
  lax_numpy = _make_jnp("jax._src.numpy.lax_numpy")
  lax_numpy._promote_dtypes_inexact = _promote_dtypes_inexact
  _module("jax._src.util", safe_zip=lambda *a: list(zip(*a)))
  _module("jax._src.numpy.util", _wraps=lambda fun, **kw: (lambda f: f))
  src.numpy, src_np.lax_numpy = src_np, lax_numpy
  jax._src = src

  linen = _module("flax.linen", Module=Module, compact=compact, Dense=Dense, Conv=Conv, GroupNorm=GroupNorm, Dropout=Dropout,
                  swish=swish, silu=swish, relu=jnn.relu, elu=jnn.elu, leaky_relu=jnn.leaky_relu, avg_pool=avg_pool,
                  softmax=softmax, initializers=inits)
  struct = _module("flax.struct", dataclass=_struct_dataclass)
  optim = _module("flax.optim", Optimizer=object)
  jax_utils = _module("flax.jax_utils", replicate=lambda t: tree_map(lambda x: wrap(np.asarray(x)[None]), t),
                      unreplicate=lambda t: tree_map(lambda x: wrap(np.asarray(x)[0]), t))
  training = _module("flax.training")
  ckpts = _stub_module("flax.training.checkpoints")
  training.checkpoints = ckpts
  ser = _stub_module("flax.serialization")
  _module("flax", linen=linen, struct=struct, optim=optim, jax_utils=jax_utils, training=training, serialization=ser)

  # jammy: pickle cache + git root (the generator always runs with used_cache=False; dumps are dropped)
  import tempfile
  scratch = tempfile.mkdtemp(prefix="refshim_cache_")
  jio = _module("jammy.io", mkdir=lambda p: os.makedirs(p, exist_ok=True), dump=lambda path, obj: None,
                load=lambda path: (_ for _ in ()).throw(FileNotFoundError(path)))
  jgit = _module("jammy.utils.git", git_rootdir=lambda sub="": os.path.join(scratch, sub))
  jutils = _module("jammy.utils", git=jgit)
  _module("jammy", io=jio, utils=jutils)

  class ConfigDict(dict):
    pass
  cdd = _module("ml_collections.config_dict.config_dict", ConfigDict=ConfigDict)
  cd = _module("ml_collections.config_dict", config_dict=cdd, ConfigDict=ConfigDict)
  _module("ml_collections", ConfigDict=ConfigDict, config_dict=cd)
  for stub in ("tensorflow", "wandb", "matplotlib", "matplotlib.pyplot", "tensorflow_gan", "tensorflow_hub", "tensorflow_datasets"):
    _stub_module(stub)
  sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
  # einops probes every framework found in sys.modules with isinstance(x, (tf.Tensor, tf.Variable)): give it real types
  sys.modules["tensorflow"].Tensor = type("Tensor", (), {})
  sys.modules["tensorflow"].Variable = type("Variable", (), {})

  sys.path.insert(0, os.path.join(REF_ROOT, tree))
  return jax
