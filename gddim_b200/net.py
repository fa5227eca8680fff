"""Model adapter: the score-network handle and the eps functions built on it.

Stands in for cld_jax/models/utils.py:32-66,109-182 (State, init_model, get_model_fn, get_eps_fn) and
blur_jax/models/utils.py:104-160 (get_eps_fn, get_yeps_fn).  The network itself lives in
libgddim_b200.so (csrc/unet.cpp); torch is used only to own device memory and streams.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _lib
from . import params as _params

_KINDS = {0: "vs", 1: "zeros", 2: "ones", 3: "normal"}


def model_cfg_from_config(config, cld=True):
  """Reads the attribute paths ncsnpp.py:43-67 reads (duck-typed config)."""
  m, d = config.model, config.data
  if str(getattr(m, "name", "ncsnpp")).lower() != "ncsnpp":
    raise ValueError(f"model {m.name} is not supported (only 'ncsnpp')")
  if str(m.resblock_type).lower() != "biggan":
    raise ValueError("only resblock_type='biggan' is supported")
  if str(m.progressive).lower() != "none":
    raise ValueError("only progressive='none' is supported")
  if str(m.nonlinearity).lower() != "swish":
    raise ValueError("only nonlinearity='swish' is supported")
  if getattr(m, "scale_by_sigma", False):
    raise ValueError("scale_by_sigma is not supported")
  pin = str(m.progressive_input).lower()
  if pin not in ("none", "residual"):
    raise ValueError(f"progressive_input={pin!r} is not supported")
  emb = str(m.embedding_type).lower()
  if emb not in ("fourier", "positional"):
    raise ValueError(f"embedding type {emb} unknown.")
  c = _lib.ModelCfg()
  c.image_size, c.data_channels, c.state_mult = int(d.image_size), int(d.num_channels), 2 if cld else 1
  c.nf, c.n_levels, c.num_res_blocks = int(m.nf), len(m.ch_mult), int(m.num_res_blocks)
  for i, v in enumerate(m.ch_mult):
    c.ch_mult[i] = int(v)
  c.n_attn = len(m.attn_resolutions)
  for i, v in enumerate(m.attn_resolutions):
    c.attn_resolutions[i] = int(v)
  c.fir, c.skip_rescale = int(bool(m.fir)), int(bool(m.skip_rescale))
  c.progressive_input = 1 if pin == "residual" else 0
  c.embedding_type = 0 if emb == "fourier" else 1
  c.conditional, c.centered = int(bool(m.conditional)), int(bool(d.centered))
  return c


def flatten_params(tree, prefix=""):
  """Flax-style nested dict (or FrozenDict) -> flat {"A/B/kernel": array}."""
  out = {}
  for k, v in tree.items():
    name = f"{prefix}{k}"
    if hasattr(v, "items"):
      out.update(flatten_params(v, name + "/"))
    else:
      out[name] = np.asarray(v, dtype=np.float32)
  return out


class ScoreNet:
  """Handle of one NCSN++/DDPM++ network on one GPU (the `model` / `pstate` stand-in)."""

  def __init__(self, config, cld=True, device=None, precise=False):
    """precise=True: convolution weights as fp16 (hi, lo) pairs, two K passes per convolution (parity mode)."""
    self.config, self.cld, self.precise = config, cld, bool(precise)
    self.device = device
    self._cfg = model_cfg_from_config(config, cld)
    self._ctx = None
    self._max_batch = 0
    self._params = None
    self._gemm_impl = 0
    self._specs = None
    self._src = None          # the pstate.params_ema object the current parameters came from
    self.generation = 0       # bumped whenever the device context is (re)created: samplers compare it, not the ctx address
    self._samplers = []       # weak references to the Python samplers built on the current context

  # -- parameters -----------------------------------------------------------------------------------
  def specs(self):
    """Ordered name -> (shape, kind, scale), Flax naming/layouts (from the library's own walk)."""
    if self._specs is None:
      L = _lib.lib()
      ctx = C.c_void_p()
      _lib.check(L.gddim_ctx_create(0, C.byref(self._cfg), 1, C.byref(ctx)), "gddim_ctx_create")
      try:
        self._specs = _read_specs(ctx)
      finally:
        L.gddim_ctx_destroy(ctx)
    return self._specs

  def plan(self, batch=1):
    """[(tag, kind)] of the static launch plan for `batch` images per call (host-side walk only: works without a GPU)."""
    L = _lib.lib()
    ctx = C.c_void_p()
    _lib.check(L.gddim_ctx_create(0, C.byref(self._cfg), int(batch), C.byref(ctx)), "gddim_ctx_create")
    try:
      out = []
      buf, kind = C.create_string_buffer(256), C.c_int()
      for i in range(L.gddim_ctx_plan_size(ctx)):
        _lib.check(L.gddim_ctx_plan_op(ctx, i, buf, 256, C.byref(kind)), "gddim_ctx_plan_op")
        out.append((buf.value.decode(), kind.value))
      return out
    finally:
      L.gddim_ctx_destroy(ctx)

  def set_params(self, params):
    """params: flat or nested mapping of float arrays in Flax layout (e.g. state.params_ema)."""
    flat = flatten_params(params) if any(hasattr(v, "items") for v in params.values()) else \
        {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in params.items()}
    self._params = flat
    self._destroy()

  def init_params(self, seed=1234, nondegenerate=False):
    self.set_params(_params.generate(self.specs(), seed=seed, nondegenerate=nondegenerate))
    return self._params

  @property
  def params(self):
    return self._params

  def set_gemm_impl(self, impl):
    """0 = tcgen05 kernels (default); 1 = CUDA-core reference kernels (validation only)."""
    self._gemm_impl = int(impl)
    if self._ctx is not None:
      _lib.check(_lib.lib().gddim_ctx_set_gemm_impl(self._ctx, self._gemm_impl))

  # -- context ------------------------------------------------------------------------------------------
  def register_sampler(self, sampler):
    import weakref
    self._samplers = [r for r in self._samplers if r() is not None and r() is not sampler] + [weakref.ref(sampler)]

  def _destroy(self):
    if self._ctx is not None:
      # samplers built on this context go first (the library would orphan them anyway, gddim_ctx_destroy)
      for r in self._samplers:
        smp = r()
        if smp is not None:
          smp._destroy()
      self._samplers = []
      _lib.lib().gddim_ctx_destroy(self._ctx)
      self._ctx = None
      self._max_batch = 0

  def __del__(self):
    try:
      self._destroy()
    except Exception:
      pass

  def ensure(self, batch):
    """Creates (or re-creates, if `batch` grew) the device context: packs weights, plans the workspace."""
    if self._ctx is not None and batch <= self._max_batch:
      return self._ctx
    _lib.require_cuda("ScoreNet")
    if self._params is None:
      raise RuntimeError("ScoreNet has no parameters: call set_params(...) or init_params(...)")
    import torch
    self._destroy()
    L = _lib.lib()
    dev = torch.cuda.current_device() if self.device is None else int(self.device)
    ctx = C.c_void_p()
    _lib.check(L.gddim_ctx_create_ex(dev, C.byref(self._cfg), int(batch), 1 if self.precise else 0, C.byref(ctx)),
               "gddim_ctx_create")
    try:
      for name in _read_specs(ctx):
        if name not in self._params:
          raise KeyError(f"missing parameter {name}")
        a = np.ascontiguousarray(self._params[name], dtype=np.float32)
        _lib.check(L.gddim_param_set(ctx, name.encode(), a.ctypes.data, a.size), f"gddim_param_set({name})")
      _lib.check(L.gddim_ctx_finalize(ctx), "gddim_ctx_finalize")
      _lib.check(L.gddim_ctx_set_gemm_impl(ctx, self._gemm_impl))
    except Exception:
      L.gddim_ctx_destroy(ctx)
      raise
    self._ctx, self._max_batch = ctx, int(batch)
    self.generation += 1
    return ctx

  @property
  def net_channels(self):
    return self._cfg.data_channels * self._cfg.state_mult

  def launch_count(self):
    return int(_lib.lib().gddim_ctx_launch_count(self._ctx)) if self._ctx is not None else 0

  def workspace_bytes(self):
    return int(_lib.lib().gddim_ctx_workspace_bytes(self._ctx)) if self._ctx is not None else 0

  # -- per-op timing (bench.py's roofline leg) ----------------------------------------------------------
  def set_profile(self, on):
    _lib.check(_lib.lib().gddim_ctx_set_profile(self._ctx, int(bool(on))))

  def get_profile(self):
    """-> (ms_by_kind dict, gemm_flops, gemm_launches) accumulated since set_profile(True)."""
    ms = (C.c_double * 8)()
    fl, nl = C.c_double(), C.c_longlong()
    _lib.check(_lib.lib().gddim_ctx_get_profile(self._ctx, ms, C.byref(fl), C.byref(nl)))
    names = ["stem", "groupnorm", "conv_gemm", "head", "im2col", "transpose_v", "attention", "softmax_rows"]
    return {n: ms[i] for i, n in enumerate(names)}, fl.value, nl.value

  def get_profile_norm_bytes(self):
    """Algorithmic HBM bytes of every GroupNorm op over the profiled forwards (gddim_ctx_get_profile_hbm)."""
    b = C.c_double()
    _lib.check(_lib.lib().gddim_ctx_get_profile_hbm(self._ctx, C.byref(b)))
    return b.value

  def dump_profile(self, path):
    _lib.check(_lib.lib().gddim_ctx_dump_profile(self._ctx, str(path).encode()))

  # -- forward --------------------------------------------------------------------------------------------
  def forward(self, x, t):
    """x: [B,H,W,Cnet] (numpy or torch.cuda float32), t: diffusion time shared by the batch
    (labels = 999 t inside, models/utils.py:172).  Returns the same array type."""
    import torch
    is_np = not torch.is_tensor(x)
    xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda() if is_np else \
        x.detach().to(torch.float32).contiguous()
    B = int(xd.shape[0])
    ctx = self.ensure(B)
    out = torch.empty_like(xd)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().gddim_unet_forward(ctx, xd.data_ptr(), float(t), out.data_ptr(), B, st), "gddim_unet_forward")
    if is_np:
      return out.cpu().numpy()
    return out

  __call__ = forward


def _read_specs(ctx):
  L = _lib.lib()
  n = L.gddim_param_count(ctx)
  specs = OrderedDict()
  buf = C.create_string_buffer(256)
  shape = (C.c_int * 4)()
  ndim, kind, scale = C.c_int(), C.c_int(), C.c_float()
  for i in range(n):
    _lib.check(L.gddim_param_spec(ctx, i, buf, 256, C.byref(shape), C.byref(ndim), C.byref(kind), C.byref(scale)))
    specs[buf.value.decode()] = (tuple(shape[j] for j in range(ndim.value)), _KINDS[kind.value], float(scale.value))
  return specs


class State:
  """Minimal stand-in for models/utils.py:32-40 `State` (only the fields the samplers read)."""

  def __init__(self, params_ema, model_state=None, step=0):
    self.params_ema, self.model_state, self.step = params_ema, model_state, step


def init_model(rng, config, cld=True, nondegenerate=False, precise=False):
  """models/utils.py:109-125: returns (model, init_model_state, initial_params).  `rng` is an int seed or
  None (JAX's threefry stream is not reproduced; see gddim_b200/params.py).  precise: see ScoreNet."""
  model = ScoreNet(config, cld=cld, precise=precise)
  seed = 1234 if rng is None else int(np.asarray(rng).ravel()[-1])
  params = model.init_params(seed=seed, nondegenerate=nondegenerate)
  return model, None, params


def resolve_net(model, pstate, cld):
  """Accepts what run_lib passes as (model, pstate): our handle, or a state carrying `params_ema`."""
  net = model if isinstance(model, ScoreNet) else (pstate if isinstance(pstate, ScoreNet) else None)
  if net is None:
    raise TypeError("model (or pstate) must be a gddim_b200.net.ScoreNet handle")
  if net.cld != cld:
    raise ValueError("ScoreNet was built for the other SDE family (state_mult mismatch)")
  p = None
  if isinstance(pstate, dict) and "params_ema" in pstate:
    p = pstate["params_ema"]
  elif hasattr(pstate, "params_ema"):
    p = pstate.params_ema
  if p is not None and p is not net._src and p is not net._params:
    net.set_params(p)
    net._src = p
  return net


def get_eps_fn(sde, model, params=None, states=None, train=False, continuous=True):
  """cld models/utils.py:168-182: eps_fn(x, t) on reference-layout x [B,H,W,C,2]; t scalar or [B] (all equal)."""
  if train:
    raise NotImplementedError("training mode (dropout) is outside the sampling hot path")
  import torch
  from .cld import sde_lib as _sl

  def eps_fn(x, t):
    tt = float(np.asarray(t.detach().cpu() if torch.is_tensor(t) else t).ravel()[0])
    is_np = not torch.is_tensor(x)
    xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda() if is_np else x.float().contiguous()
    C_ = xd.shape[-2]
    net_in = torch.cat([xd[..., 0], xd[..., 1]], dim=-1).contiguous()              # 'b ... d g -> b ... (g d)'
    out = model.forward(net_in, tt)
    out = torch.stack([out[..., :C_], out[..., C_:]], dim=-1)                      # inverse rearrange
    if sde.mixed_score:
      inv_r = torch.as_tensor(np.asarray(sde.s_invR(tt), dtype=np.float32), device=out.device)
      u0 = xd.clone()
      u0[..., 0] = 0.0
      out = out + torch.einsum("ij,...j->...i", inv_r, u0)
    return out.cpu().numpy() if is_np else out
  return eps_fn


def get_blur_eps_fn(sde, model, params=None, states=None, train=False, continuous=False, return_state=False):
  """blur models/utils.py:141-153: eps_fn(x, t) in pixel space (labels = sde.encode_t(t) = 999 t)."""
  if train:
    raise NotImplementedError("training mode (dropout) is outside the sampling hot path")
  import torch

  def eps_fn(x, t, rng=None):
    tt = float(np.asarray(t.detach().cpu() if torch.is_tensor(t) else t).ravel()[0])
    out = model.forward(sde.encode_x(x), tt)
    return (out, states) if return_state else out
  return eps_fn


def get_yeps_fn(sde, model, params=None, states=None, train=False, continuous=False):
  """blur models/utils.py:155-160: DCT o net o IDCT."""
  xeps_fn = get_blur_eps_fn(sde, model, params, states, train, continuous)

  def eps_fn(y, t, rng=None):
    return sde.x2y(xeps_fn(sde.y2x(y), t, rng))
  return eps_fn
