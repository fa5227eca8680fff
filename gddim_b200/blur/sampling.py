"""Stand-in for blur_jax/sampling.py: the order-0 (DDIM in DCT space) sampler of blurring diffusion.

`get_sampling_fn(config, sde, model, shape, inverse_scaler, is_p=True)` (sampling.py:11-40) ->
`psampler(prng, pstate, batch_size, u=None) -> (xs, nfe)`; unknown method -> bare RuntimeError (39-40).
"""
import ctypes as C

import numpy as np

from .. import _lib
from .. import net as _net
from ..cld.sampling import _affine_of, get_data_shape


def get_rev_ts(sde, ts_order, num_step):
  """sampling.py:42-51."""
  out = np.empty(num_step + 1)
  _lib.check(_lib.lib().gddim_rev_ts(float(sde.sampling_T), float(sde.sampling_eps), int(ts_order), int(num_step),
                                      out.ctypes.data))
  return out.astype(np.float32)


def get_sampling_fn(config, sde, model, shape, inverse_scaler, is_p=True):
  del shape
  if config.sampling.method.lower() == "order0":
    return get_order0_sampler(sde=sde, model=model, data_shape=get_data_shape(config),
                              ts_order=config.sampling.ts_order, nfe=config.sampling.nfe,
                              inverse_scaler=inverse_scaler, is_p=is_p)
  raise RuntimeError


class _Sampler:
  def __init__(self, sde, model, data_shape, ts_order, nfe, inverse_scaler, use_graph=True):
    self.sde, self.model, self.data_shape = sde, model, tuple(data_shape)
    self.ts_order, self.nfe = int(ts_order), int(nfe)
    self.inverse_scaler = inverse_scaler
    self.mul, self.add, self.affine = _affine_of(inverse_scaler)
    self.use_graph = use_graph
    self._h, self._gen, self._net = None, None, None

  def _destroy(self):
    if self._h is not None:
      _lib.lib().gddim_sampler_destroy(self._h)
      self._h = None
    self._gen = None

  def __del__(self):
    try:
      self._destroy()
    except Exception:
      pass

  def handle(self, net, batch):
    ctx = net.ensure(batch)
    # tied to ONE network context: reused only while that context lives (see cld/sampling.py _Sampler._current)
    if self._h is not None and self._net is net and self._gen == net.generation and _lib.lib().gddim_sampler_alive(self._h):
      return self._h
    self._destroy()
    cfg = _lib.SamplerCfg(kind=_lib.BLUR_ORDER0, nfe=self.nfe, deis_order=0, ts_order=self.ts_order, denoising=0,
                          mixed_score=0, use_graph=int(self.use_graph), x_mul=self.mul if self.affine else 1.0,
                          x_add=self.add if self.affine else 0.0)
    h = C.c_void_p()
    _lib.check(_lib.lib().gddim_sampler_create(ctx, C.byref(cfg), None, self.sde._h, C.byref(h)),
               "gddim_sampler_create")
    self._h, self._gen, self._net = h, net.generation, net
    net.register_sampler(self)
    return h

  def launch_count(self):
    return int(_lib.lib().gddim_sampler_launch_count(self._h)) if self._h is not None else 0

  def run(self, pstate, batch_size, u, trace=False):
    import torch
    _lib.require_cuda("sampler")
    net = _net.resolve_net(self.model, pstate, cld=False)
    shape = (batch_size,) + self.data_shape
    if tuple(u.shape) != shape:
      raise ValueError(f"u has shape {tuple(u.shape)}, expected {shape}")
    is_np = not torch.is_tensor(u)
    h = self.handle(net, batch_size)
    st = torch.cuda.current_stream().cuda_stream
    tr = torch.empty((self.nfe,) + shape, dtype=torch.float32, device="cuda") if trace else None
    trp = tr.data_ptr() if tr is not None else None
    if is_np:
      uh = np.ascontiguousarray(u, dtype=np.float32)
      x = np.empty(shape, np.float32)
      _lib.check(_lib.lib().gddim_sample(h, uh.ctypes.data, x.ctypes.data, None, batch_size, 1, trp, st), "gddim_sample")
    else:
      ud = u.detach().to(device="cuda", dtype=torch.float32).contiguous()
      x = torch.empty(shape, dtype=torch.float32, device="cuda")
      _lib.check(_lib.lib().gddim_sample(h, ud.data_ptr(), x.data_ptr(), None, batch_size, 0, trp, st), "gddim_sample")
    if not self.affine:
      x = self.inverse_scaler(x)
    if trace:
      return x, self.nfe, (tr.cpu().numpy() if is_np else tr)
    return x, self.nfe


def get_order0_sampler(sde, model, data_shape, ts_order, nfe, inverse_scaler, is_p=False):
  """sampling.py:53-90."""
  core = _Sampler(sde, model, data_shape, ts_order, nfe, inverse_scaler)

  def sampler(rng, state, batch_size, u=None, trace=False):
    if u is None:
      u = sde.prior_sampling(rng, (batch_size,) + tuple(data_shape))
    return core.run(state, batch_size, u, trace=trace)

  def psampler(prng, pstate, batch_size, u=None):
    if u is None:
      u = sde.prior_sampling(prng, (1, batch_size) + tuple(data_shape))
    if u.shape[0] != 1:
      raise ValueError("this process drives one GPU: the leading device axis of u must be 1")
    x, nfe = core.run(pstate, batch_size, u[0])
    return x[None], nfe

  fn = psampler if is_p else sampler
  fn.core = core
  return fn
