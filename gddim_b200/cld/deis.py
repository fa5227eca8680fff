"""Stand-in for cld_jax/deis.py (identical to blur_jax/deis.py): DEIS coefficient tables and the update op.

Tables come from the fp64 host code in libgddim_b200.so; `multistep_ab_step` runs the CUDA kernel
(csrc/update.cu) on torch.cuda tensors, or on numpy arrays through a device round trip.
"""
import numpy as np

from .. import _lib


def runge_kutta(x, t, dt, fn):
  """deis.py:5-17 (classic RK4; plain Python, used by callers outside the per-step path)."""
  g1 = fn(x, t)
  g2 = fn(x + g1 * dt / 2, t + dt / 2)
  g3 = fn(x + g2 * dt / 2, t + dt / 2)
  g4 = fn(x + g3 * dt, t + dt)
  return x + dt / 6 * (g1 + 2 * g2 + 2 * g3 + g4)


def single_poly_coef(t_val, ts_poly, coef_idx=0):
  """deis.py:30-36: Lagrange basis polynomial l_{coef_idx}(t_val) over the nodes ts_poly."""
  ts_poly = np.asarray(ts_poly, np.float64)
  num = t_val - ts_poly
  denum = ts_poly[coef_idx] - ts_poly
  num[coef_idx] = 1.0
  denum[coef_idx] = 1.0
  return np.prod(num) / np.prod(denum)


def vec_poly_coef(t_vals, ts_poly, coef_idx=0):
  return np.asarray([single_poly_coef(t, ts_poly, coef_idx) for t in np.asarray(t_vals)])


def get_ab_eps_coef(sde, highest_order, timesteps, order):
  """deis.py:71-95 -> [N, highest_order+1, 2, 2] (rows i < order run at order i; unused slots are zero)."""
  if highest_order < order:
    raise ValueError("highest_order must be >= order")
  full = np.asarray(sde.get_deis_coef(order, timesteps), np.float64)     # [N, order+3, 2, 2]
  out = np.zeros((full.shape[0], highest_order + 1, 2, 2), full.dtype)
  out[:, :order + 1] = full[:, 1:order + 2]
  return out.astype(np.float64 if getattr(sde, "x64", False) else np.float32)


def get_ab_eps_coef_order0(sde, highest_order, timesteps):
  return get_ab_eps_coef(sde, highest_order, timesteps, 0)


def get_am_eps_coef(sde, highest_order, timesteps, order):
  raise NotImplementedError("Adams-Moulton tables (deis.py:97-139) are unused by every reference sampler")


def multistep_ab_step(x, deis_coef, new_eps, eps_pred):
  """deis.py:141-151.  x, new_eps: (B, ..., d, 2); deis_coef: (order+3, 2, 2); eps_pred: (order+1, B, ..., d, 2).
  Returns (x_next, eps_pred_next) of the same array type as x."""
  import torch
  _lib.require_cuda("multistep_ab_step")
  is_np = not torch.is_tensor(x)

  def dev(a):
    if torch.is_tensor(a):
      return a.detach().to(device="cuda", dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()

  xd, ed, hd = dev(x), dev(new_eps), dev(eps_pred)
  coef = np.ascontiguousarray(np.asarray(deis_coef.detach().cpu() if torch.is_tensor(deis_coef) else deis_coef,
                                         dtype=np.float32))
  order = coef.shape[0] - 3
  if xd.shape[-1] != 2 or hd.shape[0] != order + 1 or tuple(hd.shape[1:]) != tuple(xd.shape) or ed.shape != xd.shape:
    raise ValueError("multistep_ab_step: inconsistent shapes")
  x_out, h_out = torch.empty_like(xd), torch.empty_like(hd)
  st = torch.cuda.current_stream().cuda_stream
  _lib.check(_lib.lib().gddim_multistep_ab_step(xd.data_ptr(), coef.ctypes.data, ed.data_ptr(), hd.data_ptr(),
                                                 x_out.data_ptr(), h_out.data_ptr(), order, xd.numel() // 2, st),
             "gddim_multistep_ab_step")
  if is_np:
    return x_out.cpu().numpy(), h_out.cpu().numpy()
  return x_out, h_out
