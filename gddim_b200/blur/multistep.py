"""Stand-in for blur_jax/multistep.py (scalar-coefficient DEIS from qsh-zh/deis; dead code in the reference,
kept for API parity).  `ab_step` runs on the GPU; `get_ab_eps_coef` is host numpy for
scalar SDEs exposing `psi(t_start, t_end)` and `eps_integrand(t)`."""
import numpy as np

from .. import _lib


def ab_step(x, ei_coef, new_eps, eps_pred):
  """multistep.py:94-98: x' = ei_coef[0] x + sum_i ei_coef[1+i] full_eps[i], full_eps = [new_eps, *eps_pred]."""
  import torch
  _lib.require_cuda("ab_step")
  is_np = not torch.is_tensor(x)

  def dev(a):
    if torch.is_tensor(a):
      return a.detach().to(device="cuda", dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()

  xd, ed, hd = dev(x), dev(new_eps), dev(eps_pred)
  coef = np.ascontiguousarray(np.asarray(ei_coef, dtype=np.float32)).ravel()
  n_hist = int(hd.shape[0])
  if coef.size != n_hist + 2 or tuple(hd.shape[1:]) != tuple(xd.shape):
    raise ValueError("ab_step: inconsistent shapes")
  x_out, h_out = torch.empty_like(xd), torch.empty_like(hd)
  st = torch.cuda.current_stream().cuda_stream
  _lib.check(_lib.lib().gddim_scalar_ab_step(xd.data_ptr(), coef.ctypes.data, ed.data_ptr(), hd.data_ptr(),
                                              x_out.data_ptr(), h_out.data_ptr(), n_hist, xd.numel(), st),
             "gddim_scalar_ab_step")
  if is_np:
    return x_out.cpu().numpy(), h_out.cpu().numpy()
  return x_out, h_out


def get_integrator_basis_fn(sde):
  def _worker(t_start, t_end, num_item):
    dt = (t_end - t_start) / num_item
    t_inter = np.linspace(t_start, t_end, num_item, endpoint=False)
    return sde.psi(t_inter, t_end) * sde.eps_integrand(t_inter), t_inter, dt
  return _worker


def single_poly_coef(t_val, ts_poly, coef_idx=0):
  ts_poly = np.asarray(ts_poly, np.float64)
  num, denum = t_val - ts_poly, ts_poly[coef_idx] - ts_poly
  num[coef_idx], denum[coef_idx] = 1.0, 1.0
  return np.prod(num) / np.prod(denum)


def get_one_coef_per_step_fn(sde):
  basis = get_integrator_basis_fn(sde)

  def _worker(t_start, t_end, ts_poly, coef_idx=0, num_item=10000):
    integrand, t_inter, dt = basis(t_start, t_end, num_item)
    poly = np.asarray([single_poly_coef(t, ts_poly, coef_idx) for t in t_inter])
    return np.sum(integrand * poly) * dt
  return _worker


def get_ab_eps_coef(sde, highest_order, timesteps, order):
  """multistep.py:68-92: [N, highest_order+2] rows (x_coef, eps_coef_0..), lower-order warm-up rows first."""
  timesteps = np.asarray(timesteps, np.float64)
  one = get_one_coef_per_step_fn(sde)

  def rows(ts, r):
    out = []
    for k in range(len(ts) - r - 1):
      t_start, t_end, ts_poly = ts[r + k], ts[r + k + 1], ts[k:k + r + 1]
      row = np.zeros(highest_order + 2)
      row[0] = sde.psi(t_start, t_end)
      for j, idx in enumerate(range(r, -1, -1)):
        row[1 + j] = one(t_start, t_end, ts_poly, idx)
      out.append(row)
    return out

  if order == 0:
    return np.asarray(rows(timesteps, 0))
  prev = get_ab_eps_coef(sde, highest_order, timesteps[:order + 1], order - 1)
  return np.concatenate([prev, np.asarray(rows(timesteps, order))], axis=0)
