"""Stand-in for blur_jax/blur.py (+ the lax.fft copy in blur_jax/fft.py): orthonormal 2-D DCT-II / DCT-III.

The reference evaluates the DCT through Makhoul's FFT trick (blur.py:11-97); here it is a dense 32x32
transform in shared memory (csrc/update.cu), which is exact up to fp32 rounding and cheaper than a
32-point FFT launch at this size.
"""
import numpy as np

from .. import _lib


def _run(x, forward):
  import torch
  _lib.require_cuda("batch_img_dct")
  is_np = not torch.is_tensor(x)
  xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda() if is_np else \
      x.detach().to(device="cuda", dtype=torch.float32).contiguous()
  if xd.dim() != 4 or xd.shape[1] != 32 or xd.shape[2] != 32:
    raise ValueError("batch_img_dct expects [B,32,32,C] (the reference hard-codes 32, blur_jax/sde_lib.py:24)")
  out = torch.empty_like(xd)
  st = torch.cuda.current_stream().cuda_stream
  _lib.check(_lib.lib().gddim_dct2d_32(xd.data_ptr(), out.data_ptr(), xd.shape[0], xd.shape[3], int(forward), st),
             "gddim_dct2d_32")
  return out.cpu().numpy() if is_np else out


def batch_img_dct(xs):
  """blur.py:99-102: NHWC images -> DCT-II coefficients per channel (norm='ortho')."""
  return _run(xs, True)


def batch_img_idct(ys):
  """blur.py:104-107."""
  return _run(ys, False)
