"""Deterministic synthetic parameters in Flax layout, keyed by Flax parameter names.

The reference initialises through flax `model.init` (cld_jax/models/utils.py:109-125) with
`variance_scaling(scale, 'fan_avg', 'uniform')` kernels (layers.py:60-63), zero biases, GroupNorm
scale=1 / bias=0 and Fourier `W ~ N(0, fourier_scale^2)` (layerspp.py:40).  JAX's threefry stream is
not reproduced; instead each parameter draws from `numpy.random.default_rng([seed, crc32(name)])`, so
any consumer (library, oracle, a checkpoint converter) that knows the name and shape gets the same
numbers regardless of iteration order.

`nondegenerate=True` replaces every `init_scale=0 -> 1e-10` kernel by scale 1 and perturbs biases and
GroupNorm affine parameters, so that parity tests exercise every term (with the faithful init the
network output is ~1e-10 and parity would be vacuous, SURVEY.md 7 "Degenerate random init").
"""
import zlib

import numpy as np


def _fans(shape):
  if len(shape) == 1:
    return shape[0], shape[0]
  rf = 1
  for s in shape[:-2]:
    rf *= s
  return shape[-2] * rf, shape[-1] * rf


def generate_one(name, shape, kind, scale, seed=1234, nondegenerate=False):
  rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
  shape = tuple(int(s) for s in shape)
  if kind == "vs":
    if nondegenerate and scale < 1e-6:
      scale = 1.0
    fan_in, fan_out = _fans(shape)
    lim = np.sqrt(3.0 * scale / ((fan_in + fan_out) / 2.0))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)
  if kind == "normal":
    return (rng.standard_normal(shape) * scale).astype(np.float32)
  if kind == "zeros":
    return rng.uniform(-0.1, 0.1, size=shape).astype(np.float32) if nondegenerate else np.zeros(shape, np.float32)
  if kind == "ones":
    base = np.ones(shape, np.float32)
    return base + rng.uniform(-0.1, 0.1, size=shape).astype(np.float32) if nondegenerate else base
  raise ValueError(kind)


def generate(specs, seed=1234, nondegenerate=False):
  """specs: mapping name -> (shape, kind, scale)  ->  dict name -> float32 array."""
  return {n: generate_one(n, s, k, sc, seed, nondegenerate) for n, (s, k, sc) in specs.items()}


def count(specs):
  return int(sum(int(np.prod(s)) for s, _, _ in specs.values()))
