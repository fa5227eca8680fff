"""JAX-compatible PRNG for the sampler's prior: threefry2x32 keys, `split`, `random_bits`,
`uniform`, `normal` as jax 0.2.8 computes them (jax/_src/random.py, jax/_src/prng.py), in numpy.

  PRNGKey(seed)            -> uint32[2] = (seed >> 32, seed & 0xffffffff)
  split(key, num)          -> threefry_2x32(key, iota(2 num)).reshape(num, 2)
  random_bits(key, shape)  -> threefry_2x32(key, iota(size))       (counter halves are the two Threefry words)
  uniform                  -> (bits >> 9 | 0x3f800000).view(f32) - 1, scaled to [minval, maxval)
  normal                   -> sqrt(2) * erfinv(uniform(nextafter(-1, 0), 1))   (XLA's fp32 erfinv polynomial, Giles 2010)

Pinned by the known answers printed in the JAX documentation (tests/test_jax_random.py).  The last ulp of `normal`
can differ from a given XLA backend (log / polynomial evaluation order), the bit stream cannot.
"""
import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def PRNGKey(seed):
  seed = int(seed)
  return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def _rotl(x, d):
  return (x << _U32(d)) | (x >> _U32(32 - d))


def threefry_2x32(key, count):
  """jax._src.prng.threefry_2x32: the count array is split in two halves that form the two 32-bit words."""
  key = np.asarray(key, _U32)
  count = np.asarray(count, _U32).ravel()
  odd = count.size % 2
  if odd:
    count = np.concatenate([count, np.zeros(1, _U32)])
  x0, x1 = np.split(count.copy(), 2)
  ks = [key[0], key[1], key[0] ^ key[1] ^ _U32(0x1BD11BDA)]
  with np.errstate(over="ignore"):
    x0 = x0 + ks[0]
    x1 = x1 + ks[1]
    for r in range(5):
      for d in _ROT[r % 2]:
        x0 = x0 + x1
        x1 = _rotl(x1, d)
        x1 = x0 ^ x1
      x0 = x0 + ks[(r + 1) % 3]
      x1 = x1 + ks[(r + 2) % 3] + _U32(r + 1)
  out = np.concatenate([x0, x1])
  return out[:-1] if odd else out


def split(key, num=2):
  return threefry_2x32(key, np.arange(num * 2, dtype=_U32)).reshape(num, 2)


def random_bits(key, shape):
  size = int(np.prod(shape)) if len(shape) else 1
  return threefry_2x32(key, np.arange(size, dtype=_U32)).reshape(shape)


def uniform(key, shape, minval=0.0, maxval=1.0):
  bits = random_bits(key, shape)
  floats = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
  minval, maxval = np.float32(minval), np.float32(maxval)
  return np.maximum(minval, floats * (maxval - minval) + minval).astype(np.float32)


def erfinv_f32(x):
  """XLA's ErfInv for fp32 (M. Giles, "Approximating the erfinv function"), evaluated in float32."""
  x = np.asarray(x, np.float32)
  w = -np.log((np.float32(1.0) - x) * (np.float32(1.0) + x)).astype(np.float32)
  lt = w < np.float32(5.0)
  w1 = (w - np.float32(2.5)).astype(np.float32)
  w2 = (np.sqrt(np.maximum(w, np.float32(0))) - np.float32(3.0)).astype(np.float32)
  c1 = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
        -0.00417768164, 0.246640727, 1.50140941]
  c2 = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047,
        1.00167406, 2.83297682]
  p1 = np.full_like(x, np.float32(c1[0]))
  for c in c1[1:]:
    p1 = (np.float32(c) + p1 * w1).astype(np.float32)
  p2 = np.full_like(x, np.float32(c2[0]))
  for c in c2[1:]:
    p2 = (np.float32(c) + p2 * w2).astype(np.float32)
  return (np.where(lt, p1, p2) * x).astype(np.float32)


def normal(key, shape):
  lo = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
  u = uniform(key, shape, lo, 1.0)
  return (np.float32(np.sqrt(2.0)) * erfinv_f32(u)).astype(np.float32)


def cld_prior(key, shape, m_inv):
  """CLD.prior_sampling (cld_jax/sde_lib.py:270-274) with a jax PRNGKey: x ~ N(0,1), v ~ N(0,1)/sqrt(m_inv)."""
  x_rng, v_rng = split(np.asarray(key, _U32), 2)
  xs = normal(x_rng, tuple(shape))
  vs = (normal(v_rng, tuple(shape)) / np.float32(np.sqrt(m_inv))).astype(np.float32)
  return np.stack([xs, vs], axis=-1)
