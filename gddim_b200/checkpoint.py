"""Flax checkpoint reader: `flax.training.checkpoints.save_checkpoint` writes the `State`
dataclass of cld_jax/models/utils.py:32-40 with `flax.serialization.to_bytes`, i.e. one msgpack document whose
leaves are numpy arrays in msgpack ExtType 1 = packb((shape, dtype_name, raw_bytes)).  This module restores such a
file to nested dicts of numpy arrays without flax/jax, so that `state['params_ema']` can be handed to
`ScoreNet.set_params` / passed as `pstate` (run_lib.py:707 `checkpoints.restore_checkpoint`).

The writer exists for round-trip tests and for exporting synthetic parameters in the same container format.
Format reference: flax/serialization.py (`_ndarray_to_bytes`, `_msgpack_ext_pack`, chunked arrays above 2**30 B).
"""
import numpy as np

_EXT_NDARRAY, _EXT_NATIVE_COMPLEX, _EXT_NPSCALAR = 1, 2, 3
_CHUNK_KEY = "__msgpack_chunked_array__"


def _msgpack():
  import msgpack
  return msgpack


def _ndarray_from_bytes(data):
  shape, dtype_name, buf = _msgpack().unpackb(data, raw=True)
  dtype_name = dtype_name.decode() if isinstance(dtype_name, bytes) else dtype_name
  if dtype_name == "bfloat16":          # stored as raw 2-byte words: widen to float32
    a = np.frombuffer(buf, dtype=np.uint16).astype(np.uint32) << 16
    return a.view(np.float32).reshape(shape)
  return np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape).copy()


def _ext_hook(code, data):
  if code == _EXT_NDARRAY:
    return _ndarray_from_bytes(data)
  if code == _EXT_NPSCALAR:
    return _ndarray_from_bytes(data)[()]
  if code == _EXT_NATIVE_COMPLEX:
    re, im = _msgpack().unpackb(data)
    return complex(re, im)
  return _msgpack().ExtType(code, data)


def _unchunk(tree):
  if isinstance(tree, dict):
    if tree.get(_CHUNK_KEY):
      # flax.serialization._chunk: 'shape' and 'chunks' are _tuple_to_dict(...) = {'0': v0, '1': v1, ...}
      sh = tree["shape"]
      shape = tuple(int(sh[k]) for k in sorted(sh, key=int)) if isinstance(sh, dict) else tuple(int(v) for v in sh)
      chunks = tree["chunks"]
      seq = [chunks[k] for k in sorted(chunks, key=int)] if isinstance(chunks, dict) else list(chunks)
      parts = [np.asarray(c).ravel() for c in seq]
      return np.concatenate(parts).reshape(shape)
    return {k: _unchunk(v) for k, v in tree.items()}
  return tree


def _decode_keys(tree):
  if isinstance(tree, dict):
    return {(k.decode() if isinstance(k, bytes) else str(k)): _decode_keys(v) for k, v in tree.items()}
  return tree


def restore_bytes(data):
  """bytes of a Flax msgpack checkpoint -> nested dict of numpy arrays / python scalars."""
  tree = _msgpack().unpackb(data, ext_hook=_ext_hook, raw=True, strict_map_key=False)
  return _unchunk(_decode_keys(tree))


def _natural_key(name):
  import re
  return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", name)]


def resolve_checkpoint_path(ckpt_dir, step=None, prefix="checkpoint_"):
  """flax.training.checkpoints.restore_checkpoint path rules: a file is taken as is; a directory yields
  `<dir>/<prefix><step>` or, without `step`, the latest `<prefix>*` file in natural order (tmp files skipped)."""
  import os
  if not os.path.isdir(ckpt_dir):
    return ckpt_dir
  if step is not None:
    return os.path.join(ckpt_dir, f"{prefix}{step}")
  names = sorted((n for n in os.listdir(ckpt_dir) if n.startswith(prefix) and not n.endswith("tmp")), key=_natural_key)
  if not names:
    raise FileNotFoundError(f"no '{prefix}*' checkpoint in {ckpt_dir}")
  return os.path.join(ckpt_dir, names[-1])


def load_flax_checkpoint(path, step=None, prefix="checkpoint_"):
  with open(resolve_checkpoint_path(path, step, prefix), "rb") as f:
    return restore_bytes(f.read())


def params_ema_from_checkpoint(path):
  """The tree the samplers read (sampling.py:216 `state.params_ema`), as nested dicts keyed by Flax module names."""
  state = load_flax_checkpoint(path)
  if "params_ema" not in state:
    raise KeyError(f"{path}: no 'params_ema' entry (keys: {sorted(state)})")
  return state["params_ema"]


def _ext_pack(x):
  if isinstance(x, np.ndarray):
    return _msgpack().ExtType(_EXT_NDARRAY, _msgpack().packb((list(x.shape), x.dtype.name, x.tobytes()), use_bin_type=True))
  if isinstance(x, np.generic):
    a = np.asarray(x)
    return _msgpack().ExtType(_EXT_NPSCALAR, _msgpack().packb((list(a.shape), a.dtype.name, a.tobytes()), use_bin_type=True))
  raise TypeError(f"cannot serialise {type(x)}")


def to_bytes(tree):
  return _msgpack().packb(tree, default=_ext_pack, strict_types=True, use_bin_type=True)


def nest(flat):
  """{'A/B/kernel': arr} -> {'A': {'B': {'kernel': arr}}}  (Flax params tree shape)."""
  out = {}
  for name, v in flat.items():
    d = out
    parts = name.split("/")
    for p in parts[:-1]:
      d = d.setdefault(p, {})
    d[parts[-1]] = v
  return out


def save_flax_checkpoint(path, params_ema_flat, step=0, ema_rate=0.9999):
  """Writes a `State`-shaped document (only the fields the samplers read are meaningful)."""
  tree = {"step": int(step), "ema_rate": float(ema_rate), "params_ema": nest(params_ema_flat), "model_state": {},
          "optimizer": {}, "lr": 0.0, "rng": np.zeros(2, np.uint32)}
  with open(path, "wb") as f:
    f.write(to_bytes(tree))
