"""Flax msgpack checkpoint reader (no flax/jax needed): round trip through the documented container format and
name-compatibility with the library's parameter inventory."""
import numpy as np

from gddim_b200 import checkpoint, configs, net, params


def test_roundtrip_and_names(tmp_path):
  cfg = configs.tiny(configs.cld_accr_dcifar10(), nf=64, num_res_blocks=1)
  model = net.ScoreNet(cfg, cld=True)
  specs = model.specs()
  flat = params.generate(specs, seed=3, nondegenerate=True)
  path = tmp_path / "checkpoint_1"
  checkpoint.save_flax_checkpoint(path, flat, step=7)
  state = checkpoint.load_flax_checkpoint(path)
  assert state["step"] == 7 and set(state) >= {"params_ema", "model_state", "ema_rate"}
  tree = checkpoint.params_ema_from_checkpoint(path)
  assert "ResnetBlockBigGANpp_0" in tree and "kernel" in tree["ResnetBlockBigGANpp_0"]["Conv_0"]
  back = net.flatten_params(tree)
  assert set(back) == set(specs)
  for k, v in flat.items():
    np.testing.assert_array_equal(back[k], v)
  model.set_params(tree)                                 # nested tree accepted directly (pstate.params_ema)
  assert set(model.params) == set(specs)


def test_ext_types_and_chunked_arrays():
  import msgpack
  a = np.arange(12, dtype=np.float32).reshape(3, 4)
  chunked = {"__msgpack_chunked_array__": True, "shape": [3, 4], "chunks": {"0": a.ravel()[:5], "1": a.ravel()[5:]}}
  blob = checkpoint.to_bytes({"x": a, "s": np.float32(2.5), "c": chunked, "n": {"k": np.int32(3)}})
  out = checkpoint.restore_bytes(blob)
  np.testing.assert_array_equal(out["x"], a)
  np.testing.assert_array_equal(out["c"], a)
  assert out["s"] == np.float32(2.5) and out["n"]["k"] == 3
  # bfloat16 leaves (raw 2-byte words) widen to float32
  bf = (np.array([1.0, -2.5], np.float32).view(np.uint32) >> 16).astype(np.uint16)
  ext = msgpack.ExtType(1, msgpack.packb(([2], "bfloat16", bf.tobytes()), use_bin_type=True))
  out = checkpoint.restore_bytes(msgpack.packb({"w": ext}, use_bin_type=True))
  np.testing.assert_array_equal(out["w"], np.array([1.0, -2.5], np.float32))


def _flax_pack(tree):
  """An INDEPENDENT restatement of flax.serialization.msgpack_serialize (flax 0.3.x serialization.py) written with raw
  msgpack only -- it does not call gddim_b200.checkpoint: ndarray -> ExtType(1, packb((shape, dtype.name, bytes))),
  numpy scalar -> ExtType(3, ...), arrays above `max_chunk` bytes -> {'__msgpack_chunked_array__': True,
  'shape': {'0': d0, ...}, 'chunks': {'0': flat[a:b], ...}}, outer packb(default=..., strict_types=True)."""
  import msgpack

  def nd(a):
    return msgpack.packb((a.shape, a.dtype.name, a.tobytes("C")), use_bin_type=True)

  def default(x):
    if isinstance(x, np.ndarray):
      return msgpack.ExtType(1, nd(x))
    if isinstance(x, np.generic):
      return msgpack.ExtType(3, nd(np.asarray(x)))
    raise TypeError(type(x))

  def chunk(t, max_chunk):
    if isinstance(t, dict):
      return {k: chunk(v, max_chunk) for k, v in t.items()}
    if isinstance(t, np.ndarray) and t.size * t.dtype.itemsize > max_chunk:
      n = max(1, max_chunk // t.dtype.itemsize)
      flat = t.reshape(-1)
      return {"__msgpack_chunked_array__": True, "shape": {str(i): d for i, d in enumerate(t.shape)},
              "chunks": {str(i): flat[a:a + n] for i, a in enumerate(range(0, flat.size, n))}}
    return t
  return lambda max_chunk=2 ** 30: msgpack.packb(chunk(tree, max_chunk), default=default, strict_types=True)


def _hand_built_state(flat):
  """The document flax.training.checkpoints.save_checkpoint writes for models/utils.py:32-40 `State`:
  to_state_dict of the dataclass -- optimizer = {'target': params, 'state': {'step', 'param_states'}}."""
  nested = {}
  for k, v in flat.items():
    d = nested
    parts = k.split("/")
    for q in parts[:-1]:
      d = d.setdefault(q, {})
    d[parts[-1]] = v
  zeros = {}
  for k, v in flat.items():
    d = zeros
    parts = k.split("/")
    for q in parts[:-1]:
      d = d.setdefault(q, {})
    d[parts[-1]] = {"grad_ema": np.zeros_like(v), "grad_sq_ema": np.zeros_like(v)}
  return {"step": 130001, "optimizer": {"target": nested, "state": {"step": np.int32(130000), "param_states": zeros}},
          "lr": 2e-4, "model_state": {}, "ema_rate": 0.9999, "params_ema": nested, "rng": np.array([7, 9], np.uint32)}


def test_independent_flax_blob_with_chunked_leaves_and_directory_lookup(tmp_path):
  """SURVEY 8(f) N2 / VERDICT r1 item 9: a checkpoint NOT written by the repo's own writer."""
  cfg = configs.tiny(configs.cld_accr_dcifar10(), nf=64, num_res_blocks=1)
  model = net.ScoreNet(cfg, cld=True)
  flat = params.generate(model.specs(), seed=11, nondegenerate=True)
  blob = _flax_pack(_hand_built_state(flat))(max_chunk=64 * 1024)        # every conv kernel becomes a chunked array
  assert b"__msgpack_chunked_array__" in blob
  d = tmp_path / "checkpoints"
  d.mkdir()
  (d / "checkpoint_2").write_bytes(b"stale")
  (d / "checkpoint_10").write_bytes(blob)                                # natural order: 10 is later than 2
  (d / "checkpoint_11tmp").write_bytes(b"partial")
  assert checkpoint.resolve_checkpoint_path(str(d)).endswith("checkpoint_10")
  assert checkpoint.resolve_checkpoint_path(str(d), step=2).endswith("checkpoint_2")
  state = checkpoint.load_flax_checkpoint(str(d))
  assert state["step"] == 130001 and state["optimizer"]["state"]["step"] == 130000
  back = net.flatten_params(state["params_ema"])
  assert list(sorted(back)) == list(sorted(flat))
  for k, v in flat.items():
    assert back[k].dtype == np.float32 and back[k].shape == v.shape
    np.testing.assert_array_equal(back[k], v)
  np.testing.assert_array_equal(state["rng"], np.array([7, 9], np.uint32))


def _write_independent_checkpoint(path, flat):
  with open(path, "wb") as f:
    f.write(_flax_pack(_hand_built_state(flat))(max_chunk=256 * 1024))
