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
