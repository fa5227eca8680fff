"""`sample_data` of cld_jax/run_lib.py:674-731 on top of the B200 sampler: seed -> prior ->
sampler -> uint8 npz files, with the reference's key handling (PRNGKey(seed + 1), split per round), file naming and
resume-by-skipping.  Everything else of run_lib.py (training, FID) stays out of scope."""
import gc
import io
import logging
import os

import numpy as np

from . import checkpoint, jax_random, net
from .cld import sampling, sde_lib


def get_data_inverse_scaler(config):
  """cld_jax/datasets.py:34-40."""
  if config.data.centered:
    return lambda x: (x + 1.) / 2.
  return lambda x: x


def sample_data(config, ckpt_file, result_folder, is_continue=False, params=None, max_rounds=None):
  """Mirrors run_lib.sample_data.  `ckpt_file` is a Flax msgpack checkpoint of the training `State` (its
  `params_ema` is used, run_lib.py:707); alternatively pass `params` (Flax-named dict) directly."""
  logging.critical(f"sample data for {result_folder}")
  os.makedirs(result_folder, exist_ok=True)
  rng = jax_random.PRNGKey(config.seed + 1)                                    # run_lib.py:682
  inverse_scaler = get_data_inverse_scaler(config)
  rng, model_rng = jax_random.split(rng)                                       # :688
  score_model = net.ScoreNet(config, cld=True)
  if params is None:
    if not os.path.exists(ckpt_file):
      raise RuntimeError(f"{ckpt_file} not exist")                             # :708-709
    params = checkpoint.params_ema_from_checkpoint(ckpt_file)
  score_model.set_params(params)
  sde = sde_lib.from_config(config)
  sampling_fn = sampling.get_sampling_fn(config, sde, score_model, None, inverse_scaler)
  num_sampling_rounds = config.eval.num_samples // config.eval.batch_size + 1  # :704
  # The reference drives jax.local_device_count() GPUs from one process; here one process drives one GPU and the
  # launcher (torchrun) starts WORLD_SIZE of them.  The key handling is that of a WORLD_SIZE-device pmap: the round's
  # keys are split n_dev + 1 ways, key 1 draws the prior of ALL devices in one call (sampling.py:233-235) and this rank
  # keeps its slice, so the samples do not depend on how the devices are spread over processes; rank k gets
  # sample_rng[k] for its noise.  Every rank writes its own file (no collective on the path).
  n_dev = max(1, int(os.environ.get("WORLD_SIZE", "1")))
  rank = int(os.environ.get("RANK", "0"))
  if not 0 <= rank < n_dev:
    raise RuntimeError(f"RANK={rank} outside WORLD_SIZE={n_dev}")
  data_shape = sampling.get_data_shape(config)
  written = []
  for r in range(num_sampling_rounds):
    keys = jax_random.split(rng, n_dev + 1)                                    # :715
    rng, sample_rng = keys[0], keys[1:]
    f_sample = os.path.join(result_folder, f"samples_{r}.npz" if n_dev == 1 else f"samples_{r}_rank{rank}.npz")
    if os.path.exists(f_sample) and is_continue:
      logging.critical(f"SKIP!!! Already exists {f_sample}")
      continue
    if max_rounds is not None and len(written) >= max_rounds:
      break
    per_dev = config.eval.batch_size // n_dev
    u = None
    if n_dev > 1:
      u = sde.prior_sampling(sample_rng[0], (n_dev, per_dev) + tuple(data_shape))[rank:rank + 1]
    samples_org_x, samples_v, nfe_cnt = sampling_fn(sample_rng[rank:rank + 1], score_model, per_dev, u)
    samples_org_x, samples_v = np.asarray(samples_org_x), np.asarray(samples_v)
    samples_x = np.clip(samples_org_x * 255., 0, 255).astype(np.uint8)         # :723
    samples_x = samples_x.reshape((-1, config.data.image_size, config.data.image_size, config.data.num_channels))
    io_buffer = io.BytesIO()
    np.savez_compressed(io_buffer, samples=samples_x, nfe_cnt=nfe_cnt, samples_v=samples_v, samples_x=samples_org_x)
    with open(f_sample, "wb") as fout:
      fout.write(io_buffer.getvalue())
    written.append(f_sample)
    gc.collect()
  return written
