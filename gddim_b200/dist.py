"""Multi-GPU plumbing: one process per GPU (torchrun), the batch axis sharded across ranks, parameters
broadcast once from rank 0 (NCCL over NVLink on GPUs, gloo on CPU for tests), no per-step collective.

Mirrors what `jax.pmap(sampler)` + `flax.jax_utils.replicate(state)` do in the reference
(cld_jax/sampling.py:232-237, run_lib.py:711): replicated parameters, independent per-device trajectories.
"""
import os

import numpy as np


def env_world():
  return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend=None):
  """Initialises torch.distributed from the torchrun environment (no-op for world size 1)."""
  import torch
  import torch.distributed as dist
  rank, local_rank, world = env_world()
  if world == 1 or dist.is_initialized():
    return rank, local_rank, world
  if backend is None:
    backend = "nccl" if torch.cuda.is_available() else "gloo"
  if backend == "nccl":
    torch.cuda.set_device(local_rank)
  os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
  dist.init_process_group(backend=backend, rank=rank, world_size=world)
  return rank, local_rank, world


def shard_bounds(global_batch, rank, world):
  """Contiguous even split of the leading axis: the reference layout (n_dev, B, ...) flattened."""
  if global_batch % world != 0:
    raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
  per = global_batch // world
  return rank * per, (rank + 1) * per


def shard(array, rank, world):
  lo, hi = shard_bounds(array.shape[0], rank, world)
  return array[lo:hi]


def broadcast_params(params, src=0, device=None):
  """params: {name: float32 array} on `src` (other ranks pass the same names/shapes with any content, or the
  spec mapping name -> shape).  One flat fp32 buffer, one broadcast.  Returns {name: np.ndarray}."""
  import torch
  import torch.distributed as dist
  if not dist.is_initialized() or dist.get_world_size() == 1:
    return params
  names = sorted(params)
  shapes = [tuple(np.shape(params[n])) if not isinstance(params[n], tuple) else params[n] for n in names]
  sizes = [int(np.prod(s)) for s in shapes]
  dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
  if dist.get_rank() == src:
    flat = torch.from_numpy(np.concatenate([np.asarray(params[n], np.float32).ravel() for n in names])).to(dev)
  else:
    flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
  dist.broadcast(flat, src=src)
  flat = flat.cpu().numpy()
  out, off = {}, 0
  for n, s, k in zip(names, shapes, sizes):
    out[n] = flat[off:off + k].reshape(s)
    off += k
  return out


def max_over_ranks(value, device=None):
  import torch
  import torch.distributed as dist
  if not dist.is_initialized() or dist.get_world_size() == 1:
    return float(value)
  dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")
  t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t.item())


def barrier():
  import torch.distributed as dist
  if dist.is_initialized() and dist.get_world_size() > 1:
    dist.barrier()
