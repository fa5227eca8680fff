"""World-size-2 gloo test (CPU) of the multi-process host logic: batch sharding covers the global batch exactly
once, the parameter broadcast delivers rank 0's arrays bit-exactly, and the timing reduction is a max."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from gddim_b200 import dist as gdist


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, q):
  os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                    MASTER_PORT=str(port))
  gdist.init_process_group("gloo")
  rng = np.random.default_rng(123 if rank == 0 else 999)          # only rank 0 holds the "real" parameters
  params = {"b/kernel": rng.standard_normal((3, 3, 4, 8)).astype(np.float32),
            "a/bias": rng.standard_normal((8,)).astype(np.float32)}
  got = gdist.broadcast_params(params, src=0)
  u = np.arange(8 * 5, dtype=np.float32).reshape(8, 5)             # same global prior on every rank
  mine = gdist.shard(u, rank, world)
  t = gdist.max_over_ranks(1.0 + rank)
  gdist.barrier()
  q.put((rank, {k: v.copy() for k, v in got.items()}, mine.copy(), t))


def test_shard_broadcast_max_world2():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  want = np.random.default_rng(123)
  k = want.standard_normal((3, 3, 4, 8)).astype(np.float32)
  b = want.standard_normal((8,)).astype(np.float32)
  for rank, got, mine, t in res:
    np.testing.assert_array_equal(got["b/kernel"], k)
    np.testing.assert_array_equal(got["a/bias"], b)
    assert t == 2.0
  np.testing.assert_array_equal(np.concatenate([res[0][2], res[1][2]]), np.arange(40, dtype=np.float32).reshape(8, 5))


def test_shard_bounds_errors():
  assert gdist.shard_bounds(2048, 3, 8) == (768, 1024)
  import pytest
  with pytest.raises(ValueError):
    gdist.shard_bounds(10, 0, 4)
