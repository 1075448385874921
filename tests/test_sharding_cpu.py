"""CPU tests of the N > 1 path: shard layout + final gather over `gloo`, world_size 2 (and a ragged 3)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rednose_b200.sharding import gather_filters, shard_bounds, shard_sizes


def test_shard_bounds_cover_the_batch():
  for total in (0, 1, 7, 100000, 1 << 20):
    for world in (1, 2, 3, 8):
      b = [shard_bounds(total, r, world) for r in range(world)]
      assert b[0][0] == 0 and b[-1][1] == total
      assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
      assert max(shard_sizes(total, world)) - min(shard_sizes(total, world)) <= 1


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, total, oracle_dir, ret):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from tests.util import Oracle, kinematic_batch
  # every rank builds the same global problem, steps only its shard (here with the CPU oracle standing in
  # for the GPU kernels: this test is about the host-side layout), then the shards are gathered
  x, P, Q, z, R = kinematic_batch(total, seed=3)
  lo, hi = shard_bounds(total, rank, world)
  o = Oracle(oracle_dir, "kinematic")
  xs, Ps, _ = o.batch_step(1, x[lo:hi], P[lo:hi], Q, 0.01, z[lo:hi], R[lo:hi], nthreads=1)
  gx = gather_filters(torch.as_tensor(xs), total)
  gP = gather_filters(torch.as_tensor(Ps), total)
  if rank == 0:
    xr, Pr, _ = o.batch_step(1, x, P, Q, 0.01, z, R, nthreads=1)
    ret.put((bool(np.array_equal(gx.numpy(), xr)), bool(np.array_equal(gP.numpy(), Pr)), tuple(gx.shape)))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 1000), (3, 1001)])
def test_sharded_step_plus_gather_equals_single_process(oracle_dir, world, total):
  ctx = mp.get_context("spawn")
  ret = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, total, oracle_dir, ret)) for r in range(world)]
  for p in procs:
    p.start()
  ok_x, ok_P, shape = ret.get(timeout=120)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert ok_x and ok_P and shape == (total, 2)


def test_numa_binding_is_a_no_op_without_a_gpu():
  """bind_to_gpu_numa must never break a process that has no (visible) GPU or no NUMA information: it returns None and
  leaves the affinity mask alone."""
  import os
  import torch
  from rednose_b200.sharding import bind_to_gpu_numa, gpu_numa_cpus
  if torch.cuda.is_available():
    import pytest
    pytest.skip("a GPU is present")
  before = os.sched_getaffinity(0)
  assert gpu_numa_cpus(0) == (None, None) and bind_to_gpu_numa(0) is None
  assert os.sched_getaffinity(0) == before
