"""Multi-GPU layout of a batch of independent filters.

Filters never interact (no cross-filter term anywhere in ekf_c.c / ekf_sym.cc), so the batch is cut into
contiguous shards, one per rank / GPU; stepping and smoothing need NO communication.  The only collective
of the system is the optional final gather of the state estimates (and, if asked, covariances) over
NCCL / NVLink.  One process per GPU, `torch.distributed` for the plumbing.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
  """Contiguous [lo, hi) of `total` filters owned by `rank`; sizes differ by at most one."""
  base, rem = divmod(total, world)
  lo = rank * base + min(rank, rem)
  return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> list[int]:
  return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]


def gather_filters(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
  """All-gather per-filter rows ([B_local, ...]) of every rank into [total, ...] in rank order.

  Shards may be ragged (total not divisible by world): rows are padded to the largest shard for the
  collective and trimmed afterwards.
  """
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return local
  world = dist.get_world_size(group)
  sizes = shard_sizes(total, world)
  assert local.shape[0] == sizes[dist.get_rank(group)]
  pad = max(sizes)
  buf = local
  if local.shape[0] < pad:
    buf = torch.cat([local, local.new_zeros((pad - local.shape[0],) + tuple(local.shape[1:]))])
  out = local.new_empty((world * pad,) + tuple(local.shape[1:]))
  dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
  if all(s == pad for s in sizes):
    return out
  return torch.cat([out[r * pad:r * pad + sizes[r]] for r in range(world)])


def gpu_numa_cpus(device_index: int):
  """(numa_node, sorted CPU list) of the host socket GPU `device_index` hangs off, or (None, None) if unknown.
  Read from sysfs through the GPU's PCI address (no nvidia-smi dependency)."""
  try:
    import torch
    bus = torch.cuda.get_device_properties(device_index)
    pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
    with open(f"/sys/bus/pci/devices/{pci}/numa_node", encoding="utf-8") as f:
      node = int(f.read().strip())
    if node < 0:
      return None, None
    with open(f"/sys/devices/system/node/node{node}/cpulist", encoding="utf-8") as f:
      spec = f.read().strip()
    cpus = []
    for part in spec.split(","):
      a, _, b = part.partition("-")
      cpus.extend(range(int(a), int(b or a) + 1))
    return node, sorted(cpus)
  except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
    return None, None


def bind_to_gpu_numa(device_index: int):
  """Pin this process (one process per GPU) to the CPUs of the GPU's NUMA node, so that pinned staging buffers
  allocated afterwards are first-touched on the socket the GPU's PCIe root belongs to and the per-step host<->device
  copies of 8 ranks do not cross the inter-socket link.  Returns the node bound to, or None (left unbound)."""
  import os
  node, cpus = gpu_numa_cpus(device_index)
  if node is None:
    return None
  try:
    allowed = os.sched_getaffinity(0)
    target = allowed & set(cpus)
    if not target:
      return None
    os.sched_setaffinity(0, target)
    return node
  except (OSError, AttributeError):
    return None
