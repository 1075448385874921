"""Forward filter + RTS smoother over long histories, tiled over filters.

The smoother needs x_{k|k-1}, x_{k|k}, P_{k|k-1}, P_{k|k} of every step (rednose/helpers/ekf_sym.py:651-690
walks the list of tuples returned by predict_and_update_batch): 2*EDIM^2 + 2*DIM + 1 doubles per filter-step,
8.1 kB for live_kf.  BASELINE.json config 4 (1M filters x 10k steps) would be 81 TB, so the batch is cut into
TILES of filters whose whole history fits the HBM budget; each tile runs forward-store-then-backward-consume and
hands its smoothed track to a sink before the next tile reuses the buffers (hard part 3).
Tiles are independent, so multi-GPU use is: shard the filters over ranks first, tile within a rank.
"""
from __future__ import annotations

import torch

from rednose_b200.batched import BatchedEKF


def history_bytes_per_filter(dim_x, dim_err, T, smoothed_in_place=True):
  per_step = 2 * dim_err * dim_err + 2 * dim_x
  if not smoothed_in_place:
    per_step += dim_err * dim_err + dim_x
  return 8 * per_step * T


class TiledSmoother:
  def __init__(self, folder, name, Q, dim_x, dim_err, quaternion_idxs=(), device="cuda", hbm_budget_bytes=120 << 30, tile=None):
    self.folder, self.name, self.Q = folder, name, Q
    self.dim_x, self.dim_err = dim_x, dim_err
    self.quat = tuple(quaternion_idxs)
    self.device = torch.device(device)
    self.budget = hbm_budget_bytes
    self.tile = tile
    self._engine = None
    self._hist = None

  def tile_size(self, T):
    if self.tile:
      return self.tile
    per = history_bytes_per_filter(self.dim_x, self.dim_err, T) + 8 * (self.dim_err**2 + self.dim_x)
    return max(1, int(self.budget // per))

  def run(self, x0, P0, T, obs_fn, sink, norm_quats=False, t0=0.0, passes=1):
    """x0 [B, DIM], P0 [B, EDIM, EDIM] (host or device).  obs_fn(k, lo, hi) -> (t, kind, z [hi-lo, m], R) gives the
    observation of step k for filters lo..hi.  sink(lo, hi, xs [T, n, DIM], Ps [T, n, EDIM, EDIM]) receives device
    views that are only valid during the call.  Returns the number of tiles.

    passes > 1: "multiple forward and backwards passes of the data" (reference README.md:41-45) -- each further pass
    restarts the forward filter of the tile from the previous pass's smoothed estimate at the first step
    (x_{0|N}, P_{0|N}), which removes the dependence on a poor initialisation; the sink sees the last pass."""
    B = x0.shape[0]
    n_tile = min(self.tile_size(T), B)
    tiles = 0
    for lo in range(0, B, n_tile):
      hi = min(lo + n_tile, B)
      n = hi - lo
      if self._engine is None or self._engine.B != n:
        self._engine = BatchedEKF(self.folder, self.name, self.Q, x0[lo:hi], P0[lo:hi], device=self.device, quaternion_idxs=self.quat)
        self._hist = self._engine.new_history(T) if (self._hist is None or self._hist.B != n or self._hist.T != T) else self._hist
      else:
        self._engine.init_state(x0[lo:hi], P0[lo:hi], None)
      eng, hist = self._engine, self._hist
      for p in range(max(1, int(passes))):
        if p > 0:   # the smoothed slabs alias the history buffers the next forward pass overwrites: copy step 0 out first
          eng.init_state(xs[0].clone(), Ps[0].clone(), None)
        hist.n = 0
        eng.filter_time = t0
        for k in range(T):
          t, kind, z, R = obs_fn(k, lo, hi)
          eng.step_recorded(hist, kind, t, z, R)
        xs, Ps = eng.rts_smooth(hist, norm_quats=norm_quats, quaternion_idxs=self.quat or (3,), in_place=True)
      sink(lo, hi, xs, Ps)
      tiles += 1
    return tiles
