"""Ragged observation scheduler: independent filters with DIFFERENT observation streams.

In the reference every filter instance is driven on the host, one `predict_and_update_batch(t, kind, z, R)`
call per observation (rednose/helpers/ekf_sym.py:464-531, ekf_sym.cc:83-117).  With a batch of filters the
streams interleave: at a given tick some filters see a gyro sample, some a position fix, some nothing.  The
scheduler buckets the observations of one tick by kind and issues ONE indexed fused launch per kind
(`<name>_batch_step_<kind>_idx`), keeping a per-filter clock on the device so each filter is predicted over
its own dt = t_obs - t_filter.  Filters without an observation in the tick are not touched.

Late observations (t_obs < t_filter) are dropped and counted -- the host-side rewind of the single-filter
drivers (ekf_sym.cc:125-142) has no batched counterpart yet.
"""
from __future__ import annotations

import torch


class RaggedScheduler:
  def __init__(self, engine):
    self.e = engine
    self.t_filter = torch.full((engine.B,), float("nan"), dtype=torch.float64, device=engine.device)
    self.dropped = 0

  def tick(self, filter_ids, t, kinds, z_by_kind, R_by_kind, ea_by_kind=None):
    """One scheduling tick.

    filter_ids [n] int, t [n] float64 (or scalar), kinds [n] int: the observations of this tick, at most one per
    filter.  z_by_kind[k] is [n_k, m_k] in the order the entries of kind k appear in `filter_ids`;
    R_by_kind[k] is [m_k, m_k] (shared) or [n_k, m_k, m_k].  Returns {kind: (filter_ids_k, innovations_k)}.
    """
    dev = self.e.device
    fid = torch.as_tensor(filter_ids, device=dev).to(torch.int64)
    kinds = torch.as_tensor(kinds, device=dev)
    t = torch.as_tensor(t, dtype=torch.float64, device=dev).expand(fid.shape[0])
    out = {}
    for k in sorted(z_by_kind):
      sel = (kinds == k).nonzero(as_tuple=True)[0]
      if sel.numel() == 0:
        continue
      ids, tk = fid[sel], t[sel]
      tf = self.t_filter[ids]
      tf = torch.where(torch.isnan(tf), tk, tf)          # first observation initialises the clock (ekf_sym.py:502-503)
      dt = tk - tf
      ok = dt >= 0
      z = torch.as_tensor(z_by_kind[k], device=dev, dtype=torch.float64)
      R = torch.as_tensor(R_by_kind[k], device=dev, dtype=torch.float64)
      ea = None if not ea_by_kind or k not in ea_by_kind else torch.as_tensor(ea_by_kind[k], device=dev, dtype=torch.float64)
      if not bool(ok.all()):                              # too old: ignored, like ekf_sym.py:468-471
        self.dropped += int((~ok).sum())
        ids, tk, dt, z = ids[ok], tk[ok], dt[ok], z[ok]
        if R.ndim == 3:
          R = R[ok]
        if ea is not None:
          ea = ea[ok]
      y = self.e.step_indexed(k, ids.to(torch.int32), dt, z.clone(), R, ea)
      self.t_filter[ids] = tk
      out[k] = (ids, None if y is None else y[:, 0])
    return out
