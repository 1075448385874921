"""Host <-> device streaming front-end of a BatchedEKF: observations arrive in pinned HOST memory every
step, state estimates and innovations go back to pinned HOST memory every step, the covariance never leaves
the GPU.

This is the batched analogue of the reference's calling pattern, where every
`predict_and_update_batch(t, kind, z, R)` call hands host arrays to the driver and gets host arrays back
(rednose/helpers/ekf_sym_pyx.pyx:144-180).  Three CUDA streams overlap the work of consecutive steps:

  copy-in   z_k  -> device slot k % 2                     (H2D, pinned)
  compute   fused predict+update on slot k % 2; the kernel also writes x_{k|k} into snapshot k % 2
  copy-out  snapshot k % 2 -> x_host[k % 2], innovations -> y_host[k % 2]   (D2H, pinned)

so a step costs max(H2D, kernel, D2H) instead of their sum.  Results of step k are valid after `wait(k)`.
"""
from __future__ import annotations

import torch


class HostStreamer:
  def __init__(self, engine, zdims: dict[int, int], depth: int = 2, out_cols=None, every: int = 1):
    """out_cols: state columns copied back to the host each step (None = the whole state; e.g. range(7) for pose only);
    every: copy the estimate back only every `every`-th step (innovations always come back)."""
    self.e = engine
    self.every = max(1, int(every))
    dev = engine.device
    self.depth = depth
    self.s_in = torch.cuda.Stream(dev)
    self.s_out = torch.cuda.Stream(dev)
    B, D = engine.B, engine.dim_x
    kw = dict(dtype=torch.float64, device=dev)
    self.z_dev = {k: [torch.empty(B, 1, m, **kw) for _ in range(depth)] for k, m in zdims.items()}
    self.x_snap = [torch.empty(B, D, **kw) for _ in range(depth)]
    self.cols = None if out_cols is None else torch.as_tensor(list(out_cols), dtype=torch.long, device=dev)
    Dout = D if self.cols is None else int(self.cols.numel())
    self.x_sel = None if self.cols is None else [torch.empty(B, Dout, **kw) for _ in range(depth)]
    self.x_host = [torch.empty(B, Dout, dtype=torch.float64).pin_memory() for _ in range(depth)]
    self.y_host = {k: [torch.empty(B, 1, m, dtype=torch.float64).pin_memory() for _ in range(depth)] for k, m in zdims.items()}
    self.ev_in = [torch.cuda.Event() for _ in range(depth)]
    self.ev_done = [torch.cuda.Event() for _ in range(depth)]
    self.ev_out = [torch.cuda.Event() for _ in range(depth)]
    self.k = 0
    self.h2d_bytes = 0
    self.d2h_bytes = 0

  def submit(self, t, kind, z_host: torch.Tensor, R):
    """Enqueue one step for the whole batch.  z_host: pinned [B, m]; R: [m, m] (shared) or device [B, m, m].
    Returns a ticket for wait()."""
    e, slot = self.e, self.k % self.depth
    main = torch.cuda.current_stream(e.device)
    if self.k >= self.depth:
      self.s_in.wait_event(self.ev_out[slot])      # slot's previous results must have left the device
    with torch.cuda.stream(self.s_in):
      self.z_dev[kind][slot][:, 0, :].copy_(z_host, non_blocking=True)
      self.ev_in[slot].record(self.s_in)
    main.wait_event(self.ev_in[slot])
    if e.filter_time is None:
      e.filter_time = t
    dt = t - e.filter_time
    e.step(kind, dt, self.z_dev[kind][slot], R, hist_filt=(self.x_snap[slot], None))
    e.filter_time = t
    self.ev_done[slot].record(main)
    self.s_out.wait_event(self.ev_done[slot])
    send_x = (self.k % self.every) == 0
    with torch.cuda.stream(self.s_out):
      if send_x:
        if self.cols is None:
          self.x_host[slot].copy_(self.x_snap[slot], non_blocking=True)
        else:
          torch.index_select(self.x_snap[slot], 1, self.cols, out=self.x_sel[slot])
          self.x_host[slot].copy_(self.x_sel[slot], non_blocking=True)
      self.y_host[kind][slot].copy_(self.z_dev[kind][slot], non_blocking=True)
      self.ev_out[slot].record(self.s_out)
    self.h2d_bytes += z_host.numel() * 8
    self.d2h_bytes += ((self.x_host[slot].numel() if send_x else 0) + self.y_host[kind][slot].numel()) * 8
    self.k += 1
    return self.k - 1

  def wait(self, ticket=None):
    """Block until the results of `ticket` (default: everything submitted) are in host memory."""
    if ticket is None:
      for ev in self.ev_out[:min(self.k, self.depth)]:
        ev.synchronize()
      return None
    slot = ticket % self.depth
    self.ev_out[slot].synchronize()
    return slot

  def result(self, ticket, kind):
    slot = self.wait(ticket)
    return self.x_host[slot], self.y_host[kind][slot][:, 0]
