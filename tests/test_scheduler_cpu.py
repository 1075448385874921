"""Host logic of the ragged scheduler (rednose_b200/scheduler.py) against a recording stand-in for the engine:
bucketing by kind, per-filter dt from the per-filter clock, first observation initialises the clock
(rednose/helpers/ekf_sym.py:502-503), observations older than the filter are dropped (:468-471)."""
import numpy as np
import torch

from rednose_b200.scheduler import RaggedScheduler


class _Recorder:
  def __init__(self, B):
    self.B, self.device, self.calls = B, torch.device("cpu"), []

  def step_indexed(self, kind, idx, dt, z, R, ea=None):
    self.calls.append((kind, idx.clone(), dt.clone(), z.clone(), R.clone(), None if ea is None else ea.clone()))
    return z.reshape(z.shape[0], 1, -1) * 0.0 + float(kind)


def test_buckets_clock_and_late_drops():
  eng = _Recorder(10)
  s = RaggedScheduler(eng)
  ids = np.array([0, 3, 5, 7, 9])
  kinds = np.array([4, 10, 4, 12, 4])
  z4, z10, z12 = np.arange(9.0).reshape(3, 3), np.ones((1, 3)), np.full((1, 3), 2.0)
  R = {4: np.eye(3), 10: np.eye(3)[None] * 2.0, 12: np.eye(3)}
  out = s.tick(ids, 1.0, kinds, {4: z4, 10: z10, 12: z12}, R)
  assert [c[0] for c in eng.calls] == [4, 10, 12]                       # one launch per kind, ascending
  assert eng.calls[0][1].tolist() == [0, 5, 9] and eng.calls[0][1].dtype == torch.int32
  assert all(float(c[2].abs().max()) == 0.0 for c in eng.calls)        # first observation: dt = 0, clock set
  assert torch.equal(eng.calls[0][3], torch.as_tensor(z4))
  assert set(out) == {4, 10, 12} and out[4][1].shape == (3, 3) and float(out[10][1][0, 0]) == 10.0
  assert np.isnan(s.t_filter[[1, 2, 4, 6, 8]].numpy()).all() and (s.t_filter[[0, 3, 5, 7, 9]] == 1.0).all()

  # second tick: per-filter times; filter 5's observation is older than its clock -> dropped, clock untouched
  eng.calls.clear()
  out = s.tick(np.array([0, 5, 1]), np.array([1.25, 0.5, 3.0]), np.array([4, 4, 4]), {4: np.array([[1.0] * 3, [2.0] * 3, [3.0] * 3])},
               {4: np.stack([np.eye(3) * (i + 1) for i in range(3)])})
  (kind, idx, dt, z, Rk, ea), = eng.calls
  assert kind == 4 and idx.tolist() == [0, 1] and dt.tolist() == [0.25, 0.0]
  assert z[:, 0].tolist() == [1.0, 3.0] and Rk[:, 0, 0].tolist() == [1.0, 3.0] and ea is None   # rows of the dropped entry removed everywhere
  assert s.dropped == 1 and float(s.t_filter[5]) == 1.0 and float(s.t_filter[0]) == 1.25 and float(s.t_filter[1]) == 3.0
  assert out[4][0].tolist() == [0, 1]


def test_kinds_without_entries_launch_nothing_and_extra_args_follow_the_selection():
  eng = _Recorder(4)
  s = RaggedScheduler(eng)
  s.tick([2, 0], 0.5, [13, 13], {13: np.ones((2, 3)), 19: np.zeros((0, 3))}, {13: np.eye(3), 19: np.eye(3)},
         ea_by_kind={13: np.array([[7.0], [8.0]])})
  (kind, idx, dt, z, R, ea), = eng.calls
  assert kind == 13 and idx.tolist() == [2, 0] and ea[:, 0].tolist() == [7.0, 8.0] and R.shape == (3, 3)
