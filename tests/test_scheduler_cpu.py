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


# ---- RewindingScheduler against per-filter reference-semantics drivers (rewind + fast-forward, ekf_sym.py:418-482) ----
class _OracleEngine:
  """Stand-in with BatchedEKF's surface (x, P, step_indexed) that computes on the CPU oracle library: the scheduler's
  ring / rewind / replay bookkeeping is device-agnostic torch code, this lets it run without a GPU."""

  def __init__(self, oracle, x, P, Q, quat_idxs, flags):
    self.o, self.Q, self.quat, self.flags = oracle, Q, list(quat_idxs), flags
    self.x, self.P = torch.as_tensor(x.copy()), torch.as_tensor(P.copy())
    self.B, self.device, self.launches = x.shape[0], torch.device("cpu"), 0

  def step_indexed(self, kind, idx, dt, z, R, ea=None):
    sub = idx.numpy().astype(np.int64)
    n = len(sub)
    Rn = R.numpy() if R.ndim == 3 else np.tile(R.numpy(), (n, 1, 1))
    xr, Pr, y = self.o.batch_step(kind, self.x[sub].numpy(), self.P[sub].numpy(), self.Q, dt.numpy(), z.numpy().reshape(n, -1), Rn,
                                  ea=None if ea is None else ea.numpy(), quat_idxs=self.quat, flags=self.flags, nthreads=1)
    self.x[sub], self.P[sub] = torch.as_tensor(xr), torch.as_tensor(Pr)
    self.launches += 1
    return torch.as_tensor(y).reshape(n, 1, -1)


def _run_rewinding_case(oracle_dir, depth, monkeypatch, seed, packed=False):
  import rednose_b200.ekf_sym as drv
  from rednose_b200.filters.live import LiveKalman
  from rednose_b200.scheduler import RewindingScheduler
  from tests.util import Oracle
  monkeypatch.setattr(drv, "REWIND_TO_KEEP", depth)          # the reference keeps 512 (ekf_sym.py:447); same ring size on both sides
  B, zd = 5, {3: 1, 4: 3, 10: 3, 12: 3}
  rng = np.random.default_rng(seed)
  x0 = np.tile(LiveKalman.initial_x, (B, 1)); x0[:, :3] += rng.normal(0, 10.0, (B, 3))
  P0 = np.tile(np.diag(LiveKalman.initial_P_diag), (B, 1, 1))
  Rk = {3: np.array([[0.2**2]]), 4: np.eye(3) * 0.025**2, 10: np.eye(3) * 0.5**2, 12: np.eye(3) * 25.0}
  refs = [drv.EKF_sym(oracle_dir, "live", LiveKalman.Q, x0[b], P0[b], 23, 22, quaternion_idxs=[3], max_rewind_age=0.5) for b in range(B)]
  eng = _OracleEngine(Oracle(oracle_dir, "live"), x0, P0, LiveKalman.Q, [3], flags=2)    # python-driver semantics: normalise after the update only
  s = RewindingScheduler(eng, zd, depth=depth, max_rewind_age=0.5, packed=packed)
  ref_dropped = 0
  for tick in range(70):
    now = 0.01 * (tick + 1)
    ids, ts, ks, zs = [], [], [], {k: [] for k in zd}
    for b in range(B):
      if rng.random() < 0.25:
        continue
      u = rng.random()
      tb = now + 1e-4 * b
      if tick > 5 and u < 0.15:
        tb -= rng.uniform(0.011, 0.06)            # late: rewind over 1-6 checkpoints
      elif tick > 5 and u < 0.20:
        tb -= 3.0                                 # hopelessly late: ignored
      k = int(rng.choice([4, 10, 10, 4, 3 if tick > 12 else 4, 12]))   # the speed observation is singular at v = 0
      zb = {3: np.array([0.1]), 4: rng.normal(0, 0.01, 3), 10: rng.normal(0, 0.1, 3) + [0, 0, -9.8], 12: refs[b].state()[:3] + rng.normal(0, 1.0, 3)}[k]
      ids.append(b); ts.append(tb); ks.append(k); zs[k].append(zb)
      if refs[b].predict_and_update_batch(tb, k, zb[None], Rk[k][None]) is None:
        ref_dropped += 1
    if ids:
      s.tick(np.array(ids), np.array(ts), np.array(ks), {k: np.array(v) for k, v in zs.items() if v}, Rk)
  for b in range(B):
    ex = np.max(np.abs(eng.x[b].numpy() - refs[b].state())) / np.max(np.abs(refs[b].state()))
    eP = np.max(np.abs(eng.P[b].numpy() - refs[b].covs())) / np.max(np.abs(refs[b].covs()))
    assert ex < 1e-12 and eP < 1e-10, (b, ex, eP)
    assert abs(float(s.t_filter[b]) - refs[b].filter_time) < 1e-12
    assert int(s.cnt[b]) == len(refs[b].rewind_t)
  assert s.dropped == ref_dropped
  return s


def test_rewinding_scheduler_equals_per_filter_rewind(oracle_dir, monkeypatch):
  s = _run_rewinding_case(oracle_dir, 64, monkeypatch, seed=7)
  assert s.rewinds > 10 and s.replayed > s.rewinds and s.dropped > 3      # the stream really exercised all three paths


def test_rewinding_scheduler_with_a_short_ring(oracle_dir, monkeypatch):
  """depth 4: the ring wraps many times, rewinds reach its oldest entry and observations older than it are ignored."""
  s = _run_rewinding_case(oracle_dir, 4, monkeypatch, seed=11)
  assert s.rewinds > 5 and s.dropped > 5


def test_rewinding_scheduler_with_packed_covariance_snapshots(oracle_dir, monkeypatch):
  """packed=True: the ring holds lower triangles (half the memory); same results to the tolerance of the unpacked run."""
  s = _run_rewinding_case(oracle_dir, 64, monkeypatch, seed=7, packed=True)
  assert s.ring_P.shape[-1] == 22 * 23 // 2 and s.rewinds > 10 and s.replayed > s.rewinds
