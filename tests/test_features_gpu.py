"""MSCKF front-end kernels on the GPU (through the C-ABI of libfeatures_<K>.so) against the oracle's restatement of
rednose/templates/compute_pos.c and feature_handler.c."""
import numpy as np
import pytest

from tests.test_features_cpu import K, feat_oracle, oracle_compute_pos, py_sane, synth_frame, synth_tracks  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe():
  import torch
  assert torch.cuda.is_available()
  from rednose_b200.features import FeatureFrontend
  return FeatureFrontend(K)


def test_compute_pos_batch_matches_oracle(fe, feat_oracle):
  import torch
  for noise, B in ((0.0, 1000), (2e-3, 3001)):
    to_c, poses, img, _ = synth_tracks(B, seed=11, noise=noise)
    pos_r, param_r = oracle_compute_pos(feat_oracle, to_c, poses, img)
    pos, param, iters = fe.compute_pos_batch(to_c, torch.as_tensor(poses).cuda(), torch.as_tensor(img).cuda())
    torch.cuda.synchronize()
    pos, param, iters = pos.cpu().numpy(), param.cpu().numpy(), iters.cpu().numpy()
    # float64, same algorithm, different summation order of J^T J: 1e-6 relative is the contract, measured ~1e-12
    assert np.max(np.abs(param - param_r)) / np.max(np.abs(param_r)) < 1e-8
    assert np.max(np.abs(pos - pos_r)) / np.max(np.abs(pos_r)) < 1e-11
    assert iters.min() >= 1 and iters.max() <= 30


def test_compute_pos_single_host_pointer_entry_point(fe, feat_oracle):
  to_c, poses, img, _ = synth_tracks(4, seed=3, noise=1e-3)
  pos_r, param_r = oracle_compute_pos(feat_oracle, to_c, poses, img)
  for b in range(4):
    pos, param = fe.compute_pos(to_c, poses[b], img[b])
    assert np.max(np.abs(pos - pos_r[b])) / np.max(np.abs(pos_r[b])) < 1e-11 and np.max(np.abs(param - param_r[b])) < 1e-8


def test_res_and_jac_leaf_entry_points_match_reference_printed_c(fe, feat_oracle):
  ffi_o, lib_o = feat_oracle
  to_c, poses, img, _ = synth_tracks(3, seed=5, noise=1e-3)
  x = np.array([0.03, -0.02, 0.07])
  for b in range(3):
    want_r, want_j, got_r, got_j = np.zeros(2 * K), np.zeros(6 * K), np.zeros(2 * K), np.zeros(6 * K)
    po = lambda a: ffi_o.cast("double *", a.ctypes.data)
    lib_o.res_fun(po(x), po(poses[b]), po(img[b]), po(want_r)); lib_o.jac_fun(po(x), po(poses[b]), po(img[b]), po(want_j))
    pg = lambda a: fe.ffi.cast("double *", a.ctypes.data)
    fe.lib.res_fun(pg(x), pg(poses[b]), pg(img[b]), pg(got_r)); fe.lib.jac_fun(pg(x), pg(poses[b]), pg(img[b]), pg(got_j))
    assert np.max(np.abs(got_r - want_r)) < 1e-12 * max(1.0, np.max(np.abs(want_r)))
    assert np.max(np.abs(got_j - want_j)) < 1e-11 * max(1.0, np.max(np.abs(want_j)))


@pytest.mark.parametrize("smooth,collide", [(True, False), (False, False), (True, True)])
def test_merge_features_batch_bit_exact(fe, feat_oracle, smooth, collide):
  import torch
  ffi_o, lib_o = feat_oracle
  B, nt, nf = 24, 700, 333
  tr, ft, em = zip(*[synth_frame(nt, nf, 100 + s, smooth, collide and s % 3 == 0) for s in range(B)])
  tracks, feats, empty = np.stack(tr), np.stack(ft), np.stack(em)
  want = tracks.copy()
  for b in range(B):
    lib_o.merge_features_n(ffi_o.cast("double *", want[b].ctypes.data), ffi_o.cast("double *", feats[b].ctypes.data),
                           ffi_o.cast("long long *", empty[b].ctypes.data), nf, nt)
  d_tracks = torch.as_tensor(tracks).cuda()
  fb = fe.merge_features_batch(d_tracks, torch.as_tensor(feats).cuda(), torch.as_tensor(empty).cuda())
  torch.cuda.synchronize()
  assert np.array_equal(d_tracks.cpu().numpy(), want)              # integer / index work: bit exact
  assert int(fb.item()) == (sum(1 for s in range(B) if s % 3 == 0) if collide else 0)


def test_merge_features_reference_sizes_host_pointers(fe, feat_oracle):
  ffi_o, lib_o = feat_oracle
  tracks, feats, empty = synth_frame(6000, 3000, 9, True, False)
  want = tracks.copy()
  lib_o.merge_features(ffi_o.cast("double *", want.ctypes.data), ffi_o.cast("double *", feats.ctypes.data), ffi_o.cast("long long *", empty.ctypes.data))
  got = fe.merge_features(tracks.copy(), feats, empty)
  assert np.array_equal(got, want)


def test_sane_batch(fe):
  import torch
  rng = np.random.default_rng(0)
  tracks = rng.uniform(-0.3, 0.3, (5000, K + 1, 5))
  tracks[:2500, 1:, 2] = np.linspace(0, 0.5, K)[None, :] + rng.normal(0, 1e-3, (2500, K))
  tracks[:2500, 1:, 3] = np.linspace(0, 0.2, K)[None, :]
  got = fe.sane_batch(torch.as_tensor(tracks).cuda()).cpu().numpy()
  want = np.array([py_sane(t) for t in tracks]).astype(np.int32)
  assert np.array_equal(got, want) and 0 < want.sum() < 5000


def test_compute_pos_fallback_guard(fe):
  """fallback_depth > 0 (an addition; the reference has no guard): a track whose Gauss-Newton does not converge comes back
  finite, on the last camera's optical axis, flagged by a negative iteration count; converged tracks are untouched."""
  import torch
  from rednose_b200.geometry import quat2rot
  to_c, poses, img, _ = synth_tracks(500, seed=21, noise=1e-3)
  img_bad = img.copy()
  bad = np.arange(0, 500, 25)
  rng = np.random.default_rng(3)
  img_bad[bad] = rng.normal(0, 5.0, (bad.size, 2 * K))          # observations unrelated to the geometry
  p0, _, it0 = fe.compute_pos_batch(to_c, torch.as_tensor(poses).cuda(), torch.as_tensor(img_bad).cuda())
  p1, _, it1 = fe.compute_pos_batch(to_c, torch.as_tensor(poses).cuda(), torch.as_tensor(img_bad).cuda(), fallback_depth=30.0)
  p0, p1, it0, it1 = p0.cpu().numpy(), p1.cpu().numpy(), it0.cpu().numpy(), it1.cpu().numpy()
  failed = (it0 >= 30) | ~np.isfinite(p0).all(axis=1)
  assert failed.any() and set(np.nonzero(failed)[0]) <= set(bad)
  assert np.isfinite(p1).all() and np.array_equal(it1 < 0, failed) and np.array_equal(p1[~failed], p0[~failed])
  P = poses.reshape(500, K, 7)
  for b in np.nonzero(failed)[0]:
    want = P[b, K - 1, 0:3] + quat2rot(P[b, K - 1, 3:7]) @ to_c.T @ np.array([0.0, 0.0, 30.0])
    assert np.max(np.abs(p1[b] - want)) < 1e-6
