// ORACLE -- test infrastructure only.  Eigen-free CPU restatement of the reference's MSCKF front-end templates
//   rednose/templates/compute_pos.c:10-52     (gauss_newton, compute_pos)
//   rednose/templates/feature_handler.c:1-56  (sane, merge_features)
// statement by statement, same operation order.  KDIM and K are #defined by the assembled translation unit
// (oracle/build_ref.py: build_features), which also supplies res_fun / jac_fun printed by the reference's OWN
// sympy_into_c (rednose/helpers/sympy_helpers.py:122-162) from the residual of rednose_b200/features.py:residual_sym
// (the templates' user, openpilot's lst_sq_computer.py, is not part of /root/reference).
//
// Parity: the reference holds no test, fixture or golden vector for these templates ("parity unpinned" by the
// reference itself); the restatement is pinned instead by domain properties in tests/test_features_cpu.py
// (a synthetic 3-D point is recovered from its projections; J^T E = 0 at the solution; merge_features agrees with a
// straight Python transcription of the template).
#pragma once
#include <math.h>
#include <string.h>

// (J^T J)^-1 J^T E with the closed-form 3 x 3 inverse (Eigen's fixed-size Matrix3d::inverse() is cofactors / determinant)
static void oracle_gauss_newton(double* in_x, double* in_poses, double* in_img_positions) {   // compute_pos.c:10-27
  double res[KDIM * 2] = {0};
  double jac[KDIM * 6] = {0};
  double x[3] = {in_x[0], in_x[1], in_x[2]};
  double delta[3] = {0, 0, 0};
  int counter = 0;
  while (((delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]) > 0.0001 && counter < 30) || counter == 0) {   // :18
    res_fun(in_x, in_poses, in_img_positions, res);   // :19
    jac_fun(in_x, in_poses, in_img_positions, jac);   // :20
    double A[3][3] = {{0}}, g[3] = {0};
    for (int r = 0; r < KDIM * 2; ++r)                // J^T J, J^T E (:22)
      for (int a = 0; a < 3; ++a) {
        g[a] += jac[r * 3 + a] * res[r];
        for (int b = 0; b < 3; ++b) A[a][b] += jac[r * 3 + a] * jac[r * 3 + b];
      }
    double c[3][3];
    c[0][0] = A[1][1] * A[2][2] - A[1][2] * A[2][1]; c[0][1] = A[0][2] * A[2][1] - A[0][1] * A[2][2]; c[0][2] = A[0][1] * A[1][2] - A[0][2] * A[1][1];
    c[1][0] = A[1][2] * A[2][0] - A[1][0] * A[2][2]; c[1][1] = A[0][0] * A[2][2] - A[0][2] * A[2][0]; c[1][2] = A[0][2] * A[1][0] - A[0][0] * A[1][2];
    c[2][0] = A[1][0] * A[2][1] - A[1][1] * A[2][0]; c[2][1] = A[0][1] * A[2][0] - A[0][0] * A[2][1]; c[2][2] = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    const double det = A[0][0] * c[0][0] + A[0][1] * c[1][0] + A[0][2] * c[2][0];
    for (int a = 0; a < 3; ++a) delta[a] = (c[a][0] * g[0] + c[a][1] * g[1] + c[a][2] * g[2]) / det;
    for (int a = 0; a < 3; ++a) x[a] = x[a] - delta[a];   // :23
    memcpy(in_x, x, 3 * sizeof(double));                  // :24
    counter = counter + 1;
  }
}

extern "C" void compute_pos(double* to_c, double* poses, double* img_positions, double* param, double* pos) {   // compute_pos.c:30-52
  param[0] = img_positions[KDIM * 2 - 2];
  param[1] = img_positions[KDIM * 2 - 1];
  param[2] = 0.1;
  oracle_gauss_newton(param, poses, img_positions);
  double w = poses[KDIM * 7 - 4], x = poses[KDIM * 7 - 3], y = poses[KDIM * 7 - 2], z = poses[KDIM * 7 - 1];
  const double n = sqrt(w * w + x * x + y * y + z * z);   // q.normalized()
  w /= n; x /= n; y /= n; z /= n;
  // Eigen::QuaternionBase::toRotationMatrix
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  const double R[3][3] = {{1 - (tyy + tzz), txy - twz, txz + twy}, {txy + twz, 1 - (txx + tzz), tyz - twx}, {txz - twy, tyz + twx, 1 - (txx + tyy)}};
  double rot[3][3];   // R * RC^T, RC = to_c row-major (:45-46)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) rot[i][j] = R[i][0] * to_c[j * 3 + 0] + R[i][1] * to_c[j * 3 + 1] + R[i][2] * to_c[j * 3 + 2];
  pos[0] = param[0] / param[2];
  pos[1] = param[1] / param[2];
  pos[2] = 1.0 / param[2];
  double out[3];
  for (int i = 0; i < 3; ++i) out[i] = rot[i][0] * pos[0] + rot[i][1] * pos[1] + rot[i][2] * pos[2] + poses[KDIM * 7 - 7 + i];   // :51
  memcpy(pos, out, 3 * sizeof(double));
}

static bool oracle_sane(double track[K + 1][5]) {   // feature_handler.c:1-21
  double diffs_x[K - 1];
  double diffs_y[K - 1];
  int i;
  for (i = 0; i < K - 1; i++) {
    diffs_x[i] = fabs(track[i + 2][2] - track[i + 1][2]);
    diffs_y[i] = fabs(track[i + 2][3] - track[i + 1][3]);
  }
  for (i = 1; i < K - 1; i++) {
    if (((diffs_x[i] > 0.05 || diffs_x[i - 1] > 0.05) && (diffs_x[i] > 2 * diffs_x[i - 1] || diffs_x[i] < .5 * diffs_x[i - 1])) ||
        ((diffs_y[i] > 0.05 || diffs_y[i - 1] > 0.05) && (diffs_y[i] > 2 * diffs_y[i - 1] || diffs_y[i] < .5 * diffs_y[i - 1]))) {
      return false;
    }
  }
  return true;
}
extern "C" int sane(double* track) { return oracle_sane((double(*)[5])track) ? 1 : 0; }

// feature_handler.c:23-56 with the table sizes (3000 features, 6000 tracks in the template) as arguments; works in
// place instead of through the template's stack copies (:25-28,55).
extern "C" void merge_features_n(double* tracks, double* features, long long* empty_idxs, int n_features, int n_tracks) {
  double (*feature_arr)[5] = (double(*)[5])features;
  double (*track_arr)[K + 1][5] = (double(*)[K + 1][5])tracks;
  (void)n_tracks;
  int match;
  int empty_idx = 0;
  int idx;
  for (int i = 0; i < n_features; i++) {
    match = (int)feature_arr[i][4];
    if (track_arr[match][0][1] == match && track_arr[match][0][2] == 0) {
      track_arr[match][0][0] = track_arr[match][0][0] + 1;
      track_arr[match][0][1] = feature_arr[i][1];
      track_arr[match][0][2] = 1;
      idx = (int)track_arr[match][0][0];
      memcpy(track_arr[match][idx], feature_arr[i], 5 * sizeof(double));
      if (idx == K) {
        track_arr[match][0][3] = 1;             // label complete
        if (oracle_sane(track_arr[match])) {
          track_arr[match][0][4] = 1;           // label valid
        }
      }
    } else {                                    // gen new track with this feature
      track_arr[empty_idxs[empty_idx]][0][0] = 1;
      track_arr[empty_idxs[empty_idx]][0][1] = feature_arr[i][1];
      track_arr[empty_idxs[empty_idx]][0][2] = 1;
      memcpy(track_arr[empty_idxs[empty_idx]][1], feature_arr[i], 5 * sizeof(double));
      empty_idx = empty_idx + 1;
    }
  }
}
extern "C" void merge_features(double* tracks, double* features, long long* empty_idxs) { merge_features_n(tracks, features, empty_idxs, 3000, 6000); }
