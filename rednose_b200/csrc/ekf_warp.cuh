// rednose_b200 -- warp-per-filter fused predict+update kernel (6 < EDIM <= 32,
// e.g. live_kf: DIM 23 / EDIM 22, examples/live_kf.py:97-124).
//
// Mapping.  One warp owns one filter.  Lane j holds COLUMN j of the symmetric
// covariance P in registers (EDIM doubles).  With that ownership
//   * (F P)[:,j]   = F * P[:,j]            is lane-local  (generated sparse MODEL::F_apply)
//   * (H P)[:,j]   = Herr * P[:,j]         is lane-local  (generated sparse KIND::Herr_apply)
//   * W[:,j]       = S^-1 (H P)[:,j]       is lane-local  (row j of the gain K)
//   * P'[:,j]      = P[:,j] - (H P)^T W[:,j]               needs (H P) broadcast: ZDIM*EDIM doubles
//                                                           through per-warp shared memory
//   * F P F^T needs a transposition only for the few rows of F that differ from
//     the identity (9 of 22 for live_kf): column j of F P F^T = F * (row j of F P)^T,
//     and row j of F P equals column j of P (symmetry) unless j is such a row.
//     Those rows go through a [NFROWS][33] shared-memory exchange.
// The small state x, the innovation and S (ZDIM x ZDIM) are warp-uniform: every
// lane evaluates the generated leaf code redundantly from shared memory.
//
// Reference semantics: ekf_c.c:8-33 (predict), :37-121 (update, He==NULL path),
// normalisation ekf_sym.cc:69-77,207,213.  P is assumed symmetric on entry (it is a
// covariance; the reference's own arithmetic keeps it symmetric to rounding).
#pragma once
#include "ekf_common.cuh"

namespace rnb {

constexpr int WARPS_PER_CTA = 4;

template <class M, int Z>
struct WarpScratch {
  static constexpr int DP = (M::DIM + 1) & ~1;
  double x[2][DP];                                  // state, double buffered (leaf in -> out)
  double ex[(M::NFROWS > 0 ? M::NFROWS : 1) * 33];  // row exchange for F P F^T
  double hp[Z * 32];                                // (H P)[c][k]
  double dx[32];                                    // error-state correction K y
};

template <class M>
__device__ __forceinline__ void warp_normalize(double* xs, const StepArgs<M::NG>& a, int lane) {
  for (int q = 0; q < a.n_quat; ++q) {
    double* qp = xs + a.quat_idx[q];
    const double n = sqrt(qp[0] * qp[0] + qp[1] * qp[1] + qp[2] * qp[2] + qp[3] * qp[3]);
    __syncwarp();
    if (lane < 4) qp[lane] = qp[lane] / n;
    __syncwarp();
  }
}

template <class M, class K, bool PRED, bool UPD>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) ekf_step_warp(const StepArgs<M::NG> a) {
  constexpr int D = M::DIM, E = M::EDIM, Z = K::ZDIM;
  static_assert(E <= 32, "warp-per-filter kernel needs EDIM <= 32");
  __shared__ WarpScratch<M, Z> s_all[WARPS_PER_CTA];

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * WARPS_PER_CTA + wib;
  if (b >= a.B) return;  // whole warp exits together
  WarpScratch<M, Z>& s = s_all[wib];
  const bool act = lane < E;
  const int col = act ? lane : 0;

  // ---- loads: column `lane` of P (coalesced 8*E-byte rows), state into smem ----
  double p[E];
  {
    const double* Pg = a.P + b * (long long)(E * E) + col;
#pragma unroll
    for (int i = 0; i < E; ++i) p[i] = Pg[i * E];
  }
  for (int i = lane; i < D; i += 32) s.x[0][i] = a.x[b * D + i];
  __syncwarp();
  double* xs = s.x[0];
  double* xo = s.x[1];

  if constexpr (PRED) {
    const double dt = a.dt_arr ? a.dt_arr[b] : a.dt;
    double fv[M::NF > 0 ? M::NF : 1];
    M::predict_leaf(xs, dt, a.gv, xo, fv);  // all lanes write identical values to xo
    { double* t = xs; xs = xo; xo = t; }

    // m = F p  (lane-local);  rows of F P that are not rows of P go to the exchange
    double m[E];
#pragma unroll
    for (int i = 0; i < E; ++i) m[i] = p[i];
    M::F_apply(fv, m);
    if constexpr (M::NFROWS > 0) {
      M::frows_store(m, s.ex + lane, 33);
      __syncwarp();
      const bool in_rf = (M::FROW_MASK >> lane) & 1u;
      const int slot = __popc(M::FROW_MASK & ((1u << lane) - 1u));
      const double* row = s.ex + (in_rf ? slot : 0) * 33;
      double r[E];
#pragma unroll
      for (int i = 0; i < E; ++i) r[i] = row[i];
      M::F_apply(fv, r);
#pragma unroll
      for (int i = 0; i < E; ++i) p[i] = in_rf ? r[i] : m[i];
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i) p[i] = m[i];
    }
    {
      const double* Qg = a.Q + col;
#pragma unroll
      for (int i = 0; i < E; ++i) p[i] = fma(dt, __ldg(Qg + i * E), p[i]);
    }
    __syncwarp();
    if (a.flags & FLAG_NORM_AFTER_PREDICT) warp_normalize<M>(xs, a, lane);
    if (a.hx_pred) for (int i = lane; i < D; i += 32) a.hx_pred[b * D + i] = xs[i];
    if (a.hP_pred && act) {
      double* Hg = a.hP_pred + b * (long long)(E * E) + col;
#pragma unroll
      for (int i = 0; i < E; ++i) Hg[i * E] = p[i];
    }
  }

  if constexpr (UPD) {
    for (int o = 0; o < a.n_obs; ++o) {
      const long long bo = b * a.n_obs + o;
      // z, R: identical address in every lane -> one broadcast transaction each
      double y[Z], R[Z][Z];
#pragma unroll
      for (int i = 0; i < Z; ++i) y[i] = a.z[bo * Z + i];
#pragma unroll
      for (int i = 0; i < Z; ++i)
#pragma unroll
        for (int j = 0; j < Z; ++j) R[i][j] = a.R[bo * (Z * Z) + i * Z + j];
      const double* ea = a.ea ? a.ea + bo * a.ea_dim : nullptr;

      double hx[Z];
      double hv[K::NH > 0 ? K::NH : 1];
      K::obs_leaf(xs, ea, a.gv, hx, hv);
#pragma unroll
      for (int i = 0; i < Z; ++i) y[i] -= hx[i];

      // (H P)[:,lane]
      double hp[Z];
      K::Herr_apply(hv, p, hp);
#pragma unroll
      for (int c = 0; c < Z; ++c) s.hp[c * 32 + lane] = act ? hp[c] : 0.0;
      __syncwarp();

      // S = Herr (H P)^T + R, warp-uniform
      double S[Z][Z];
#pragma unroll
      for (int i = 0; i < Z; ++i)
#pragma unroll
        for (int j = 0; j < Z; ++j) S[i][j] = 0.0;
      K::S_accum(hv, [&](int c, int k) { return s.hp[c * 32 + k]; }, S);

      LDL<Z> ldl;
      if constexpr (K::MAHA) {
        double Sg[Z][Z];
#pragma unroll
        for (int i = 0; i < Z; ++i)
#pragma unroll
          for (int j = 0; j < Z; ++j) Sg[i][j] = S[i][j] + R[i][j];
        ldl.factor(Sg);
        double u[Z];
#pragma unroll
        for (int i = 0; i < Z; ++i) u[i] = y[i];
        ldl.solve(u);
        double d = 0.0;
#pragma unroll
        for (int i = 0; i < Z; ++i) d += y[i] * u[i];
        if (d > K::MAHA_THRESH) {  // warp-uniform predicate (ekf_c.c:91-93)
#pragma unroll
          for (int i = 0; i < Z; ++i)
#pragma unroll
            for (int j = 0; j < Z; ++j) R[i][j] *= 1.0e16;
        }
      }
#pragma unroll
      for (int i = 0; i < Z; ++i)
#pragma unroll
        for (int j = 0; j < Z; ++j) S[i][j] += R[i][j];
      ldl.factor(S);

      // w = S^-1 hp : row `lane` of the Kalman gain;  dx[lane] = K[lane,:] y
      ldl.solve(hp);
      double dxl = 0.0;
#pragma unroll
      for (int c = 0; c < Z; ++c) dxl = fma(hp[c], y[c], dxl);
      s.dx[lane] = act ? dxl : 0.0;

      // P[:,lane] -= (H P)^T w
#pragma unroll
      for (int i = 0; i < E; ++i) {
        double acc = p[i];
#pragma unroll
        for (int c = 0; c < Z; ++c) acc = fma(-s.hp[c * 32 + i], hp[c], acc);
        p[i] = acc;
      }
      __syncwarp();

      M::err_fun(xs, s.dx, a.gv, xo);
      { double* t = xs; xs = xo; xo = t; }
      __syncwarp();
      if (a.flags & FLAG_NORM_AFTER_UPDATE) warp_normalize<M>(xs, a, lane);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < Z; ++i) a.z[bo * Z + i] = y[i];
      }
    }
    if (a.hx_filt) for (int i = lane; i < D; i += 32) a.hx_filt[b * D + i] = xs[i];
    if (a.hP_filt && act) {
      double* Hg = a.hP_filt + b * (long long)(E * E) + col;
#pragma unroll
      for (int i = 0; i < E; ++i) Hg[i * E] = p[i];
    }
  }

  for (int i = lane; i < D; i += 32) a.x[b * D + i] = xs[i];
  if (act) {
    double* Pg = a.P + b * (long long)(E * E) + col;
#pragma unroll
    for (int i = 0; i < E; ++i) Pg[i * E] = p[i];
  }
}

}  // namespace rnb
