// rednose_b200 -- CTA-per-filter predict / update for large error states (EDIM > 32), e.g. an MSCKF
// with 10 cloned camera poses: DIM 93, EDIM 82, MEDIM 22, feature-track kind ZDIM 20 / EADIM 3.
//
// The covariance (EDIM^2 doubles = 53.8 kB at EDIM 82) does not fit a warp's registers, so one CTA owns one
// filter with P resident in shared memory (odd leading dimension: rows and columns are both conflict-free)
// and thread j working on column j.  A step is two launches:
//
//   ekf_leaf_thread   thread-per-filter: x_pred = f(x, dt), the F value slots, y = z - h(x_pred), the H_err
//                     value slots and (feature kinds) He = dh/d(ea) -> a small per-filter workspace in HBM.
//                     (Same reason as the warp kernel's phase A: scalar generated code wants one filter per
//                     lane, not one filter per CTA.)
//   ekf_step_cta      CTA-per-filter: P <- F P F^T + dt Q on the MEDIM main block only (ekf_c.c:23-26),
//                     left-null-space projection of (y, H_err, R) with Householder reflectors of He
//                     (ekf_c.c:66-76 uses fullPivLu().kernel(); x and P do not depend on the basis),
//                     Mahalanobis gate (ekf_c.c:88-94), S = LDL^T, gain, P <- P - (HP)^T S^-1 (HP), inject.
//
// Exploited structure: H_err of a feature kind has 6 non-zeros per row (one clone each), so H_err P goes
// through the generated sparse KIND::Herr_apply, and the projection is applied to the small ZDIM-vectors
// (Q^T (H_err P) = (Q^T H_err) P) instead of densifying H_err.
#pragma once
#include "ekf_common.cuh"
#include "ekf_warp.cuh"
#include "ekf_rts_mma.cuh"   // dmma884

namespace rnb {

template <class M, class K>
struct CtaWs {  // per-filter workspace record (doubles)
  static constexpr int OFF_X = 0;
  static constexpr int OFF_FV = OFF_X + M::DIM;
  static constexpr int OFF_DT = OFF_FV + (M::NF > 0 ? M::NF : 1);
  static constexpr int OFF_Y = OFF_DT + 1;
  static constexpr int OFF_HV = OFF_Y + K::ZDIM;
  static constexpr int OFF_HE = OFF_HV + (K::NH > 0 ? K::NH : 1);
  // feature kinds: Householder vectors V [NR][Z] of He (NR = EADIM) followed by the NR scalars beta
  static constexpr int SIZE = OFF_HE + (K::HAS_HE ? K::ZDIM * K::EADIM + K::EADIM : 0);
};

// ------------------------------------------------------------------ leaf kernel ---
template <class M, class K, bool PRED, bool UPD>
__global__ void __launch_bounds__(64) ekf_leaf_thread(const StepArgs<M::NG> a, int o, double* __restrict__ ws_all) {
  constexpr int D = M::DIM, Z = K::ZDIM;
  using W = CtaWs<M, K>;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  double* ws = ws_all + b * W::SIZE;
  const long long fb = a.idx ? (long long)a.idx[b] : b;   // filter this entry works on
  double x[D];
  if (PRED || o == 0) {
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = a.x[fb * D + i];
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = ws[W::OFF_X + i];
  }
  if constexpr (PRED) {
    const double dt = a.dt_arr ? a.dt_arr[b] : a.dt;
    double xn[D];
    double fv[M::NF > 0 ? M::NF : 1];
    M::predict_leaf(x, dt, a.gv, xn, fv);
#pragma unroll
    for (int i = 0; i < (M::NF > 0 ? M::NF : 1); ++i) ws[W::OFF_FV + i] = fv[i];
    ws[W::OFF_DT] = dt;
#pragma unroll
    for (int i = 0; i < D; ++i) ws[W::OFF_X + i] = xn[i];
    if ((a.flags & FLAG_NORM_AFTER_PREDICT) && a.n_quat > 0) {
      for (int q = 0; q < a.n_quat; ++q) normalize4(ws + W::OFF_X + a.quat_idx[q]);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = ws[W::OFF_X + i];
    if (a.hx_pred) {
#pragma unroll
      for (int i = 0; i < D; ++i) a.hx_pred[fb * D + i] = x[i];
    }
    if (!UPD) {
#pragma unroll
      for (int i = 0; i < D; ++i) a.x[fb * D + i] = x[i];
    }
  } else if (o == 0) {
#pragma unroll
    for (int i = 0; i < D; ++i) ws[W::OFF_X + i] = x[i];
  }
  if constexpr (UPD) {
    const long long bo = b * a.n_obs + o;
    const double* ea = a.ea ? a.ea + bo * a.ea_dim : nullptr;
    double hx[Z];
    K::obs_leaf(x, ea, a.gv, hx, *reinterpret_cast<double(*)[K::NH > 0 ? K::NH : 1]>(ws + W::OFF_HV));
#pragma unroll
    for (int i = 0; i < Z; ++i) ws[W::OFF_Y + i] = a.z[bo * Z + i] - hx[i];
    if constexpr (K::HAS_HE) {
      // Householder QR of He (Z x EA) done here, one filter per thread: Q = H_1 .. H_EA, the left null space of
      // He is spanned by the last Z - EA columns of Q (ekf_c.c:66-76 takes fullPivLu().kernel() of He^T instead)
      constexpr int EA = K::EADIM;
      double He[Z * EA];
      K::He_dense(x, ea, a.gv, He);
#pragma unroll
      for (int r = 0; r < EA; ++r) {
        double nrm = 0.0;
#pragma unroll
        for (int i = r; i < Z; ++i) nrm = fma(He[i * EA + r], He[i * EA + r], nrm);
        nrm = sqrt(nrm);
        const double a0 = He[r * EA + r];
        const double alpha = (a0 >= 0.0) ? -nrm : nrm;
        double v[Z];
#pragma unroll
        for (int i = 0; i < Z; ++i) v[i] = (i < r) ? 0.0 : He[i * EA + r];
        v[r] = a0 - alpha;
        double vn = 0.0;
#pragma unroll
        for (int i = r; i < Z; ++i) vn = fma(v[i], v[i], vn);
        const double beta = (vn > 0.0) ? 2.0 / vn : 0.0;
#pragma unroll
        for (int c = r + 1; c < EA; ++c) {
          double w = 0.0;
#pragma unroll
          for (int i = r; i < Z; ++i) w = fma(v[i], He[i * EA + c], w);
          w *= beta;
#pragma unroll
          for (int i = r; i < Z; ++i) He[i * EA + c] = fma(-w, v[i], He[i * EA + c]);
        }
#pragma unroll
        for (int i = 0; i < Z; ++i) ws[W::OFF_HE + r * Z + i] = v[i];
        ws[W::OFF_HE + EA * Z + r] = beta;
      }
    }
  }
}

// ------------------------------------------------------------------ CTA kernel ---
template <class M, class K>
struct CtaSmem {
  static constexpr int E = M::EDIM, Z = K::ZDIM, Y = K::YDIM;
  static constexpr int LD = E | 1;                 // odd: row and column sweeps are both conflict-free
  static constexpr int HL = (Z + 3) & ~3;          // leading dimension of the HP / W buffers ([column][c]); multiple of the mma k = 4
  static constexpr int SL = Z | 1;
  double P[E * LD];
  double HP[E * HL];                               // (H_err P)[c][k] stored as HP[k * HL + c]
  double W[E * HL];                                // S^-1 (H_err P), same layout (B operand of the rank-m update)
  double S[Z * SL];                                // H_err P H_err^T (projected in place)
  double Rm[Z * SL];                               // R (projected in place)
  double LT[Z * SL];                               // LDL^T factor of S, transposed
  double dinv[Z];
  double V[(K::HAS_HE ? K::EADIM : 1) * Z];        // Householder vectors of He
  double beta[K::HAS_HE ? K::EADIM : 1];
  double y[Z];
  double hv[K::NH > 0 ? K::NH : 1];
  double fv[M::NF > 0 ? M::NF : 1];
  double x[M::DIM], xo[M::DIM], dx[E];
  double dt;
  int gated;
};

template <int Z>
struct SmemCol {  // read-only view of one shared-memory column / row as a vector
  const double* p;
  int stride;
  __device__ __forceinline__ double operator[](int i) const { return p[i * stride]; }
};

// apply Q^T = H_r ... H_1 (Householder reflectors, vectors V[r][:], scalars beta[r]) to a Z-vector in registers
template <int Z, int NR>
__device__ __forceinline__ void apply_reflectors(const double* V, const double* beta, double (&u)[Z]) {
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double w = 0.0;
#pragma unroll
    for (int i = 0; i < Z; ++i) w = fma(V[r * Z + i], u[i], w);
    w *= beta[r];
#pragma unroll
    for (int i = 0; i < Z; ++i) u[i] = fma(-w, V[r * Z + i], u[i]);
  }
}

template <class M>
constexpr int cta_tpg() { return ((M::EDIM + 31) / 32) * 32; }   // threads per group: one per column
constexpr int CTA_GROUPS = 2;                                      // groups split the rows of the rank-m covariance update

template <class M, class K, bool PRED, bool UPD>
__global__ void __launch_bounds__(cta_tpg<M>() * CTA_GROUPS, 2) ekf_step_cta(const StepArgs<M::NG> a, int o, const double* __restrict__ ws_all) {
  constexpr int D = M::DIM, E = M::EDIM, ME = M::MEDIM, Z = K::ZDIM, Y = K::YDIM, NR = Z - Y;
  using SM = CtaSmem<M, K>;
  using W = CtaWs<M, K>;
  constexpr int LD = SM::LD, HL = SM::HL, SL = SM::SL;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  SM& s = *reinterpret_cast<SM*>(smem_raw);
  constexpr int TPG = cta_tpg<M>();
  const int tid = threadIdx.x, nth = blockDim.x;
  const int grp = tid / TPG, col = tid - grp * TPG;   // group 0 does the per-column work, all groups share the big row loops
  const long long b = blockIdx.x;
  const double* ws = ws_all + b * W::SIZE;
  const long long fb = a.idx ? (long long)a.idx[b] : b;   // filter this entry works on
  double* Pg = a.P + fb * (long long)(E * E);
  const bool own = col < E;            // this thread is attached to column `col`
  const bool own0 = own && grp == 0;   // ... and is the one that writes per-column results

  // ---- stage: covariance tile (coalesced, re-pitched), leaf values ----
  {
    // all loads of the tile are issued before the first shared-memory store (one global round trip, not one per pass)
    constexpr int NIT = (E * E + TPG * CTA_GROUPS - 1) / (TPG * CTA_GROUPS);
    double v[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = it * nth + tid;
      v[it] = (idx < E * E) ? Pg[idx] : 0.0;
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int idx = it * nth + tid;
      const int i = idx / E, j = idx - i * E;
      if (idx < E * E) s.P[i * LD + j] = v[it];
    }
  }
  for (int i = tid; i < D; i += nth) s.x[i] = ws[W::OFF_X + i];
  if constexpr (PRED) {
    for (int i = tid; i < (M::NF > 0 ? M::NF : 1); i += nth) s.fv[i] = ws[W::OFF_FV + i];
    if (tid == 0) s.dt = ws[W::OFF_DT];
  }
  if constexpr (UPD) {
    for (int i = tid; i < (K::NH > 0 ? K::NH : 1); i += nth) s.hv[i] = ws[W::OFF_HV + i];
    for (int i = tid; i < Z; i += nth) s.y[i] = ws[W::OFF_Y + i];
    if constexpr (K::HAS_HE) {
      for (int i = tid; i < Z * K::EADIM; i += nth) s.V[i] = ws[W::OFF_HE + i];
      for (int i = tid; i < K::EADIM; i += nth) s.beta[i] = ws[W::OFF_HE + Z * K::EADIM + i];
    }
    const double* Rg = a.R + ((a.flags & FLAG_SHARED_R) ? 0 : (b * a.n_obs + o) * (long long)(Z * Z));
    for (int idx = tid; idx < Z * Z; idx += nth) s.Rm[(idx / Z) * SL + idx % Z] = Rg[idx];
    if (tid == 0) s.gated = 0;
  }
  __syncthreads();

  // =============================== predict: P <- F P F^T + dt Q (main block) ===============================
  if constexpr (PRED) {
    if (own0) {  // column col: (F P)[:, col]
      double v[ME];
#pragma unroll
      for (int i = 0; i < ME; ++i) v[i] = s.P[i * LD + col];
      M::F_apply(s.fv, v);
      M::frows_scatter(v, &s.P[col], LD);  // only the rows of F that differ from the identity change
    }
    __syncthreads();
    if (own0) {  // row col: ((F P) F^T)[col, :]
      double v[ME];
#pragma unroll
      for (int k = 0; k < ME; ++k) v[k] = s.P[col * LD + k];
      M::F_apply(s.fv, v);
      M::frows_scatter(v, &s.P[col * LD], 1);
    }
    __syncthreads();
    {
      const double dt = s.dt;
      if (a.flags & FLAG_Q_DIAG) {
        if (own0) s.P[col * LD + col] += dt * __ldg(a.Q + col * E + col);
      } else {
        for (int idx = tid; idx < E * E; idx += nth) {
          const int i = idx / E, j = idx - i * E;
          s.P[i * LD + j] = fma(dt, __ldg(a.Q + idx), s.P[i * LD + j]);
        }
      }
    }
    __syncthreads();
    if (a.hP_pred) {
      double* Hg = a.hP_pred + fb * (long long)(E * E);
      for (int idx = tid; idx < E * E; idx += nth) Hg[idx] = s.P[(idx / E) * LD + idx % E];
    }
  }

  if constexpr (UPD) {
    // ---- HP_raw[:, col] = H_err P[:, col] (sparse); every group keeps its own copy in registers ----
    double hp[Z];
    if (own) {
      SmemCol<Z> pc{&s.P[col], LD};
      K::Herr_apply(s.hv, pc, hp);
      if (grp == 0) {
#pragma unroll
        for (int c = 0; c < Z; ++c) s.HP[col * HL + c] = hp[c];  // unprojected, for S
      }
    }
    __syncthreads();
    // S_raw[:, t] = H_err (HP_raw[t, :])^T   (P symmetric)
    if (tid < Z) {
      SmemCol<Z> hr{&s.HP[tid], HL};
      double sc[Z];
      K::Herr_apply(s.hv, hr, sc);
#pragma unroll
      for (int c = 0; c < Z; ++c) s.S[c * SL + tid] = sc[c];
    }
    __syncthreads();
    if constexpr (K::HAS_HE) {
      // project S and R on both sides, y and HP on the left; S and R are handled by two different warps
      const int w = tid >> 5, t = tid & 31;
      if (w < 2 && t < Z) {  // columns
        double* Mx = (w == 0) ? s.S : s.Rm;
        double u[Z];
#pragma unroll
        for (int c = 0; c < Z; ++c) u[c] = Mx[c * SL + t];
        apply_reflectors<Z, NR>(s.V, s.beta, u);
#pragma unroll
        for (int c = 0; c < Z; ++c) Mx[c * SL + t] = u[c];
      }
      if (own) apply_reflectors<Z, NR>(s.V, s.beta, hp);
      __syncthreads();
      if (w < 2 && t < Z) {  // rows
        double* Mx = (w == 0) ? s.S : s.Rm;
        double u[Z];
#pragma unroll
        for (int c = 0; c < Z; ++c) u[c] = Mx[t * SL + c];
        apply_reflectors<Z, NR>(s.V, s.beta, u);
#pragma unroll
        for (int c = 0; c < Z; ++c) Mx[t * SL + c] = u[c];
      }
      if (w == 2 && t == 0) {  // the innovation
        double u[Z];
#pragma unroll
        for (int c = 0; c < Z; ++c) u[c] = s.y[c];
        apply_reflectors<Z, NR>(s.V, s.beta, u);
#pragma unroll
        for (int c = 0; c < Z; ++c) s.y[c] = u[c];
      }
      __syncthreads();
    }
    // from here on only the trailing Y x Y block / Y entries are used (offset NR)
    if (own0) {
#pragma unroll
      for (int c = 0; c < Y; ++c) s.HP[col * HL + c] = hp[NR + c];
    }

    // ---- factor S = S_raw + R (warp 0, lane j = column j), gate, refactor if gated ----
    for (int pass = 0; pass < (K::MAHA ? 2 : 1); ++pass) {
      if (tid < 32) {
        const int j = tid < Y ? tid : 0;
        double A[Y];
        const double rs = (K::MAHA && pass == 1) ? 1.0e16 : 1.0;  // ekf_c.c:92
#pragma unroll
        for (int i = 0; i < Y; ++i) A[i] = s.S[(NR + i) * SL + NR + j] + rs * s.Rm[(NR + i) * SL + NR + j];
#pragma unroll 1
        for (int kk = 0; kk < Y; ++kk) {
          // lane kk publishes its (unscaled) column; every lane needs c[lane] (= its own A[kk], symmetry) and c[i]
          if (tid == kk) {
#pragma unroll
            for (int i = 0; i < Y; ++i) s.LT[kk * SL + i] = A[i];
          }
          __syncwarp();
          const double di = 1.0 / s.LT[kk * SL + kk];
          if (tid == 0) s.dinv[kk] = di;
          const double cj = s.LT[kk * SL + j] * di;
#pragma unroll
          for (int i = 0; i < Y; ++i)
            if (i > kk) A[i] = fma(-s.LT[kk * SL + i], cj, A[i]);
        }
        __syncwarp();
        if (K::MAHA && pass == 0 && tid == 0) {
          double u[Y];
#pragma unroll
          for (int i = 0; i < Y; ++i) u[i] = s.y[NR + i];
#pragma unroll
          for (int kk = 0; kk < Y; ++kk) {
            const double uk = u[kk] * s.dinv[kk];
#pragma unroll
            for (int i = kk + 1; i < Y; ++i) u[i] = fma(-s.LT[kk * SL + i], uk, u[i]);
          }
          double d = 0.0;   // y^T S^-1 y = sum_k u_k^2 / D_k  (u = L^-1 y)
#pragma unroll
          for (int i = 0; i < Y; ++i) d = fma(u[i] * s.dinv[i], u[i], d);
          s.gated = d > K::MAHA_THRESH;
        }
      }
      __syncthreads();
      if (!(K::MAHA && pass == 0 && s.gated)) break;
    }

    // ---- gain row: w = S^-1 HP[:, col] (every group, in registers); dx; covariance ----
    double w[Y];
    if (own) {
#pragma unroll
      for (int c = 0; c < Y; ++c) w[c] = hp[NR + c];
#pragma unroll
      for (int kk = 0; kk < Y; ++kk) {
        const double wk = w[kk] * s.dinv[kk];
#pragma unroll
        for (int i = kk + 1; i < Y; ++i) w[i] = fma(-s.LT[kk * SL + i], wk, w[i]);
      }
#pragma unroll
      for (int i = 0; i < Y; ++i) w[i] *= s.dinv[i];
      const volatile double* LTv = s.LT;
#pragma unroll
      for (int kk = Y - 2; kk >= 0; --kk) {
        double acc = 0.0;
#pragma unroll
        for (int i = kk + 1; i < Y; ++i) acc = fma(LTv[kk * SL + i], w[i], acc);
        w[kk] = fma(-acc, s.dinv[kk], w[kk]);
      }
      if (grp == 0) {
        double dxl = 0.0;
#pragma unroll
        for (int c = 0; c < Y; ++c) dxl = fma(w[c], s.y[NR + c], dxl);
        s.dx[col] = dxl;
      }
    }
    // ---- P -= HP^T W on the FP64 tensor path: 8 x 8 output tiles, k = Y padded to a multiple of 4 ----
    if (own0) {
#pragma unroll
      for (int c = 0; c < HL; ++c) s.W[col * HL + c] = (c < Y) ? w[c < Y ? c : 0] : 0.0;
#pragma unroll
      for (int c = Y; c < HL; ++c) s.HP[col * HL + c] = 0.0;
    }
    __syncthreads();  // HP (projected) and W complete in shared memory
    {
      constexpr int NTE = (E + 7) / 8, NKY = HL / 4;
      const int lane = tid & 31, warp = tid >> 5, nwarps = nth >> 5;
      const int fg = lane >> 2, ft = lane & 3;
      for (int tile = warp; tile < NTE * NTE; tile += nwarps) {
        const int mi = tile / NTE, ni = tile - mi * NTE;
        const int r = mi * 8 + fg, c = ni * 8 + 2 * ft, n = ni * 8 + fg;
        double c0 = (r < E && c < E) ? s.P[r * LD + c] : 0.0;
        double c1 = (r < E && c + 1 < E) ? s.P[r * LD + c + 1] : 0.0;
#pragma unroll
        for (int kq = 0; kq < NKY; ++kq) {
          const double av = (r < E) ? -s.HP[r * HL + kq * 4 + ft] : 0.0;   // A[m][k] = -(HP)^T
          const double bv = (n < E) ? s.W[n * HL + kq * 4 + ft] : 0.0;     // B[k][n] = W
          dmma884(c0, c1, av, bv);
        }
        if (r < E && c < E) s.P[r * LD + c] = c0;
        if (r < E && c + 1 < E) s.P[r * LD + c + 1] = c1;
      }
    }
    // state injection (every thread evaluates the small generated function; identical values)
    M::err_fun(s.x, s.dx, a.gv, s.xo);
    __syncthreads();
    if ((a.flags & FLAG_NORM_AFTER_UPDATE) && a.n_quat > 0) {
      if (tid == 0) for (int q = 0; q < a.n_quat; ++q) normalize4(s.xo + a.quat_idx[q]);
      __syncthreads();
    }
    const bool last = (o == a.n_obs - 1);
    for (int i = tid; i < D; i += nth) {
      if (last) a.x[fb * D + i] = s.xo[i];
      const_cast<double*>(ws_all)[b * W::SIZE + W::OFF_X + i] = s.xo[i];  // next observation of this batch starts here
      if (last && a.hx_filt) a.hx_filt[fb * D + i] = s.xo[i];
    }
    // innovation overwrites z (ekf_c.c:120): the first YDIM entries
    for (int i = tid; i < Y; i += nth) a.z[(b * a.n_obs + o) * Z + i] = s.y[NR + i];
    if (last && a.hP_filt) {
      double* Hg = a.hP_filt + fb * (long long)(E * E);
      for (int idx = tid; idx < E * E; idx += nth) Hg[idx] = s.P[(idx / E) * LD + idx % E];
    }
  }

  __syncthreads();
  for (int idx = tid; idx < E * E; idx += nth) {
    const int i = idx / E, j = idx - i * E;
    Pg[idx] = s.P[i * LD + j];
  }
}

// per-instantiation device workspace
inline double* cta_workspace(size_t doubles) {
  static double* buf = nullptr;
  static size_t cap = 0;
  if (doubles > cap) {
    if (buf) cudaFree(buf);
    buf = nullptr; cap = 0;
    if (!check(cudaMalloc(&buf, doubles * sizeof(double)), "cudaMalloc(cta workspace)")) return nullptr;
    cap = doubles;
  }
  return buf;
}

template <class M, class K, bool PRED, bool UPD>
inline void launch_step_cta(const StepArgs<M::NG>& a, cudaStream_t st) {
  using W = CtaWs<M, K>;
  double* ws = cta_workspace((size_t)a.B * W::SIZE);
  if (!ws) return;
  constexpr size_t smem = sizeof(CtaSmem<M, K>);
  static bool configured = false;
  if (!configured) {
    check(cudaFuncSetAttribute(ekf_step_cta<M, K, PRED, UPD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attribute");
    check(cudaFuncSetAttribute(ekf_step_cta<M, K, false, UPD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attribute");
    configured = true;
  }
  constexpr int threads = cta_tpg<M>() * CTA_GROUPS;
  const int n_obs = UPD ? a.n_obs : 1;
  for (int o = 0; o < n_obs; ++o) {
    const unsigned lgrid = (unsigned)((a.B + 63) / 64);
    if (o == 0) {
      ekf_leaf_thread<M, K, PRED, UPD><<<lgrid, 64, 0, st>>>(a, o, ws);
      ekf_step_cta<M, K, PRED, UPD><<<(unsigned)a.B, threads, smem, st>>>(a, o, ws);
    } else {
      ekf_leaf_thread<M, K, false, UPD><<<lgrid, 64, 0, st>>>(a, o, ws);
      ekf_step_cta<M, K, false, UPD><<<(unsigned)a.B, threads, smem, st>>>(a, o, ws);
    }
  }
}

}  // namespace rnb
