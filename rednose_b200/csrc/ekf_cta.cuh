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

namespace rnb {

template <class M, class K>
struct CtaWs {  // per-filter workspace record (doubles)
  static constexpr int OFF_X = 0;
  static constexpr int OFF_FV = OFF_X + M::DIM;
  static constexpr int OFF_DT = OFF_FV + (M::NF > 0 ? M::NF : 1);
  static constexpr int OFF_Y = OFF_DT + 1;
  static constexpr int OFF_HV = OFF_Y + K::ZDIM;
  static constexpr int OFF_HE = OFF_HV + (K::NH > 0 ? K::NH : 1);
  static constexpr int SIZE = OFF_HE + (K::HAS_HE ? K::ZDIM * K::EADIM : 0);
};

// ------------------------------------------------------------------ leaf kernel ---
template <class M, class K, bool PRED, bool UPD>
__global__ void __launch_bounds__(64) ekf_leaf_thread(const StepArgs<M::NG> a, int o, double* __restrict__ ws_all) {
  constexpr int D = M::DIM, Z = K::ZDIM;
  using W = CtaWs<M, K>;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  double* ws = ws_all + b * W::SIZE;
  double x[D];
  if (PRED || o == 0) {
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = a.x[b * D + i];
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
      for (int i = 0; i < D; ++i) a.hx_pred[b * D + i] = x[i];
    }
    if (!UPD) {
#pragma unroll
      for (int i = 0; i < D; ++i) a.x[b * D + i] = x[i];
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
    if constexpr (K::HAS_HE) K::He_dense(x, ea, a.gv, ws + W::OFF_HE);
  }
}

// ------------------------------------------------------------------ CTA kernel ---
template <class M, class K>
struct CtaSmem {
  static constexpr int E = M::EDIM, Z = K::ZDIM, Y = K::YDIM;
  static constexpr int LD = E | 1;                 // odd: row and column sweeps are both conflict-free
  static constexpr int HL = (Z + 1) & ~1;          // leading dimension of the HP buffer ([column][c])
  static constexpr int SL = Z | 1;
  double P[E * LD];
  double HP[E * HL];                               // (H_err P)[c][k] stored as HP[k * HL + c]
  double S[Z * SL];                                // H_err P H_err^T (projected in place)
  double Rm[Z * SL];                               // R (projected in place)
  double LT[Z * SL];                               // LDL^T factor of S, transposed
  double dinv[Z];
  double V[(K::HAS_HE ? K::EADIM : 1) * Z];        // Householder vectors of He
  double beta[K::HAS_HE ? K::EADIM : 1];
  double He[K::HAS_HE ? Z * K::EADIM : 1];
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

template <class M, class K, bool PRED, bool UPD>
__global__ void __launch_bounds__(((M::EDIM + 31) / 32) * 32) ekf_step_cta(const StepArgs<M::NG> a, int o, const double* __restrict__ ws_all) {
  constexpr int D = M::DIM, E = M::EDIM, ME = M::MEDIM, Z = K::ZDIM, Y = K::YDIM, NR = Z - Y;
  using SM = CtaSmem<M, K>;
  using W = CtaWs<M, K>;
  constexpr int LD = SM::LD, HL = SM::HL, SL = SM::SL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM& s = *reinterpret_cast<SM*>(smem_raw);
  const int tid = threadIdx.x, nth = blockDim.x;
  const long long b = blockIdx.x;
  const double* ws = ws_all + b * W::SIZE;
  double* Pg = a.P + b * (long long)(E * E);
  const bool own = tid < E;

  // ---- stage: covariance tile (coalesced, re-pitched), leaf values ----
  for (int idx = tid; idx < E * E; idx += nth) {
    const int i = idx / E, j = idx - i * E;
    s.P[i * LD + j] = Pg[idx];
  }
  for (int i = tid; i < D; i += nth) s.x[i] = ws[W::OFF_X + i];
  if constexpr (PRED) {
    for (int i = tid; i < (M::NF > 0 ? M::NF : 1); i += nth) s.fv[i] = ws[W::OFF_FV + i];
    if (tid == 0) s.dt = ws[W::OFF_DT];
  }
  if constexpr (UPD) {
    for (int i = tid; i < (K::NH > 0 ? K::NH : 1); i += nth) s.hv[i] = ws[W::OFF_HV + i];
    for (int i = tid; i < Z; i += nth) s.y[i] = ws[W::OFF_Y + i];
    if constexpr (K::HAS_HE) for (int i = tid; i < Z * K::EADIM; i += nth) s.He[i] = ws[W::OFF_HE + i];
    const double* Rg = a.R + ((a.flags & FLAG_SHARED_R) ? 0 : (b * a.n_obs + o) * (long long)(Z * Z));
    for (int idx = tid; idx < Z * Z; idx += nth) s.Rm[(idx / Z) * SL + idx % Z] = Rg[idx];
    if (tid == 0) s.gated = 0;
  }
  __syncthreads();

  // =============================== predict: P <- F P F^T + dt Q (main block) ===============================
  if constexpr (PRED) {
    if (own) {  // column tid: (F P)[:, tid]
      double v[ME];
#pragma unroll
      for (int i = 0; i < ME; ++i) v[i] = s.P[i * LD + tid];
      M::F_apply(s.fv, v);
      M::frows_scatter(v, &s.P[tid], LD);  // only the rows of F that differ from the identity change
    }
    __syncthreads();
    if (own) {  // row tid: ((F P) F^T)[tid, :]
      double v[ME];
#pragma unroll
      for (int k = 0; k < ME; ++k) v[k] = s.P[tid * LD + k];
      M::F_apply(s.fv, v);
      M::frows_scatter(v, &s.P[tid * LD], 1);
    }
    __syncthreads();
    if (own) {
      const double dt = s.dt;
      if (a.flags & FLAG_Q_DIAG) {
        s.P[tid * LD + tid] += dt * __ldg(a.Q + tid * E + tid);
      } else {
        for (int i = 0; i < E; ++i) s.P[i * LD + tid] = fma(dt, __ldg(a.Q + i * E + tid), s.P[i * LD + tid]);
      }
    }
    __syncthreads();
    if (a.hP_pred) {
      double* Hg = a.hP_pred + b * (long long)(E * E);
      for (int idx = tid; idx < E * E; idx += nth) Hg[idx] = s.P[(idx / E) * LD + idx % E];
    }
  }

  if constexpr (UPD) {
    // ---- Householder QR of He (Z x NR): Q = H_1 ... H_NR, left null space = last Y columns of Q ----
    if constexpr (K::HAS_HE) {
      if (tid == 0) {
        constexpr int EA = K::EADIM;
        for (int r = 0; r < NR; ++r) {
          double nrm = 0.0;
          for (int i = r; i < Z; ++i) nrm += s.He[i * EA + r] * s.He[i * EA + r];
          nrm = sqrt(nrm);
          const double a0 = s.He[r * EA + r];
          const double alpha = (a0 >= 0.0) ? -nrm : nrm;
          for (int i = 0; i < Z; ++i) s.V[r * Z + i] = (i < r) ? 0.0 : s.He[i * EA + r];
          s.V[r * Z + r] = a0 - alpha;
          double vn = 0.0;
          for (int i = r; i < Z; ++i) vn += s.V[r * Z + i] * s.V[r * Z + i];
          s.beta[r] = (vn > 0.0) ? 2.0 / vn : 0.0;
          for (int c = r + 1; c < EA; ++c) {  // update the remaining columns of He
            double w = 0.0;
            for (int i = r; i < Z; ++i) w += s.V[r * Z + i] * s.He[i * EA + c];
            w *= s.beta[r];
            for (int i = r; i < Z; ++i) s.He[i * EA + c] -= w * s.V[r * Z + i];
          }
        }
      }
      __syncthreads();
    }

    // ---- HP_raw[:, tid] = H_err P[:, tid] (sparse), projected: hp <- (Q^T hp)[NR:] ----
    double hp[Z];
    if (own) {
      SmemCol<Z> pc{&s.P[tid], LD};
      K::Herr_apply(s.hv, pc, hp);
#pragma unroll
      for (int c = 0; c < Z; ++c) s.HP[tid * HL + c] = hp[c];  // unprojected, for S
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
      // project S and R on both sides, y and HP on the left
      if (tid < Z) {  // columns
        double u[Z], r[Z];
#pragma unroll
        for (int c = 0; c < Z; ++c) { u[c] = s.S[c * SL + tid]; r[c] = s.Rm[c * SL + tid]; }
        apply_reflectors<Z, NR>(s.V, s.beta, u);
        apply_reflectors<Z, NR>(s.V, s.beta, r);
#pragma unroll
        for (int c = 0; c < Z; ++c) { s.S[c * SL + tid] = u[c]; s.Rm[c * SL + tid] = r[c]; }
      }
      __syncthreads();
      if (tid < Z) {  // rows
        double u[Z], r[Z];
#pragma unroll
        for (int c = 0; c < Z; ++c) { u[c] = s.S[tid * SL + c]; r[c] = s.Rm[tid * SL + c]; }
        apply_reflectors<Z, NR>(s.V, s.beta, u);
        apply_reflectors<Z, NR>(s.V, s.beta, r);
#pragma unroll
        for (int c = 0; c < Z; ++c) { s.S[tid * SL + c] = u[c]; s.Rm[tid * SL + c] = r[c]; }
      }
      if (tid == Z) {  // one spare thread projects the innovation
        double u[Z];
#pragma unroll
        for (int c = 0; c < Z; ++c) u[c] = s.y[c];
        apply_reflectors<Z, NR>(s.V, s.beta, u);
#pragma unroll
        for (int c = 0; c < Z; ++c) s.y[c] = u[c];
      }
      if (own) {
        apply_reflectors<Z, NR>(s.V, s.beta, hp);
      }
      __syncthreads();
    }
    // from here on only the trailing Y x Y block / Y entries are used (offset NR)
    if (own) {
#pragma unroll
      for (int c = 0; c < Y; ++c) s.HP[tid * HL + c] = hp[NR + c];
    }

    // ---- factor S = S_raw + R (warp 0, lane j = column j), gate, refactor if gated ----
    for (int pass = 0; pass < (K::MAHA ? 2 : 1); ++pass) {
      if (tid < 32) {
        const int j = tid < Y ? tid : 0;
        double A[Y];
        const double rs = (K::MAHA && pass == 1) ? 1.0e16 : 1.0;  // ekf_c.c:92
#pragma unroll
        for (int i = 0; i < Y; ++i) A[i] = s.S[(NR + i) * SL + NR + j] + rs * s.Rm[(NR + i) * SL + NR + j];
#pragma unroll
        for (int kk = 0; kk < Y; ++kk) {
          if (tid == kk) {
            const double di = 1.0 / A[kk];
            s.dinv[kk] = di;
#pragma unroll
            for (int i = kk + 1; i < Y; ++i) s.LT[kk * SL + i] = A[i] * di;
          }
          __syncwarp();
          const double akk = A[kk];
#pragma unroll
          for (int i = kk + 1; i < Y; ++i) A[i] = fma(-s.LT[kk * SL + i], akk, A[i]);
        }
        if (K::MAHA && pass == 0 && tid == 0) {
          double u[Y];
#pragma unroll
          for (int i = 0; i < Y; ++i) u[i] = s.y[NR + i];
#pragma unroll
          for (int kk = 0; kk < Y; ++kk)
#pragma unroll
            for (int i = kk + 1; i < Y; ++i) u[i] = fma(-s.LT[kk * SL + i], u[kk], u[i]);
#pragma unroll
          for (int i = 0; i < Y; ++i) u[i] *= s.dinv[i];
#pragma unroll
          for (int kk = Y - 2; kk >= 0; --kk)
#pragma unroll
            for (int i = kk + 1; i < Y; ++i) u[kk] = fma(-s.LT[kk * SL + i], u[i], u[kk]);
          double d = 0.0;
#pragma unroll
          for (int i = 0; i < Y; ++i) d = fma(s.y[NR + i], u[i], d);
          s.gated = d > K::MAHA_THRESH;
        }
      }
      __syncthreads();
      if (!(K::MAHA && pass == 0 && s.gated)) break;
    }

    // ---- gain row: w = S^-1 HP[:, tid]; dx; covariance ----
    double w[Y];
    if (own) {
#pragma unroll
      for (int c = 0; c < Y; ++c) w[c] = hp[NR + c];
#pragma unroll
      for (int kk = 0; kk < Y; ++kk)
#pragma unroll
        for (int i = kk + 1; i < Y; ++i) w[i] = fma(-s.LT[kk * SL + i], w[kk], w[i]);
#pragma unroll
      for (int i = 0; i < Y; ++i) w[i] *= s.dinv[i];
#pragma unroll
      for (int kk = Y - 2; kk >= 0; --kk)
#pragma unroll
        for (int i = kk + 1; i < Y; ++i) w[kk] = fma(-s.LT[kk * SL + i], w[i], w[kk]);
      double dxl = 0.0;
#pragma unroll
      for (int c = 0; c < Y; ++c) dxl = fma(w[c], s.y[NR + c], dxl);
      s.dx[tid] = dxl;
    }
    __syncthreads();  // HP (projected) complete in shared memory
    if (own) {
      for (int i = 0; i < E; ++i) {
        double acc = s.P[i * LD + tid];
#pragma unroll
        for (int c = 0; c < Y; ++c) acc = fma(-s.HP[i * HL + c], w[c], acc);
        s.P[i * LD + tid] = acc;
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
      if (last) a.x[b * D + i] = s.xo[i];
      const_cast<double*>(ws_all)[b * W::SIZE + W::OFF_X + i] = s.xo[i];  // next observation of this batch starts here
      if (last && a.hx_filt) a.hx_filt[b * D + i] = s.xo[i];
    }
    // innovation overwrites z (ekf_c.c:120): the first YDIM entries
    for (int i = tid; i < Y; i += nth) a.z[(b * a.n_obs + o) * Z + i] = s.y[NR + i];
    if (last && a.hP_filt) {
      double* Hg = a.hP_filt + b * (long long)(E * E);
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
  constexpr int threads = ((M::EDIM + 31) / 32) * 32;
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
