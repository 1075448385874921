// rednose_b200 -- RTS backward pass, variant whose two dense n x n x n products run on the FP64 tensor path
// (mma.sync.aligned.m8n8k4.f64, SASS DMMA).  Same recursion, same factorisation / substitution code as
// ekf_rts_warp (ekf_rts.cuh); only  P_{k|N} = P_{k|k} + X^T (dP X)  changes:
//
//   * the scalar version broadcasts every row of dP and of X to all lanes from shared memory (2 wavefronts per
//     128-bit broadcast load): ~1000 of the ~2000 L1TEX wavefronts per step, the kernel's limiter (72 % of the
//     shared-memory pipe at 15 % of the HBM roofline);
//   * here dP, X and Y = dP X are read as m8n8k4 fragments (one 64-bit element per lane, 36 loads per product
//     instead of 242) and the smoothed covariance is carried between steps in accumulator-fragment layout
//     (row = lane / 4, columns 2 (lane % 4) + {0, 1} of every 8 x 8 tile), which is also the layout it is
//     loaded from / stored to global memory in (aligned 128-bit accesses).
//
// n is padded to a multiple of 8 with zeros (22 -> 24 for live_kf).  FP64 mma is IEEE fused multiply-add, so
// results agree with the scalar kernel to rounding (different summation order).
#pragma once
#include "ekf_rts.cuh"

namespace rnb {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

#ifndef RNB_RTS_PREFETCH
#define RNB_RTS_PREFETCH 1   // pull the next step's history slabs into L2 while the current step computes
#endif
#ifndef RNB_RTS_PF_EARLY
#define RNB_RTS_PF_EARLY 1   // P_{k|k} accumulator fragments loaded with the step's other global loads (one round trip)
#endif

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <class M>
struct RtsMmaScratch {
  static constexpr int N = M::MEDIM;
  static constexpr int NP = (N + 7) & ~7;             // padded extent
  static constexpr int LD = (N + 3) & ~1;             // factor buffer (as in the scalar kernel)
  static constexpr int LP = NP + 2;                   // fragment buffers: even (128-bit rows), spreads rows over banks
  alignas(16) double LT[(N * LD > NP * LP) ? N * LD : NP * LP];   // L during the solve, then Y = dP X
  alignas(16) double DP[NP * LP];                     // dP, row-major, zero padded
  alignas(16) double XS[NP * LP];                     // X, row-major, zero padded
  alignas(16) double xf[(M::DIM + 1) & ~1];
  alignas(16) double xp[(M::DIM + 1) & ~1];
  alignas(16) double xn[(M::DIM + 1) & ~1];
  alignas(16) double xt[(M::DIM + 1) & ~1];
  alignas(16) double dl[(M::EDIM + 1) & ~1];
  alignas(16) double dinv[(N + 1) & ~1];
};

template <class M>
__global__ void __launch_bounds__(RTS_WARPS * 32, RTS_MIN_CTAS) ekf_rts_warp_mma(const RtsArgs<M::NG> a) {
  constexpr int D = M::DIM, E = M::EDIM, N = M::MEDIM, D1 = M::DMAIN;
  using SC = RtsMmaScratch<M>;
  constexpr int LD = SC::LD, NP = SC::NP, LP = SC::LP, NT = NP / 8, NK = NP / 4;
  static_assert(E <= 32 && E % 2 == 0, "fragment I/O needs an even EDIM <= 32");
  __shared__ SC s_all[RTS_WARPS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * RTS_WARPS + wib;
  if (b >= a.B) return;
  SC& s = s_all[wib];
  const bool act = lane < N, actE = lane < E;
  const int col = actE ? lane : 0;
  const int fg = lane >> 2, ft = lane & 3;   // fragment coordinates: row fg, k / column pair ft
  const long long BP = a.B * (long long)(E * E), BX = a.B * (long long)D;

  auto normalize_xn = [&]() {
    for (int q = 0; q < a.n_quat; ++q) {
      double* qp = s.xn + a.quat_idx[q];
      const double nrm = sqrt(qp[0] * qp[0] + qp[1] * qp[1] + qp[2] * qp[2] + qp[3] * qp[3]);
      __syncwarp();
      if (lane < 4) qp[lane] = qp[lane] / nrm;
      __syncwarp();
    }
  };
  // element pair (r, c), (r, c+1) of tile (mi, ni) in accumulator layout
  auto frag_rc = [&](int mi, int ni, int& r, int& c) { r = mi * 8 + fg; c = ni * 8 + 2 * ft; };

  // ---- start: smoothed = predicted at T-1 (ekf_sym.py:658-659); carried in fragment layout ----
  double pn[NT * NT * 2];
  {
    const long long k = a.T - 1;
    const bool seg = a.x_term != nullptr;   // segment continuation: start from the smoothed estimate handed in
    const double* Pg = seg ? a.P_term + b * (long long)(E * E) : a.hP_pred + k * BP + b * (long long)(E * E);
    double* Po = a.Ps + k * BP + b * (long long)(E * E);
    if (!seg) for (int idx = lane; idx < E * E; idx += 32) Po[idx] = Pg[idx];
#pragma unroll
    for (int mi = 0; mi < NT; ++mi)
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) {
        int r, c; frag_rc(mi, ni, r, c);
        double2 v = make_double2(0.0, 0.0);
        if (r < N && c < N) v = *reinterpret_cast<const double2*>(Pg + r * E + c);
        pn[(mi * NT + ni) * 2] = v.x; pn[(mi * NT + ni) * 2 + 1] = v.y;
      }
    for (int i = lane; i < D; i += 32) s.xn[i] = seg ? a.x_term[b * D + i] : a.hx_pred[k * BX + b * D + i];
    __syncwarp();
    if (!seg) {
      if (a.norm_quats && a.T >= 2) normalize_xn();
      for (int i = lane; i < D; i += 32) a.xs[k * BX + b * D + i] = s.xn[i];
    }
  }

#pragma unroll 1
  for (long long k = a.T - 2; k >= 0; --k) {
    const double* Pf_b = a.hP_filt + k * BP + b * (long long)(E * E);
    const double* Pp_b = a.hP_pred + (k + 1) * BP + b * (long long)(E * E);
    const double* Pf_g = Pf_b + col;
    const double* Pp_g = Pp_b + col;
    // every global load of the step is issued here, before any dependent work (one latency round trip per step)
    double g[N], A[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { g[i] = Pf_g[i * E]; A[i] = Pp_g[i * E]; }
    // P_{k|k} once more, in accumulator layout (the C operand of the last product): fetched here with everything
    // else -- inside the product loop each tile's load sat behind the previous tile's store to Ps (may alias) and
    // paid its own L2 round trip
    double pf[NT * NT * 2];
    if constexpr (RNB_RTS_PF_EARLY) {
#pragma unroll
      for (int mi = 0; mi < NT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
          int r, c; frag_rc(mi, ni, r, c);
          double2 v = make_double2(0.0, 0.0);
          if (r < N && c < N) v = *reinterpret_cast<const double2*>(Pf_b + r * E + c);
          pf[(mi * NT + ni) * 2] = v.x; pf[(mi * NT + ni) * 2 + 1] = v.y;
        }
    }
    if constexpr (RNB_RTS_PREFETCH) {
      if (k > 0) {   // step k-1 reads P_{k-1|k-1}, P_{k|k-1}, x_{k-1|k-1}, x_{k|k-1}: one 128-byte line per lane
        constexpr int TB = E * E * (int)sizeof(double);
        const int off = (lane * 128 < TB - 8) ? lane * 128 : TB - 8;
        prefetch_l2(reinterpret_cast<const char*>(Pf_b - BP) + off);
        prefetch_l2(reinterpret_cast<const char*>(a.hP_pred + k * BP + b * (long long)(E * E)) + off);
        if (lane < 2) {
          constexpr int XB = D * (int)sizeof(double);
          const int xo = (lane * 128 < XB - 8) ? lane * 128 : XB - 8;
          prefetch_l2(reinterpret_cast<const char*>(a.hx_filt + (k - 1) * BX + b * D) + xo);
          prefetch_l2(reinterpret_cast<const char*>(a.hx_pred + k * BX + b * D) + xo);
        }
      }
    }
    // dP = P_{k+1|N} - P_{k+1|k} in fragment layout -> shared memory (zero padded)
#pragma unroll
    for (int mi = 0; mi < NT; ++mi)
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) {
        int r, c; frag_rc(mi, ni, r, c);
        double2 v = make_double2(0.0, 0.0);
        if (r < N && c < N) {
          const double2 pp = *reinterpret_cast<const double2*>(Pp_b + r * E + c);
          v.x = pn[(mi * NT + ni) * 2] - pp.x;
          v.y = pn[(mi * NT + ni) * 2 + 1] - pp.y;
        }
        *reinterpret_cast<double2*>(&s.DP[r * LP + c]) = v;
      }

    for (int i = lane; i < D; i += 32) {
      s.xf[i] = a.hx_filt[k * BX + b * D + i];
      s.xp[i] = a.hx_pred[(k + 1) * BX + b * D + i];
    }
    const double dt = a.t_per_filter ? (a.t[(k + 1) * a.B + b] - a.t[k * a.B + b]) : (a.t[k + 1] - a.t[k]);
    __syncwarp();
    {
      double fv[M::NF > 0 ? M::NF : 1];
      M::F_vals(s.xf, dt, a.gv, fv);
      M::F_apply(fv, g);   // G[:,lane] = F P_{k|k}[:,lane]
    }

    // ---- P_{k+1|k} = L D L^T, right-looking, FULLY UNROLLED: register arrays are indexed statically, only the rows
    //      below the pivot are published / updated (231 instead of 484 FMAs), and the reciprocal of the pivot -- the long
    //      pole of every step -- is started from a register shuffle BEFORE the column makes its round trip through shared
    //      memory, so the two latencies overlap instead of adding up. ----
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      const double piv = __shfl_sync(0xffffffffu, A[kk], kk);   // lane kk's diagonal entry is final
      const double di = 1.0 / piv;
      if (lane == kk) {   // publish the (unscaled) column below the pivot; rows <= kk of this buffer row are never read
#pragma unroll
        for (int i = kk + 1; i < N; ++i) s.LT[kk * LD + i] = A[i];
        s.dinv[kk] = di;
      }
      __syncwarp();
      // L[lane][kk] = c[lane] / D[kk]: read from the published column (not the lane's own A[kk]) so that L is exactly
      // the factor the substitutions below use; lanes <= kk own finished columns and update nothing
      const double cj = ((lane > kk && act) ? s.LT[kk * LD + lane] : 0.0) * di;
#pragma unroll
      for (int i = kk + 1; i < N; ++i) A[i] = fma(-s.LT[kk * LD + i], cj, A[i]);
    }
    __syncwarp();
    // ---- X[:,lane] = (L D L^T)^-1 G[:,lane];  L[i][kk] = LT[kk][i] * dinv[kk] ----
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      const double gk = g[kk] * s.dinv[kk];
#pragma unroll
      for (int i = kk + 1; i < N; ++i) g[i] = fma(-s.LT[kk * LD + i], gk, g[i]);
      asm volatile("" ::: "memory");
    }
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] *= s.dinv[i];
    const double* LTv = s.LT;
    asm volatile("" : "+l"(LTv));
#pragma unroll
    for (int kk = N - 2; kk >= 0; --kk) {
      double acc0 = 0.0, acc1 = 0.0;   // two partial sums: half the dependent-FMA chain of the dot product
#pragma unroll
      for (int i = kk + 1; i < N; i += 2) {
        acc0 = fma(LTv[kk * LD + i], g[i], acc0);
        if (i + 1 < N) acc1 = fma(LTv[kk * LD + i + 1], g[i + 1], acc1);
      }
      g[kk] = fma(-(acc0 + acc1), s.dinv[kk], g[kk]);
      asm volatile("" ::: "memory");
    }
    // g = X[:,lane]

    // ---- state ----
    M::inv_err_fun(s.xp, s.xn, a.gv, s.dl);
    __syncwarp();
    double cd = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) cd = fma(g[i], s.dl[i], cd);
    __syncwarp();
    if (act) s.dl[lane] = cd;
    __syncwarp();
    M::err_fun(s.xf, s.dl, a.gv, s.xt);
    __syncwarp();
    for (int i = lane; i < D; i += 32) s.xn[i] = (i < D1) ? s.xt[i] : s.xf[i];
    __syncwarp();
    if (a.norm_quats && k + a.k0 >= 1) normalize_xn();
    for (int i = lane; i < D; i += 32) a.xs[k * BX + b * D + i] = s.xn[i];

    // ---- X into shared memory, row-major, zero padded (lane j writes column j) ----
    if (lane < NP) {
#pragma unroll
      for (int i = 0; i < NP; ++i) s.XS[i * LP + lane] = (act && i < N) ? g[i < N ? i : 0] : 0.0;
    }
    __syncwarp();

    // X fragments (B operand of dP X, and A operand of X^T Y): xb[kq][t] = X[kq*4 + ft][t*8 + fg]
    double xb[NK][NT];
#pragma unroll
    for (int kq = 0; kq < NK; ++kq)
#pragma unroll
      for (int t = 0; t < NT; ++t) xb[kq][t] = s.XS[(kq * 4 + ft) * LP + t * 8 + fg];

    // ---- Y = dP X  (accumulator fragments), then to shared memory (L's buffer: L is dead) ----
#pragma unroll
    for (int mi = 0; mi < NT; ++mi) {
      double ad[NK];
#pragma unroll
      for (int kq = 0; kq < NK; ++kq) ad[kq] = s.DP[(mi * 8 + fg) * LP + kq * 4 + ft];
#pragma unroll
      for (int ni = 0; ni < NT; ++ni) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int kq = 0; kq < NK; ++kq) dmma884(c0, c1, ad[kq], xb[kq][ni]);
        int r, c; frag_rc(mi, ni, r, c);
        *reinterpret_cast<double2*>(&s.LT[r * LP + c]) = make_double2(c0, c1);
      }
    }
    __syncwarp();

    // ---- P_{k|N} = P_{k|k} + X^T Y : A[m][k] = X[k][m] = xb[kq][mi], B[k][n] = Y[k][n] ----
#pragma unroll
    for (int ni = 0; ni < NT; ++ni) {
      double yb[NK];
#pragma unroll
      for (int kq = 0; kq < NK; ++kq) yb[kq] = s.LT[(kq * 4 + ft) * LP + ni * 8 + fg];
#pragma unroll
      for (int mi = 0; mi < NT; ++mi) {
        int r, c; frag_rc(mi, ni, r, c);
        double c0 = 0.0, c1 = 0.0;
        if constexpr (RNB_RTS_PF_EARLY) {
          c0 = pf[(mi * NT + ni) * 2]; c1 = pf[(mi * NT + ni) * 2 + 1];
        } else if (r < N && c < N) {
          const double2 t = *reinterpret_cast<const double2*>(Pf_b + r * E + c);
          c0 = t.x; c1 = t.y;
        }
#pragma unroll
        for (int kq = 0; kq < NK; ++kq) dmma884(c0, c1, xb[kq][mi], yb[kq]);
        pn[(mi * NT + ni) * 2] = c0; pn[(mi * NT + ni) * 2 + 1] = c1;
        if (r < N && c < N) *reinterpret_cast<double2*>(a.Ps + k * BP + b * (long long)(E * E) + r * E + c) = make_double2(c0, c1);
      }
    }
    // rows / columns outside the main block keep P_{k|k} (ekf_sym.py:686 smooths the main block only)
    if constexpr (E > N) {
      double* Po = a.Ps + k * BP + b * (long long)(E * E);
      for (int idx = lane; idx < E * E; idx += 32) {
        const int i = idx / E, j = idx - i * E;
        if (i >= N || j >= N) Po[idx] = Pf_b[idx];
      }
    }
    __syncwarp();
  }
}

#ifndef RNB_RTS_MMA
#define RNB_RTS_MMA 1   // 1: dense products of the smoother on the FP64 tensor path (DMMA); 0: scalar broadcast version
#endif

template <class M>
inline void launch_rts_auto(const RtsArgs<M::NG>& a, cudaStream_t st) {
  if (a.B <= 0 || a.T <= 0) return;
  if constexpr (RNB_RTS_MMA && M::EDIM <= 32 && M::EDIM % 2 == 0 && M::MEDIM >= 8) {
    const unsigned grid = (unsigned)((a.B + RTS_WARPS - 1) / RTS_WARPS);
    ekf_rts_warp_mma<M><<<grid, RTS_WARPS * 32, 0, st>>>(a);
    check(cudaGetLastError(), "ekf_rts_mma launch");
  } else {
    launch_rts<M>(a, st);
  }
}

}  // namespace rnb
