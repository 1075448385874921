// ORACLE -- test infrastructure only.  Never linked, imported or executed by the product
// path (rednose_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it.
//
// Eigen-free restatement of the reference's numeric template rednose/templates/ekf_c.c:
//   predict()                      ekf_c.c:8-33
//   update<ZDIM, EADIM, MAHA_TEST> ekf_c.c:37-121
// The reference delegates its dense arithmetic to Eigen (un-vendored third-party dependency,
// `comma-deps-eigen`, version unpinned in pyproject.toml:13; absent from this image), so the
// template cannot be compiled as is.  This file follows it statement by statement with plain
// row-major loops, keeping the operation ORDER of the reference (dense F P F^T, dense Joseph
// form, full-pivot LU solve / kernel / inverse) so that it is a numerically faithful stand-in.
// Eigen's published algorithms restated here: FullPivLU (complete pivoting, solve(), kernel()).
//
// It is textually included where the reference pastes ekf_c.c (ekf_sym.py:207-208): inside the
// anonymous namespace of a generated <name>.cpp, after the sympy leaf functions, with DIM / EDIM /
// MEDIM #defined and f_fun, F_fun, H_mod_fun, err_fun visible.  oracle/build_ref.py assembles that
// translation unit from the output of the reference's own unmodified generator.
#include <math.h>
#include <string.h>
#include <vector>

namespace oracle_la {

// C[m x n] = A[m x k] * B[k x n], row-major
static inline void matmul(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int m, int k, int n) {
  // i-l-j order: same per-element summation order as the textbook triple loop, but the inner
  // loop streams rows so gcc -O2 vectorises it (the reference gets the same from Eigen)
  for (int i = 0; i < m; ++i) {
    double* c = C + (size_t)i * n;
    for (int j = 0; j < n; ++j) c[j] = 0.0;
    for (int l = 0; l < k; ++l) {
      const double a = A[i * k + l];
      const double* b = B + (size_t)l * n;
      for (int j = 0; j < n; ++j) c[j] += a * b[j];
    }
  }
}

// same loop with compile-time extents (what Eigen's fixed-size matrices give the reference, ekf_c.c:4-6,20-26):
// the compiler can unroll / vectorise; summation order per element is unchanged
template <int M_, int K_, int N_>
static inline void matmul_fixed(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C) {
  for (int i = 0; i < M_; ++i) {
    double* c = C + (size_t)i * N_;
    for (int j = 0; j < N_; ++j) c[j] = 0.0;
    for (int l = 0; l < K_; ++l) {
      const double a = A[i * K_ + l];
      const double* b = B + (size_t)l * N_;
      for (int j = 0; j < N_; ++j) c[j] += a * b[j];
    }
  }
}

static inline void transpose(const double* A, double* At, int m, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) At[j * m + i] = A[i * n + j];
}

// Full-pivoting LU of an m x n matrix (Eigen::FullPivLU): P A Q = L U.
struct FullPivLU {
  int m, n, rank;
  std::vector<double> lu;      // packed L (unit lower) and U
  std::vector<int> rowperm;    // P as a permutation of rows: row i of PA is row rowperm[i] of A
  std::vector<int> colperm;    // Q: column j of AQ is column colperm[j] of A
  double maxpivot;

  FullPivLU(const double* A, int m_, int n_) : m(m_), n(n_), lu(A, A + m_ * n_), rowperm(m_), colperm(n_) {
    for (int i = 0; i < m; ++i) rowperm[i] = i;
    for (int j = 0; j < n; ++j) colperm[j] = j;
    const int size = m < n ? m : n;
    maxpivot = 0.0;
    int nonzero = size;
    for (int k = 0; k < size; ++k) {
      // biggest coefficient of the remaining bottom-right corner
      int pr = k, pc = k;
      double big = -1.0;
      for (int i = k; i < m; ++i)
        for (int j = k; j < n; ++j) {
          const double a = fabs(lu[i * n + j]);
          if (a > big) { big = a; pr = i; pc = j; }
        }
      if (big == 0.0) { nonzero = k; break; }
      if (big > maxpivot) maxpivot = big;
      if (pr != k) {
        for (int j = 0; j < n; ++j) { double t = lu[k * n + j]; lu[k * n + j] = lu[pr * n + j]; lu[pr * n + j] = t; }
        int t = rowperm[k]; rowperm[k] = rowperm[pr]; rowperm[pr] = t;
      }
      if (pc != k) {
        for (int i = 0; i < m; ++i) { double t = lu[i * n + k]; lu[i * n + k] = lu[i * n + pc]; lu[i * n + pc] = t; }
        int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t;
      }
      const double piv = lu[k * n + k];
      for (int i = k + 1; i < m; ++i) lu[i * n + k] /= piv;
      for (int i = k + 1; i < m; ++i) {
        const double l = lu[i * n + k];
        for (int j = k + 1; j < n; ++j) lu[i * n + j] -= l * lu[k * n + j];
      }
    }
    // rank with Eigen's default threshold: |pivot| > maxpivot * eps * min(m, n)
    const double thr = maxpivot * 2.220446049250313e-16 * (double)size;
    rank = 0;
    for (int k = 0; k < nonzero; ++k) if (fabs(lu[k * n + k]) > thr) ++rank;
  }

  // X = A^-1 B for square invertible A; B, X are n x nrhs row-major
  void solve(const double* B, double* X, int nrhs) const {
    std::vector<double> c((size_t)n * nrhs);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < nrhs; ++j) c[i * nrhs + j] = B[rowperm[i] * nrhs + j];
    for (int i = 0; i < n; ++i)  // L c = P b
      for (int k = 0; k < i; ++k)
        for (int j = 0; j < nrhs; ++j) c[i * nrhs + j] -= lu[i * n + k] * c[k * nrhs + j];
    for (int i = n - 1; i >= 0; --i) {  // U y = c
      for (int k = i + 1; k < n; ++k)
        for (int j = 0; j < nrhs; ++j) c[i * nrhs + j] -= lu[i * n + k] * c[k * nrhs + j];
      for (int j = 0; j < nrhs; ++j) c[i * nrhs + j] /= lu[i * n + i];
    }
    for (int i = 0; i < n; ++i)  // x = Q y
      for (int j = 0; j < nrhs; ++j) X[colperm[i] * nrhs + j] = c[i * nrhs + j];
  }

  // Basis of ker A as columns of an n x (n - rank) row-major matrix (Eigen kernel()):
  // ker A = Q ker U; with U = [U1 U2] (U1 rank x rank upper triangular) the basis is Q [-U1^-1 U2; I].
  std::vector<double> kernel(int* dimker_out) const {
    const int dimker = n - rank;
    *dimker_out = dimker;
    std::vector<double> K((size_t)n * (dimker > 0 ? dimker : 1), 0.0);
    if (dimker == 0) return K;
    std::vector<double> X((size_t)rank * dimker);
    for (int i = 0; i < rank; ++i)
      for (int j = 0; j < dimker; ++j) X[i * dimker + j] = -lu[i * n + rank + j];
    for (int i = rank - 1; i >= 0; --i) {
      for (int k = i + 1; k < rank; ++k)
        for (int j = 0; j < dimker; ++j) X[i * dimker + j] -= lu[i * n + k] * X[k * dimker + j];
      for (int j = 0; j < dimker; ++j) X[i * dimker + j] /= lu[i * n + i];
    }
    for (int i = 0; i < rank; ++i)
      for (int j = 0; j < dimker; ++j) K[colperm[i] * dimker + j] = X[i * dimker + j];
    for (int j = 0; j < dimker; ++j) K[colperm[rank + j] * dimker + j] = 1.0;
    return K;
  }
};

}  // namespace oracle_la

// ekf_c.c:8-33
void predict(double* in_x, double* in_P, double* in_Q, double dt) {
  using namespace oracle_la;
  double nx[DIM] = {0};
  double in_F[EDIM * EDIM] = {0};
  f_fun(in_x, dt, nx);      // ekf_c.c:15
  F_fun(in_x, dt, in_F);    // ekf_c.c:16

  std::vector<double> P(in_P, in_P + EDIM * EDIM);
  const int M = MEDIM, A = EDIM - MEDIM;
  std::vector<double> Fm((size_t)M * M), FmT((size_t)M * M);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) Fm[i * M + j] = in_F[i * EDIM + j];
  transpose(Fm.data(), FmT.data(), M, M);

  // P_mm <- (F_mm P_mm) F_mm^T      ekf_c.c:24
  std::vector<double> Pmm((size_t)M * M), T((size_t)M * M), T2((size_t)M * M);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) Pmm[i * M + j] = P[i * EDIM + j];
  matmul_fixed<MEDIM, MEDIM, MEDIM>(Fm.data(), Pmm.data(), T.data());
  matmul_fixed<MEDIM, MEDIM, MEDIM>(T.data(), FmT.data(), T2.data());
  if (A > 0) {
    // P_ma <- F_mm P_ma ; P_am <- P_am F_mm^T    ekf_c.c:25-26
    std::vector<double> Pma((size_t)M * A), Pam((size_t)A * M), Ra((size_t)M * A), Rb((size_t)A * M);
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < A; ++j) Pma[i * A + j] = P[i * EDIM + M + j];
    for (int i = 0; i < A; ++i)
      for (int j = 0; j < M; ++j) Pam[i * M + j] = P[(M + i) * EDIM + j];
    matmul(Fm.data(), Pma.data(), Ra.data(), M, M, A);
    matmul(Pam.data(), FmT.data(), Rb.data(), A, M, M);
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < A; ++j) P[i * EDIM + M + j] = Ra[i * A + j];
    for (int i = 0; i < A; ++i)
      for (int j = 0; j < M; ++j) P[(M + i) * EDIM + j] = Rb[i * M + j];
  }
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) P[i * EDIM + j] = T2[i * M + j];

  for (int i = 0; i < EDIM * EDIM; ++i) P[i] = P[i] + dt * in_Q[i];  // ekf_c.c:28

  memcpy(in_x, nx, DIM * sizeof(double));
  memcpy(in_P, P.data(), EDIM * EDIM * sizeof(double));
}

// ekf_c.c:37-121
template <int ZDIM, int EADIM, bool MAHA_TEST>
void update(double* in_x, double* in_P, Hfun h_fun, Hfun H_fun, Hfun Hea_fun, double* in_z, double* in_R, double* in_ea, double MAHA_THRESHOLD) {
  using namespace oracle_la;
  double in_hx[ZDIM] = {0};
  double in_H[ZDIM * DIM] = {0};
  double in_H_mod[EDIM * DIM] = {0};
  double delta_x[EDIM] = {0};
  double x_new[DIM] = {0};

  h_fun(in_x, in_ea, in_hx);   // ekf_c.c:59
  H_fun(in_x, in_ea, in_H);    // ekf_c.c:60

  double pre_y[ZDIM];
  for (int i = 0; i < ZDIM; ++i) pre_y[i] = in_z[i] - in_hx[i];  // ekf_c.c:64

  int ydim = ZDIM;
  std::vector<double> y, H, R;
  if (Hea_fun) {  // ekf_c.c:66-76: project onto the left null space of He
    double in_Hea[ZDIM * EADIM] = {0};
    Hea_fun(in_x, in_ea, in_Hea);
    std::vector<double> HeaT((size_t)EADIM * ZDIM);
    transpose(in_Hea, HeaT.data(), ZDIM, EADIM);
    FullPivLU lu(HeaT.data(), EADIM, ZDIM);
    int dk = 0;
    std::vector<double> A = lu.kernel(&dk);  // ZDIM x dk
    ydim = dk;
    std::vector<double> At((size_t)dk * ZDIM);
    transpose(A.data(), At.data(), ZDIM, dk);
    y.resize(dk); H.resize((size_t)dk * DIM); R.resize((size_t)dk * dk);
    matmul(At.data(), pre_y, y.data(), dk, ZDIM, 1);
    matmul(At.data(), in_H, H.data(), dk, ZDIM, DIM);
    std::vector<double> AtR((size_t)dk * ZDIM);
    matmul(At.data(), in_R, AtR.data(), dk, ZDIM, ZDIM);
    matmul(AtR.data(), A.data(), R.data(), dk, ZDIM, dk);
  } else {
    y.assign(pre_y, pre_y + ZDIM);
    H.assign(in_H, in_H + ZDIM * DIM);
    R.assign(in_R, in_R + ZDIM * ZDIM);
  }
  const int m = ydim;

  H_mod_fun(in_x, in_H_mod);   // ekf_c.c:83  (DIM x EDIM)
  std::vector<double> H_err((size_t)m * EDIM);
  matmul(H.data(), in_H_mod, H_err.data(), m, DIM, EDIM);  // ekf_c.c:85

  std::vector<double> P(in_P, in_P + EDIM * EDIM), Pt((size_t)EDIM * EDIM);
  std::vector<double> H_errT((size_t)EDIM * m), HP((size_t)m * EDIM), S((size_t)m * m);
  transpose(H_err.data(), H_errT.data(), m, EDIM);
  matmul(H_err.data(), P.data(), HP.data(), m, EDIM, EDIM);

  if (MAHA_TEST) {  // ekf_c.c:88-94
    matmul(HP.data(), H_errT.data(), S.data(), m, EDIM, m);
    for (int i = 0; i < m * m; ++i) S[i] += R[i];
    std::vector<double> I((size_t)m * m, 0.0), a((size_t)m * m);
    for (int i = 0; i < m; ++i) I[i * m + i] = 1.0;
    FullPivLU lu(S.data(), m, m);   // Eigen's general inverse() is LU based
    lu.solve(I.data(), a.data(), m);
    double maha_dist = 0.0;
    for (int i = 0; i < m; ++i) {
      double s = 0.0;
      for (int j = 0; j < m; ++j) s += a[i * m + j] * y[j];
      maha_dist += y[i] * s;
    }
    if (maha_dist > MAHA_THRESHOLD) {
      for (int i = 0; i < m * m; ++i) R[i] = 1.0e16 * R[i];
    }
  }

  const double weight = 1;  // ekf_c.c:97

  // S = (H_err P) H_err^T + R / weight ; KT = S^-1 (H_err P^T)      ekf_c.c:100-101
  matmul(HP.data(), H_errT.data(), S.data(), m, EDIM, m);
  for (int i = 0; i < m * m; ++i) S[i] += R[i] / weight;
  transpose(P.data(), Pt.data(), EDIM, EDIM);
  std::vector<double> HPt((size_t)m * EDIM), KT((size_t)m * EDIM), K((size_t)EDIM * m);
  matmul(H_err.data(), Pt.data(), HPt.data(), m, EDIM, EDIM);
  {
    FullPivLU lu(S.data(), m, m);
    lu.solve(HPt.data(), KT.data(), EDIM);
  }
  transpose(KT.data(), K.data(), m, EDIM);

  // I_KH = I - K H_err      ekf_c.c:105
  std::vector<double> I_KH((size_t)EDIM * EDIM), I_KHt((size_t)EDIM * EDIM);
  matmul(K.data(), H_err.data(), I_KH.data(), EDIM, m, EDIM);
  for (int i = 0; i < EDIM; ++i)
    for (int j = 0; j < EDIM; ++j) I_KH[i * EDIM + j] = (i == j ? 1.0 : 0.0) - I_KH[i * EDIM + j];

  // dx = K y ; x <- err_fun(x, dx)      ekf_c.c:108-112
  matmul(K.data(), y.data(), delta_x, EDIM, m, 1);
  err_fun(in_x, delta_x, x_new);

  // P <- (I_KH P) I_KH^T + (K R) K^T      ekf_c.c:115
  std::vector<double> T((size_t)EDIM * EDIM), Pn((size_t)EDIM * EDIM), KR((size_t)EDIM * m), KRKt((size_t)EDIM * EDIM);
  matmul_fixed<EDIM, EDIM, EDIM>(I_KH.data(), P.data(), T.data());
  transpose(I_KH.data(), I_KHt.data(), EDIM, EDIM);
  matmul_fixed<EDIM, EDIM, EDIM>(T.data(), I_KHt.data(), Pn.data());
  matmul(K.data(), R.data(), KR.data(), EDIM, m, m);
  matmul(KR.data(), KT.data(), KRKt.data(), EDIM, m, EDIM);
  for (int i = 0; i < EDIM * EDIM; ++i) Pn[i] += KRKt[i];

  memcpy(in_x, x_new, DIM * sizeof(double));
  memcpy(in_P, Pn.data(), EDIM * EDIM * sizeof(double));
  memcpy(in_z, y.data(), m * sizeof(double));  // ekf_c.c:120
}
