// rednose_b200 -- batched MSCKF state augmentation (K3): drop the oldest clone, append a copy of the
// first dim_augment main states; the same selection on rows and columns of P.
// Reference: EKF_sym.augment, rednose/helpers/ekf_sym.py:365-391, which builds a selection matrix and does
// two dense products; here it is the permutation/copy it really is.  One CTA per filter, P staged in
// shared memory so the in-place rewrite is safe.
#pragma once
#include "ekf_common.cuh"

namespace rnb {

template <class M>
__global__ void __launch_bounds__(128) ekf_augment_cta(double* __restrict__ x, double* __restrict__ P, long long B, int d3, int d4) {
  constexpr int D = M::DIM, E = M::EDIM, d1 = M::DMAIN, d2 = M::MEDIM;
  extern __shared__ __align__(16) double tile[];  // E*E + D
  const long long b = blockIdx.x;
  if (b >= B) return;
  double* Pg = P + b * (long long)(E * E);
  double* xg = x + b * D;
  double* xs = tile + E * E;
  for (int i = threadIdx.x; i < E * E; i += blockDim.x) tile[i] = Pg[i];
  for (int i = threadIdx.x; i < D; i += blockDim.x) xs[i] = xg[i];
  __syncthreads();
  // state: [main | clone_1 .. clone_N] -> [main | clone_2 .. clone_N | main[:d3]]
  for (int i = threadIdx.x + d1; i < D; i += blockDim.x) xg[i] = (i < D - d3) ? xs[i + d3] : xs[i - (D - d3)];
  // covariance: src(i) = i for the main block, i + d4 for the surviving clones, i - (E - d4) for the new clone
  auto src = [&](int i) { return i < d2 ? i : (i < E - d4 ? i + d4 : i - (E - d4)); };
  for (int idx = threadIdx.x; idx < E * E; idx += blockDim.x) {
    const int i = idx / E, j = idx - i * E;
    Pg[idx] = tile[src(i) * E + src(j)];
  }
}

template <class M>
inline void launch_augment(double* x, double* P, long long B, int dim_augment, int dim_augment_err, cudaStream_t st) {
  if (B <= 0) return;
  const size_t smem = sizeof(double) * (M::EDIM * M::EDIM + M::DIM);
  if (first_launch_of((const void*)ekf_augment_cta<M>))   // per (device, kernel)
    cudaFuncSetAttribute(ekf_augment_cta<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ekf_augment_cta<M><<<(unsigned)B, 128, smem, st>>>(x, P, B, dim_augment, dim_augment_err);
  check(cudaGetLastError(), "ekf_augment launch");
}

}  // namespace rnb
