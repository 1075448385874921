// rednose_b200 -- CTA-per-filter kernel for large error states (EDIM > 32, MSCKF).
// Placeholder until the shared-memory-resident covariance kernel lands.
#pragma once
#include "ekf_common.cuh"

namespace rnb {

template <class M, class K, bool PRED, bool UPD>
inline void launch_step_cta(const StepArgs<M::NG>&, cudaStream_t) {
  fprintf(stderr, "[rednose_b200] EDIM=%d > 32: CTA-per-filter kernel not built into this library\n", M::EDIM);
  last_status() = (int)cudaErrorNotSupported;
}

}  // namespace rnb
