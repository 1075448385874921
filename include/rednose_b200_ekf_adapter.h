/* rednose_b200 -- adapter from the plugin descriptor of a generated CUDA filter library (`rednose_ekf_desc`, returned by
 * its `ekf_get()`) to the reference's C++ plugin table `struct EKF` (rednose/helpers/ekf.h:16-33).
 *
 * Why an adapter: the reference's `struct EKF` has std::string / std::vector / std::unordered_map members, so its layout
 * depends on the compiler and standard library that built the plugin -- it cannot be filled in by a library built with
 * another toolchain (nvcc's host compiler here).  The library therefore exports a plain-C descriptor, and the HOST
 * program, which knows its own `struct EKF`, builds the table from it.  With this header the reference's loader
 * (rednose/helpers/ekf_load.cc:33-38) needs one changed line:
 *
 *     const EKF* ekf = rednose_b200_adapt((const rednose_ekf_desc*)ekf_get());     // was: (const EKF*)ekf_get()
 *     ekf_register(ekf);
 *
 * and the reference's own C++ driver `EKFSym` (rednose/helpers/ekf_sym.cc:12,80,206,212,222) and its Cython wrapper run
 * the CUDA library unchanged.
 *
 * Usage: include the reference's "rednose/helpers/ekf.h" (or any header that declares `struct EKF` and
 * `extra_routine_t` with that field list) BEFORE this header.  C++ only.  Compiled and driven in
 * tests/test_abi_cpu.py::test_struct_ekf_adapter_drives_a_library.
 */
#ifndef REDNOSE_B200_EKF_ADAPTER_H
#define REDNOSE_B200_EKF_ADAPTER_H
#ifndef __cplusplus
#error "rednose_b200_ekf_adapter.h builds the reference's C++ struct EKF: include it from C++"
#endif
#include "rednose_b200.h"

/* Builds a heap-allocated table that lives as long as the process (like the reference's static `const EKF <name>`,
 * rednose/helpers/ekf_sym.py:186-203).  Returns nullptr for an unknown descriptor version. */
static inline EKF* rednose_b200_adapt(const rednose_ekf_desc* d) {
  if (!d || d->abi_version != 1) return nullptr;
  EKF* e = new EKF();
  e->name = d->name;
  e->f_fun = d->f_fun;
  e->F_fun = d->F_fun;
  e->err_fun = d->err_fun;
  e->inv_err_fun = d->inv_err_fun;
  e->H_mod_fun = d->H_mod_fun;
  e->predict = d->predict;
  for (int i = 0; i < d->n_kinds; ++i) {
    const int k = d->kinds[i];
    e->kinds.push_back(k);
    e->hs[k] = d->hs[i];
    e->Hs[k] = d->Hs[i];
    e->updates[k] = d->updates[i];
    if (d->feature_kind[i]) {
      e->feature_kinds.push_back(k);
      e->Hes[k] = d->Hes[i];
    }
  }
  for (int i = 0; i < d->n_sets; ++i) e->sets[d->set_names[i]] = d->sets[i];
  for (int i = 0; i < d->n_extra; ++i) e->extra_routines[d->extra_names[i]] = (extra_routine_t)d->extra_fns[i];
  return e;
}

#endif /* REDNOSE_B200_EKF_ADAPTER_H */
