// rednose_b200 runtime library (librednose_b200.so): plugin registry + native single-filter driver.
//
// Native counterparts of the reference's C++ runtime around the numeric kernel:
//   registry / loader   rednose/helpers/ekf_load.{h,cc}   (ekf_get_all, ekf_register, ekf_lookup,
//                       ekf_load_and_register = dlopen + dlsym("ekf_get"))
//   driver EKFSym       rednose/helpers/ekf_sym.{h,cc}    (time handling, per-observation update loop,
//                       quaternion normalisation, rewind ring of 512 checkpoints + fast-forward)
// Eigen-free: matrices are plain row-major std::vector<double>.  All numerics are delegated to the
// filter library's C-ABI (<name>_predict / <name>_update_<kind>), i.e. to the CUDA kernels; this file
// contains no filter arithmetic except quaternion normalisation and the augment() permutation.
#include "rednose_b200.h"

#include <dlfcn.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

namespace {

constexpr size_t REWIND_TO_KEEP = 512;  // ekf_sym.h:18

// One entry per loaded plugin.  The reference keys its table by filter name only (ekf_load.cc:13-25: first match
// wins, a second library of the same name is never loaded); here the directory a plugin was loaded from is kept as
// well, so two builds of one filter (e.g. the CUDA library and a CPU build of the same model in a test process) can
// coexist and a driver gets the one from the directory it asked for.
struct Plugin {
  const rednose_ekf_desc* desc;
  std::string directory;   // empty: self-registered by its constructor, origin not (yet) known
};
std::vector<Plugin>& registry() {
  static std::vector<Plugin> v;
  return v;
}
std::mutex& registry_mu() {
  static std::mutex m;  // the reference's vector is unsynchronised (ekf_load.cc:4-7); this one is not
  return m;
}

struct Observation {  // ekf_sym.h:24-30
  double t;
  int kind;
  int n, zdim, eadim;
  std::vector<double> z, R, ea;
};

struct Checkpoint {
  double t;
  std::vector<double> x, P;
  Observation obs;
};

struct EKFSym {
  const rednose_ekf_desc* ekf = nullptr;
  int dim_x = 0, dim_err = 0, dim_main = 0, dim_main_err = 0, N = 0, dim_augment = 0, dim_augment_err = 0;
  bool msckf = false;
  std::vector<double> x, P, Q;
  double filter_time = NAN;  // ekf_sym.cc:42
  double max_rewind_age = 1.0;
  std::vector<int> maha_test_kinds, quaternion_idxs;
  std::vector<double> augment_times;
  std::deque<Checkpoint> rewind_buf;

  int kind_index(int kind) const {
    for (int i = 0; i < ekf->n_kinds; ++i)
      if (ekf->kinds[i] == kind) return i;
    return -1;
  }

  void normalize_quaternions() {  // ekf_sym.cc:69-77
    for (int idx : quaternion_idxs) {
      double n = 0.0;
      for (int c = 0; c < 4; ++c) n += x[idx + c] * x[idx + c];
      n = std::sqrt(n);
      for (int c = 0; c < 4; ++c) x[idx + c] /= n;
    }
  }

  void init_state(const double* state, const double* covs, double t) {  // ekf_sym.cc:45-51
    x.assign(state, state + dim_x);
    P.assign(covs, covs + (size_t)dim_err * dim_err);
    filter_time = t;
    augment_times.assign(N, 0.0);
    rewind_buf.clear();
  }

  void predict(double t) {  // ekf_sym.cc:196-209
    if (std::isnan(filter_time)) filter_time = t;
    const double dt = t - filter_time;
    assert(dt >= 0.0);
    ekf->predict(x.data(), P.data(), Q.data(), dt);
    normalize_quaternions();
    filter_time = t;
  }

  // one observation; returns the innovation length written into y (ekf_sym.cc:211-219)
  int update(int kind, const double* z, const double* R, const double* ea, int zdim, int eadim, double* y) {
    const int ki = kind_index(kind);
    assert(ki >= 0);
    std::vector<double> zbuf(z, z + zdim), Rbuf(R, R + (size_t)zdim * zdim), eabuf(ea ? ea : z, (ea ? ea : z) + (ea ? eadim : 0));
    if (eabuf.empty()) eabuf.push_back(0.0);
    ekf->updates[ki](x.data(), P.data(), zbuf.data(), Rbuf.data(), eabuf.data());
    normalize_quaternions();
    const int ydim = ekf->feature_kind[ki] ? zdim - eadim : zdim;
    std::memcpy(y, zbuf.data(), sizeof(double) * ydim);
    return ydim;
  }

  void augment() {  // ekf_sym.py:365-391 (the C++ reference asserts !augment, ekf_sym.cc:186)
    assert(msckf);
    const int d1 = dim_main, d2 = dim_main_err, d3 = dim_augment, d4 = dim_augment_err;
    // state: drop the oldest clone, append a copy of the first d3 main states
    std::memmove(&x[d1], &x[d1 + d3], sizeof(double) * (dim_x - d1 - d3));
    std::memcpy(&x[dim_x - d3], &x[0], sizeof(double) * d3);
    // covariance: same selection on rows and columns
    std::vector<int> src;
    for (int i = 0; i < dim_err; ++i)
      if (i < d2 || i >= d2 + d4) src.push_back(i);
    for (int i = 0; i < d4; ++i) src.push_back(i);
    std::vector<double> Pn((size_t)dim_err * dim_err);
    for (int i = 0; i < dim_err; ++i)
      for (int j = 0; j < dim_err; ++j) Pn[(size_t)i * dim_err + j] = P[(size_t)src[i] * dim_err + src[j]];
    std::copy(Pn.begin(), Pn.end(), P.begin());   // into the existing storage: views handed out by P_ptr() stay valid
    if (!augment_times.empty()) {
      augment_times.erase(augment_times.begin());
      augment_times.push_back(filter_time);
    }
  }

  void checkpoint(const Observation& obs) {  // ekf_sym.cc:144-156
    rewind_buf.push_back(Checkpoint{filter_time, x, P, obs});
    if (rewind_buf.size() > REWIND_TO_KEEP) rewind_buf.pop_front();
  }

  std::deque<Observation> rewind(double t) {  // ekf_sym.cc:125-142
    std::deque<Observation> rewound;
    while (rewind_buf.back().t > t) {
      rewound.push_front(rewind_buf.back().obs);
      rewind_buf.pop_back();
    }
    filter_time = rewind_buf.back().t;
    x = rewind_buf.back().x;
    P = rewind_buf.back().P;
    return rewound;
  }

  // ekf_sym.cc:158-194; outputs may be null (fast-forward replays)
  void predict_and_update(const Observation& obs, bool do_augment, double* xk1, double* xk, double* Pk1, double* Pk, double* y) {
    predict(obs.t);
    if (xk1) std::memcpy(xk1, x.data(), sizeof(double) * dim_x);
    if (Pk1) std::memcpy(Pk1, P.data(), sizeof(double) * dim_err * dim_err);
    std::vector<double> ytmp(obs.zdim > 0 ? obs.zdim : 1);
    for (int i = 0; i < obs.n; ++i) {
      const double* ea = obs.eadim > 0 ? &obs.ea[(size_t)i * obs.eadim] : nullptr;
      const int ydim = update(obs.kind, &obs.z[(size_t)i * obs.zdim], &obs.R[(size_t)i * obs.zdim * obs.zdim], ea, obs.zdim, obs.eadim, ytmp.data());
      if (y) std::memcpy(y + (size_t)i * obs.zdim, ytmp.data(), sizeof(double) * ydim);
    }
    if (xk) std::memcpy(xk, x.data(), sizeof(double) * dim_x);
    if (Pk) std::memcpy(Pk, P.data(), sizeof(double) * dim_err * dim_err);
    if (do_augment) augment();
    checkpoint(obs);
  }
};

}  // namespace

extern "C" {

// ------------------------------------------------------------------ registry (ekf_load.cc) ---
void rednose_b200_register(const rednose_ekf_desc* desc) {
  std::lock_guard<std::mutex> lk(registry_mu());
  for (const auto& p : registry())
    if (p.desc == desc) return;
  registry().push_back({desc, std::string()});
}

const rednose_ekf_desc* rednose_b200_lookup(const char* name) {
  std::lock_guard<std::mutex> lk(registry_mu());
  for (const auto& p : registry())
    if (std::strcmp(p.desc->name, name) == 0) return p.desc;  // first match wins (ekf_load.cc:13-20)
  return nullptr;
}

// the plugin `name` that was loaded from `directory` (nullptr if none)
const rednose_ekf_desc* rednose_b200_lookup_in(const char* directory, const char* name) {
  std::lock_guard<std::mutex> lk(registry_mu());
  for (const auto& p : registry())
    if (p.directory == directory && std::strcmp(p.desc->name, name) == 0) return p.desc;
  return nullptr;
}

int rednose_b200_load_and_register(const char* directory, const char* name) {
  if (rednose_b200_lookup_in(directory, name)) return 0;  // ekf_load.cc:23-25, per directory
  const std::string path = std::string(directory) + "/lib" + name + ".so";
  void* handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!handle) {
    fprintf(stderr, "[rednose_b200] dlopen(%s) failed: %s\n", path.c_str(), dlerror());
    return -1;
  }
  auto get = reinterpret_cast<void* (*)()>(dlsym(handle, "ekf_get"));
  if (!get) {
    fprintf(stderr, "[rednose_b200] %s has no ekf_get()\n", path.c_str());
    return -2;
  }
  const auto* desc = static_cast<const rednose_ekf_desc*>(get());
  if (desc->abi_version != REDNOSE_B200_ABI_VERSION) {
    fprintf(stderr, "[rednose_b200] %s: ABI %d, runtime expects %d\n", path.c_str(), desc->abi_version, REDNOSE_B200_ABI_VERSION);
    return -3;
  }
  std::lock_guard<std::mutex> lk(registry_mu());
  for (auto& p : registry())
    if (p.desc == desc) {   // the library's constructor registered it already: record where it came from
      if (p.directory.empty()) p.directory = directory;
      return 0;
    }
  registry().push_back({desc, std::string(directory)});
  return 0;
}

// ------------------------------------------------------------------ driver (ekf_sym.cc) ---
void* rednose_ekfsym_create(const char* directory, const char* name, const double* Q, const double* x0, const double* P0,
                            int dim_x, int dim_err, int dim_main, int dim_main_err, int N, int dim_augment, int dim_augment_err,
                            const int* maha_test_kinds, int n_maha, const int* quaternion_idxs, int n_quat, double max_rewind_age) {
  if (rednose_b200_load_and_register(directory, name) != 0) return nullptr;
  auto* e = new EKFSym();
  e->ekf = rednose_b200_lookup_in(directory, name);
  if (!e->ekf) e->ekf = rednose_b200_lookup(name);
  e->msckf = N > 0;
  e->N = N; e->dim_augment = dim_augment; e->dim_augment_err = dim_augment_err;
  e->dim_main = dim_main; e->dim_main_err = dim_main_err;
  e->dim_x = dim_x; e->dim_err = dim_err;
  // ekf_sym.cc:25-27 (+ agreement with what the library was generated for)
  if (dim_main + dim_augment * N != dim_x || dim_main_err + dim_augment_err * N != dim_err || e->ekf->dim != dim_x || e->ekf->edim != dim_err) {
    fprintf(stderr, "[rednose_b200] dimension mismatch for filter %s\n", name);
    delete e;
    return nullptr;
  }
  e->maha_test_kinds.assign(maha_test_kinds, maha_test_kinds + n_maha);
  e->quaternion_idxs.assign(quaternion_idxs, quaternion_idxs + n_quat);
  e->Q.assign(Q, Q + (size_t)dim_err * dim_err);
  e->max_rewind_age = max_rewind_age;
  e->init_state(x0, P0, NAN);
  return e;
}

void rednose_ekfsym_destroy(void* h) { delete static_cast<EKFSym*>(h); }

void rednose_ekfsym_init_state(void* h, const double* x, const double* P, double filter_time) {
  static_cast<EKFSym*>(h)->init_state(x, P, filter_time);
}
double* rednose_ekfsym_x_ptr(void* h) { return static_cast<EKFSym*>(h)->x.data(); }
double* rednose_ekfsym_P_ptr(void* h) { return static_cast<EKFSym*>(h)->P.data(); }
double rednose_ekfsym_get_filter_time(void* h) { return static_cast<EKFSym*>(h)->filter_time; }
void rednose_ekfsym_set_filter_time(void* h, double t) { static_cast<EKFSym*>(h)->filter_time = t; }
void rednose_ekfsym_reset_rewind(void* h) { static_cast<EKFSym*>(h)->rewind_buf.clear(); }
int rednose_ekfsym_rewind_depth(void* h) { return (int)static_cast<EKFSym*>(h)->rewind_buf.size(); }
void rednose_ekfsym_normalize_quaternions(void* h) { static_cast<EKFSym*>(h)->normalize_quaternions(); }
void rednose_ekfsym_augment(void* h) { static_cast<EKFSym*>(h)->augment(); }
void rednose_ekfsym_get_augment_times(void* h, double* out) {
  auto* e = static_cast<EKFSym*>(h);
  std::copy(e->augment_times.begin(), e->augment_times.end(), out);
}

int rednose_ekfsym_set_global(void* h, const char* var, double val) {  // ekf_sym.cc:79-81
  auto* e = static_cast<EKFSym*>(h);
  for (int i = 0; i < e->ekf->n_sets; ++i)
    if (std::strcmp(e->ekf->set_names[i], var) == 0) { e->ekf->sets[i](val); return 0; }
  return -1;
}

void* rednose_ekfsym_get_extra_routine(void* h, const char* routine) {  // ekf_sym.cc:221-223
  auto* e = static_cast<EKFSym*>(h);
  for (int i = 0; i < e->ekf->n_extra; ++i)
    if (std::strcmp(e->ekf->extra_names[i], routine) == 0) return e->ekf->extra_fns[i];
  return nullptr;
}

void rednose_ekfsym_predict(void* h, double t) { static_cast<EKFSym*>(h)->predict(t); }

// ekf_sym.cc:83-117.  z [n, zdim], R [n, zdim, zdim], ea [n, eadim] row-major.  Returns 1 and fills the outputs
// (xk1 [dim_x], xk [dim_x], Pk1/Pk [dim_err^2], y [n, zdim]) or 0 if the observation was too old and ignored.
int rednose_ekfsym_predict_and_update_batch(void* h, double t, int kind, const double* z, const double* R, const double* ea,
                                            int n, int zdim, int eadim, int augment,
                                            double* xk1, double* xk, double* Pk1, double* Pk, double* y) {
  auto* e = static_cast<EKFSym*>(h);
  if (e->kind_index(kind) < 0) return -1;  // .at() would throw std::out_of_range in the reference (ekf_sym.cc:212)
  std::deque<Observation> rewound;
  if (!std::isnan(e->filter_time) && t < e->filter_time) {
    if (e->rewind_buf.empty() || t < e->rewind_buf.front().t || t < e->rewind_buf.back().t - e->max_rewind_age) {
      fprintf(stderr, "observation too old at %f with filter at %f, ignoring!\n", t, e->filter_time);  // logger.h LOGD
      return 0;
    }
    rewound = e->rewind(t);
  }
  Observation obs;
  obs.t = t; obs.kind = kind; obs.n = n; obs.zdim = zdim; obs.eadim = eadim;
  obs.z.assign(z, z + (size_t)n * zdim);
  obs.R.assign(R, R + (size_t)n * zdim * zdim);
  if (eadim > 0 && ea) obs.ea.assign(ea, ea + (size_t)n * eadim);
  e->predict_and_update(obs, augment != 0, xk1, xk, Pk1, Pk, y);
  while (!rewound.empty()) {  // fast-forward (ekf_sym.cc:111-114)
    e->predict_and_update(rewound.front(), false, nullptr, nullptr, nullptr, nullptr, nullptr);
    rewound.pop_front();
  }
  return 1;
}

}  // extern "C"
