// rednose_b200 -- host-side templates behind the generated C-ABI of one filter library.
//
// The generated lib<name>.so exports (a) the reference's own symbol set
// (rednose/helpers/ekf_sym.py:149-171: <name>_predict, <name>_update_<kind>, leaf
// functions, <name>_set_<var>) operating on HOST pointers for ONE filter, and (b)
// batched additions (<name>_batch_*) operating on DEVICE pointers.  Both run the same
// CUDA kernels; there is no CPU implementation in this library.
#pragma once
#include "ekf_common.cuh"
#include "ekf_thread.cuh"
#include "ekf_warp2.cuh"
#include "ekf_cta.cuh"
#include "ekf_rts.cuh"
#include "ekf_rts_mma.cuh"
#include "ekf_augment.cuh"
#include <cstring>
#include <mutex>
#include <unordered_set>

namespace rnb {

// placeholder kind for predict-only instantiations
struct NullKind {
  static constexpr int KIND = -1, ZDIM = 1, YDIM = 1, EADIM = 0, NH = 0;
  static constexpr bool MAHA = false, HAS_HE = false;
  static constexpr double MAHA_THRESH = 0.0;
  template <class HV> static __device__ __forceinline__ void obs_leaf(const double*, const double*, const double*, double (&)[1], HV&) {}
  template <class HV, class V> static __device__ __forceinline__ void Herr_apply(const HV&, const V&, double (&hp)[1]) { hp[0] = 0.0; }
  template <class HV, class A> static __device__ __forceinline__ void S_accum(const HV&, A, double (&)[1][1]) {}
};

template <int NG>
struct GV { double v[NG > 0 ? NG : 1]; };

// Per-library host context: global_vars (ekf_sym.py:129-132 keeps them as file
// statics; here they are copied into every launch's argument block) and a small
// device scratch used by the single-filter host-pointer entry points.
constexpr int MAX_DEVICES = 32;
inline int current_device() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < MAX_DEVICES) ? d : 0; }

template <class M>
struct HostCtx {
  GV<M::NG> gv{};
  std::mutex mu;
  // device scratch / pinned staging / private stream of the single-filter entry points, ONE SET PER DEVICE (a process may
  // drive several GPUs; a pointer or stream created on device 0 must never be used while device 1 is current)
  struct PerDevice {
    double* d_scratch = nullptr;
    size_t scratch_doubles = 0;
    double* h_pinned = nullptr;      // pinned host staging: one copy in, one copy out
    size_t pinned_doubles = 0;
    cudaStream_t stream = nullptr;
  };
  PerDevice dev[MAX_DEVICES];

  double* scratch(size_t n) {
    PerDevice& p = dev[current_device()];
    if (n > p.scratch_doubles) {
      if (p.d_scratch) cudaFree(p.d_scratch);
      p.d_scratch = nullptr; p.scratch_doubles = 0;
      if (!check(cudaMalloc(&p.d_scratch, n * sizeof(double)), "cudaMalloc(scratch)")) return nullptr;
      p.scratch_doubles = n;
    }
    return p.d_scratch;
  }
  double* pinned(size_t n) {
    PerDevice& p = dev[current_device()];
    if (n > p.pinned_doubles) {
      if (p.h_pinned) cudaFreeHost(p.h_pinned);
      p.h_pinned = nullptr; p.pinned_doubles = 0;
      if (!check(cudaMallocHost((void**)&p.h_pinned, n * sizeof(double)), "cudaMallocHost(staging)")) return nullptr;
      p.pinned_doubles = n;
    }
    return p.h_pinned;
  }
  cudaStream_t single_stream() {
    PerDevice& p = dev[current_device()];
    if (!p.stream) check(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking), "cudaStreamCreate(single)");
    return p.stream;
  }
};

constexpr int THREAD_MAX_EDIM = 6;

template <class M, class K, bool PRED, bool UPD>
inline void launch_step(const StepArgs<M::NG>& a, cudaStream_t st) {
  if (a.B <= 0) return;
  if ((a.flags & FLAG_AUGMENT) && !(M::EDIM > 32 || K::HAS_HE)) {
    fprintf(stderr, "[rednose_b200] the fused augment exists only in the CTA-per-filter kernel (EDIM > 32): call <name>_batch_augment\n");
    last_status() = (int)cudaErrorNotSupported;
    return;
  }
  // feature-track kinds (left-null-space projection with He, ekf_c.c:66-76) exist only in the CTA kernel: they go there
  // whatever the state size
  if constexpr (M::EDIM <= THREAD_MAX_EDIM && !K::HAS_HE) {
    const unsigned grid = (unsigned)((a.B + 127) / 128);
    ekf_step_thread<M, K, PRED, UPD><<<grid, 128, 0, st>>>(a);
  } else if constexpr (M::EDIM <= 32 && !K::HAS_HE) {
    if (use_tma<M>() && (reinterpret_cast<uintptr_t>(a.P) & 15u)) {
      fprintf(stderr, "[rednose_b200] P must be 16-byte aligned (bulk-copy staging of covariance tiles)\n");
      last_status() = (int)cudaErrorMisalignedAddress;
      return;
    }
    bool paired = false;
    if constexpr (use_pair<M>()) paired = pair_enabled();
    if constexpr (use_pair<M>()) if (paired) {
      // two filters per warp (ekf_warp2.cuh): 128-bit accesses to every covariance array
      if ((reinterpret_cast<uintptr_t>(a.hP_pred) | reinterpret_cast<uintptr_t>(a.hP_filt)) & 15u) {
        fprintf(stderr, "[rednose_b200] covariance history slabs must be 16-byte aligned\n");
        last_status() = (int)cudaErrorMisalignedAddress;
        return;
      }
      constexpr int G = RNB_PAIR_GROUP;
      constexpr size_t smem = pair_smem_bytes<M, K, G>();
      const unsigned grid = (unsigned)((a.B + G - 1) / G);
      auto run = [&](void (*kern)(const StepArgs<M::NG>)) {
        if (first_launch_of((const void*)kern)) {
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        kern<<<grid, 32, smem, st>>>(a);
      };
      if constexpr (PRED && UPD) {
        if (a.idx) run(ekf_step_pair<M, K, PRED, UPD, G, true>);
        else run(ekf_step_pair<M, K, PRED, UPD, G, false>);
      } else {
        if (a.idx) {
          fprintf(stderr, "[rednose_b200] gather lists are only supported by the fused predict+update step\n");
          last_status() = (int)cudaErrorNotSupported;
          return;
        }
        run(ekf_step_pair<M, K, PRED, UPD, G, false>);
      }
    }
    if (!paired) {
      constexpr int G = RNB_GROUP, W = RNB_WARPS;
      constexpr size_t smem = warp_smem_bytes<M, K, G, W>();
      const long long per_cta = (long long)G * W;
      const unsigned grid = (unsigned)((a.B + per_cta - 1) / per_cta);
      auto run = [&](void (*kern)(const StepArgs<M::NG>)) {
        if (first_launch_of((const void*)kern)) {
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        kern<<<grid, W * 32, smem, st>>>(a);
      };
      if constexpr (PRED && UPD) {
        // the gather-list variant exists only for the fused step (what the ragged scheduler issues)
        if (a.idx) run(ekf_step_warp<M, K, PRED, UPD, G, W, true>);
        else run(ekf_step_warp<M, K, PRED, UPD, G, W, false>);
      } else {
        if (a.idx) {
          fprintf(stderr, "[rednose_b200] gather lists are only supported by the fused predict+update step\n");
          last_status() = (int)cudaErrorNotSupported;
          return;
        }
        run(ekf_step_warp<M, K, PRED, UPD, G, W, false>);
      }
    }
  } else {
    launch_step_cta<M, K, PRED, UPD>(a, st);
  }
  check(cudaGetLastError(), "ekf_step launch");
}

template <class M>
inline void fill_common(StepArgs<M::NG>& a, HostCtx<M>& ctx, long long B, const int* quat_idxs, int n_quat, int flags) {
  memset(&a, 0, sizeof(a));
  a.B = B;
  a.flags = flags;
  a.n_quat = n_quat < 0 ? 0 : (n_quat > MAX_QUAT ? MAX_QUAT : n_quat);
  for (int i = 0; i < a.n_quat; ++i) a.quat_idx[i] = quat_idxs[i];
  for (int i = 0; i < (M::NG > 0 ? M::NG : 1); ++i) a.gv[i] = ctx.gv.v[i];
  a.n_obs = 1;
}

// ----------------------------------------------------------- batched (device) ---
template <class M>
inline void batch_predict(HostCtx<M>& ctx, double* x, double* P, const double* Q, const double* dt_arr, double dt,
                          long long B, const int* quat_idxs, int n_quat, int flags,
                          double* hx_pred, double* hP_pred, void* stream) {
  StepArgs<M::NG> a;
  fill_common<M>(a, ctx, B, quat_idxs, n_quat, flags);
  a.x = x; a.P = P; a.Q = Q; a.dt_arr = dt_arr; a.dt = dt;
  a.hx_pred = hx_pred; a.hP_pred = hP_pred;
  launch_step<M, NullKind, true, false>(a, (cudaStream_t)stream);
}

template <class M, class K, bool PRED>
inline void batch_step(HostCtx<M>& ctx, double* x, double* P, const double* Q, const double* dt_arr, double dt,
                       double* z, const double* R, const double* ea, int n_obs, long long B,
                       const int* quat_idxs, int n_quat, int flags,
                       double* hx_pred, double* hP_pred, double* hx_filt, double* hP_filt, void* stream,
                       const int* idx = nullptr) {
  StepArgs<M::NG> a;
  fill_common<M>(a, ctx, B, quat_idxs, n_quat, flags);
  a.idx = idx;
  a.x = x; a.P = P; a.Q = Q; a.dt_arr = dt_arr; a.dt = dt;
  a.z = z; a.R = R; a.ea = (K::EADIM > 0) ? ea : nullptr; a.ea_dim = K::EADIM; a.n_obs = n_obs;
  a.hx_pred = hx_pred; a.hP_pred = hP_pred; a.hx_filt = hx_filt; a.hP_filt = hP_filt;
  launch_step<M, K, PRED, true>(a, (cudaStream_t)stream);
}

}  // namespace rnb
#include "ekf_maha.cuh"
namespace rnb {

// Mahalanobis distances of B observations of one kind (ekf_sym.py:626-649); out [B] on the device
template <class M, class K>
inline void batch_maha(HostCtx<M>& ctx, const double* x, const double* P, const double* z, const double* R, const double* ea,
                       long long B, int flags, double* out, void* stream) {
  if (B <= 0) return;
  // per-call scratch, allocated and released in stream order (no buffer shared between streams or devices)
  double* scratch = (double*)stream_alloc(sizeof(double) * (size_t)B * K::ZDIM * M::EDIM, (cudaStream_t)stream, "cudaMallocAsync(maha scratch)");
  if (!scratch) return;
  ekf_maha_thread<M, K><<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, P, z, R, ea, B, flags, ctx.gv, out, scratch);
  check(cudaGetLastError(), "ekf_maha launch");
  check(cudaFreeAsync(scratch, (cudaStream_t)stream), "cudaFreeAsync(maha scratch)");
}

template <class M>
inline void batch_rts(HostCtx<M>& ctx, const double* hx_pred, const double* hP_pred, const double* hx_filt, const double* hP_filt,
                      const double* t, int t_per_filter, double* xs, double* Ps, int T, long long B,
                      const int* quat_idxs, int n_quat, int norm_quats, void* stream,
                      const double* x_term = nullptr, const double* P_term = nullptr, long long k0 = 0) {
  RtsArgs<M::NG> a;
  memset(&a, 0, sizeof(a));
  a.x_term = (x_term && P_term) ? x_term : nullptr; a.P_term = (x_term && P_term) ? P_term : nullptr; a.k0 = k0;
  a.hx_pred = hx_pred; a.hP_pred = hP_pred; a.hx_filt = hx_filt; a.hP_filt = hP_filt;
  a.t = t; a.t_per_filter = t_per_filter; a.xs = xs; a.Ps = Ps; a.T = T; a.B = B; a.norm_quats = norm_quats;
  a.n_quat = n_quat < 0 ? 0 : (n_quat > MAX_QUAT ? MAX_QUAT : n_quat);
  for (int i = 0; i < a.n_quat; ++i) a.quat_idx[i] = quat_idxs[i];
  for (int i = 0; i < (M::NG > 0 ? M::NG : 1); ++i) a.gv[i] = ctx.gv.v[i];
  launch_rts_auto<M>(a, (cudaStream_t)stream);
}

// --------------------------------------------- batched, HOST buffers (stateless) ---
// Full round trip: x,P,z,R(,ea) host -> device, fused step, x,P,y device -> host, in
// chunks on alternating streams so copies overlap the kernel when the host
// buffers are pinned.  This is the batched analogue of calling the reference's
// <name>_predict + <name>_update_<kind> on caller-owned host arrays.
template <class M, class K>
inline void host_step(HostCtx<M>& ctx, double* x, double* P, const double* Q, const double* dt_arr, double dt,
                      double* z, const double* R, const double* ea, int n_obs, long long B,
                      const int* quat_idxs, int n_quat, int flags) {
  constexpr int D = M::DIM, E = M::EDIM, Z = K::ZDIM, EA = K::EADIM;
  if (B <= 0) return;
  const long long per = D + E * E + 1 + (long long)n_obs * (Z + Z * Z + EA) + 1;
  constexpr long long PAD = 8;   // each of the six sub-buffers below is rounded up to an even number of doubles
  long long chunk = (64ll << 20) / (per * 8);  // ~64 MiB of device staging per stream
  if (chunk < 1) chunk = 1;
  if (chunk > B) chunk = B;
  constexpr int NS = 3;
  struct Stage { cudaStream_t streams[NS] = {nullptr, nullptr, nullptr}; double* dbuf[NS] = {nullptr, nullptr, nullptr}; long long dcap = 0; double* dQ = nullptr; };
  static Stage stages[MAX_DEVICES];   // streams and staging buffers belong to the device that is current
  std::lock_guard<std::mutex> lk(ctx.mu);
  Stage& sg = stages[current_device()];
  cudaStream_t* streams = sg.streams;
  double** dbuf = sg.dbuf;
  long long& dcap = sg.dcap;
  double*& dQ = sg.dQ;
  for (int i = 0; i < NS; ++i)
    if (!streams[i] && !check(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking), "cudaStreamCreate")) return;
  if (!dQ && !check(cudaMalloc(&dQ, sizeof(double) * E * E), "cudaMalloc(Q)")) return;
  if (chunk * per + PAD > dcap) {
    for (int i = 0; i < NS; ++i) {
      if (dbuf[i]) cudaFree(dbuf[i]);
      dbuf[i] = nullptr;
      if (!check(cudaMalloc(&dbuf[i], sizeof(double) * (chunk * per + PAD)), "cudaMalloc(stage)")) return;
    }
    dcap = chunk * per + PAD;
  }
  if (!check(cudaMemcpy(dQ, Q, sizeof(double) * E * E, cudaMemcpyHostToDevice), "memcpy Q")) return;
  int si = 0;
  for (long long b0 = 0; b0 < B; b0 += chunk, si = (si + 1) % NS) {
    const long long nb = (B - b0 < chunk) ? (B - b0) : chunk;
    cudaStream_t st = streams[si];
    auto up2 = [](long long n) { return (n + 1) & ~1ll; };  // keep every sub-buffer 16-byte aligned (bulk copies)
    double* dx = dbuf[si];
    double* dP = dx + up2(nb * D);
    double* ddt = dP + up2(nb * E * E);
    double* dz = ddt + up2(nb);
    double* dR = dz + up2(nb * n_obs * Z);
    double* dea = dR + up2(nb * n_obs * Z * Z);
    cudaMemcpyAsync(dx, x + b0 * D, sizeof(double) * nb * D, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dP, P + b0 * E * E, sizeof(double) * nb * E * E, cudaMemcpyHostToDevice, st);
    if (dt_arr) cudaMemcpyAsync(ddt, dt_arr + b0, sizeof(double) * nb, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(dz, z + b0 * n_obs * Z, sizeof(double) * nb * n_obs * Z, cudaMemcpyHostToDevice, st);
    if (flags & FLAG_SHARED_R) cudaMemcpyAsync(dR, R, sizeof(double) * Z * Z, cudaMemcpyHostToDevice, st);
    else cudaMemcpyAsync(dR, R + b0 * n_obs * Z * Z, sizeof(double) * nb * n_obs * Z * Z, cudaMemcpyHostToDevice, st);
    if (EA > 0 && ea) cudaMemcpyAsync(dea, ea + b0 * n_obs * EA, sizeof(double) * nb * n_obs * EA, cudaMemcpyHostToDevice, st);
    StepArgs<M::NG> a;
    fill_common<M>(a, ctx, nb, quat_idxs, n_quat, flags);
    a.x = dx; a.P = dP; a.Q = dQ; a.dt_arr = dt_arr ? ddt : nullptr; a.dt = dt;
    a.z = dz; a.R = dR; a.ea = (EA > 0 && ea) ? dea : nullptr; a.ea_dim = EA; a.n_obs = n_obs;
    launch_step<M, K, true, true>(a, st);
    cudaMemcpyAsync(x + b0 * D, dx, sizeof(double) * nb * D, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(P + b0 * E * E, dP, sizeof(double) * nb * E * E, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(z + b0 * n_obs * Z, dz, sizeof(double) * nb * n_obs * Z, cudaMemcpyDeviceToHost, st);
  }
  for (int i = 0; i < NS; ++i) check(cudaStreamSynchronize(streams[i]), "host_step sync");
}

// --------------------------------------------- single filter, HOST pointers ---
// The reference's own entry points (<name>_predict / <name>_update_<kind> on caller-owned host arrays, one filter).  All
// inputs are packed into ONE pinned staging buffer and travel in one asynchronous copy each way on a private stream
// (round 1 issued 3-5 synchronous pageable cudaMemcpy per direction): copy in, launch with B = 1, copy out, one wait.
template <class M>
inline void single_predict(HostCtx<M>& ctx, double* x, double* P, const double* Q, double dt) {
  constexpr int D = M::DIM, E = M::EDIM;
  std::lock_guard<std::mutex> lk(ctx.mu);
  constexpr int DA = (D + 1) & ~1;   // P starts 16-byte aligned behind x (covariance tiles move by bulk copy / 128-bit accesses)
  constexpr size_t N = DA + 2 * E * E;
  double* d = ctx.scratch(N);
  double* h = ctx.pinned(N);
  cudaStream_t st = ctx.single_stream();
  if (!d || !h || !st) return;
  memcpy(h, x, sizeof(double) * D); memcpy(h + DA, P, sizeof(double) * E * E); memcpy(h + DA + E * E, Q, sizeof(double) * E * E);
  if (!check(cudaMemcpyAsync(d, h, sizeof(double) * N, cudaMemcpyHostToDevice, st), "memcpy in")) return;
  StepArgs<M::NG> a;
  fill_common<M>(a, ctx, 1, nullptr, 0, 0);
  a.x = d; a.P = d + DA; a.Q = d + DA + E * E; a.dt = dt;
  launch_step<M, NullKind, true, false>(a, st);
  cudaMemcpyAsync(h, d, sizeof(double) * (DA + E * E), cudaMemcpyDeviceToHost, st);
  if (!check(cudaStreamSynchronize(st), "single_predict")) return;
  memcpy(x, h, sizeof(double) * D); memcpy(P, h + DA, sizeof(double) * E * E);
}

template <class M, class K>
inline void single_update(HostCtx<M>& ctx, double* x, double* P, double* z, const double* R, const double* ea) {
  constexpr int D = M::DIM, E = M::EDIM, Z = K::ZDIM, EA = K::EADIM;
  std::lock_guard<std::mutex> lk(ctx.mu);
  constexpr int DA = (D + 1) & ~1;   // see single_predict
  constexpr size_t OZ = DA + E * E, OR_ = OZ + Z, OEA = OR_ + Z * Z, N = OEA + (EA > 0 ? EA : 1);
  double* d = ctx.scratch(N);
  double* h = ctx.pinned(N);
  cudaStream_t st = ctx.single_stream();
  if (!d || !h || !st) return;
  memcpy(h, x, sizeof(double) * D); memcpy(h + DA, P, sizeof(double) * E * E); memcpy(h + OZ, z, sizeof(double) * Z);
  memcpy(h + OR_, R, sizeof(double) * Z * Z);
  if (EA > 0 && ea) memcpy(h + OEA, ea, sizeof(double) * EA);
  if (!check(cudaMemcpyAsync(d, h, sizeof(double) * N, cudaMemcpyHostToDevice, st), "memcpy in")) return;
  StepArgs<M::NG> a;
  fill_common<M>(a, ctx, 1, nullptr, 0, 0);
  a.x = d; a.P = d + DA; a.z = d + OZ; a.R = d + OR_; a.ea = (EA > 0 && ea) ? d + OEA : nullptr; a.ea_dim = EA;
  launch_step<M, K, false, true>(a, st);
  cudaMemcpyAsync(h, d, sizeof(double) * (OZ + Z), cudaMemcpyDeviceToHost, st);
  if (!check(cudaStreamSynchronize(st), "single_update")) return;
  memcpy(x, h, sizeof(double) * D); memcpy(P, h + DA, sizeof(double) * E * E);
  // ekf_c.c:120: the innovation overwrites z (K::YDIM entries: ZDIM, or ZDIM-EADIM after projection)
  memcpy(z, h + OZ, sizeof(double) * K::YDIM);
}

// leaf functions exported with host pointers (ekf_sym.py:155-161): one-thread kernels
template <class M, class Kern>
inline void run_leaf(HostCtx<M>& ctx, Kern kern, const double* in0, int n0, const double* in1, int n1, double s,
                     double* out, int nout) {
  std::lock_guard<std::mutex> lk(ctx.mu);
  double* d = ctx.scratch((size_t)n0 + n1 + nout + 2);
  if (!d) return;
  double* d0 = d; double* d1 = d + n0; double* dout = d1 + n1;
  if (n0 > 0 && !check(cudaMemcpy(d0, in0, sizeof(double) * n0, cudaMemcpyHostToDevice), "leaf memcpy")) return;
  if (n1 > 0 && in1) cudaMemcpy(d1, in1, sizeof(double) * n1, cudaMemcpyHostToDevice);
  kern<<<1, 32>>>(d0, d1, s, ctx.gv, dout);
  check(cudaGetLastError(), "leaf launch");
  check(cudaMemcpy(out, dout, sizeof(double) * nout, cudaMemcpyDeviceToHost), "leaf memcpy back");
}

}  // namespace rnb

#define RNB_LEAF_KERNEL(KNAME, NGV, CALL)                                                              \
  __global__ void KNAME(const double* in0, const double* in1, double s, rnb::GV<NGV> gvs, double* out) { \
    const double* gv = gvs.v; (void)gv; (void)in1; (void)s;                                            \
    if (threadIdx.x == 0) { CALL; }                                                                    \
  }
