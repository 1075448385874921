/* Plain-C check of the reference-symbol single-filter entry points of the CUDA library (no Python, no torch):
 * live_predict / live_update_<k> on caller-owned host arrays, compared with the same calls on the CPU oracle library.
 *   gcc -O1 -o build/abi_single_check scripts/abi_single_check.c -ldl -lm && ./build/abi_single_check
 * (rednose/helpers/ekf_sym.py:258-343 is what normally issues these calls through cffi.) */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

typedef void (*predict_fn)(double*, double*, double*, double);
typedef void (*update_fn)(double*, double*, double*, double*, double*);
typedef int (*status_fn)(void);

static double rel_err(const double* a, const double* b, int n) {
  double d = 0, m = 0;
  for (int i = 0; i < n; ++i) { if (fabs(a[i] - b[i]) > d) d = fabs(a[i] - b[i]); if (fabs(b[i]) > m) m = fabs(b[i]); }
  return d / (m > 0 ? m : 1);
}

int main(void) {
  void* g = dlopen("rednose_b200/generated/liblive.so", RTLD_NOW | RTLD_LOCAL);
  void* c = dlopen("oracle/_ref/liblive.so", RTLD_NOW | RTLD_LOCAL);
  if (!g || !c) { printf("dlopen failed: %s\n", dlerror()); return 2; }
  predict_fn gp = (predict_fn)dlsym(g, "live_predict"), cp = (predict_fn)dlsym(c, "live_predict");
  update_fn gu12 = (update_fn)dlsym(g, "live_update_12"), cu12 = (update_fn)dlsym(c, "live_update_12");
  update_fn gu4 = (update_fn)dlsym(g, "live_update_4"), cu4 = (update_fn)dlsym(c, "live_update_4");
  status_fn st = (status_fn)dlsym(g, "live_cuda_status");
  if (!gp || !cp || !gu12 || !cu12 || !gu4 || !cu4 || !st) { printf("missing symbol\n"); return 2; }
  enum { D = 23, E = 22 };
  double x0[D] = {-2.7e6, 4.2e6, 3.8e6, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0};
  double pd[E] = {1e8, 1e8, 1e8, 100, 100, 100, 100, 100, 100, 1, 1, 1, 2.5e-3, 2.5e-3, 2.5e-3, 4e-4, 1, 1, 1, 1e-4, 1e-4, 1e-4};
  double qd[E] = {9e-4, 9e-4, 9e-4, 0, 0, 0, 0, 0, 0, 1e-2, 1e-2, 1e-2, 2.5e-9, 2.5e-9, 2.5e-9, 4e-8, 9, 9, 9, 6.9e-7, 6.9e-7, 6.9e-7};
  double xg[D], xc[D], Pg[E * E], Pc[E * E], Q[E * E];
  memset(Pg, 0, sizeof(Pg)); memset(Q, 0, sizeof(Q));
  for (int i = 0; i < E; ++i) { Pg[i * E + i] = pd[i]; Q[i * E + i] = qd[i]; }
  Q[1] = Q[E] = 1e-5;   /* one off-diagonal pair: the dense-Q path */
  memcpy(Pc, Pg, sizeof(Pg)); memcpy(xg, x0, sizeof(x0)); memcpy(xc, x0, sizeof(x0));
  double R12[9] = {25, 0, 0, 0, 25, 0, 0, 0, 25}, R4[9] = {6.25e-4, 0, 0, 0, 6.25e-4, 0, 0, 0, 6.25e-4}, ea[1] = {0};
  for (int k = 0; k < 6; ++k) {
    gp(xg, Pg, Q, 0.01); cp(xc, Pc, Q, 0.01);
    double zg[3], zc[3];
    if (k % 3 == 0) { for (int i = 0; i < 3; ++i) zg[i] = zc[i] = xc[i] + 1.0 + i; gu12(xg, Pg, zg, R12, ea); cu12(xc, Pc, zc, R12, ea); }
    else { for (int i = 0; i < 3; ++i) zg[i] = zc[i] = 0.01 * (i + 1 + k); gu4(xg, Pg, zg, R4, ea); cu4(xc, Pc, zc, R4, ea); }
    printf("step %d: cuda_status %d  rel err x %.2e  P %.2e  y %.2e\n", k, st(), rel_err(xg, xc, D), rel_err(Pg, Pc, E * E), rel_err(zg, zc, 3));
  }
  return 0;
}
