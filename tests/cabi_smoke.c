/* cabi_smoke.c -- reaches libagb200.so through nothing but include/agb200.h (no Python, no C++):
 * alloc -> h2d -> ag_gemm (+ ag_gemm_epilogue) -> d2h, checked against a plain C triple loop.
 * Build: gcc -std=c99 -I include tests/cabi_smoke.c -L autograd.jl_b200 -lagb200 -Wl,-rpath,autograd.jl_b200 -lm
 * Exit code 0 = pass.  The reference call sites these entry points replace: src/core.jl:66 (`*` forward through
 * LinearAlgebra), src/base.jl:26 (its gradient products). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "agb200.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    ag_status st_ = (call);                                                \
    if (st_ != AG_OK) {                                                    \
      char msg[512];                                                       \
      ag_last_error(msg, sizeof msg);                                      \
      fprintf(stderr, "%s -> status %d: %s\n", #call, (int)st_, msg);      \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(void) {
  const int64_t M = 96, K = 200, N = 1500;        /* 2MNK = 5.8e7: the tcgen05 engine takes it */
  float *a = malloc(sizeof(float) * M * K), *b = malloc(sizeof(float) * K * N), *c = malloc(sizeof(float) * M * N);
  float *bias = malloc(sizeof(float) * M), *z = malloc(sizeof(float) * M * N);
  unsigned s = 12345u;
  for (int64_t i = 0; i < M * K; ++i) { s = s * 1664525u + 1013904223u; a[i] = (float)(s >> 8) / 8388608.0f - 1.0f; }
  for (int64_t i = 0; i < K * N; ++i) { s = s * 1664525u + 1013904223u; b[i] = (float)(s >> 8) / 8388608.0f - 1.0f; }
  for (int64_t i = 0; i < M; ++i) bias[i] = 0.01f * (float)i;

  CHECK(ag_init(0));
  void *da, *db, *dc, *dz, *dbias;
  CHECK(ag_alloc(sizeof(float) * M * K, &da));
  CHECK(ag_alloc(sizeof(float) * K * N, &db));
  CHECK(ag_alloc(sizeof(float) * M * N, &dc));
  CHECK(ag_alloc(sizeof(float) * M * N, &dz));
  CHECK(ag_alloc(sizeof(float) * M, &dbias));
  CHECK(ag_copy_h2d(da, a, sizeof(float) * M * K));
  CHECK(ag_copy_h2d(db, b, sizeof(float) * K * N));
  CHECK(ag_copy_h2d(dbias, bias, sizeof(float) * M));
  CHECK(ag_gemm(AG_F32, 0, 0, M, N, K, 1.0, da, M, db, K, 0.0, dc, M));                /* column-major, BLAS style */
  CHECK(ag_gemm_epilogue(AG_F32, 0, 0, M, N, K, da, M, db, K, dbias, AG_U_RELU, dz, NULL, M));
  CHECK(ag_copy_d2h(c, dc, sizeof(float) * M * N));
  CHECK(ag_copy_d2h(z, dz, sizeof(float) * M * N));
  char engine[32];
  CHECK(ag_gemm_last_engine(engine, sizeof engine));

  double worst = 0.0, worst_z = 0.0, scale = 0.0;
  for (int64_t n = 0; n < N; ++n)
    for (int64_t m = 0; m < M; ++m) {
      double acc = 0.0;
      for (int64_t k = 0; k < K; ++k) acc += (double)a[m + k * M] * (double)b[k + n * K];
      double r = acc + (double)bias[m];
      r = r > 0.0 ? r : 0.0;
      if (fabs(acc) > scale) scale = fabs(acc);
      if (fabs(acc - c[m + n * M]) > worst) worst = fabs(acc - c[m + n * M]);
      if (fabs(r - z[m + n * M]) > worst_z) worst_z = fabs(r - z[m + n * M]);
    }
  CHECK(ag_free(da));
  CHECK(ag_free(db));
  CHECK(ag_free(dc));
  CHECK(ag_free(dz));
  CHECK(ag_free(dbias));
  CHECK(ag_shutdown());
  printf("cabi_smoke: engine %s, max|err| %.3e (product) %.3e (relu(product + bias)), max|C| %.3e\n", engine, worst,
         worst_z, scale);
  free(a); free(b); free(c); free(bias); free(z);
  return (worst <= 1e-5 * scale && worst_z <= 1e-5 * scale + 1e-6) ? 0 : 2;
}
