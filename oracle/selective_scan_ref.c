/*
 * ORACLE — test infrastructure only.  Never imported by the product (sigma_b200/); used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * CPU restatement of the reference's selective scan:
 *   forward  : models/encoders/selective_scan/selective_scan/selective_scan_interface.py:86-131
 *              (selective_scan_ref) == csrc/selective_scan/selective_scan_fwd_kernel.cuh:126-189
 *   backward : csrc/selective_scan/selective_scan_bwd_kernel.cuh:141-273 (same maths as torch
 *              autograd through selective_scan_ref, which is what the reference's own test compares
 *              against, test_selective_scan.py:181-224)
 * Pinned by tests/test_oracle.py against tests/golden/scan_*.npz, which were produced by the
 * reference's own selective_scan_ref (tests/golden/make_golden.py).
 *
 * All arrays contiguous fp32:  u, delta, out (batch, dim, L);  A (dim, N);  B, C (batch, G, N, L);
 * D, bias (dim) or NULL.  State and accumulation are double so that the oracle is strictly more
 * accurate than either implementation it judges.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline double softplus_d(double x) { return x <= 20.0 ? log1p(exp(x)) : x; } /* interface.py:107 */

void sigma_oracle_scan_fwd(const float *u, const float *delta, const float *A, const float *B, const float *C,
                           const float *D, const float *bias, float *out, int batch, int dim, int L, int N,
                           int G, int softplus, int nthreads) {
  const int dpg = dim / G;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
  for (long bd = 0; bd < (long)batch * dim; ++bd) {
    const int b = (int)(bd / dim), d = (int)(bd % dim), g = d / dpg; /* group: fwd_kernel.cuh:84 */
    const float *ur = u + bd * (long)L, *dr = delta + bd * (long)L;
    const float *Br = B + ((long)b * G + g) * N * (long)L, *Cr = C + ((long)b * G + g) * N * (long)L;
    double *h = (double *)calloc((size_t)N, sizeof(double));
    const double bs = bias ? bias[d] : 0.0, Dd = D ? D[d] : 0.0;
    for (int l = 0; l < L; ++l) {
      double dl = (double)dr[l] + bs;                       /* interface.py:104-105 */
      if (softplus) dl = softplus_d(dl);
      const double dlu = dl * ur[l];
      double y = 0.0;
      for (int n = 0; n < N; ++n) {
        h[n] = exp(dl * A[(long)d * N + n]) * h[n] + dlu * Br[(long)n * L + l]; /* :113,125 */
        y += h[n] * Cr[(long)n * L + l];                                       /* :127-130 */
      }
      out[bd * (long)L + l] = (float)(y + Dd * ur[l]);       /* :133 */
    }
    free(h);
  }
}

/* du, ddelta (batch,dim,L); dA (dim,N); dB, dC (batch,G,N,L); dD, dbias (dim) — all overwritten. */
void sigma_oracle_scan_bwd(const float *u, const float *delta, const float *A, const float *B, const float *C,
                           const float *D, const float *bias, const float *dout, float *du, float *ddelta,
                           float *dA, float *dB, float *dC, float *dD, float *dbias, int batch, int dim, int L,
                           int N, int G, int softplus, int nthreads) {
  const int dpg = dim / G;
  double *dA_p = (double *)calloc((size_t)batch * dim * N, sizeof(double));
  double *dD_p = (double *)calloc((size_t)batch * dim, sizeof(double));
  double *db_p = (double *)calloc((size_t)batch * dim, sizeof(double));
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int bg = 0; bg < batch * G; ++bg) {
    const int b = bg / G, g = bg % G;
    const float *Br = B + (long)bg * N * L, *Cr = C + (long)bg * N * L;
    double *dBa = (double *)calloc((size_t)N * L, sizeof(double));
    double *dCa = (double *)calloc((size_t)N * L, sizeof(double));
    double *hs = (double *)malloc((size_t)N * L * sizeof(double));
    double *dls = (double *)malloc((size_t)L * sizeof(double));
    double *dh = (double *)malloc((size_t)N * sizeof(double));
    for (int dd = 0; dd < dpg; ++dd) {
      const int d = g * dpg + dd;
      const long bd = (long)b * dim + d;
      const float *ur = u + bd * L, *dr = delta + bd * L, *dor = dout + bd * L;
      const double bs = bias ? bias[d] : 0.0, Dd = D ? D[d] : 0.0;
      /* forward recompute, keeping every state (bwd_kernel.cuh:141-171 recomputes per chunk) */
      for (int n = 0; n < N; ++n) dh[n] = 0.0;
      for (int l = 0; l < L; ++l) {
        double dl = (double)dr[l] + bs;
        if (softplus) dl = softplus_d(dl);
        dls[l] = dl;
        for (int n = 0; n < N; ++n) {
          dh[n] = exp(dl * A[(long)d * N + n]) * dh[n] + dl * ur[l] * Br[(long)n * L + l];
          hs[(long)n * L + l] = dh[n];
        }
      }
      for (int n = 0; n < N; ++n) dh[n] = 0.0;
      double dDacc = 0.0, dbacc = 0.0;
      for (int l = L - 1; l >= 0; --l) {
        const double dy = dor[l], dl = dls[l], ul = ur[l];
        double ddl = 0.0, dul = dy * Dd;                     /* bwd_kernel.cuh:143,250 */
        dDacc += dy * ul;                                     /* :144 */
        for (int n = 0; n < N; ++n) {
          const double An = A[(long)d * N + n], a = exp(dl * An);
          const double Bn = Br[(long)n * L + l], Cn = Cr[(long)n * L + l];
          const double hprev = l > 0 ? hs[(long)n * L + l - 1] : 0.0;
          dh[n] += dy * Cn;                                   /* reverse scan seed dout*C (:173-199) */
          dCa[(long)n * L + l] += dy * hs[(long)n * L + l];   /* :225 */
          const double da = dh[n] * hprev;                    /* d/da of a*h_{l-1} */
          ddl += da * a * An + dh[n] * Bn * ul;               /* :206 */
          dA_p[bd * N + n] += da * a * dl;                    /* :208 */
          dBa[(long)n * L + l] += dh[n] * dl * ul;            /* :224 */
          dul += dh[n] * dl * Bn;                             /* :205 */
          dh[n] *= a;
        }
        if (softplus) {                                       /* :241-245 */
          const double raw = (double)dr[l] + bs;
          if (raw <= 20.0) ddl *= 1.0 / (1.0 + exp(-raw));
        }
        du[bd * L + l] = (float)dul;
        ddelta[bd * L + l] = (float)ddl;
        dbacc += ddl;                                         /* :262-273 */
      }
      dD_p[bd] = dDacc;
      db_p[bd] = dbacc;
    }
    for (long i = 0; i < (long)N * L; ++i) {
      dB[(long)bg * N * L + i] = (float)dBa[i];
      dC[(long)bg * N * L + i] = (float)dCa[i];
    }
    free(dBa); free(dCa); free(hs); free(dls); free(dh);
  }
  for (int d = 0; d < dim; ++d) {
    double sD = 0.0, sb = 0.0;
    for (int b = 0; b < batch; ++b) { sD += dD_p[(long)b * dim + d]; sb += db_p[(long)b * dim + d]; }
    if (dD) dD[d] = (float)sD;
    if (dbias) dbias[d] = (float)sb;
    for (int n = 0; n < N; ++n) {
      double s = 0.0;
      for (int b = 0; b < batch; ++b) s += dA_p[((long)b * dim + d) * N + n];
      dA[(long)d * N + n] = (float)s;
    }
  }
  free(dA_p); free(dD_p); free(db_p);
}
