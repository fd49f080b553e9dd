/*
 * oracle/slq_port.c -- TEST INFRASTRUCTURE (CPU baseline), not product.
 *
 * Plain-C restatement ("port") of the reference's SLQ hot path for a CSR
 * operator, used (a) as the multi-threaded CPU baseline `bench.py` times on the
 * host cores and (b) as a second, independent checker in the tests.  It is
 * validated against `oracle/ref.py` (the NumPy restatement) in
 * `tests/test_oracle_port.py`.
 *
 * What it follows (paths under /root/reference):
 *   - probes: jax.random.rademacher via matfree/stochtrace.py:957-977 and
 *     matfree/backend/prng.py:26-29; Threefry-2x32, partitionable counters
 *     (counter = p * n + r), +1 iff the MSB of x0^x1 is 0;
 *   - Lanczos without re-orthogonalisation: matfree/decomp.py:220-292 in the
 *     exact operation order (v0 = v/|v|; a = v.Av; r = Av - a v - b v_prev;
 *     b = |r|; v_next = r / b), fp32 storage and arithmetic, all probes of a
 *     block advancing together the way jax.vmap batches them
 *     (matfree/stochtrace.py:49);
 *   - quadrature: matfree/funm.py:239-241,330-333 -- eigen-decomposition of
 *     the k x k tridiagonal (implicit QL here, LAPACK in the reference) and
 *     |v|^2 * sum_j log(theta_j) S[0,j]^2;
 *   - Hutchinson: matfree/stochtrace.py:859-863.
 * Threads: OpenMP over rows (the reference's XLA CPU backend also threads its
 * loops); compile with `gcc -O3 -fopenmp -shared -fPIC` (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static inline void threefry2x32(uint32_t k0, uint32_t k1, uint32_t* px0, uint32_t* px1) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = *px0 + ks[0], x1 = *px1 + ks[1];
  for (int i = 0; i < 5; ++i) {
    for (int j = 0; j < 4; ++j) {
      x0 += x1;
      x1 = rotl32(x1, R[i & 1][j]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
  *px0 = x0;
  *px1 = x1;
}

/* OpenMP thread count for the following calls (torchrun exports OMP_NUM_THREADS=1 to its
 * workers; the CPU arm of the benchmark must still use all the cores it can). */
void slq_port_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int slq_port_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Rademacher probes p0..p0+B-1 in blocked layout X[n][B]. */
void slq_port_probes(float* X, int64_t n, int64_t B, int64_t p0, uint32_t key0, uint32_t key1) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    for (int64_t b = 0; b < B; ++b) {
      uint64_t ctr = (uint64_t)(p0 + b) * (uint64_t)n + (uint64_t)r;
      uint32_t x0 = (uint32_t)(ctr >> 32), x1 = (uint32_t)ctr;
      threefry2x32(key0, key1, &x0, &x1);
      X[r * B + b] = ((x0 ^ x1) >> 31) ? -1.0f : 1.0f;
    }
  }
}

/* implicit QL on (d, e) carrying the first eigenvector row z; all double */
static void tridiag_ql_first_row(int k, double* d, double* e, double* z) {
  const double eps = 2.220446049250313e-16;
  for (int i = 0; i < k; ++i) z[i] = (i == 0) ? 1.0 : 0.0;
  e[k - 1] = 0.0;
  for (int l = 0; l < k; ++l) {
    int iter = 0;
    for (;;) {
      int m = l;
      for (; m < k - 1; ++m) {
        double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= eps * dd) break;
      }
      if (m == l) break;
      if (++iter > 80) break;
      double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
      double r = hypot(g, 1.0);
      g = d[m] - d[l] + e[l] / (g + copysign(r, g));
      double s = 1.0, c = 1.0, p = 0.0;
      int i, under = 0;
      for (i = m - 1; i >= l; --i) {
        double f = s * e[i], b = c * e[i];
        r = hypot(f, g);
        e[i + 1] = r;
        if (r == 0.0) {
          d[i + 1] -= p;
          e[m] = 0.0;
          under = 1;
          break;
        }
        s = f / r;
        c = g / r;
        g = d[i + 1] - p;
        r = (d[i] - g) * s + 2.0 * c * b;
        p = s * r;
        d[i + 1] = g + p;
        g = c * r - b;
        double zf = z[i + 1], zi = z[i];
        z[i + 1] = s * zi + c * zf;
        z[i] = c * zi - s * zf;
      }
      if (under) continue;
      d[l] -= p;
      e[l] = g;
      e[m] = 0.0;
    }
  }
}

/*
 * SLQ log-det quadratic forms for probes p0..p0+B-1 of key (key0, key1).
 *   quad_out[B]; alphas_out/betas_out optional [B][k].
 * Workspace is allocated here (4 block vectors of n*B floats).  Returns 0 / -1.
 */
int slq_port_csr_logdet(const int32_t* indptr, const int32_t* indices, const float* data,
                        int64_t n, int64_t B, int64_t p0, int64_t k, uint32_t key0,
                        uint32_t key1, float* quad_out, float* alphas_out, float* betas_out) {
  float* V = (float*)malloc(sizeof(float) * n * B);
  float* Vp = (float*)malloc(sizeof(float) * n * B);
  float* W = (float*)malloc(sizeof(float) * n * B);
  float* a = (float*)calloc(B, sizeof(float));
  float* b = (float*)calloc(B, sizeof(float));
  float* len = (float*)calloc(B, sizeof(float));
  double* acc = (double*)calloc(B, sizeof(double));
  float* al = (float*)malloc(sizeof(float) * B * k);
  float* be = (float*)malloc(sizeof(float) * B * k);
  if (!V || !Vp || !W || !a || !b || !len || !acc || !al || !be) return -1;
  slq_port_probes(V, n, B, p0, key0, key1);
  memset(Vp, 0, sizeof(float) * n * B);

  /* length = |v0| (funm.py:228), v0 /= length; decomp.py:227 normalises again */
  for (int pass = 0; pass < 2; ++pass) {
    memset(acc, 0, sizeof(double) * B);
#pragma omp parallel
    {
      double* loc = (double*)calloc(B, sizeof(double));
#pragma omp for schedule(static) nowait
      for (int64_t r = 0; r < n; ++r)
        for (int64_t c = 0; c < B; ++c) loc[c] += (double)(V[r * B + c] * V[r * B + c]);
#pragma omp critical
      for (int64_t c = 0; c < B; ++c) acc[c] += loc[c];
      free(loc);
    }
    float* nrm = (float*)malloc(sizeof(float) * B);
    for (int64_t c = 0; c < B; ++c) {
      nrm[c] = (float)sqrt(acc[c]);
      if (pass == 0) len[c] = nrm[c];
    }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r)
      for (int64_t c = 0; c < B; ++c) V[r * B + c] = V[r * B + c] / nrm[c];
    free(nrm);
  }

  for (int64_t j = 0; j < k; ++j) {
    /* w = A v ; a = v . w */
    memset(acc, 0, sizeof(double) * B);
#pragma omp parallel
    {
      double* loc = (double*)calloc(B, sizeof(double));
#pragma omp for schedule(static) nowait
      for (int64_t r = 0; r < n; ++r) {
        float* w = W + r * B;
        for (int64_t c = 0; c < B; ++c) w[c] = 0.0f;
        for (int32_t q = indptr[r]; q < indptr[r + 1]; ++q) {
          const float av = data[q];
          const float* x = V + (int64_t)indices[q] * B;
          for (int64_t c = 0; c < B; ++c) w[c] += av * x[c];
        }
        const float* v = V + r * B;
        for (int64_t c = 0; c < B; ++c) loc[c] += (double)(v[c] * w[c]);
      }
#pragma omp critical
      for (int64_t c = 0; c < B; ++c) acc[c] += loc[c];
      free(loc);
    }
    for (int64_t c = 0; c < B; ++c) a[c] = (float)acc[c];
    /* r = w - a v - b v_prev ; b = |r| */
    memset(acc, 0, sizeof(double) * B);
#pragma omp parallel
    {
      double* loc = (double*)calloc(B, sizeof(double));
#pragma omp for schedule(static) nowait
      for (int64_t r = 0; r < n; ++r) {
        float* w = W + r * B;
        const float* v = V + r * B;
        const float* vp = Vp + r * B;
        for (int64_t c = 0; c < B; ++c) {
          float t = w[c] - a[c] * v[c];
          t = t - b[c] * vp[c];
          w[c] = t;
          loc[c] += (double)(t * t);
        }
      }
#pragma omp critical
      for (int64_t c = 0; c < B; ++c) acc[c] += loc[c];
      free(loc);
    }
    for (int64_t c = 0; c < B; ++c) {
      b[c] = (float)sqrt(acc[c]);
      al[c * k + j] = a[c];
      be[c * k + j] = b[c];
    }
    /* v_prev <- v ; v <- r / b */
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r)
      for (int64_t c = 0; c < B; ++c) {
        Vp[r * B + c] = V[r * B + c];
        V[r * B + c] = W[r * B + c] / b[c];
      }
  }

  /* quadrature per probe */
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < B; ++c) {
    double* d = (double*)malloc(sizeof(double) * 3 * k);
    double* e = d + k;
    double* z = e + k;
    for (int64_t j = 0; j < k; ++j) {
      d[j] = (double)al[c * k + j];
      e[j] = (j < k - 1) ? (double)be[c * k + j] : 0.0;
    }
    tridiag_ql_first_row((int)k, d, e, z);
    double s = 0.0;
    for (int64_t j = 0; j < k; ++j) s += log(d[j]) * z[j] * z[j];
    quad_out[c] = (float)((double)len[c] * (double)len[c] * s);
    free(d);
  }
  if (alphas_out) memcpy(alphas_out, al, sizeof(float) * B * k);
  if (betas_out) memcpy(betas_out, be, sizeof(float) * B * k);
  free(V); free(Vp); free(W); free(a); free(b); free(len); free(acc); free(al); free(be);
  return 0;
}

/* Hutchinson trace samples v^T A v for probes p0..p0+B-1. */
int slq_port_csr_trace(const int32_t* indptr, const int32_t* indices, const float* data,
                       int64_t n, int64_t B, int64_t p0, uint32_t key0, uint32_t key1,
                       float* out) {
  float* V = (float*)malloc(sizeof(float) * n * B);
  double* acc = (double*)calloc(B, sizeof(double));
  if (!V || !acc) return -1;
  slq_port_probes(V, n, B, p0, key0, key1);
#pragma omp parallel
  {
    double* loc = (double*)calloc(B, sizeof(double));
    float* w = (float*)malloc(sizeof(float) * B);
#pragma omp for schedule(static) nowait
    for (int64_t r = 0; r < n; ++r) {
      for (int64_t c = 0; c < B; ++c) w[c] = 0.0f;
      for (int32_t q = indptr[r]; q < indptr[r + 1]; ++q) {
        const float av = data[q];
        const float* x = V + (int64_t)indices[q] * B;
        for (int64_t c = 0; c < B; ++c) w[c] += av * x[c];
      }
      const float* v = V + r * B;
      for (int64_t c = 0; c < B; ++c) loc[c] += (double)(v[c] * w[c]);
    }
#pragma omp critical
    for (int64_t c = 0; c < B; ++c) acc[c] += loc[c];
    free(loc);
    free(w);
  }
  for (int64_t c = 0; c < B; ++c) out[c] = (float)acc[c];
  free(V);
  free(acc);
  return 0;
}
