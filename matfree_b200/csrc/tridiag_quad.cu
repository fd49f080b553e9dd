// K5 / K6 -- Gauss quadrature of the Lanczos tridiagonals and the Monte-Carlo
// reduction.
//
// Replaces `dense_funm_sym_eigh` + `e1^T f(T) e1` (matfree/funm.py:239-241,
// 330-333: eigh of the k x k matrix, V diag(f(theta)) V^T, first entry) and
// `mean` / `std` over the probe axis (matfree/stochtrace.py:50,85-86).
//
// One lane per probe: a warp handles 32 probes, each lane runs the implicit-QL
// iteration with Wilkinson shifts on its own (alpha, beta) in fp64 and carries
// only the first row of the eigenvector matrix (Golub-Welsch), because the
// integrand needs sum_j f(theta_j) * S[0,j]^2 and nothing else.  The rotation
// chain of a QL sweep is strictly sequential, so there is nothing for the other
// lanes of a warp to do for ONE probe; giving every lane its own probe keeps
// all scratch accesses ([idx][probe] layout) coalesced.
#include "internal.h"

namespace mf {
namespace {

__device__ __forceinline__ double apply_fn(int fn, double param, double x) {
  switch (fn) {
    case MF_FN_LOG: return log(x);
    case MF_FN_EXP: return exp(param * x);
    case MF_FN_INV: return 1.0 / x;
    case MF_FN_SQRT: return sqrt(x);
    case MF_FN_POW: return pow(x, param);
    case MF_FN_IDENTITY: return x;
    case MF_FN_SIN: return sin(param * x);
    default: return 0.0;
  }
}

// FULL == false: z is the first eigenvector row, [k][ld].
// FULL == true : z is the whole eigenvector matrix, Z[r][j] at z[(r*k + j)*ld].
template <typename T, bool FULL>
__global__ void tridiag_ql_kernel(const T* __restrict__ alphas, const T* __restrict__ betas,
                                  const T* __restrict__ init_len, int ld, int num_probes, int k,
                                  int fn, double fn_param, T* __restrict__ quad,
                                  double* __restrict__ nodes, double* __restrict__ weights,
                                  T* __restrict__ coeffs, double* __restrict__ work,
                                  int product) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= num_probes) return;
  double* d = work + p;
  double* e = work + (int64_t)k * ld + p;
  double* z = work + (int64_t)2 * k * ld + p;
#define D(i) d[(int64_t)(i) * ld]
#define E(i) e[(int64_t)(i) * ld]
#define Z0(i) z[(int64_t)(i) * ld]
#define ZF(r, j) z[((int64_t)(r) * k + (j)) * ld]
  for (int i = 0; i < k; ++i) {
    if (product) {
      // (alphas, betas) are the diagonal / superdiagonal of an upper-bidiagonal B (Golub-Kahan,
      // matfree/decomp.py:608-750); the quadrature runs on T = B^T B, formed here in fp64:
      // T[i][i] = a_i^2 + e_{i-1}^2, T[i][i+1] = a_i e_i  (matfree/funm.py:305-319 takes the
      // SVD of B and squares the singular values -- the same spectrum)
      const double a = (double)alphas[(int64_t)i * ld + p];
      const double em = i > 0 ? (double)betas[(int64_t)(i - 1) * ld + p] : 0.0;
      D(i) = a * a + em * em;
      E(i) = (i < k - 1) ? a * (double)betas[(int64_t)i * ld + p] : 0.0;
    } else {
      D(i) = (double)alphas[(int64_t)i * ld + p];
      E(i) = (i < k - 1) ? (double)betas[(int64_t)i * ld + p] : 0.0;
    }
    if (FULL) {
      for (int r = 0; r < k; ++r) ZF(r, i) = (r == i) ? 1.0 : 0.0;
    } else {
      Z0(i) = (i == 0) ? 1.0 : 0.0;
    }
  }
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < k; ++l) {
    int iter = 0;
    while (true) {
      int m = l;
      for (; m < k - 1; ++m) {
        const double dd = fabs(D(m)) + fabs(D(m + 1));
        if (fabs(E(m)) <= eps * dd) break;
      }
      if (m == l) break;
      if (++iter > 80) break;  // no convergence: leave as is (values stay finite or NaN)
      double g = (D(l + 1) - D(l)) / (2.0 * E(l));
      double r = hypot(g, 1.0);
      g = D(m) - D(l) + E(l) / (g + copysign(r, g));
      double s = 1.0, c = 1.0, pp = 0.0;
      int i = m - 1;
      bool underflow = false;
      for (; i >= l; --i) {
        double f = s * E(i);
        const double b = c * E(i);
        r = hypot(f, g);
        E(i + 1) = r;
        if (r == 0.0) {
          D(i + 1) -= pp;
          E(m) = 0.0;
          underflow = true;
          break;
        }
        s = f / r;
        c = g / r;
        g = D(i + 1) - pp;
        r = (D(i) - g) * s + 2.0 * c * b;
        pp = s * r;
        D(i + 1) = g + pp;
        g = c * r - b;
        if (FULL) {
          for (int rr = 0; rr < k; ++rr) {
            const double zf = ZF(rr, i + 1);
            const double zi = ZF(rr, i);
            ZF(rr, i + 1) = s * zi + c * zf;
            ZF(rr, i) = c * zi - s * zf;
          }
        } else {
          const double zf = Z0(i + 1);
          const double zi = Z0(i);
          Z0(i + 1) = s * zi + c * zf;
          Z0(i) = c * zi - s * zf;
        }
      }
      if (underflow) continue;
      D(l) -= pp;
      E(l) = g;
      E(m) = 0.0;
    }
  }
  const double len = init_len ? (double)init_len[p] : 1.0;
  if (!FULL) {
    if (fn != MF_FN_NONE && quad != nullptr) {
      double acc = 0.0;
      for (int j = 0; j < k; ++j) {
        const double w = Z0(j);
        acc += apply_fn(fn, fn_param, D(j)) * (w * w);
      }
      quad[p] = (T)(len * len * acc);
    }
    if (nodes != nullptr || weights != nullptr) {
      // insertion sort by node (ascending), carrying the weights
      for (int i = 1; i < k; ++i) {
        const double di = D(i), zi = Z0(i);
        int j = i - 1;
        while (j >= 0 && D(j) > di) {
          D(j + 1) = D(j);
          Z0(j + 1) = Z0(j);
          --j;
        }
        D(j + 1) = di;
        Z0(j + 1) = zi;
      }
      for (int j = 0; j < k; ++j) {
        if (nodes) nodes[(int64_t)j * ld + p] = D(j);
        if (weights) weights[(int64_t)j * ld + p] = Z0(j) * Z0(j);
      }
    }
  } else {
    // y = f(T) e1 = sum_j f(theta_j) S[0,j] S[:,j]
    for (int r = 0; r < k; ++r) {
      double acc = 0.0;
      for (int j = 0; j < k; ++j) acc += ZF(r, j) * apply_fn(fn, fn_param, D(j)) * ZF(0, j);
      coeffs[(int64_t)r * ld + p] = (T)acc;
    }
  }
#undef D
#undef E
#undef Z0
#undef ZF
}

template <typename T>
__global__ void mc_reduce_kernel(const T* __restrict__ v, int64_t num, double* __restrict__ out) {
  // single CTA, deterministic: thread t sums elements t, t+B, ...; then a fixed tree
  __shared__ double sh[kBlock];
  __shared__ double mean_sh;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < num; i += kBlock) s += (double)v[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = kBlock / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) mean_sh = sh[0] / (double)num;
  __syncthreads();
  const double mean = mean_sh;
  double q = 0.0;
  for (int64_t i = threadIdx.x; i < num; i += kBlock) {
    const double dlt = (double)v[i] - mean;
    q += dlt * dlt;
  }
  __syncthreads();
  sh[threadIdx.x] = q;
  __syncthreads();
  for (int off = kBlock / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double sd = sqrt(sh[0] / (double)num);
    out[0] = mean;
    out[1] = sd;
    out[2] = sd / sqrt((double)num);
    out[3] = (double)num;
  }
}

}  // namespace

int32_t launch_tridiag_quad(const void* alphas, const void* betas, const void* init_len,
                            int32_t dtype, int64_t ld, int64_t num_probes, int64_t k,
                            int32_t fn, double fn_param, void* quad, double* nodes,
                            double* weights, void* coeffs, double* work, cudaStream_t st,
                            int product) {
  MF_KSCOPE(MF_KC_TRIDIAG_QUAD, st);
  if (num_probes <= 0) return MF_OK;
  const int threads = 32;
  const int blocks = (int)((num_probes + threads - 1) / threads);
  const bool full = coeffs != nullptr;
#define MF_QL(T, FULL)                                                                        \
  tridiag_ql_kernel<T, FULL><<<blocks, threads, 0, st>>>(                                     \
      (const T*)alphas, (const T*)betas, (const T*)init_len, (int)ld, (int)num_probes, (int)k, \
      fn, fn_param, (T*)quad, nodes, weights, (T*)coeffs, work, product)
  if (dtype == MF_F32) {
    if (full) MF_QL(float, true); else MF_QL(float, false);
  } else {
    if (full) MF_QL(double, true); else MF_QL(double, false);
  }
#undef MF_QL
  return check_launch("tridiag_quad");
}

int32_t launch_mc_reduce(const void* values, int32_t dtype, int64_t num, double* stats,
                         cudaStream_t st) {
  MF_KSCOPE(MF_KC_MC_REDUCE, st);
  if (dtype == MF_F32)
    mc_reduce_kernel<float><<<1, kBlock, 0, st>>>((const float*)values, num, stats);
  else
    mc_reduce_kernel<double><<<1, kBlock, 0, st>>>((const double*)values, num, stats);
  return check_launch("mc_reduce");
}

}  // namespace mf
