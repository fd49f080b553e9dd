// Host-side launchers shared between translation units (no device code here).
#pragma once

#include "common.cuh"

namespace mf {

// grid used by all reducing block-vector kernels for a problem of `total`
// elements processed `per_cta_sweep` elements per CTA sweep.
int reduce_grid(int64_t total_elems, int vec);

// ---- probe_gen.cu
int32_t launch_probe_gen(void* out, int32_t dtype, int32_t layout, int64_t n, int64_t ld,
                         int64_t p0, int64_t num_probes, uint32_t key0, uint32_t key1,
                         int32_t sampler, int32_t prng_flags, double* partial,
                         double* sqnorm_out, cudaStream_t st);

// ---- blockvec.cu : all operate on blocked vectors X[n][ld]
// partial buffers are double[kMaxPartialCtas * ld] (times nacc where noted)

// out[c] = sum_r (X[r][c] * sx[c]) * Y[r][c]         (sx may be null)
int32_t launch_dot(const void* X, const void* sx, const void* Y, int32_t dtype, int64_t n,
                   int64_t ld, double* partial, int* grid_out, cudaStream_t st);
// reduce partial[grid][ld] -> value (dtype) ; mode 0: value = sum ; mode 1: value = sqrt(sum),
// inv = 1/value.  dbl_out (optional) gets the fp64 sum.
int32_t launch_finalize(const double* partial, int grid, int64_t ld, int32_t dtype, int mode,
                        void* value_out, void* inv_out, double* dbl_out, cudaStream_t st);
// Lanczos three-term update (matfree/decomp.py:286-292, lazily normalised):
//   out = (W - a * (Rc * sc)) - bprev * (Rp * sp);  partial <- column sums of out^2
// Rp/bprev/sp may be null (first step).  out may alias Rp.
int32_t launch_lanczos_update(const void* W, const void* Rc, const void* sc, const void* a,
                              const void* Rp, const void* sp, const void* bprev, void* out,
                              int32_t dtype, int64_t n, int64_t ld, double* partial,
                              int* grid_out, cudaStream_t st);
// out = X * s (mode 0) or X / s (mode 1), per column; s may be null (copy)
int32_t launch_scale(const void* X, const void* s, void* out, int mode, int32_t dtype,
                     int64_t n, int64_t ld, cudaStream_t st);
// CGS pass, dots: h[j][c] = sum_r Q[j][r][c] * V[r][c], j = 0..nq-1
// (matfree/decomp.py:463,468).  partial: double[nq][kMaxPartialCtas*ld]
int32_t launch_reorth_dots(const void* Q, int64_t nq, const void* V, int32_t dtype, int64_t n,
                           int64_t ld, double* partial, void* h_out, cudaStream_t st);
// CGS pass, update: V <- V - sum_j Q[j] * h[j]  (decomp.py:464,468); optional
// column sums of the new V^2 into partial (norm fused).
int32_t launch_reorth_update(const void* Q, int64_t nq, const void* h, void* V, int32_t dtype,
                             int64_t n, int64_t ld, double* partial, int* grid_out,
                             cudaStream_t st);
// out[r][c] = scale[c] * sum_j Q[j][r][c] * coeff[j][c]
int32_t launch_basis_combine(const void* Q, const void* coeffs, const void* scale,
                             int32_t dtype, int64_t n, int64_t ld, int64_t k, void* out,
                             cudaStream_t st);
int32_t launch_transpose(const void* src, void* dst, int32_t dtype, int64_t n,
                         int64_t num_probes, int64_t ld, bool to_blocked, cudaStream_t st);
// tiny helpers on [ld] scalar rows
int32_t launch_full_offdiag(void* betas_prev_row, const void* h_row, int32_t dtype, int64_t ld,
                            cudaStream_t st);

// ---- spmm_csr.cu
// W = s * (A @ X) per column (s may be null); if partial != null also
// partial <- column sums of (X*s) * W   (the Lanczos alpha, decomp.py:288)
int32_t launch_spmm_csr(const int32_t* indptr, const int32_t* indices, const void* data,
                        int64_t n, int64_t nnz, int32_t dtype, const void* X, const void* s,
                        void* W, int64_t ld, double* partial, int* grid_out, cudaStream_t st);

// ---- gemm.cu
// C[M][ld] = op(A) @ B[K][ld]; A is [M][K] (trans=0, lda>=K) or [K][M] (trans=1, lda>=M)
// colscale (optional, [ld]): C[:, c] is multiplied by colscale[c] in the epilogue
int32_t launch_gemm_blocked(const void* A, int64_t lda, bool trans, int64_t M, int64_t K,
                            const void* B, const void* colscale, void* C, int64_t ld,
                            int32_t dtype, cudaStream_t st);
int32_t launch_gemm_simt(const void* A, int64_t lda, bool trans, int64_t M, int64_t K,
                         const void* B, const void* colscale, void* C, int64_t ld,
                         int32_t dtype, cudaStream_t st);

// ---- tridiag_quad.cu
int32_t launch_tridiag_quad(const void* alphas, const void* betas, const void* init_len,
                            int32_t dtype, int64_t ld, int64_t num_probes, int64_t k,
                            int32_t fn, double fn_param, void* quad, double* nodes,
                            double* weights, void* coeffs, double* work, cudaStream_t st);
int32_t launch_mc_reduce(const void* values, int32_t dtype, int64_t num, double* stats,
                         cudaStream_t st);

}  // namespace mf
