/*
 * matfree_b200.h -- C ABI of libmatfree_b200.so
 *
 * B200 (sm_100a) kernels for matfree's stochastic-Lanczos-quadrature hot path.
 * The reference (pnkraemer/matfree) is pure Python over JAX and has no FFI of
 * its own; each entry point below replaces the XLA lowering of the reference
 * lines cited next to it (paths relative to the reference checkout) and is what
 * a `jax.ffi.ffi_call` target for that function binds.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name starts with `h_`;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as
 *     void*); nothing here synchronises the device or the stream, allocates or
 *     frees device memory, or keeps a pointer after returning -- with the
 *     documented exceptions of the profiling aid mf_timing_collect and of the
 *     set-up / tear-down calls of the multi-GPU communicator (mf_comm_*);
 *   - scratch memory is caller-provided (`workspace`, sized by the matching
 *     `*_workspace_bytes` function);
 *   - return value: 0 on success, a negative MF_ERR_* code otherwise;
 *     `mf_last_error()` returns a thread-local message for the last failure;
 *   - `dtype`: MF_F32 or MF_F64 (the arithmetic type of vectors and operator
 *     values; reductions always accumulate in fp64);
 *   - block vectors ("blocked layout") are row-major `X[n][ld]`: the `ld`
 *     probes of a tile are contiguous for each of the n rows.  `ld` must be a
 *     power of two in [1, 256].  The reference layout is probe-major `(P, n)`
 *     (matfree/stochtrace.py:961-962).
 */
#ifndef MATFREE_B200_H_
#define MATFREE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MF_ABI_VERSION 7

/* status codes */
#define MF_OK 0
#define MF_ERR_INVALID_ARGUMENT (-1)
#define MF_ERR_CUDA (-2)
#define MF_ERR_UNSUPPORTED (-3)
#define MF_ERR_WORKSPACE (-4)
#define MF_ERR_PEER_TIMEOUT (-5) /* a rank of a communicator did not arrive (mf_comm_status) */

/* dtypes */
#define MF_F32 0
#define MF_F64 1

/* sampler kinds (matfree/stochtrace.py:927-937) */
#define MF_SAMPLER_SIGNS 0  /* sampler_signs  -> jax.random.rademacher */
#define MF_SAMPLER_NORMAL 1 /* sampler_normal -> jax.random.normal     */

/* prng flags */
#define MF_PRNG_X64_BITS 1 /* jax_enable_x64: rademacher consumes a 64-bit draw */

/* probe layouts */
#define MF_LAYOUT_PROBE_MAJOR 0 /* out[p * ld + r], reference layout (P, n)   */
#define MF_LAYOUT_BLOCKED 1     /* out[r * ld + b], blocked layout   [n][ld] */

/* operator kinds (new surface; the reference takes any callable,
 * matfree/stochtrace.py:47-49) */
#define MF_OP_DENSE 0 /* A[n][n] row-major, symmetric            */
#define MF_OP_CSR 1   /* indptr[n+1] i32, indices[nnz] i32, data */
#define MF_OP_GRAM 2  /* A[m][n] row-major; operator is A^T A    */

/* reorthogonalisation (matfree/decomp.py:30-37) */
#define MF_REORTHO_NONE 0 /* three-term Lanczos, decomp.py:220-292          */
#define MF_REORTHO_FULL 1 /* Arnoldi + CGS2, T=(H+H^T)/2, decomp.py:125-145 */

/* matrix functions fused into the quadrature (matfree/funm.py:186-202,322-335) */
#define MF_FN_NONE 0     /* only nodes/weights                  */
#define MF_FN_LOG 1      /* log(x)        (logdet)              */
#define MF_FN_EXP 2      /* exp(param*x)                        */
#define MF_FN_INV 3      /* 1/x                                 */
#define MF_FN_SQRT 4     /* sqrt(x)                             */
#define MF_FN_POW 5      /* x^param                             */
#define MF_FN_IDENTITY 6 /* x                                   */
#define MF_FN_SIN 7      /* sin(param*x)                        */

/* integrands for mf_estimate_* */
#define MF_INTEGRAND_SLQ 0   /* funm.monte_carlo_funm_sym[_logdet], funm.py:186-243 */
#define MF_INTEGRAND_TRACE 1 /* stochtrace.monte_carlo_trace, stochtrace.py:853-865 */

/* An operator; plain pointers and sizes only. */
typedef struct mf_operator {
  int32_t kind;           /* MF_OP_*                                     */
  int32_t dtype;          /* MF_F32 / MF_F64                             */
  int64_t n;              /* operator is n x n                           */
  int64_t m;              /* GRAM: rows of A; otherwise ignored          */
  int64_t nnz;            /* CSR only                                    */
  const void* values;     /* DENSE A[n][n]; GRAM A[m][n]; CSR data[nnz]  */
  const int32_t* indptr;  /* CSR only                                    */
  const int32_t* indices; /* CSR only                                    */
  int64_t lda;            /* DENSE/GRAM leading dimension (elements)     */
  const void* split_planes; /* DENSE/GRAM fp32, optional: TF32 planes [2][rows][lda]
                             * of `values` made by mf_operator_split.  With them the
                             * operator runs on the tcgen05 tensor cores (3xTF32);
                             * without them on the CUDA-core kernel.                */
  int32_t csr_max_row_nnz; /* CSR, optional hint: longest row.  Above 128 the product takes
                            * the load-balanced route for irregular matrices (long rows cut
                            * into 512-non-zero segments, one CTA each); 0 = unknown/regular */
  int64_t csr_bandwidth;   /* CSR, optional hint: max |column - row|; 0 = unknown.  When the rows
                            * `bandwidth` apart (the planes of a 3-D stencil) would not survive
                            * in L2 between their uses, the product walks the rows in a blocked
                            * order (see csrc/spmm_csr.cu) -- same results, fewer DRAM re-reads */
  int32_t csr_num_diagonals; /* CSR, optional hint: number of distinct diagonals (column - row) the
                              * entries lie on, 255 = more than 8; 0 = unknown.  The TMA-staged
                              * band kernel (csrc/spmm_tma.cu) is chosen from the average row
                              * length when this is 0, and only for matrices with exactly 5 (or 7)
                              * diagonals when it is known -- an irregular matrix that merely
                              * averages 5 entries per row then stays on the row-group kernel.
                              * Either way the kernel verifies every chunk and is correct for
                              * any CSR matrix. */
  int64_t csr_line_stride;   /* CSR stencils, optional hint: the smallest diagonal offset above 1,
                              * i.e. the length of a grid line (2-D: = csr_bandwidth; 3-D: the
                              * +-line neighbour); 0 = unknown.  With it the TMA kernels walk
                              * down strips of the grid (csrc/spmm_tma.cu) */
} mf_operator_t;

const char* mf_last_error(void);
int32_t mf_abi_version(void);
/* Number of CUDA kernels this library has launched in the calling process. */
int64_t mf_launch_count(void);

/* Profiling aid (not on the hot path): when enabled, every kernel launch of this
 * library is bracketed by CUDA events recorded on the launching stream.
 * mf_timing_collect SYNCHRONISES on those events (the one documented exception to
 * the no-sync rule; call it only after the stream has been synchronised), adds the
 * elapsed milliseconds and launch counts per kernel class into the HOST arrays
 * h_ms[MF_KC_COUNT] / h_launches[MF_KC_COUNT], and resets the recorder. */
#define MF_KC_PROBE_GEN 0
#define MF_KC_SPMM_CSR 1
#define MF_KC_GEMM 2
#define MF_KC_DOT 3
#define MF_KC_FINALIZE 4
#define MF_KC_LANCZOS_UPDATE 5
#define MF_KC_SCALE 6
#define MF_KC_REORTH_DOTS 7
#define MF_KC_REORTH_UPDATE 8
#define MF_KC_TRIDIAG_QUAD 9
#define MF_KC_MC_REDUCE 10
#define MF_KC_OTHER 11
#define MF_KC_COUNT 12
int32_t mf_timing_enable(int32_t on);
int32_t mf_timing_collect(double* h_ms, int64_t* h_launches);

/* K1 -- probe generator.  Replaces jax.random.rademacher / jax.random.normal
 * as called by matfree/stochtrace.py:957-964 (via matfree/backend/prng.py:14-29):
 * Threefry-2x32 in JAX's partitionable counter mode, counter = p * n + r for
 * probe p, component r.  Writes probes p0 .. p0+num_probes-1.  In blocked
 * layout columns num_probes .. ld-1 are zero-filled.  `sqnorm_out` (optional,
 * double[ld or num_probes]) receives the squared 2-norm of every probe. */
int32_t mf_probe_gen(void* out, int32_t dtype, int32_t layout, int64_t n, int64_t ld,
                     int64_t p0, int64_t num_probes, uint32_t key0, uint32_t key1,
                     int32_t sampler, int32_t prng_flags, double* sqnorm_out,
                     void* stream);

/* The same stream for ONE RANK's rows of a row-sharded sample array (blocked layout only):
 * out[rows][ld] holds components row0 .. row0+rows-1 of probes p0 .. p0+num_probes-1 of length
 * n_total, i.e. counter = p * n_total + row0 + r -- bit-identical to the corresponding rows of
 * the single-device array, so a row-sharded estimate does not depend on the number of ranks. */
int32_t mf_probe_gen_rows(void* out, int32_t dtype, int64_t n_total, int64_t row0, int64_t rows,
                          int64_t ld, int64_t p0, int64_t num_probes, uint32_t key0,
                          uint32_t key1, int32_t sampler, int32_t prng_flags, void* stream);

/* One-time preprocessing of an fp32 dense / Gram operator for the tensor-core
 * path: the two TF32 planes of its matrix, hi = rna_tf32(A), lo = rna_tf32(A - hi),
 * written to planes[2][rows][lda] (rows = n, or m for GRAM).  The fp32 product is
 * then formed as A_hi X_hi + A_hi X_lo + A_lo X_hi with fp32 accumulation in TMEM
 * ("3xTF32"), which is what replaces XLA's fp32 `dot_general` of the user matvec
 * (matfree/stochtrace.py:47-49; tutorials/1_log_determinants.py:19-21).
 * mf_operator_split_bytes returns 0 for operators that have no planes (CSR, fp64). */
int64_t mf_operator_split_bytes(const mf_operator_t* op);
/* Tuning / cross-check knob of the dense path (process-wide).  variant = how many
 * 32-k stages the tensor core sums in TMEM before the partial sum moves to the fp32
 * register accumulators (0: two stages, the default; 1: one; 2: four -- fewer TMEM
 * drains, more truncation drift).  use_tensor_cores = 0 routes dense / Gram operators to the CUDA-core
 * kernel even when TF32 planes are present (used by the tests to cross-check). */
int32_t mf_gemm_config(int32_t variant, int32_t use_tensor_cores);
int32_t mf_operator_split(const mf_operator_t* op, void* planes, void* stream);
/* Tuning / cross-check knobs of the CSR product (process-wide; negative / zero = keep).
 * use_band_kernel: 0 = the row-group gather kernel; 1 = banded / stencil matrices on tiles of at
 * least one warp per row take the band kernel (csrc/spmm_strip.cu: register window over adjacent
 * diagonals, TMA bulk copies of the CSR metadata); 2 = 5-diagonal band matrices on the 256-wide
 * fp32 tile take the TMA-staged kernels (csrc/spmm_tma.cu: the X rows of a chunk land in shared
 * memory by cp.async.bulk ahead of their use; 2-D stencils whose line length is known from
 * csr_bandwidth are walked down strips of the grid so that every X row is staged once; 3-D
 * stencils in the blocked row order take the 7-diagonal variant); 3 = as 2, every 7-diagonal
 * matrix takes the TMA kernel; 4 = as 3, and the strip walk whatever the problem size; 5 = as 3
 * without the strip walk (3-5: tests and A/B runs).  All produce the same bits; the default is 2,
 * the one measured fastest on B200 (profiles/); matrices / tiles a kernel does not take fall
 * through to the row-group gather kernel.
 * rows_per_chunk (default 64), prefetch_rows (> 0: L2 prefetch distance, < 0: L1 prefetch
 * distance, 0: off; <= -100 keeps the current value), min_ctas_per_sm (3 or 4: register budget
 * 80 / 64) configure the band kernel. */
int32_t mf_spmm_config(int32_t use_band_kernel, int32_t rows_per_chunk, int32_t prefetch_rows,
                       int32_t min_ctas_per_sm);

/* The user matvec (matfree/stochtrace.py:47-49, funm.py:231-235,
 * decomp.py:163-164) applied to a whole probe block:
 * W[n][ld] = A @ X[n][ld] (blocked layout).  Dense and Gram operators need
 * scratch (TF32 planes of X, the intermediate A X): mf_matmat_workspace_bytes. */
int64_t mf_matmat_workspace_bytes(const mf_operator_t* op, int64_t ld);
int32_t mf_matmat(const mf_operator_t* op, const void* X, void* W, int64_t ld, void* workspace,
                  int64_t workspace_bytes, void* stream);
int32_t mf_matmat_dense(const void* A, const void* A_planes, int64_t n, int64_t lda,
                        int32_t dtype, const void* X, void* W, int64_t ld, void* workspace,
                        int64_t workspace_bytes, void* stream);
int32_t mf_matmat_csr(const int32_t* indptr, const int32_t* indices, const void* data,
                      int64_t n, int64_t nnz, int32_t dtype, const void* X, void* W,
                      int64_t ld, void* stream);
int32_t mf_matmat_gram(const void* A, const void* A_planes, int64_t m, int64_t n, int64_t lda,
                       int32_t dtype, const void* X, void* W, int64_t ld, void* workspace,
                       int64_t workspace_bytes, void* stream);

/* Layout helpers: probe-major (P, n) <-> blocked [n][ld]. */
int32_t mf_to_blocked(const void* src_pn, void* dst_blocked, int32_t dtype, int64_t n,
                      int64_t num_probes, int64_t ld, void* stream);
int32_t mf_from_blocked(const void* src_blocked, void* dst_pn, int32_t dtype, int64_t n,
                        int64_t num_probes, int64_t ld, void* stream);

/* decomp.tridiag_sym (matfree/decomp.py:30-122) on a probe block.
 *   V0        in : start block [n][ld] (blocked), NOT normalised; preserved
 *   alphas    out: [k][ld]   diagonal of T           (row j = step j)
 *   betas     out: [k][ld]   row j < k-1: off-diagonal j; row k-1: residual norm
 *                            (reortho=NONE) / norm of the last residual (FULL)
 *   init_len  out: [ld]      |V0[:, b]|  (init_length_inv = 1/init_len)
 *   Q         out: optional [k][n][ld] basis (required for reortho=FULL);
 *                  Q[j] is the j-th Lanczos vector of every probe
 *   residual  out: optional [n][ld]; NONE: b_{k-1} * v_k  (decomp.py:167);
 *                  FULL: last un-normalised vector (decomp.py:141-143)
 * All scalar outputs are in `dtype`. */
int64_t mf_lanczos_workspace_bytes(const mf_operator_t* op, int64_t ld, int64_t k,
                                   int32_t reortho, int32_t want_Q);
int32_t mf_lanczos(const mf_operator_t* op, const void* V0, int64_t ld, int64_t k,
                   int32_t reortho, void* alphas, void* betas, void* init_len, void* Q,
                   void* residual, void* workspace, int64_t workspace_bytes, void* stream);

/* decomp.hessenberg (matfree/decomp.py:351-477): Arnoldi factorisation A Q^T ~ Q^T H of an
 * ARBITRARY square operator on a block of start vectors -- the loop of mf_lanczos(reortho=FULL)
 * keeping the whole upper-Hessenberg matrix instead of its symmetrised tridiagonal part.
 *   H   out: [k][k][ld], H[r][i][c] (column i = first-pass coefficients Q^T A q_i, the norm of
 *            the orthogonalised vector below the diagonal), zero elsewhere
 *   reortho: MF_REORTHO_NONE = one Gram-Schmidt pass, MF_REORTHO_FULL = two (decomp.py:467-468)
 *   Q   out: [k][n][ld] basis;  residual out (optional): the un-normalised last vector */
int64_t mf_hessenberg_workspace_bytes(const mf_operator_t* op, int64_t ld, int64_t k);
int32_t mf_hessenberg(const mf_operator_t* op, const void* V0, int64_t ld, int64_t k,
                      int32_t reortho, void* H, void* init_len, void* Q, void* residual,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* Building blocks of a ROW-SHARDED decomposition (operators too large for one GPU, or
 * BASELINE config 4): the same kernels mf_lanczos chains, one call each, with every
 * reduction stopping at THIS device's fp64 partial sums so the driver can all-reduce
 * them across the row shards (NCCL) before the next call.  `n` is the number of LOCAL
 * rows; block vectors are [n][ld] as everywhere.  Replaces, per shard,
 * matfree/decomp.py:454-477 (Arnoldi step with CGS twice) and :286-292 (three-term step).
 *   mf_block_dot      sums[c]    = sum_r X[r][c] * Y[r][c]
 *   mf_reorth_dots    sums[j][c] = sum_r Q[j][r][c] * V[r][c],  j < nq   (decomp.py:463,468)
 *   mf_reorth_update  V -= sum_j Q[j] * h[j];  optional sqnorm[c] = sum_r V[r][c]^2 (:464,468,471)
 *   mf_lanczos_update out = (W - a (Rc*sc)) - bprev (Rp*sp);  sqnorm[c] = sum_r out^2 (:289-290)
 *   mf_block_scale    out = X * s  or  X / s  per column                (:456-457)
 *   mf_sums_finalize  value = sums or sqrt(sums), inv = 1/value, cast to dtype
 *   mf_full_offdiag   offdiag = (offdiag + h) / 2          (T = (H + H^T)/2, :133-135) */
int64_t mf_blockvec_workspace_bytes(int64_t ld, int64_t max_nq);
int32_t mf_block_dot(const void* X, const void* Y, int32_t dtype, int64_t n, int64_t ld,
                     double* sums, void* workspace, int64_t workspace_bytes, void* stream);
int32_t mf_reorth_dots(const void* Q, int64_t nq, const void* V, int32_t dtype, int64_t n,
                       int64_t ld, double* sums, void* workspace, int64_t workspace_bytes,
                       void* stream);
int32_t mf_reorth_update(const void* Q, int64_t nq, const void* h, void* V, int32_t dtype,
                         int64_t n, int64_t ld, double* sqnorm, void* workspace,
                         int64_t workspace_bytes, void* stream);
int32_t mf_lanczos_update(const void* W, const void* Rc, const void* sc, const void* a,
                          const void* Rp, const void* sp, const void* bprev, void* out,
                          int32_t dtype, int64_t n, int64_t ld, double* sqnorm, void* workspace,
                          int64_t workspace_bytes, void* stream);
int32_t mf_block_scale(const void* X, const void* s, void* out, int32_t divide, int32_t dtype,
                       int64_t n, int64_t ld, void* stream);
int32_t mf_sums_finalize(const double* sums, int64_t count, int32_t take_sqrt, void* value,
                         void* inv, int32_t dtype, void* stream);
int32_t mf_full_offdiag(void* offdiag_row, const void* h_row, int32_t dtype, int64_t ld,
                        void* stream);

/* K5 -- Gauss quadrature of the tridiagonal matrices of a probe block:
 * replaces eigh + V f(L) V^T + e1^T(.)e1 of matfree/funm.py:239-241,330-333.
 * One lane per probe, implicit-QL with the first eigenvector row only.
 *   alphas [k][ld], betas [k][ld] as produced by mf_lanczos
 *   quad   out: [ld] (dtype)  init_len^2 * sum_j f(theta_j) w_j  (unless fn==MF_FN_NONE)
 *   nodes  out: optional double [k][ld] Ritz values, ascending
 *   weights out: optional double [k][ld] squared first eigenvector components */
int64_t mf_tridiag_quad_workspace_bytes(int64_t ld, int64_t k);
int32_t mf_tridiag_quad(const void* alphas, const void* betas, const void* init_len,
                        int32_t dtype, int64_t ld, int64_t num_probes, int64_t k, int32_t fn,
                        double fn_param, void* quad, double* nodes, double* weights,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* K6 -- Monte-Carlo reduction (matfree/stochtrace.py:50,85-86):
 * stats_out = double[4] {mean, std(ddof=0), sem = std/sqrt(P), P}. */
int32_t mf_mc_reduce(const void* values, int32_t dtype, int64_t num, double* stats_out,
                     void* stream);

/* Fused estimator: everything `estimate(matvec, key)` does for one probe range
 * (matfree/stochtrace.py:47-50 with the SLQ or Hutchinson integrand):
 * generate probes p0..p0+num_probes-1 of the (P_total, n) sample array of
 * `key`, run k Lanczos steps per probe (tile by tile, `ld` probes per tile),
 * quadrature, and write one value per probe to `quad_out[num_probes]` (dtype).
 * Probes are never materialised in the reference layout.  The caller reduces
 * `quad_out` with mf_mc_reduce (after gathering the shards of other GPUs).
 * Optional per-tile outputs (tile t = probes t*ld .. t*ld+ld-1 of the range):
 * alphas_out / betas_out [tiles][k][ld], lens_out [tiles][ld] (|v0| per probe). */
int64_t mf_estimate_workspace_bytes(const mf_operator_t* op, int64_t ld, int64_t k,
                                    int32_t reortho, int32_t integrand);
int32_t mf_estimate(const mf_operator_t* op, int32_t integrand, int32_t sampler,
                    int32_t prng_flags, uint32_t key0, uint32_t key1, int64_t p0,
                    int64_t num_probes, int64_t ld, int64_t k, int32_t reortho, int32_t fn,
                    double fn_param, void* quad_out, void* alphas_out, void* betas_out,
                    void* lens_out, void* workspace, int64_t workspace_bytes, void* stream);

/* Per-kind spellings of the fused estimator (what the per-kind FFI targets bind). */
int32_t mf_slq_estimate_dense(const void* A, const void* A_planes, int64_t n, int64_t lda,
                              int32_t dtype, int32_t sampler, int32_t prng_flags, uint32_t key0,
                              uint32_t key1, int64_t p0, int64_t num_probes, int64_t ld,
                              int64_t k, int32_t reortho, int32_t fn, double fn_param,
                              void* quad_out, void* workspace, int64_t workspace_bytes,
                              void* stream);
int32_t mf_slq_estimate_csr(const int32_t* indptr, const int32_t* indices, const void* data,
                            int64_t n, int64_t nnz, int32_t dtype, int32_t sampler,
                            int32_t prng_flags, uint32_t key0, uint32_t key1, int64_t p0,
                            int64_t num_probes, int64_t ld, int64_t k, int32_t reortho,
                            int32_t fn, double fn_param, void* quad_out, void* workspace,
                            int64_t workspace_bytes, void* stream);
int32_t mf_slq_estimate_gram(const void* A, const void* A_planes, int64_t m, int64_t n,
                             int64_t lda, int32_t dtype, int32_t sampler, int32_t prng_flags, uint32_t key0,
                             uint32_t key1, int64_t p0, int64_t num_probes, int64_t ld,
                             int64_t k, int32_t reortho, int32_t fn, double fn_param,
                             void* quad_out, void* workspace, int64_t workspace_bytes,
                             void* stream);

/* Golub-Kahan bidiagonalisation support (decomp.bidiag, matfree/decomp.py:608-750, and
 * funm.monte_carlo_funm_product[_logdet | _schatten_norm], matfree/funm.py:246-319).
 *   mf_matmat_rect: W = A X (trans = 0: X [n][ld] -> W [m][ld]) or W = A^T X (trans = 1:
 *     X [m][ld] -> W [n][ld]) for a dense row-major A [m][lda >= n] -- the matvec and its `vjp`
 *     (decomp.py:703,712) on a probe block; fp32 with A_planes (mf_operator_split of a GRAM
 *     operator over the same A) runs on tcgen05 (3xTF32), fp64 on the FP64 tensor cores.
 *   mf_bidiag_quad: as mf_tridiag_quad, for an upper-bidiagonal B given by its diagonal
 *     `alphas` [k][ld] and superdiagonal `betas` [k][ld] (row i = B[i][i+1], row k-1 unused):
 *     init_len^2 * e1^T f(B^T B) e1 -- the reference takes the SVD of B and squares the singular
 *     values (funm.py:305-319); here T = B^T B is formed in fp64 inside the kernel.
 * The recurrence itself is driven with the building blocks above (mf_lanczos_update,
 * mf_reorth_dots / _update, mf_block_scale, mf_sums_finalize). */
int64_t mf_matmat_rect_workspace_bytes(int64_t m, int64_t n, int64_t lda, int32_t trans,
                                       int32_t dtype, int32_t have_planes, int64_t ld);
int32_t mf_matmat_rect(const void* A, const void* A_planes, int64_t m, int64_t n, int64_t lda,
                       int32_t trans, int32_t dtype, const void* X, void* W, int64_t ld,
                       void* workspace, int64_t workspace_bytes, void* stream);
int32_t mf_bidiag_quad(const void* alphas, const void* betas, const void* init_len,
                       int32_t dtype, int64_t ld, int64_t num_probes, int64_t k, int32_t fn,
                       double fn_param, void* quad, double* nodes, double* weights,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* funm.funm_lanczos_sym (matfree/funm.py:114-147) given a stored basis:
 * out[n][ld] = init_len * sum_j Q[j] * y[j],  y = f(T) e1 per probe
 * (coeffs [k][ld], dtype).  `mf_tridiag_funm_e1` computes y. */
int32_t mf_tridiag_funm_e1(const void* alphas, const void* betas, int32_t dtype, int64_t ld,
                           int64_t num_probes, int64_t k, int32_t fn, double fn_param,
                           void* coeffs, void* workspace, int64_t workspace_bytes,
                           void* stream);
int32_t mf_basis_combine(const void* Q, const void* coeffs, const void* scale, int32_t dtype,
                         int64_t n, int64_t ld, int64_t k, void* out, void* stream);

/* funm.funm_lanczos_sym (matfree/funm.py:114-147) WITHOUT a stored basis, for sizes where
 * Q[k][n][ld] does not fit (BASELINE config 5: n = 1e7, k = 30, 256 probes per tile would
 * need 307 GB): the three-term recurrence (reortho = "none") is run twice -- pass 1 yields
 * T, from which y = f(T) e1 per probe; pass 2 repeats the (deterministic, hence bit-identical)
 * recurrence and accumulates out[n][ld] = |v0| * sum_j y_j v_j on the way.  Twice the
 * matvecs, three block vectors of memory.  Columns num_probes..ld-1 of V0 must be valid
 * (e.g. zero) but their output is unspecified. */
int64_t mf_funm_lanczos_workspace_bytes(const mf_operator_t* op, int64_t ld, int64_t k);
int32_t mf_funm_lanczos(const mf_operator_t* op, const void* V0, int64_t ld, int64_t num_probes,
                        int64_t k, int32_t fn, double fn_param, void* out, void* workspace,
                        int64_t workspace_bytes, void* stream);

/* Hutchinson integrands that return one value PER ROW (matfree/stochtrace.py:836-849 diagonal,
 * :868-883 trace_and_diagonal, :886-898 rownorms_squared), accumulated over the probes of a tile:
 *   t[r][c] = A[r][c] * B[r][c]     diagonal: A = probe block, B = operator applied to it;
 *                                   squared row norms: A = B = operator applied to the probes
 *   rowsum[r] (+)= sum_{c < num_probes} t[r][c];   rowsumsq[r] (+)= sum_c t[r][c]^2   (fp64)
 * `accumulate` = 0 overwrites (first tile), 1 adds (later tiles).  mean = rowsum / P and
 * sem = sqrt(rowsumsq / P - mean^2) / sqrt(P) are then what estimator_monte_carlo[_mean_and_sem]
 * returns (stochtrace.py:47-50,83-87).  rowsumsq may be NULL. */
int32_t mf_hutch_rows(const void* A, const void* B, int32_t dtype, int64_t n, int64_t ld,
                      int64_t num_probes, int32_t accumulate, double* rowsum, double* rowsumsq,
                      void* stream);

/* ------------------------------------------------------------------ adjoints (gradients)
 * The custom VJPs of the reference's decompositions -- matfree/decomp.py:184-217,295-348
 * (`_tridiag_adjoint`) and :398-423,480-600 (`_hessenberg_adjoint`), Kraemer et al. 2024 -- are
 * backward recurrences of k more operator products plus the same dots / projections / basis
 * combinations as the forward pass.  They are enqueued from the host layer
 * (matfree_b200/adjoint.py) with the building blocks above (mf_block_dot, mf_reorth_dots,
 * mf_reorth_update, mf_basis_combine, mf_matmat, mf_matmat_rect) and these two:
 *   mf_lincomb    out[n][ld] = sum_t h_scales[t] * coeffs[t][col] * vectors[t][n][ld], t < nterms
 *                 <= 6, summed left to right (coeffs[t] may be NULL = 1; `vectors`, `coeffs`,
 *                 `h_scales` are HOST arrays of device pointers / host doubles): one adjoint step's
 *                 `lambda = -xi + mu x+ + nu x` or `xi = -dx - A lambda + a lambda + b lambda+ -
 *                 b nu x+` (decomp.py:342,349) in one pass;
 *   mf_sddmm_csr  parameter gradient of a CSR operator: the reference accumulates
 *                 vjp(p -> matvec(v, p)) per step (decomp.py:345-346,588-590); for a CSR operator
 *                 that is the outer product cot arg^T on the sparsity pattern:
 *                 out_data[j] (+)= sum_{i<k} C[row_j][i] * G[col_j][i], C / G blocked [n][ld]
 *                 holding the k cotangent / argument vectors of the k steps as columns. */
int32_t mf_lincomb(const void* const* vectors, const void* const* coeffs, const double* h_scales,
                   int32_t nterms, void* out, int32_t dtype, int64_t n, int64_t ld, void* stream);
int32_t mf_sddmm_csr(const int32_t* indptr, const int32_t* indices, int64_t n, const void* C,
                     const void* G, int64_t ld, int64_t k, int32_t accumulate, void* out_data,
                     int32_t dtype, void* stream);

/* ------------------------------------------------------------------ multi-GPU (row sharding)
 * One process per GPU.  A communicator owns one device region per rank -- a control block plus
 * a "heap" the caller places extended Lanczos blocks in -- that every rank maps into its own
 * address space (CUDA IPC), so that kernels of this library reach the other GPUs with plain
 * loads / stores over NVLink:
 *   - every reduction of a sharded decomposition is all-reduced INSIDE the reducing kernel
 *     (the last CTA pushes its fp64 sums to all ranks, flag handshake, sum in rank order:
 *     deterministic and bit-identical on all ranks), replacing the all-reduce XLA inserts after
 *     `linalg.inner` / `vector_norm` on a row-sharded array (matfree/decomp.py:288,290,463,468,471);
 *   - the halo rows a rank's CSR rows reference are stored straight into the neighbours'
 *     extended blocks before every product (the collective-permute before decomp.py:287,460).
 * These are the only entry points that own device memory (mf_comm_create / mf_comm_destroy)
 * or synchronise (mf_comm_status); they are set-up / tear-down calls, not on the per-step path.
 *
 * Set-up: mf_comm_create on every rank; exchange the MF_COMM_HANDLE_BYTES-byte handles of
 * mf_comm_handle between the ranks by any host channel, in rank order; mf_comm_connect.
 * All ranks must issue the same sequence of communicating calls (as with NCCL). */
#define MF_COMM_HANDLE_BYTES 64
typedef struct mf_comm mf_comm_t;
int32_t mf_comm_create(int32_t world, int32_t rank, int64_t heap_bytes, mf_comm_t** comm);
int32_t mf_comm_handle(const mf_comm_t* comm, void* h_handle);
int32_t mf_comm_connect(mf_comm_t* comm, const void* h_handles /* [world][64] */);
void* mf_comm_heap(const mf_comm_t* comm);
int64_t mf_comm_heap_bytes(const mf_comm_t* comm);
/* Synchronises `stream`; MF_ERR_PEER_TIMEOUT if an in-kernel wait gave up (20 s). */
int32_t mf_comm_status(const mf_comm_t* comm, void* stream);
int32_t mf_comm_barrier(const mf_comm_t* comm, void* stream);
/* Tear-down in two phases: every rank unmaps its peers' regions (mf_comm_disconnect), the host
 * layer synchronises the ranks, then every rank frees its own region (mf_comm_destroy). */
int32_t mf_comm_disconnect(mf_comm_t* comm);
int32_t mf_comm_destroy(mf_comm_t* comm);

/* Halo plan of one rank.  Its extended block is `rows_alloc` rows [rows_alloc][ld]: padding,
 * lower halo, the n owned rows starting at `mid_row`, upper halo.  `sends[i]`: `rows` rows
 * starting at row `src_row` of MY extended block go to row `dst_row` of rank `peer`'s extended
 * block (which has `dst_rows_alloc` rows).  `recv_peers`: the ranks that send to me. */
typedef struct mf_halo_send {
  int32_t peer;
  int64_t src_row, rows, dst_row, dst_rows_alloc;
} mf_halo_send_t;
typedef struct mf_halo_plan {
  int64_t rows_alloc, mid_row;
  int32_t num_sends;
  const mf_halo_send_t* sends; /* HOST array */
  int32_t num_recv_peers;
  const int32_t* recv_peers; /* HOST array */
} mf_halo_plan_t;

/* Fill the halo rows of extended block number `block_index` of the array of extended blocks
 * that starts `heap_offset` bytes into every rank's heap.  `barrier_first`: pass 1 when the
 * block's previous content may still be read by a peer's product that no reduction separates
 * from this call. */
int32_t mf_halo_exchange(const mf_comm_t* comm, const mf_halo_plan_t* plan, int64_t heap_offset,
                         int64_t block_index, int64_t ld, int32_t dtype, int32_t barrier_first,
                         void* stream);

/* decomp.tridiag_sym on a ROW-SHARDED CSR operator, the whole k-step loop enqueued by one call
 * (matfree/decomp.py:125-145,426-477 for MF_REORTHO_FULL, :220-292 for MF_REORTHO_NONE; same
 * outputs as mf_lanczos, all scalars identical on every rank).  `op_local`: this rank's CSR
 * rows, n = local rows, column ids relative to its extended block.  V0: the rank's rows of
 * the start block [n][ld] (anywhere).  The extended blocks live in the communicator heap at
 * `heap_offset` (the same on all ranks):
 *   MF_REORTHO_FULL: k extended blocks -- the basis; vector i of the basis is rows
 *                    [mid_row, mid_row + n) of block i;
 *   MF_REORTHO_NONE: 2 extended blocks (ping-pong), plus k more when want_Q != 0.
 * mf_lanczos_sharded_heap_bytes gives the size.  comm may be NULL (or a 1-rank communicator)
 * together with ext != NULL pointing at ordinary device memory of that size. */
int64_t mf_lanczos_sharded_heap_bytes(const mf_halo_plan_t* plan, int64_t ld, int64_t k,
                                      int32_t reortho, int32_t want_Q, int32_t dtype);
int64_t mf_lanczos_sharded_workspace_bytes(const mf_operator_t* op_local, int64_t ld, int64_t k,
                                           int32_t reortho);
int32_t mf_lanczos_sharded(const mf_comm_t* comm, const mf_operator_t* op_local,
                           const mf_halo_plan_t* plan, const void* V0, int64_t ld, int64_t k,
                           int32_t reortho, int32_t want_Q, int64_t heap_offset, void* ext,
                           void* alphas, void* betas, void* init_len, void* residual,
                           void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MATFREE_B200_H_ */
