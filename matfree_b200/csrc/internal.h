// Host-side launchers shared between translation units (no device code here).
#pragma once

#include "common.cuh"

namespace mf {

// Number of CTAs of `kernel` (block size `block`, `smem` dynamic bytes) that are
// co-resident on the device, capped at `want` and at kMaxPartialCtas.  Reducing
// and streaming kernels are launched with exactly this many CTAs (one wave) and
// loop over their work, so work assignment is sequential in memory.
int resident_grid(const void* kernel, int block, size_t smem, int64_t want);

// A reduction request handed to the reducing launchers: the per-CTA partial rows
// plus what the last CTA writes (common.cuh, `Finalize`).
struct Reduce {
  double* partial;  // double[nacc][kMaxPartialCtas * ld]
  Finalize fin;
};
inline int64_t partial_bytes(int64_t ld, int nacc = 1) {
  return (int64_t)nacc * kMaxPartialCtas * ld * 8;
}

// ---- probe_gen.cu
// red (optional): column sums of squares -> red->fin (mode 1: |probe|, 1/|probe|)
// row0 / n_total (blocked layout): the block holds rows [row0, row0 + n) of probes of length
// n_total (a rank's slab of a row-sharded sample array); n_total = 0 means n_total = n.
int32_t launch_probe_gen(void* out, int32_t dtype, int32_t layout, int64_t n, int64_t ld,
                         int64_t p0, int64_t num_probes, uint32_t key0, uint32_t key1,
                         int32_t sampler, int32_t prng_flags, const Reduce* red,
                         cudaStream_t st, int64_t row0 = 0, int64_t n_total = 0);

// ---- blockvec.cu : all operate on blocked vectors X[n][ld]
// sum_r (X[r][c] * sx[c]) * Y[r][c] -> red.fin         (sx may be null)
int32_t launch_dot(const void* X, const void* sx, const void* Y, int32_t dtype, int64_t n,
                   int64_t ld, const Reduce& red, cudaStream_t st);
// Lanczos three-term update (matfree/decomp.py:286-292, lazily normalised):
//   out = (W - a * (Rc * sc)) - bprev * (Rp * sp);  column sums of out^2 -> red.fin
// Rp/bprev/sp may be null (first step).  out may alias Rp.
int32_t launch_lanczos_update(const void* W, const void* Rc, const void* sc, const void* a,
                              const void* Rp, const void* sp, const void* bprev, void* out,
                              int32_t dtype, int64_t n, int64_t ld, const Reduce& red,
                              cudaStream_t st);
// out = X * s (mode 0) or X / s (mode 1), per column; s may be null (copy)
int32_t launch_scale(const void* X, const void* s, void* out, int mode, int32_t dtype,
                     int64_t n, int64_t ld, cudaStream_t st);
// out (+)= (coef .* s1 .* s2)[column] * X   (null scalars = 1; first: out is overwritten)
int32_t launch_axpy_cols(const void* X, const void* coef, const void* s1, const void* s2, void* out,
                         bool first, int32_t dtype, int64_t n, int64_t ld, cudaStream_t st);
// CGS pass, dots: h[j][c] = sum_r Q[j][r][c] * V[r][c], j = 0..nq-1
// (matfree/decomp.py:463,468).  partial: double[4][kMaxPartialCtas*ld]
// dbl_out (optional): the fp64 sums [nq][ld] (row-sharded drivers all-reduce these)
// q_stride: elements between consecutive basis vectors (0: n*ld; larger when the basis lives in
// extended blocks with halo rows); peer: all-reduce the sums over peer memory inside the kernel
int32_t launch_reorth_dots(const void* Q, int64_t nq, const void* V, int32_t dtype, int64_t n,
                           int64_t ld, double* partial, unsigned int* counter, void* h_out,
                           cudaStream_t st, double* dbl_out = nullptr, int64_t partial_rows = 4,
                           int64_t q_stride = 0, const PeerCtx* peer = nullptr,
                           void* alpha_dst = nullptr, void* offdiag = nullptr);
// alpha_dst / offdiag (optional, T[ld] each): the Arnoldi bookkeeping of common.cuh's Finalize --
// h[nq-1] is also written to alpha_dst and offdiag <- (offdiag + h[nq-2]) / 2.
// Accumulator rows the partial buffer of the CGS dots gets: all k sums in one launch when
// that costs at most 8 MB (narrow tiles), otherwise 4 (groups of four basis vectors per launch).
inline int64_t reorth_partial_rows(int64_t ld, int64_t k) {
  const int64_t k4 = (k + 3) / 4 * 4;
  // narrow tiles only (ld <= 64): k4 * ld <= 768 keeps the partial rows under 8 MB
  return (k4 > 4 && k4 * ld <= 768 && ld <= 64) ? k4 : 4;
}
// value[i] = sums[i] (mode 0) or sqrt(sums[i]) (mode 1); inv[i] = 1 / value[i]
int32_t launch_sums_finalize(const double* sums, int64_t count, int mode, void* value, void* inv,
                             int32_t dtype, cudaStream_t st);
// CGS pass, update: V <- V - sum_j Q[j] * h[j]  (decomp.py:464,468); optional
// column sums of the new V^2 -> red->fin (norm fused).
int32_t launch_reorth_update(const void* Q, int64_t nq, const void* h, void* V, int32_t dtype,
                             int64_t n, int64_t ld, const Reduce* red, cudaStream_t st,
                             int64_t q_stride = 0);
// Fused CGS pass for narrow fp32 tiles: V <- V - sum_j Q[j] h[j], then h_out[j] = Q[j] . V of
// the NEW V, in one sweep over the basis (decomp.py:464 + :468).  cgs_fused_supported says
// whether the shape qualifies (fp32, ld == 1, nq <= 104, one partial row per basis vector).
bool cgs_fused_supported(const void* Q, int64_t q_stride, int64_t nq, const void* V, int32_t dtype,
                         int64_t n, int64_t ld, int64_t partial_rows);
int32_t launch_reorth_update_dots(const void* Q, int64_t nq, const void* h, void* V, int64_t n,
                                  int64_t ld, double* partial, unsigned int* counter, void* h_out,
                                  cudaStream_t st, int64_t q_stride = 0,
                                  const PeerCtx* peer = nullptr);
// out[r][c] = scale[c] * sum_j Q[j][r][c] * coeff[j][c]
int32_t launch_basis_combine(const void* Q, const void* coeffs, const void* scale,
                             int32_t dtype, int64_t n, int64_t ld, int64_t k, void* out,
                             cudaStream_t st);
int32_t launch_transpose(const void* src, void* dst, int32_t dtype, int64_t n,
                         int64_t num_probes, int64_t ld, bool to_blocked, cudaStream_t st);
// rowsum[r] (+)= sum_c A[r][c] B[r][c], rowsumsq[r] (+)= sum_c (A B)^2 over the first num_probes
// columns (Hutchinson diagonal / row norms, stochtrace.py:836-849,868-898)
int32_t launch_hutch_rows(const void* A, const void* B, int32_t dtype, int64_t n, int64_t ld,
                          int64_t num_probes, bool accumulate, double* rowsum, double* rowsumsq,
                          cudaStream_t st);
// tiny helpers on [ld] scalar rows
int32_t launch_full_offdiag(void* betas_prev_row, const void* h_row, int32_t dtype, int64_t ld,
                            cudaStream_t st);

// ---- spmm_csr.cu
// W = s * (A @ X) per column (s may be null); if red != null also the column sums
// of (X*s) * W -> red->fin   (the Lanczos alpha, decomp.py:288).
// tickets (optional): two zeroed unsigned ints; chunks of rows are then handed out in
// global row order (re-armed by the kernel itself).
// irregular_scratch (optional, spmm_irregular_scratch_bytes): the route for matrices with very
// long rows (mf_operator_t::csr_max_row_nnz > kSpmmLongRow): rows above that length are cut into
// segments of kSpmmSegNnz non-zeros, one CTA each; no fused dot on that route (red must be null).
constexpr int kSpmmLongRow = 128;
constexpr int kSpmmSegNnz = 512;
int64_t spmm_irregular_scratch_bytes(int64_t nnz, int64_t ld, int32_t dtype);
int32_t launch_spmm_csr(const int32_t* indptr, const int32_t* indices, const void* data,
                        int64_t n, int64_t nnz, int32_t dtype, const void* X, const void* s,
                        void* W, int64_t ld, const Reduce* red, unsigned int* tickets,
                        cudaStream_t st, void* irregular_scratch = nullptr, int64_t bandwidth = 0,
                        int32_t num_diagonals = 0, int64_t line_stride = 0);

// ---- spmm_strip.cu : the band route of launch_spmm_csr (stencil / banded matrices, tiles wide
// enough that a row is at least one warp).  *taken = false: not applicable, nothing launched.
int32_t launch_spmm_strip(const int32_t* indptr, const int32_t* indices, const void* data,
                          int64_t n, int64_t nnz, int32_t dtype, const void* X, const void* s,
                          void* W, int64_t ld, const Reduce* red, unsigned int* progress,
                          cudaStream_t st, bool* taken, int64_t bandwidth = 0);

void spmm_strip_config(int use_strip, int rows, int pfd, int minb);
// ---- spmm_tma.cu : 5-diagonal band matrices, fp32, ld = 256: X rows staged in shared memory by
// TMA bulk copies one chunk ahead.  *taken = false: not applicable / switched off.
int32_t launch_spmm_tma(const int32_t* indptr, const int32_t* indices, const void* data, int64_t n,
                        int64_t nnz, int32_t dtype, const void* X, const void* s, void* W,
                        int64_t ld, const Reduce* red, unsigned int* progress, cudaStream_t st,
                        bool* taken, int64_t bandwidth = 0, int32_t num_diagonals = 0,
                        int64_t line_stride = 0);
void spmm_tma_config(int use_tma);

// ---- gemm_simt.cu
// C[M][ld] = op(A) @ B[K][ld]; A is [M][K] (trans=0, lda>=K) or [K][M] (trans=1, lda>=M)
// colscale (optional, [ld]): C[:, c] is multiplied by colscale[c] in the epilogue
// launch_gemm_blocked: fp64 -> DMMA kernel when the shape qualifies, else the CUDA-core kernel
int32_t launch_gemm_blocked(const void* A, int64_t lda, bool trans, int64_t M, int64_t K,
                            const void* B, const void* colscale, void* C, int64_t ld,
                            int32_t dtype, cudaStream_t st);
int32_t launch_gemm_simt(const void* A, int64_t lda, bool trans, int64_t M, int64_t K,
                         const void* B, const void* colscale, void* C, int64_t ld,
                         int32_t dtype, cudaStream_t st);

// ---- gemm_dmma.cu : fp64 on the FP64 tensor cores (mma.sync m8n8k4 -> DMMA.8x8x4)
// fp64, ld a multiple of 32, lda even, 16-byte aligned pointers
bool dmma_gemm_supported(const void* A, int64_t lda, int64_t M, int64_t K, const void* B,
                         const void* C, int64_t ld, int32_t dtype);
int32_t launch_gemm_dmma(const void* A, int64_t lda, bool trans, int64_t M, int64_t K,
                         const void* B, const void* colscale, void* C, int64_t ld,
                         cudaStream_t st);

// ---- gemm_tcgen05.cu : fp32 via 3xTF32 on the tcgen05 tensor cores
// Can the tensor-core kernel take this problem?  (fp32, ld in {32,64,128,256},
// lda a multiple of 4 elements so TMA strides are multiples of 16 bytes)
bool tc_gemm_supported(int64_t lda, int64_t M, int64_t K, int64_t ld, int32_t dtype);
// planes[0][i] = rna_tf32(src[i]); planes[1][i] = rna_tf32(src[i] - planes[0][i])
int32_t launch_split_tf32(const void* src, void* planes, int64_t count, cudaStream_t st);
// C[batch][M][ld] = colscale .* (op(A) @ B) from TF32 planes Aplanes [2][rows][lda] and
// Bplanes [batch][2][K][ld]; writes C and/or the planes of the result, Csplit [batch][2][M][ld]
// plain 2-D fp32 tensor map (gemm_tcgen05.cu); `map` points at a CUtensorMap
int32_t encode_plain_map_2d(void* map, const void* base, uint64_t dim0, uint64_t dim1,
                            uint64_t stride1_bytes, uint32_t box0, uint32_t box1);
int32_t launch_gemm_tcgen05(const void* Aplanes, int64_t lda, bool trans, int64_t M, int64_t K,
                            const void* Bplanes, int64_t nbatch, const void* colscale, void* C,
                            void* Csplit, int64_t ld, int variant, cudaStream_t st);

// ---- peer.cu : peer-memory communicator (mf_comm_*), halo push, barrier
constexpr int kMaxHaloSends = 16;
const PeerCtx* comm_ctx(const mf_comm* c);  // device descriptor, null for 1 rank / no comm
void* comm_heap(const mf_comm* c);
int32_t launch_halo_exchange(const mf_comm* c, const mf_halo_plan_t* plan, int64_t heap_offset,
                             int64_t block_index, int64_t ld, int32_t dtype, cudaStream_t st);
int32_t launch_peer_barrier(const mf_comm* c, cudaStream_t st);

// ---- tridiag_quad.cu
int32_t launch_tridiag_quad(const void* alphas, const void* betas, const void* init_len,
                            int32_t dtype, int64_t ld, int64_t num_probes, int64_t k,
                            int32_t fn, double fn_param, void* quad, double* nodes,
                            double* weights, void* coeffs, double* work, cudaStream_t st,
                            int product = 0);
int32_t launch_mc_reduce(const void* values, int32_t dtype, int64_t num, double* stats,
                         cudaStream_t st);

}  // namespace mf
