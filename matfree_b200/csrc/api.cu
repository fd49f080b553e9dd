// C-ABI entry points of libmatfree_b200.so and the stream-ordered drivers
// (Lanczos loops, fused estimator) built from the kernels in this directory.
// Nothing in this file synchronises, allocates or frees device memory.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "internal.h"

namespace mf {

std::atomic<long long> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static std::mutex mu;
  static int cache[64];
  static bool have[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 148;
  }
  if (dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(mu);
  if (!have[dev]) {
    int v = 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      v = 148;
    }
    cache[dev] = v;
    have[dev] = true;
  }
  return cache[dev];
}

// ---------------------------------------------------------------- launch timing
namespace {
struct TimingRec {
  int cls;
  cudaEvent_t beg, end;
};
struct TimingState {
  std::mutex mu;
  bool enabled = false;
  std::vector<TimingRec> recs;           // in flight
  std::vector<cudaEvent_t> free_events;  // recycled
} g_timing;
std::atomic<bool> g_timing_on{false};

cudaEvent_t take_event() {
  if (!g_timing.free_events.empty()) {
    cudaEvent_t e = g_timing.free_events.back();
    g_timing.free_events.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

KernelScope::KernelScope(int cls_, cudaStream_t st_) : cls(cls_), st(st_), slot(-1) {
  if (!g_timing_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_timing.mu);
  TimingRec r{cls, take_event(), take_event()};
  cudaEventRecord(r.beg, st);
  g_timing.recs.push_back(r);
  slot = (int)g_timing.recs.size() - 1;
}
KernelScope::~KernelScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_timing.mu);
  if (slot < (int)g_timing.recs.size()) cudaEventRecord(g_timing.recs[slot].end, st);
}

namespace {

// bump allocator over the caller's workspace
struct Arena {
  char* base;
  int64_t size;
  int64_t used = 0;
  bool dry;  // sizing pass: never dereferenced
  Arena(void* b, int64_t s, bool d) : base((char*)b), size(s), dry(d) {}
  void* take(int64_t bytes) {
    const int64_t off = used;
    used = align_up(used + bytes, 256);
    if (dry) return (void*)(uintptr_t)(off + 256);  // non-null dummy
    if (used > size) return nullptr;
    return base + off;
  }
};

[[maybe_unused]] inline int vec_of(int32_t dtype, int64_t ld) {
  const int nv = dtype == MF_F64 ? 2 : 4;
  return ld >= nv ? nv : 1;
}

int32_t validate_op(const mf_operator_t* op) {
  if (op == nullptr) {
    set_error("operator is null");
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (op->dtype != MF_F32 && op->dtype != MF_F64) {
    set_error("operator dtype %d unsupported", op->dtype);
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (op->n <= 0) {
    set_error("operator dimension n=%lld must be positive", (long long)op->n);
    return MF_ERR_INVALID_ARGUMENT;
  }
  switch (op->kind) {
    case MF_OP_DENSE:
      if (!op->values || op->lda < op->n) {
        set_error("dense operator needs values and lda >= n");
        return MF_ERR_INVALID_ARGUMENT;
      }
      return MF_OK;
    case MF_OP_CSR:
      if (!op->values || !op->indptr || !op->indices) {
        set_error("csr operator needs indptr, indices and values");
        return MF_ERR_INVALID_ARGUMENT;
      }
      return MF_OK;
    case MF_OP_GRAM:
      if (!op->values || op->m <= 0 || op->lda < op->n) {
        set_error("gram operator needs values, m > 0 and lda >= n");
        return MF_ERR_INVALID_ARGUMENT;
      }
      return MF_OK;
    default:
      set_error("unknown operator kind %d", op->kind);
      return MF_ERR_INVALID_ARGUMENT;
  }
}

// Scratch an operator application needs (carved from the caller's workspace).
struct OpScratch {
  void* spmm;    // CSR with very long rows: segment list + partial sums of the irregular route
  void* gram;    // SIMT Gram path: Y = A X, m*ld elements
  void* xsplit;  // tcgen05 path: TF32 planes of X, 2*n*ld floats
  void* tsplit;  // tcgen05 Gram path: TF32 planes of A X, 2*m*ld floats
};

// Tuning knobs of the dense path (mf_gemm_config; environment MF_GEMM_VARIANT /
// MF_GEMM_FORCE_SIMT give the initial values).
std::atomic<int> g_gemm_variant{[] {
  const char* e = getenv("MF_GEMM_VARIANT");
  return e ? atoi(e) : 0;
}()};
std::atomic<int> g_gemm_tc{getenv("MF_GEMM_FORCE_SIMT") == nullptr ? 1 : 0};
int gemm_variant() { return g_gemm_variant.load(std::memory_order_relaxed); }

// The fp32 dense / Gram operators run on the tcgen05 tensor cores (3xTF32) when the
// operator carries its TF32 planes (mf_operator_split) and the shape qualifies;
// otherwise (fp64, tiny ld, unaligned lda) on the CUDA-core kernel.
bool use_tc(const mf_operator_t* op, int64_t ld) {
  if (op->kind != MF_OP_DENSE && op->kind != MF_OP_GRAM) return false;
  if (op->split_planes == nullptr) return false;
  if (g_gemm_tc.load(std::memory_order_relaxed) == 0) return false;
  const int64_t rows = op->kind == MF_OP_GRAM ? op->m : op->n;
  return tc_gemm_supported(op->lda, rows, op->n, ld, op->dtype);
}

bool csr_irregular(const mf_operator_t* op) {
  return op->kind == MF_OP_CSR && op->csr_max_row_nnz > kSpmmLongRow;
}

void carve_op_scratch(Arena& a, const mf_operator_t* op, int64_t ld, OpScratch* s) {
  memset(s, 0, sizeof(*s));
  const int64_t es = (int64_t)dtype_size(op->dtype);
  if (csr_irregular(op)) {
    s->spmm = a.take(spmm_irregular_scratch_bytes(op->nnz, ld, op->dtype));
    return;
  }
  if (use_tc(op, ld)) {
    s->xsplit = a.take(2 * op->n * ld * 4);
    if (op->kind == MF_OP_GRAM) s->tsplit = a.take(2 * op->m * ld * 4);
  } else if (op->kind == MF_OP_GRAM) {
    s->gram = a.take(op->m * ld * es);
  }
}

// W = s * (A @ X); if red != null and the operator can fuse it, also the column
// sums of (X*s) .* W -> red->fin.  *fused tells the caller whether it happened.
int32_t apply_op(const mf_operator_t* op, const void* X, const void* s, void* W, int64_t ld,
                 const OpScratch& scr, const Reduce* red, unsigned int* tickets, bool* fused,
                 cudaStream_t st) {
  *fused = false;
  if (op->kind == MF_OP_CSR) {
    if (csr_irregular(op)) {
      if (scr.spmm == nullptr) {
        set_error("csr operator with long rows: workspace for the segment list is missing");
        return MF_ERR_WORKSPACE;
      }
      // the caller forms the alpha dot with a separate kernel on this route
      return launch_spmm_csr(op->indptr, op->indices, op->values, op->n, op->nnz, op->dtype, X, s,
                             W, ld, nullptr, nullptr, st, scr.spmm);
    }
    MF_TRY(launch_spmm_csr(op->indptr, op->indices, op->values, op->n, op->nnz, op->dtype, X, s, W,
                           ld, red, tickets, st, nullptr, op->csr_bandwidth, op->csr_num_diagonals,
                           op->csr_line_stride));
    *fused = red != nullptr;
    return MF_OK;
  }
  if (use_tc(op, ld)) {
    if (scr.xsplit == nullptr || (op->kind == MF_OP_GRAM && scr.tsplit == nullptr)) {
      set_error("dense/gram operator: workspace for the TF32 planes is missing");
      return MF_ERR_WORKSPACE;
    }
    const int variant = gemm_variant();
    MF_TRY(launch_split_tf32(X, scr.xsplit, op->n * ld, st));
    if (op->kind == MF_OP_DENSE)
      return launch_gemm_tcgen05(op->split_planes, op->lda, false, op->n, op->n, scr.xsplit, 1, s,
                                 W, nullptr, ld, variant, st);
    // Gram: T = A X lands directly as TF32 planes (epilogue split), then W = A^T T
    MF_TRY(launch_gemm_tcgen05(op->split_planes, op->lda, false, op->m, op->n, scr.xsplit, 1,
                               nullptr, nullptr, scr.tsplit, ld, variant, st));
    return launch_gemm_tcgen05(op->split_planes, op->lda, true, op->n, op->m, scr.tsplit, 1, s, W,
                               nullptr, ld, variant, st);
  }
  // fp64 (DMMA when the shape qualifies) and everything the tcgen05 kernel does not take;
  // mf_gemm_config(.., use_tensor_cores = 0) forces the CUDA-core kernel for cross-checks
  const auto gemm = g_gemm_tc.load(std::memory_order_relaxed) ? launch_gemm_blocked : launch_gemm_simt;
  if (op->kind == MF_OP_DENSE)
    return gemm(op->values, op->lda, false, op->n, op->n, X, s, W, ld, op->dtype, st);
  if (op->kind == MF_OP_GRAM) {
    if (scr.gram == nullptr) {
      set_error("gram operator needs a scratch block of m*ld elements");
      return MF_ERR_WORKSPACE;
    }
    MF_TRY(gemm(op->values, op->lda, false, op->m, op->n, X, nullptr, scr.gram, ld, op->dtype, st));
    return gemm(op->values, op->lda, true, op->n, op->m, scr.gram, s, W, ld, op->dtype, st);
  }
  return MF_ERR_INVALID_ARGUMENT;
}

struct LanczosBufs {
  void *R0, *R1, *W, *V;       // block vectors
  void *inv;                   // [k+1][ld] 1/len, 1/beta_j
  void *h, *h2;                // [k][ld] CGS coefficients
  double* partial;             // per-CTA partial rows of the reductions
  int64_t partial_rows;        // accumulator rows `partial` holds
  unsigned int* counter;       // ticket counter of the fused finalize (zeroed per call)
  OpScratch scr;               // operator scratch
};

int32_t carve_lanczos(Arena& a, const mf_operator_t* op, int64_t ld, int64_t k, int32_t reortho,
                      LanczosBufs* b) {
  const int64_t es = (int64_t)dtype_size(op->dtype);
  const int64_t blk = op->n * ld * es;
  memset(b, 0, sizeof(*b));
  b->counter = (unsigned int*)a.take(256);
  if (reortho == MF_REORTHO_NONE) {
    b->R0 = a.take(blk);
    b->R1 = a.take(blk);
    b->W = a.take(blk);
    b->inv = a.take((k + 1) * ld * es);
    b->partial = (double*)a.take(partial_bytes(ld, 1));
  } else {
    b->V = a.take(blk);
    b->h = a.take((k + 1) * ld * es);
    b->h2 = a.take((k + 1) * ld * es);
    b->partial_rows = reorth_partial_rows(ld, k);
    b->partial = (double*)a.take(partial_bytes(ld, (int)b->partial_rows));
  }
  carve_op_scratch(a, op, ld, &b->scr);
  if (!a.dry && a.used > a.size) {
    set_error("workspace too small: need %lld bytes, have %lld", (long long)a.used,
              (long long)a.size);
    return MF_ERR_WORKSPACE;
  }
  return MF_OK;
}

int32_t zero_counter(const LanczosBufs& b, cudaStream_t st) {
  if (cudaMemsetAsync(b.counter, 0, 256, st) != cudaSuccess) {
    set_error("counter memset failed");
    return MF_ERR_CUDA;
  }
  return MF_OK;
}

inline void* row(void* base, int64_t j, int64_t ld, int32_t dtype) {
  return (char*)base + j * ld * (int64_t)dtype_size(dtype);
}

// Three-term Lanczos on a probe block (matfree/decomp.py:220-292), lazily
// normalised: the un-normalised residual block r_j and 1/beta_{j-1} are kept
// and v_j = r_j / beta_{j-1} is formed on the fly by the consumers.
//   v0_owned: V0 may be overwritten (fused estimator) -> one buffer less.
//   have_len: init_len and inv[0] were already written (by the probe generator).
// What the second pass of the two-pass f(A)v accumulates while the recurrence is re-run:
// out += (coeffs[j] * len) * v_j  (matfree/funm.py:145 without the stored basis).
struct BasisAccum {
  const void* coeffs;  // [k][ld]  y = f(T) e1 per probe
  void* out;           // [n][ld]
};

int32_t lanczos_none(const mf_operator_t* op, void* V0, bool v0_owned, bool have_len, int64_t ld,
                     int64_t k, void* alphas, void* betas, void* init_len, void* Q,
                     void* residual, const LanczosBufs& b, cudaStream_t st,
                     const BasisAccum* accum = nullptr) {
  const int32_t dt = op->dtype;
  const int64_t n = op->n;
  const int64_t blk = n * ld * (int64_t)dtype_size(dt);
  if (!have_len) {
    const Reduce red{b.partial, Finalize{b.counter, 1, init_len, row(b.inv, 0, ld, dt), nullptr}};
    MF_TRY(launch_dot(V0, nullptr, V0, dt, n, ld, red, st));
  }
  void* Rc = V0;
  void* Rp = nullptr;
  for (int64_t j = 0; j < k; ++j) {
    void* sc = row(b.inv, j, ld, dt);
    const void* X = Rc;
    const void* sx = sc;
    if (Q != nullptr) {  // materialise v_j (decomp.py:279) and read it back un-scaled
      void* Qj = (char*)Q + j * blk;
      MF_TRY(launch_scale(Rc, sc, Qj, 0, dt, n, ld, st));
      X = Qj;
      sx = nullptr;
    }
    if (accum != nullptr)  // v_j = Rc * sc
      MF_TRY(launch_axpy_cols(Rc, row(const_cast<void*>(accum->coeffs), j, ld, dt), sc, init_len,
                              accum->out, j == 0, dt, n, ld, st));
    void* aj = row(alphas, j, ld, dt);
    const Reduce red_a{b.partial, Finalize{b.counter, 0, aj, nullptr, nullptr}};
    bool fused = false;
    MF_TRY(apply_op(op, X, sx, b.W, ld, b.scr, &red_a, b.counter + 8, &fused, st));
    if (!fused) MF_TRY(launch_dot(X, sx, b.W, dt, n, ld, red_a, st));
    // pick the output buffer: alias Rp when we own it
    void* out;
    if (Rp == nullptr) out = b.R0;
    else if (Rp == V0 && !v0_owned) out = b.R1;
    else out = Rp;
    const Reduce red_b{b.partial, Finalize{b.counter, 1, row(betas, j, ld, dt),
                                           row(b.inv, j + 1, ld, dt), nullptr}};
    MF_TRY(launch_lanczos_update(b.W, Rc, sc, aj, Rp, j > 0 ? row(b.inv, j - 1, ld, dt) : nullptr,
                                 j > 0 ? row(betas, j - 1, ld, dt) : nullptr, out, dt, n, ld,
                                 red_b, st));
    Rp = Rc;
    Rc = out;
  }
  if (residual != nullptr) {
    // b_{k-1} * v_k is the un-normalised r_k itself (decomp.py:167)
    if (k > 0) {
      if (cudaMemcpyAsync(residual, Rc, blk, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("residual copy failed");
        return MF_ERR_CUDA;
      }
    }
  }
  return MF_OK;
}

// Arnoldi with classical Gram-Schmidt applied twice (matfree/decomp.py:426-477)
// and T = (H + H^T)/2 (decomp.py:133-135), on a probe block.
// Hess (optional): the public Arnoldi factorisation (decomp.hessenberg, decomp.py:351-477) on
// the same loop: H[r][i][ld] receives column i = the first-pass coefficients h (+ the norm below
// the diagonal); second_pass = false is the reference's reortho="none" (one CGS pass).
struct HessOut {
  void* H;           // [k][k][ld], zeroed by the caller
  bool second_pass;
};

int32_t lanczos_full(const mf_operator_t* op, const void* V0, bool have_len, int64_t ld,
                     int64_t k, void* alphas, void* betas, void* init_len, void* Q,
                     void* residual, const LanczosBufs& b, cudaStream_t st,
                     const HessOut* hess = nullptr) {
  const int32_t dt = op->dtype;
  const int64_t n = op->n;
  const int64_t es = (int64_t)dtype_size(dt);
  const int64_t blk = n * ld * es;
  if (!have_len) {
    const Reduce red{b.partial, Finalize{b.counter, 1, init_len, nullptr, nullptr}};
    MF_TRY(launch_dot(V0, nullptr, V0, dt, n, ld, red, st));
  }
  const void* length = init_len;
  for (int64_t i = 0; i < k; ++i) {
    void* Qi = (char*)Q + i * blk;
    MF_TRY(launch_scale(i == 0 ? V0 : b.V, length, Qi, 1, dt, n, ld, st));  // :456-457
    bool fused = false;
    MF_TRY(apply_op(op, Qi, nullptr, b.V, ld, b.scr, nullptr, b.counter + 8, &fused, st));  // :460
    // :463; the same launch files alpha_i = h_i and, for T = (H + H^T)/2 (:133-135), folds h_{i-1}
    // into the previous off-diagonal
    MF_TRY(launch_reorth_dots(Q, i + 1, b.V, dt, n, ld, b.partial, b.counter, b.h, st, nullptr,
                              b.partial_rows, 0, nullptr, row(alphas, i, ld, dt),
                              hess == nullptr && i > 0 ? row(betas, i - 1, ld, dt) : nullptr));
    if (hess != nullptr) {
      // H[0..i][i] = h (decomp.py:463,474-475)
      if (cudaMemcpy2DAsync((char*)hess->H + i * ld * es, (size_t)(k * ld * es), b.h,
                            (size_t)(ld * es), (size_t)(ld * es), (size_t)(i + 1),
                            cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("hessenberg: copy of column %lld failed", (long long)i);
        return MF_ERR_CUDA;
      }
      if (!hess->second_pass) {
        const Reduce red_1{b.partial,
                           Finalize{b.counter, 1, row(betas, i, ld, dt), nullptr, nullptr}};
        MF_TRY(launch_reorth_update(Q, i + 1, b.h, b.V, dt, n, ld, &red_1, st));  // :464,471
        length = row(betas, i, ld, dt);
        if (i + 1 < k &&
            cudaMemcpyAsync((char*)hess->H + ((i + 1) * k + i) * ld * es, length, ld * es,
                            cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
          set_error("hessenberg: copy of the subdiagonal failed");
          return MF_ERR_CUDA;
        }
        continue;
      }
    }
    if (cgs_fused_supported(Q, 0, i + 1, b.V, dt, n, ld, b.partial_rows)) {
      // :464 and the dots of :468 in one sweep over the basis
      MF_TRY(launch_reorth_update_dots(Q, i + 1, b.h, b.V, n, ld, b.partial, b.counter, b.h2, st));
    } else {
      MF_TRY(launch_reorth_update(Q, i + 1, b.h, b.V, dt, n, ld, nullptr, st));  // :464
      MF_TRY(launch_reorth_dots(Q, i + 1, b.V, dt, n, ld, b.partial, b.counter, b.h2, st, nullptr,
                                b.partial_rows));  // :468
    }
    const Reduce red_n{b.partial, Finalize{b.counter, 1, row(betas, i, ld, dt), nullptr, nullptr}};
    MF_TRY(launch_reorth_update(Q, i + 1, b.h2, b.V, dt, n, ld, &red_n, st));  // :468,471
    length = row(betas, i, ld, dt);
    if (hess != nullptr && i + 1 < k &&
        cudaMemcpyAsync((char*)hess->H + ((i + 1) * k + i) * ld * es, length, ld * es,
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("hessenberg: copy of the subdiagonal failed");
      return MF_ERR_CUDA;
    }
  }
  if (residual != nullptr) {
    const void* src = k > 0 ? b.V : V0;
    if (cudaMemcpyAsync(residual, src, blk, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("residual copy failed");
      return MF_ERR_CUDA;
    }
  }
  return MF_OK;
}

// ---------------------------------------------------------------- row-sharded drivers
// The same two recurrences on ONE RANK's rows of a row-sharded CSR operator.  Differences to
// the single-GPU drivers above: (1) the vector the operator is applied to lives in an
// "extended block" (padding | lower halo | owned rows | upper halo) in the communicator's
// peer-mapped heap, and its halo rows are filled by the neighbours' stores before every product
// (launch_halo_exchange); (2) every reduction carries the communicator's descriptor, so its last
// CTA all-reduces the fp64 sums over peer memory before writing the scalars (common.cuh) --
// all ranks hold bit-identical alphas / betas / CGS coefficients; (3) vectors are normalised by
// a true division as in the reference (decomp.py:227,291,456-457): v_j must exist in memory
// anyway because the neighbours read its boundary rows.
struct Shard {
  const mf_comm* comm;
  const PeerCtx* peer;
  const mf_halo_plan_t* plan;
  int64_t heap_offset;
  unsigned char* ext;  // local address of the extended blocks
};

int32_t lanczos_full_sharded(const Shard& sh, const mf_operator_t* op, const void* V0, int64_t ld,
                             int64_t k, void* alphas, void* betas, void* init_len, void* residual,
                             const LanczosBufs& b, cudaStream_t st) {
  const int32_t dt = op->dtype;
  const int64_t n = op->n;  // local rows
  const int64_t es = (int64_t)dtype_size(dt);
  const int64_t blk = n * ld * es;
  const int64_t ext_blk = sh.plan->rows_alloc * ld * es;
  const int64_t q_stride = sh.plan->rows_alloc * ld;
  unsigned char* Qmid = sh.ext + sh.plan->mid_row * ld * es;  // basis vector 0, owned rows
  {
    const Reduce red{b.partial, Finalize{b.counter, 1, init_len, nullptr, nullptr, sh.peer}};
    MF_TRY(launch_dot(V0, nullptr, V0, dt, n, ld, red, st));
  }
  const void* length = init_len;
  for (int64_t i = 0; i < k; ++i) {
    void* Qi = Qmid + i * ext_blk;
    MF_TRY(launch_scale(i == 0 ? V0 : b.V, length, Qi, 1, dt, n, ld, st));  // :456-457
    MF_TRY(launch_halo_exchange(sh.comm, sh.plan, sh.heap_offset, i, ld, dt, st));
    bool fused = false;
    MF_TRY(apply_op(op, sh.ext + i * ext_blk, nullptr, b.V, ld, b.scr, nullptr, b.counter + 8,
                    &fused, st));  // :460
    MF_TRY(launch_reorth_dots(Qmid, i + 1, b.V, dt, n, ld, b.partial, b.counter, b.h, st, nullptr,
                              b.partial_rows, q_stride, sh.peer, row(alphas, i, ld, dt),
                              i > 0 ? row(betas, i - 1, ld, dt) : nullptr));  // :463, :133-135
    if (cgs_fused_supported(Qmid, q_stride, i + 1, b.V, dt, n, ld, b.partial_rows)) {
      MF_TRY(launch_reorth_update_dots(Qmid, i + 1, b.h, b.V, n, ld, b.partial, b.counter, b.h2,
                                       st, q_stride, sh.peer));  // :464 + dots of :468
    } else {
      MF_TRY(launch_reorth_update(Qmid, i + 1, b.h, b.V, dt, n, ld, nullptr, st, q_stride));  // :464
      MF_TRY(launch_reorth_dots(Qmid, i + 1, b.V, dt, n, ld, b.partial, b.counter, b.h2, st,
                                nullptr, b.partial_rows, q_stride, sh.peer));  // :468
    }
    const Reduce red_n{b.partial,
                       Finalize{b.counter, 1, row(betas, i, ld, dt), nullptr, nullptr, sh.peer}};
    MF_TRY(launch_reorth_update(Qmid, i + 1, b.h2, b.V, dt, n, ld, &red_n, st, q_stride));  // :468,471
    length = row(betas, i, ld, dt);
  }
  if (residual != nullptr) {
    const void* src = k > 0 ? b.V : V0;
    if (cudaMemcpyAsync(residual, src, blk, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("residual copy failed");
      return MF_ERR_CUDA;
    }
  }
  return MF_OK;
}

int32_t lanczos_none_sharded(const Shard& sh, const mf_operator_t* op, const void* V0, int64_t ld,
                             int64_t k, bool want_Q, void* alphas, void* betas, void* init_len,
                             void* residual, const LanczosBufs& b, cudaStream_t st) {
  const int32_t dt = op->dtype;
  const int64_t n = op->n;
  const int64_t es = (int64_t)dtype_size(dt);
  const int64_t blk = n * ld * es;
  const int64_t ext_blk = sh.plan->rows_alloc * ld * es;
  const int64_t mid_off = sh.plan->mid_row * ld * es;
  {
    const Reduce red{b.partial, Finalize{b.counter, 1, init_len, nullptr, nullptr, sh.peer}};
    MF_TRY(launch_dot(V0, nullptr, V0, dt, n, ld, red, st));
  }
  const void* cur = V0;
  const void* length = init_len;
  const void* prev_mid = nullptr;
  for (int64_t j = 0; j < k; ++j) {
    const int64_t blkno = want_Q ? j : (j & 1);  // ping-pong unless the basis is kept
    unsigned char* xe = sh.ext + blkno * ext_blk;
    void* mid = xe + mid_off;
    MF_TRY(launch_scale(cur, length, mid, 1, dt, n, ld, st));  // v_j = r / b (decomp.py:227,291)
    MF_TRY(launch_halo_exchange(sh.comm, sh.plan, sh.heap_offset, blkno, ld, dt, st));
    bool fused = false;
    MF_TRY(apply_op(op, xe, nullptr, b.W, ld, b.scr, nullptr, b.counter + 8, &fused, st));  // :287
    void* aj = row(alphas, j, ld, dt);
    const Reduce red_a{b.partial, Finalize{b.counter, 0, aj, nullptr, nullptr, sh.peer}};
    MF_TRY(launch_dot(mid, nullptr, b.W, dt, n, ld, red_a, st));  // :288
    const Reduce red_b{b.partial,
                       Finalize{b.counter, 1, row(betas, j, ld, dt), nullptr, nullptr, sh.peer}};
    MF_TRY(launch_lanczos_update(b.W, mid, nullptr, aj, prev_mid, nullptr,
                                 j > 0 ? row(betas, j - 1, ld, dt) : nullptr, b.R0, dt, n, ld,
                                 red_b, st));  // :289-290
    cur = b.R0;
    length = row(betas, j, ld, dt);
    prev_mid = mid;
  }
  if (residual != nullptr) {
    if (cudaMemcpyAsync(residual, k > 0 ? b.R0 : V0, blk, cudaMemcpyDeviceToDevice, st) !=
        cudaSuccess) {
      set_error("residual copy failed");
      return MF_ERR_CUDA;
    }
  }
  return MF_OK;
}

int64_t sharded_blocks(int64_t k, int32_t reortho, bool want_Q) {
  if (reortho == MF_REORTHO_FULL || want_Q) return k > 0 ? k : 1;
  return 2;
}

}  // namespace
}  // namespace mf

using namespace mf;

extern "C" {

const char* mf_last_error(void) { return g_err; }
int32_t mf_abi_version(void) { return MF_ABI_VERSION; }
int64_t mf_launch_count(void) { return (int64_t)g_launches.load(); }

int32_t mf_timing_enable(int32_t on) {
  g_timing_on.store(on != 0);
  return MF_OK;
}

int32_t mf_timing_collect(double* h_ms, int64_t* h_launches) {
  std::lock_guard<std::mutex> lk(g_timing.mu);
  for (auto& r : g_timing.recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.end) == cudaSuccess &&
        cudaEventElapsedTime(&ms, r.beg, r.end) == cudaSuccess) {
      if (h_ms) h_ms[r.cls] += (double)ms;
      if (h_launches) h_launches[r.cls] += 1;
    } else {
      cudaGetLastError();
    }
    g_timing.free_events.push_back(r.beg);
    g_timing.free_events.push_back(r.end);
  }
  g_timing.recs.clear();
  return MF_OK;
}

int32_t mf_probe_gen(void* out, int32_t dtype, int32_t layout, int64_t n, int64_t ld,
                     int64_t p0, int64_t num_probes, uint32_t key0, uint32_t key1,
                     int32_t sampler, int32_t prng_flags, double* sqnorm_out, void* stream) {
  if (out == nullptr || n < 0 || num_probes < 0 || p0 < 0) {
    set_error("probe_gen: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (dtype != MF_F32 && dtype != MF_F64) {
    set_error("probe_gen: dtype %d unsupported", dtype);
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (sampler != MF_SAMPLER_SIGNS && sampler != MF_SAMPLER_NORMAL) {
    set_error("probe_gen: sampler %d unsupported", sampler);
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (sqnorm_out != nullptr) {
    set_error("probe_gen: sqnorm_out requires the fused estimator (use mf_lanczos init_len)");
    return MF_ERR_UNSUPPORTED;
  }
  return launch_probe_gen(out, dtype, layout, n, ld, p0, num_probes, key0, key1, sampler,
                          prng_flags, nullptr, (cudaStream_t)stream);
}

int32_t mf_probe_gen_rows(void* out, int32_t dtype, int64_t n_total, int64_t row0, int64_t rows,
                          int64_t ld, int64_t p0, int64_t num_probes, uint32_t key0,
                          uint32_t key1, int32_t sampler, int32_t prng_flags, void* stream) {
  if (out == nullptr || rows < 0 || row0 < 0 || row0 + rows > n_total || num_probes < 0 || p0 < 0) {
    set_error("probe_gen_rows: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  if ((dtype != MF_F32 && dtype != MF_F64) ||
      (sampler != MF_SAMPLER_SIGNS && sampler != MF_SAMPLER_NORMAL)) {
    set_error("probe_gen_rows: dtype %d / sampler %d unsupported", dtype, sampler);
    return MF_ERR_INVALID_ARGUMENT;
  }
  return launch_probe_gen(out, dtype, MF_LAYOUT_BLOCKED, rows, ld, p0, num_probes, key0, key1,
                          sampler, prng_flags, nullptr, (cudaStream_t)stream, row0, n_total);
}

int32_t mf_gemm_config(int32_t variant, int32_t use_tensor_cores) {
  if (variant < 0 || variant > 2) {
    set_error("gemm_config: variant must be 0, 1 or 2");
    return MF_ERR_INVALID_ARGUMENT;
  }
  g_gemm_variant.store(variant);
  g_gemm_tc.store(use_tensor_cores != 0);
  return MF_OK;
}

int32_t mf_spmm_config(int32_t use_band_kernel, int32_t rows_per_chunk, int32_t prefetch_rows,
                       int32_t min_ctas_per_sm) {
  // 0: row-group gather kernel; 1: band kernel (register window); 2 (default): band kernel with
  // the X rows staged in shared memory by TMA
  // 3: as 2, and every 7-diagonal matrix takes the TMA kernel (by default only 3-D stencils in the
  //    blocked row order); 4: as 3, and the column-walk kernel whatever the problem size (tests);
  // 5: as 3 without the column walk (the chunked TMA kernel on 2-D stencils: A/B runs, tests)
  if (use_band_kernel >= 0) spmm_tma_config(use_band_kernel >= 2 ? use_band_kernel - 1 : 0);
  spmm_strip_config(use_band_kernel < 0 ? -1 : (use_band_kernel == 1 ? 1 : 0), rows_per_chunk,
                    prefetch_rows, min_ctas_per_sm);
  return MF_OK;
}

int64_t mf_operator_split_bytes(const mf_operator_t* op) {
  if (validate_op(op) != MF_OK) return -1;
  if (op->dtype != MF_F32 || (op->kind != MF_OP_DENSE && op->kind != MF_OP_GRAM)) return 0;
  const int64_t rows = op->kind == MF_OP_GRAM ? op->m : op->n;
  return 2 * rows * op->lda * 4;
}

int32_t mf_operator_split(const mf_operator_t* op, void* planes, void* stream) {
  MF_TRY(validate_op(op));
  if (op->dtype != MF_F32 || (op->kind != MF_OP_DENSE && op->kind != MF_OP_GRAM)) {
    set_error("operator_split: only fp32 dense / gram operators have TF32 planes");
    return MF_ERR_UNSUPPORTED;
  }
  if (planes == nullptr) {
    set_error("operator_split: planes is null");
    return MF_ERR_INVALID_ARGUMENT;
  }
  const int64_t rows = op->kind == MF_OP_GRAM ? op->m : op->n;
  return launch_split_tf32(op->values, planes, rows * op->lda, (cudaStream_t)stream);
}

int64_t mf_matmat_workspace_bytes(const mf_operator_t* op, int64_t ld) {
  if (validate_op(op) != MF_OK || !valid_ld(ld)) return -1;
  Arena a(nullptr, 0, true);
  OpScratch scr;
  carve_op_scratch(a, op, ld, &scr);
  return a.used + 256;
}

int32_t mf_matmat(const mf_operator_t* op, const void* X, void* W, int64_t ld, void* workspace,
                  int64_t workspace_bytes, void* stream) {
  MF_TRY(validate_op(op));
  if (!valid_ld(ld) || X == nullptr || W == nullptr) {
    set_error("matmat: ld must be a power of two <= 256 and X, W non-null");
    return MF_ERR_INVALID_ARGUMENT;
  }
  Arena a(workspace, workspace_bytes, false);
  OpScratch scr;
  carve_op_scratch(a, op, ld, &scr);
  if (a.used > a.size) {
    set_error("workspace too small: need %lld bytes, have %lld", (long long)a.used,
              (long long)a.size);
    return MF_ERR_WORKSPACE;
  }
  bool fused;
  return apply_op(op, X, nullptr, W, ld, scr, nullptr, nullptr, &fused, (cudaStream_t)stream);
}

int32_t mf_matmat_dense(const void* A, const void* A_planes, int64_t n, int64_t lda,
                        int32_t dtype, const void* X, void* W, int64_t ld, void* workspace,
                        int64_t workspace_bytes, void* stream) {
  mf_operator_t op{};
  op.kind = MF_OP_DENSE; op.dtype = dtype; op.n = n; op.values = A; op.lda = lda;
  op.split_planes = A_planes;
  return mf_matmat(&op, X, W, ld, workspace, workspace_bytes, stream);
}

int32_t mf_matmat_csr(const int32_t* indptr, const int32_t* indices, const void* data,
                      int64_t n, int64_t nnz, int32_t dtype, const void* X, void* W,
                      int64_t ld, void* stream) {
  mf_operator_t op{};
  op.kind = MF_OP_CSR; op.dtype = dtype; op.n = n; op.nnz = nnz; op.values = data;
  op.indptr = indptr; op.indices = indices;
  return mf_matmat(&op, X, W, ld, nullptr, 0, stream);
}

int32_t mf_matmat_gram(const void* A, const void* A_planes, int64_t m, int64_t n, int64_t lda,
                       int32_t dtype, const void* X, void* W, int64_t ld, void* workspace,
                       int64_t workspace_bytes, void* stream) {
  mf_operator_t op{};
  op.kind = MF_OP_GRAM; op.dtype = dtype; op.n = n; op.m = m; op.values = A; op.lda = lda;
  op.split_planes = A_planes;
  return mf_matmat(&op, X, W, ld, workspace, workspace_bytes, stream);
}

int32_t mf_to_blocked(const void* src_pn, void* dst_blocked, int32_t dtype, int64_t n,
                      int64_t num_probes, int64_t ld, void* stream) {
  if (!valid_ld(ld) || num_probes > ld) {
    set_error("to_blocked: need num_probes <= ld, ld a power of two <= 256");
    return MF_ERR_INVALID_ARGUMENT;
  }
  return launch_transpose(src_pn, dst_blocked, dtype, n, num_probes, ld, true,
                          (cudaStream_t)stream);
}

int32_t mf_from_blocked(const void* src_blocked, void* dst_pn, int32_t dtype, int64_t n,
                        int64_t num_probes, int64_t ld, void* stream) {
  if (!valid_ld(ld) || num_probes > ld) {
    set_error("from_blocked: need num_probes <= ld, ld a power of two <= 256");
    return MF_ERR_INVALID_ARGUMENT;
  }
  return launch_transpose(src_blocked, dst_pn, dtype, n, num_probes, ld, false,
                          (cudaStream_t)stream);
}

static int32_t check_lanczos_args(const mf_operator_t* op, int64_t ld, int64_t k,
                                  int32_t reortho) {
  MF_TRY(validate_op(op));
  if (!valid_ld(ld)) {
    set_error("ld=%lld must be a power of two in [1, 256]", (long long)ld);
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (k < 0 || k > op->n) {
    // same wording as matfree/decomp.py:753-756
    set_error("Parameter 'num_matvecs'=%lld exceeds the acceptable range. "
              "Expected: 0 <= num_matvecs <= %lld.", (long long)k, (long long)op->n);
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (reortho != MF_REORTHO_NONE && reortho != MF_REORTHO_FULL) {
    set_error("reortho=%d unsupported", reortho);
    return MF_ERR_INVALID_ARGUMENT;
  }
  return MF_OK;
}

int64_t mf_lanczos_workspace_bytes(const mf_operator_t* op, int64_t ld, int64_t k,
                                   int32_t reortho, int32_t want_Q) {
  (void)want_Q;
  if (check_lanczos_args(op, ld, k, reortho) != MF_OK) return -1;
  Arena a(nullptr, 0, true);
  LanczosBufs b;
  carve_lanczos(a, op, ld, k, reortho, &b);
  return a.used + 256;
}

int32_t mf_lanczos(const mf_operator_t* op, const void* V0, int64_t ld, int64_t k,
                   int32_t reortho, void* alphas, void* betas, void* init_len, void* Q,
                   void* residual, void* workspace, int64_t workspace_bytes, void* stream) {
  MF_TRY(check_lanczos_args(op, ld, k, reortho));
  if (V0 == nullptr || init_len == nullptr || (k > 0 && (alphas == nullptr || betas == nullptr))) {
    set_error("lanczos: V0, init_len, alphas, betas must be non-null");
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (reortho == MF_REORTHO_FULL && Q == nullptr && k > 0) {
    set_error("lanczos: reortho=full needs the basis buffer Q[k][n][ld]");
    return MF_ERR_INVALID_ARGUMENT;
  }
  Arena a(workspace, workspace_bytes, false);
  LanczosBufs b;
  MF_TRY(carve_lanczos(a, op, ld, k, reortho, &b));
  cudaStream_t st = (cudaStream_t)stream;
  MF_TRY(zero_counter(b, st));
  if (reortho == MF_REORTHO_NONE)
    return lanczos_none(op, (void*)V0, false, false, ld, k, alphas, betas, init_len, Q, residual,
                        b, st);
  return lanczos_full(op, V0, false, ld, k, alphas, betas, init_len, Q, residual, b, st);
}

static int32_t check_sharded_args(const mf_operator_t* op, const mf_halo_plan_t* plan, int64_t ld,
                                  int64_t k, int32_t reortho) {
  MF_TRY(validate_op(op));
  if (op->kind != MF_OP_CSR) {
    set_error("lanczos_sharded: only CSR operators are row-sharded");
    return MF_ERR_UNSUPPORTED;
  }
  if (!valid_ld(ld) || k < 0 ||
      (reortho != MF_REORTHO_NONE && reortho != MF_REORTHO_FULL)) {
    set_error("lanczos_sharded: bad ld / k / reortho");
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (plan != nullptr && (plan->mid_row < 0 || plan->mid_row + op->n > plan->rows_alloc)) {
    set_error("lanczos_sharded: the halo plan does not cover the %lld local rows",
              (long long)op->n);
    return MF_ERR_INVALID_ARGUMENT;
  }
  return MF_OK;
}

int64_t mf_lanczos_sharded_heap_bytes(const mf_halo_plan_t* plan, int64_t ld, int64_t k,
                                      int32_t reortho, int32_t want_Q, int32_t dtype) {
  if (plan == nullptr || !valid_ld(ld) || k < 0) return -1;
  return sharded_blocks(k, reortho, want_Q != 0) * plan->rows_alloc * ld *
         (int64_t)dtype_size(dtype);
}

int64_t mf_lanczos_sharded_workspace_bytes(const mf_operator_t* op, int64_t ld, int64_t k,
                                           int32_t reortho) {
  if (check_sharded_args(op, nullptr, ld, k, reortho) != MF_OK) return -1;
  Arena a(nullptr, 0, true);
  LanczosBufs b;
  carve_lanczos(a, op, ld, k, reortho, &b);
  return a.used + 256;
}

int32_t mf_lanczos_sharded(const mf_comm_t* comm, const mf_operator_t* op,
                           const mf_halo_plan_t* plan, const void* V0, int64_t ld, int64_t k,
                           int32_t reortho, int32_t want_Q, int64_t heap_offset, void* ext,
                           void* alphas, void* betas, void* init_len, void* residual,
                           void* workspace, int64_t workspace_bytes, void* stream) {
  if (plan == nullptr) {
    set_error("lanczos_sharded: the halo plan is missing");
    return MF_ERR_INVALID_ARGUMENT;
  }
  MF_TRY(check_sharded_args(op, plan, ld, k, reortho));
  if (V0 == nullptr || init_len == nullptr || (k > 0 && (alphas == nullptr || betas == nullptr))) {
    set_error("lanczos_sharded: V0, init_len, alphas, betas must be non-null");
    return MF_ERR_INVALID_ARGUMENT;
  }
  const int64_t need = mf_lanczos_sharded_heap_bytes(plan, ld, k, reortho, want_Q, op->dtype);
  Shard sh{comm, comm_ctx(comm), plan, heap_offset, nullptr};
  if (sh.peer != nullptr) {
    if (heap_offset < 0 || heap_offset % 16 != 0 ||
        heap_offset + need > mf_comm_heap_bytes(comm)) {
      set_error("lanczos_sharded: the communicator heap holds %lld bytes, the extended blocks "
                "need %lld at offset %lld", (long long)mf_comm_heap_bytes(comm), (long long)need,
                (long long)heap_offset);
      return MF_ERR_WORKSPACE;
    }
    sh.ext = (unsigned char*)comm_heap(comm) + heap_offset;
  } else {
    // one rank: no exchange; the blocks may live anywhere
    sh.ext = ext != nullptr ? (unsigned char*)ext
                            : (comm != nullptr ? (unsigned char*)comm_heap(comm) + heap_offset
                                               : nullptr);
    if (sh.ext == nullptr) {
      set_error("lanczos_sharded: without a communicator `ext` must point at the extended blocks");
      return MF_ERR_INVALID_ARGUMENT;
    }
  }
  Arena a(workspace, workspace_bytes, false);
  LanczosBufs b;
  MF_TRY(carve_lanczos(a, op, ld, k, reortho, &b));
  cudaStream_t st = (cudaStream_t)stream;
  MF_TRY(zero_counter(b, st));
  if (reortho == MF_REORTHO_NONE)
    return lanczos_none_sharded(sh, op, V0, ld, k, want_Q != 0, alphas, betas, init_len, residual,
                                b, st);
  return lanczos_full_sharded(sh, op, V0, ld, k, alphas, betas, init_len, residual, b, st);
}

int32_t mf_hessenberg(const mf_operator_t* op, const void* V0, int64_t ld, int64_t k,
                      int32_t reortho, void* H, void* init_len, void* Q, void* residual,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  MF_TRY(check_lanczos_args(op, ld, k, reortho));
  if (V0 == nullptr || init_len == nullptr || (k > 0 && (H == nullptr || Q == nullptr))) {
    set_error("hessenberg: V0, init_len, H and Q must be non-null");
    return MF_ERR_INVALID_ARGUMENT;
  }
  Arena a(workspace, workspace_bytes, false);
  LanczosBufs b;
  MF_TRY(carve_lanczos(a, op, ld, k, MF_REORTHO_FULL, &b));
  // scratch rows for the norms (the symmetric driver's alphas / betas): after the Lanczos buffers
  const int64_t es = (int64_t)dtype_size(op->dtype);
  void* al = a.take((k + 1) * ld * es);
  void* be = a.take((k + 1) * ld * es);
  if (a.used > a.size) {
    set_error("workspace too small: need %lld bytes, have %lld", (long long)a.used,
              (long long)a.size);
    return MF_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  MF_TRY(zero_counter(b, st));
  if (k > 0 && cudaMemsetAsync(H, 0, (size_t)(k * k * ld * es), st) != cudaSuccess) {
    set_error("hessenberg: memset failed");
    return MF_ERR_CUDA;
  }
  const HessOut hess{H, reortho == MF_REORTHO_FULL};
  return lanczos_full(op, V0, false, ld, k, al, be, init_len, Q, residual, b, st, &hess);
}

int64_t mf_hessenberg_workspace_bytes(const mf_operator_t* op, int64_t ld, int64_t k) {
  if (check_lanczos_args(op, ld, k, MF_REORTHO_FULL) != MF_OK) return -1;
  Arena a(nullptr, 0, true);
  LanczosBufs b;
  carve_lanczos(a, op, ld, k, MF_REORTHO_FULL, &b);
  return a.used + 2 * (k + 1) * ld * (int64_t)dtype_size(op->dtype) + 1024;
}

int64_t mf_tridiag_quad_workspace_bytes(int64_t ld, int64_t k) {
  // d, e, first-row z : 3*k*ld doubles ; full eigenvector variant: (2 + k)*k*ld
  return (2 + k) * k * ld * 8 + 256;
}

int32_t mf_tridiag_quad(const void* alphas, const void* betas, const void* init_len,
                        int32_t dtype, int64_t ld, int64_t num_probes, int64_t k, int32_t fn,
                        double fn_param, void* quad, double* nodes, double* weights,
                        void* workspace, int64_t workspace_bytes, void* stream) {
  if (k <= 0 || ld <= 0 || num_probes > ld || alphas == nullptr || betas == nullptr) {
    set_error("tridiag_quad: bad arguments (k=%lld ld=%lld)", (long long)k, (long long)ld);
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (workspace == nullptr || workspace_bytes < 3 * k * ld * 8) {
    set_error("tridiag_quad: workspace too small");
    return MF_ERR_WORKSPACE;
  }
  return launch_tridiag_quad(alphas, betas, init_len, dtype, ld, num_probes, k, fn, fn_param,
                             quad, nodes, weights, nullptr, (double*)workspace,
                             (cudaStream_t)stream);
}

int32_t mf_bidiag_quad(const void* alphas, const void* betas, const void* init_len,
                       int32_t dtype, int64_t ld, int64_t num_probes, int64_t k, int32_t fn,
                       double fn_param, void* quad, double* nodes, double* weights,
                       void* workspace, int64_t workspace_bytes, void* stream) {
  if (k <= 0 || ld <= 0 || num_probes > ld || alphas == nullptr || betas == nullptr) {
    set_error("bidiag_quad: bad arguments (k=%lld ld=%lld)", (long long)k, (long long)ld);
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (workspace == nullptr || workspace_bytes < 3 * k * ld * 8) {
    set_error("bidiag_quad: workspace too small");
    return MF_ERR_WORKSPACE;
  }
  return launch_tridiag_quad(alphas, betas, init_len, dtype, ld, num_probes, k, fn, fn_param,
                             quad, nodes, weights, nullptr, (double*)workspace,
                             (cudaStream_t)stream, 1);
}

// One product with a rectangular dense operator or its transpose (Golub-Kahan needs both).
int64_t mf_matmat_rect_workspace_bytes(int64_t m, int64_t n, int64_t lda, int32_t trans,
                                       int32_t dtype, int32_t have_planes, int64_t ld) {
  const int64_t M = trans ? n : m, K = trans ? m : n;
  if (have_planes && g_gemm_tc.load(std::memory_order_relaxed) && tc_gemm_supported(lda, M, K, ld, dtype))
    return 2 * K * ld * 4 + 256;
  return 256;
}

int32_t mf_matmat_rect(const void* A, const void* A_planes, int64_t m, int64_t n, int64_t lda,
                       int32_t trans, int32_t dtype, const void* X, void* W, int64_t ld,
                       void* workspace, int64_t workspace_bytes, void* stream) {
  if (!A || !X || !W || m <= 0 || n <= 0 || lda < n || !valid_ld(ld) ||
      (dtype != MF_F32 && dtype != MF_F64)) {
    set_error("matmat_rect: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = trans ? n : m, K = trans ? m : n;
  if (A_planes != nullptr && g_gemm_tc.load(std::memory_order_relaxed) &&
      tc_gemm_supported(lda, M, K, ld, dtype)) {
    const int64_t need = 2 * K * ld * 4;
    void* xs = (void*)align_up((int64_t)(uintptr_t)workspace, 256);
    if (workspace == nullptr || (char*)xs + need > (char*)workspace + workspace_bytes) {
      set_error("matmat_rect: workspace for the TF32 planes of X is missing");
      return MF_ERR_WORKSPACE;
    }
    MF_TRY(launch_split_tf32(X, xs, K * ld, st));
    return launch_gemm_tcgen05(A_planes, lda, trans != 0, M, K, xs, 1, nullptr, W, nullptr, ld,
                               gemm_variant(), st);
  }
  const auto gemm = g_gemm_tc.load(std::memory_order_relaxed) ? launch_gemm_blocked : launch_gemm_simt;
  return gemm(A, lda, trans != 0, M, K, X, nullptr, W, ld, dtype, st);
}

int32_t mf_tridiag_funm_e1(const void* alphas, const void* betas, int32_t dtype, int64_t ld,
                           int64_t num_probes, int64_t k, int32_t fn, double fn_param,
                           void* coeffs, void* workspace, int64_t workspace_bytes,
                           void* stream) {
  if (k <= 0 || ld <= 0 || num_probes > ld || !alphas || !betas || !coeffs) {
    set_error("tridiag_funm_e1: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (workspace == nullptr || workspace_bytes < (2 + k) * k * ld * 8) {
    set_error("tridiag_funm_e1: workspace too small");
    return MF_ERR_WORKSPACE;
  }
  return launch_tridiag_quad(alphas, betas, nullptr, dtype, ld, num_probes, k, fn, fn_param,
                             nullptr, nullptr, nullptr, coeffs, (double*)workspace,
                             (cudaStream_t)stream);
}

int32_t mf_basis_combine(const void* Q, const void* coeffs, const void* scale, int32_t dtype,
                         int64_t n, int64_t ld, int64_t k, void* out, void* stream) {
  if (!valid_ld(ld) || !Q || !coeffs || !out) {
    set_error("basis_combine: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  return launch_basis_combine(Q, coeffs, scale, dtype, n, ld, k, out, (cudaStream_t)stream);
}

int32_t mf_mc_reduce(const void* values, int32_t dtype, int64_t num, double* stats_out,
                     void* stream) {
  if (values == nullptr || stats_out == nullptr || num <= 0) {
    set_error("mc_reduce: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  return launch_mc_reduce(values, dtype, num, stats_out, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ row-sharded building blocks
// The kernels of mf_lanczos, one call each, with every reduction stopping at this device's
// fp64 partial sums: a row-sharded driver all-reduces them between the calls.

int64_t mf_blockvec_workspace_bytes(int64_t ld, int64_t max_nq) {
  if (!valid_ld(ld) || max_nq < 0) return -1;
  return 256 + partial_bytes(ld, (int)reorth_partial_rows(ld, max_nq)) + 256;
}

namespace {
struct BvScratch {
  unsigned int* counter;
  double* partial;
  int64_t partial_rows;
};
int32_t carve_bv(void* ws, int64_t bytes, int64_t ld, BvScratch* b, cudaStream_t st) {
  if (!valid_ld(ld)) {
    set_error("ld=%lld must be a power of two in [1, 256]", (long long)ld);
    return MF_ERR_INVALID_ARGUMENT;
  }
  Arena a(ws, bytes, false);
  b->counter = (unsigned int*)a.take(256);
  b->partial = (double*)a.take(partial_bytes(ld, 4));
  // whatever the caller provided beyond the minimum holds more accumulator rows
  b->partial_rows = 4;
  if (ws != nullptr && bytes > a.used)
    b->partial_rows = 4 + (bytes - a.used) / partial_bytes(ld, 1);
  if (ws == nullptr || a.used > a.size) {
    set_error("workspace too small: need %lld bytes, have %lld", (long long)a.used, (long long)bytes);
    return MF_ERR_WORKSPACE;
  }
  if (cudaMemsetAsync(b->counter, 0, 256, st) != cudaSuccess) {
    set_error("counter memset failed");
    return MF_ERR_CUDA;
  }
  return MF_OK;
}
}  // namespace

int32_t mf_block_dot(const void* X, const void* Y, int32_t dtype, int64_t n, int64_t ld,
                     double* sums, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!X || !Y || !sums || n < 0) {
    set_error("block_dot: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  BvScratch b;
  MF_TRY(carve_bv(workspace, workspace_bytes, ld, &b, st));
  const Reduce red{b.partial, Finalize{b.counter, 0, nullptr, nullptr, sums}};
  return launch_dot(X, nullptr, Y, dtype, n, ld, red, st);
}

int32_t mf_reorth_dots(const void* Q, int64_t nq, const void* V, int32_t dtype, int64_t n,
                       int64_t ld, double* sums, void* workspace, int64_t workspace_bytes,
                       void* stream) {
  if (!Q || !V || !sums || nq <= 0 || n < 0) {
    set_error("reorth_dots: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  BvScratch b;
  MF_TRY(carve_bv(workspace, workspace_bytes, ld, &b, st));
  return launch_reorth_dots(Q, nq, V, dtype, n, ld, b.partial, b.counter, nullptr, st, sums,
                            b.partial_rows);
}

int32_t mf_reorth_update(const void* Q, int64_t nq, const void* h, void* V, int32_t dtype,
                         int64_t n, int64_t ld, double* sqnorm, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  if (!Q || !V || !h || nq <= 0 || n < 0) {
    set_error("reorth_update: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (sqnorm == nullptr) {
    if (!valid_ld(ld)) {
      set_error("ld=%lld must be a power of two in [1, 256]", (long long)ld);
      return MF_ERR_INVALID_ARGUMENT;
    }
    return launch_reorth_update(Q, nq, h, V, dtype, n, ld, nullptr, st);
  }
  BvScratch b;
  MF_TRY(carve_bv(workspace, workspace_bytes, ld, &b, st));
  const Reduce red{b.partial, Finalize{b.counter, 0, nullptr, nullptr, sqnorm}};
  return launch_reorth_update(Q, nq, h, V, dtype, n, ld, &red, st);
}

int32_t mf_lanczos_update(const void* W, const void* Rc, const void* sc, const void* a,
                          const void* Rp, const void* sp, const void* bprev, void* out,
                          int32_t dtype, int64_t n, int64_t ld, double* sqnorm, void* workspace,
                          int64_t workspace_bytes, void* stream) {
  if (!W || !Rc || !a || !out || !sqnorm || n < 0) {
    set_error("lanczos_update: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  BvScratch b;
  MF_TRY(carve_bv(workspace, workspace_bytes, ld, &b, st));
  const Reduce red{b.partial, Finalize{b.counter, 0, nullptr, nullptr, sqnorm}};
  return launch_lanczos_update(W, Rc, sc, a, Rp, sp, bprev, out, dtype, n, ld, red, st);
}

int32_t mf_block_scale(const void* X, const void* s, void* out, int32_t divide, int32_t dtype,
                       int64_t n, int64_t ld, void* stream) {
  if (!X || !out || !valid_ld(ld) || n < 0) {
    set_error("block_scale: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  return launch_scale(X, s, out, divide ? 1 : 0, dtype, n, ld, (cudaStream_t)stream);
}

int32_t mf_sums_finalize(const double* sums, int64_t count, int32_t take_sqrt, void* value,
                         void* inv, int32_t dtype, void* stream) {
  if (!sums || count < 0 || (dtype != MF_F32 && dtype != MF_F64)) {
    set_error("sums_finalize: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  return launch_sums_finalize(sums, count, take_sqrt ? 1 : 0, value, inv, dtype,
                              (cudaStream_t)stream);
}

int32_t mf_hutch_rows(const void* A, const void* B, int32_t dtype, int64_t n, int64_t ld,
                      int64_t num_probes, int32_t accumulate, double* rowsum, double* rowsumsq,
                      void* stream) {
  if (!A || !B || !rowsum || n < 0 || !valid_ld(ld) || num_probes < 0 || num_probes > ld ||
      (dtype != MF_F32 && dtype != MF_F64)) {
    set_error("hutch_rows: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  return launch_hutch_rows(A, B, dtype, n, ld, num_probes, accumulate != 0, rowsum, rowsumsq,
                           (cudaStream_t)stream);
}

int32_t mf_full_offdiag(void* offdiag_row, const void* h_row, int32_t dtype, int64_t ld,
                        void* stream) {
  if (!offdiag_row || !h_row || ld <= 0) {
    set_error("full_offdiag: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  return launch_full_offdiag(offdiag_row, h_row, dtype, ld, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ two-pass f(A) v

struct FunmBufs {
  void *alphas, *betas, *len, *coeffs;
  double* qwork;
  LanczosBufs lb;
};

static int32_t carve_funm(Arena& a, const mf_operator_t* op, int64_t ld, int64_t k, FunmBufs* f) {
  const int64_t es = (int64_t)dtype_size(op->dtype);
  memset(f, 0, sizeof(*f));
  f->alphas = a.take((k + 1) * ld * es);
  f->betas = a.take((k + 1) * ld * es);
  f->len = a.take(ld * es);
  f->coeffs = a.take((k + 1) * ld * es);
  f->qwork = (double*)a.take((2 + k) * k * ld * 8);
  MF_TRY(carve_lanczos(a, op, ld, k, MF_REORTHO_NONE, &f->lb));
  if (!a.dry && a.used > a.size) {
    set_error("workspace too small: need %lld bytes, have %lld", (long long)a.used,
              (long long)a.size);
    return MF_ERR_WORKSPACE;
  }
  return MF_OK;
}

int64_t mf_funm_lanczos_workspace_bytes(const mf_operator_t* op, int64_t ld, int64_t k) {
  if (check_lanczos_args(op, ld, k, MF_REORTHO_NONE) != MF_OK) return -1;
  Arena a(nullptr, 0, true);
  FunmBufs f;
  carve_funm(a, op, ld, k, &f);
  return a.used + 256;
}

int32_t mf_funm_lanczos(const mf_operator_t* op, const void* V0, int64_t ld, int64_t num_probes,
                        int64_t k, int32_t fn, double fn_param, void* out, void* workspace,
                        int64_t workspace_bytes, void* stream) {
  MF_TRY(check_lanczos_args(op, ld, k, MF_REORTHO_NONE));
  if (V0 == nullptr || out == nullptr || k < 1 || num_probes < 1 || num_probes > ld ||
      fn == MF_FN_NONE) {
    set_error("funm_lanczos: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Arena a(workspace, workspace_bytes, false);
  FunmBufs f;
  MF_TRY(carve_funm(a, op, ld, k, &f));
  MF_TRY(zero_counter(f.lb, st));
  const int32_t dt = op->dtype;
  // pass 1: the tridiagonal matrix (no basis kept)
  MF_TRY(lanczos_none(op, (void*)V0, false, false, ld, k, f.alphas, f.betas, f.len, nullptr,
                      nullptr, f.lb, st));
  // y = f(T) e1 per probe
  MF_TRY(launch_tridiag_quad(f.alphas, f.betas, nullptr, dt, ld, num_probes, k, fn, fn_param,
                             nullptr, nullptr, nullptr, f.coeffs, f.qwork, st));
  // pass 2: the same recurrence again (bit-identical: the kernels are deterministic),
  // accumulating |v0| * sum_j y_j v_j on the way
  const BasisAccum acc{f.coeffs, out};
  return lanczos_none(op, (void*)V0, false, true, ld, k, f.alphas, f.betas, f.len, nullptr, nullptr,
                      f.lb, st, &acc);
}

// ------------------------------------------------------------------ fused estimator

struct EstimateBufs {
  void* Z;        // probe block
  void* Qbasis;   // FULL only: [k][n][ld]
  void* alphas;   // [k][ld]
  void* betas;    // [k][ld]
  void* len;      // [ld]
  double* qwork;  // 3*k*ld doubles
  LanczosBufs lb;
};

static int32_t carve_estimate(Arena& a, const mf_operator_t* op, int64_t ld, int64_t k,
                              int32_t reortho, int32_t integrand, EstimateBufs* e) {
  const int64_t es = (int64_t)dtype_size(op->dtype);
  const int64_t blk = op->n * ld * es;
  memset(e, 0, sizeof(*e));
  e->Z = a.take(blk);
  if (integrand == MF_INTEGRAND_TRACE) {
    e->lb.counter = (unsigned int*)a.take(256);
    e->lb.W = a.take(blk);
    e->lb.partial = (double*)a.take(partial_bytes(ld, 1));
    e->len = a.take(ld * es);
    carve_op_scratch(a, op, ld, &e->lb.scr);
    if (!a.dry && a.used > a.size) {
      set_error("workspace too small: need %lld bytes, have %lld", (long long)a.used,
                (long long)a.size);
      return MF_ERR_WORKSPACE;
    }
    return MF_OK;
  }
  if (reortho == MF_REORTHO_FULL) e->Qbasis = a.take(k * blk);
  e->alphas = a.take((k + 1) * ld * es);
  e->betas = a.take((k + 1) * ld * es);
  e->len = a.take(ld * es);
  e->qwork = (double*)a.take(3 * k * ld * 8);
  MF_TRY(carve_lanczos(a, op, ld, k, reortho, &e->lb));
  if (!a.dry && a.used > a.size) {
    set_error("workspace too small: need %lld bytes, have %lld", (long long)a.used,
              (long long)a.size);
    return MF_ERR_WORKSPACE;
  }
  return MF_OK;
}

int64_t mf_estimate_workspace_bytes(const mf_operator_t* op, int64_t ld, int64_t k,
                                    int32_t reortho, int32_t integrand) {
  if (integrand == MF_INTEGRAND_TRACE) {
    if (validate_op(op) != MF_OK || !valid_ld(ld)) return -1;
  } else if (check_lanczos_args(op, ld, k, reortho) != MF_OK) {
    return -1;
  }
  Arena a(nullptr, 0, true);
  EstimateBufs e;
  carve_estimate(a, op, ld, k, reortho, integrand, &e);
  return a.used + 256;
}

int32_t mf_estimate(const mf_operator_t* op, int32_t integrand, int32_t sampler,
                    int32_t prng_flags, uint32_t key0, uint32_t key1, int64_t p0,
                    int64_t num_probes, int64_t ld, int64_t k, int32_t reortho, int32_t fn,
                    double fn_param, void* quad_out, void* alphas_out, void* betas_out,
                    void* lens_out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (integrand == MF_INTEGRAND_TRACE) {
    MF_TRY(validate_op(op));
    if (!valid_ld(ld)) {
      set_error("ld=%lld must be a power of two in [1, 256]", (long long)ld);
      return MF_ERR_INVALID_ARGUMENT;
    }
  } else if (integrand == MF_INTEGRAND_SLQ) {
    MF_TRY(check_lanczos_args(op, ld, k, reortho));
    if (k < 1) {
      set_error("estimate: the SLQ integrand needs num_matvecs >= 1");
      return MF_ERR_INVALID_ARGUMENT;
    }
  } else {
    set_error("estimate: unknown integrand %d", integrand);
    return MF_ERR_INVALID_ARGUMENT;
  }
  if (quad_out == nullptr || num_probes <= 0 || p0 < 0) {
    set_error("estimate: bad arguments");
    return MF_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t dt = op->dtype;
  const int64_t es = (int64_t)dtype_size(dt);
  Arena a(workspace, workspace_bytes, false);
  EstimateBufs e;
  MF_TRY(carve_estimate(a, op, ld, k, reortho, integrand, &e));

  MF_TRY(zero_counter(e.lb, st));
  int64_t tile = 0;
  for (int64_t t0 = 0; t0 < num_probes; t0 += ld, ++tile) {
    const int64_t np = (num_probes - t0) < ld ? (num_probes - t0) : ld;
    void* q_tile = (char*)quad_out + t0 * es;
    if (integrand == MF_INTEGRAND_TRACE) {
      // v^T (A v)  (matfree/stochtrace.py:859-863)
      MF_TRY(launch_probe_gen(e.Z, dt, MF_LAYOUT_BLOCKED, op->n, ld, p0 + t0, np, key0, key1,
                              sampler, prng_flags, nullptr, st));
      // full tile: the last CTA writes straight to the output; partial tile: via scratch
      void* dst = np == ld ? q_tile : e.len;
      const Reduce red{e.lb.partial, Finalize{e.lb.counter, 0, dst, nullptr, nullptr}};
      bool fused = false;
      MF_TRY(apply_op(op, e.Z, nullptr, e.lb.W, ld, e.lb.scr, &red, e.lb.counter + 8, &fused, st));
      if (!fused) MF_TRY(launch_dot(e.Z, nullptr, e.lb.W, dt, op->n, ld, red, st));
      if (np != ld &&
          cudaMemcpyAsync(q_tile, e.len, np * es, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("estimate: copy failed");
        return MF_ERR_CUDA;
      }
      continue;
    }
    // probes + |v0| and 1/|v0| in one kernel (funm.py:228-229, decomp.py:227 / :435)
    {
      void* inv0 = reortho == MF_REORTHO_NONE ? e.lb.inv : nullptr;
      const Reduce red{e.lb.partial, Finalize{e.lb.counter, 1, e.len, inv0, nullptr}};
      MF_TRY(launch_probe_gen(e.Z, dt, MF_LAYOUT_BLOCKED, op->n, ld, p0 + t0, np, key0, key1,
                              sampler, prng_flags, &red, st));
    }
    if (reortho == MF_REORTHO_NONE)
      MF_TRY(lanczos_none(op, e.Z, true, true, ld, k, e.alphas, e.betas, e.len, nullptr, nullptr,
                          e.lb, st));
    else
      MF_TRY(lanczos_full(op, e.Z, true, ld, k, e.alphas, e.betas, e.len, e.Qbasis, nullptr, e.lb,
                          st));
    // quadrature: quad for the np live probes goes straight to the output
    MF_TRY(launch_tridiag_quad(e.alphas, e.betas, e.len, dt, ld, np, k, fn, fn_param, q_tile,
                               nullptr, nullptr, nullptr, e.qwork, st));
    if (alphas_out != nullptr &&
        cudaMemcpyAsync((char*)alphas_out + tile * k * ld * es, e.alphas, k * ld * es,
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("estimate: copy failed");
      return MF_ERR_CUDA;
    }
    if (betas_out != nullptr &&
        cudaMemcpyAsync((char*)betas_out + tile * k * ld * es, e.betas, k * ld * es,
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("estimate: copy failed");
      return MF_ERR_CUDA;
    }
    if (lens_out != nullptr &&
        cudaMemcpyAsync((char*)lens_out + tile * ld * es, e.len, ld * es,
                        cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      set_error("estimate: copy failed");
      return MF_ERR_CUDA;
    }
  }
  return MF_OK;
}

int32_t mf_slq_estimate_dense(const void* A, const void* A_planes, int64_t n, int64_t lda,
                              int32_t dtype,
                              int32_t sampler, int32_t prng_flags, uint32_t key0,
                              uint32_t key1, int64_t p0, int64_t num_probes, int64_t ld,
                              int64_t k, int32_t reortho, int32_t fn, double fn_param,
                              void* quad_out, void* workspace, int64_t workspace_bytes,
                              void* stream) {
  mf_operator_t op{};
  op.kind = MF_OP_DENSE; op.dtype = dtype; op.n = n; op.values = A; op.lda = lda;
  op.split_planes = A_planes;
  return mf_estimate(&op, MF_INTEGRAND_SLQ, sampler, prng_flags, key0, key1, p0, num_probes, ld,
                     k, reortho, fn, fn_param, quad_out, nullptr, nullptr, nullptr, workspace,
                     workspace_bytes, stream);
}

int32_t mf_slq_estimate_csr(const int32_t* indptr, const int32_t* indices, const void* data,
                            int64_t n, int64_t nnz, int32_t dtype, int32_t sampler,
                            int32_t prng_flags, uint32_t key0, uint32_t key1, int64_t p0,
                            int64_t num_probes, int64_t ld, int64_t k, int32_t reortho,
                            int32_t fn, double fn_param, void* quad_out, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  mf_operator_t op{};
  op.kind = MF_OP_CSR; op.dtype = dtype; op.n = n; op.nnz = nnz; op.values = data;
  op.indptr = indptr; op.indices = indices;
  return mf_estimate(&op, MF_INTEGRAND_SLQ, sampler, prng_flags, key0, key1, p0, num_probes, ld,
                     k, reortho, fn, fn_param, quad_out, nullptr, nullptr, nullptr, workspace,
                     workspace_bytes, stream);
}

int32_t mf_slq_estimate_gram(const void* A, const void* A_planes, int64_t m, int64_t n,
                             int64_t lda, int32_t dtype,
                             int32_t sampler, int32_t prng_flags, uint32_t key0,
                             uint32_t key1, int64_t p0, int64_t num_probes, int64_t ld,
                             int64_t k, int32_t reortho, int32_t fn, double fn_param,
                             void* quad_out, void* workspace, int64_t workspace_bytes,
                             void* stream) {
  mf_operator_t op{};
  op.kind = MF_OP_GRAM; op.dtype = dtype; op.n = n; op.m = m; op.values = A; op.lda = lda;
  op.split_planes = A_planes;
  return mf_estimate(&op, MF_INTEGRAND_SLQ, sampler, prng_flags, key0, key1, p0, num_probes, ld,
                     k, reortho, fn, fn_param, quad_out, nullptr, nullptr, nullptr, workspace,
                     workspace_bytes, stream);
}

}  // extern "C"
