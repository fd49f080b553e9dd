// XLA FFI handler shim: the symbols `jax.ffi.register_ffi_target` binds.
//
// The XLA FFI headers (xla/ffi/api/ffi.h) ship inside jaxlib, which is neither installed nor
// installable in this image (no network, no wheel): the Makefile adds this file to the library
// only when MF_XLA_INCLUDE points at the header tree.  In this repository the file is compiled by
// the CPU test suite against a minimal stand-in of the API surface it uses
// (tests/ffi_stub/xla/ffi/api/ffi.h), which checks every handler's signature against its binding
// and every mf_* call against include/matfree_b200.h; it has not been executed under XLA.
//
// The handlers are thin: they unpack XLA buffers into the plain C ABI, move between the
// reference's probe-major layout `(B, n)` and the library's blocked layout `[n][ld]`
// (mf_to_blocked / mf_from_blocked), take all scratch from XLA's ScratchAllocator and enqueue on
// XLA's stream -- no synchronisation, no allocation of their own, no retained pointers.
// Conventions shared by all handlers:
//   * PRNG keys are two uint32 attributes `key0`, `key1` (= jax.random.key_data(key), concrete at
//     trace time as in matfree's own use of `prng_key`); `x64_bits` selects the jax_enable_x64
//     Rademacher stream;
//   * vectors arrive probe-major with a leading batch axis, `(B, n)` (what `jax.vmap` with
//     vmap_method="expand_dims" hands over; B = 1 for an un-batched call), B <= 256;
//   * dense / Gram operators take `(A, planes)`: `planes` is the output of mf_operator_split_ffi
//     (TF32 hi/lo planes, computed once per operator) or a zero-element buffer (CUDA-core GEMM).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define MF_HAVE_XLA_FFI 1
#endif
#endif

#ifdef MF_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/matfree_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error to_error(int32_t rc) {
  if (rc == MF_OK) return ffi::Error::Success();
  return ffi::Error(rc == MF_ERR_INVALID_ARGUMENT ? ffi::ErrorCode::kInvalidArgument
                                                  : ffi::ErrorCode::kInternal,
                    mf_last_error());
}
#define MF_FFI_TRY(expr)                          \
  do {                                            \
    const int32_t mf_rc_ = (expr);                \
    if (mf_rc_ != MF_OK) return to_error(mf_rc_); \
  } while (0)

int32_t mf_dtype_of(ffi::DataType t) { return t == ffi::F64 ? MF_F64 : MF_F32; }
int64_t elem_size(int32_t dtype) { return dtype == MF_F64 ? 8 : 4; }
int64_t tile_for(int64_t b) {
  int64_t ld = 1;
  while (ld < b) ld *= 2;
  return ld;
}
// `(B, n)`: the last axis is the vector, everything before it the batch
void batch_shape(const ffi::AnyBuffer& v, int64_t* B, int64_t* n) {
  auto dims = v.dimensions();
  *n = dims.size() ? (int64_t)dims[dims.size() - 1] : 1;
  *B = *n ? (int64_t)v.element_count() / *n : 0;
}

// bandwidth / num_diagonals / line_stride: the structure hints of the operator struct, 0 = unknown, computed
// once per operator by the binding; they select the TMA-staged kernels for stencils
mf_operator_t csr_op(const ffi::Buffer<ffi::S32>& indptr, const ffi::Buffer<ffi::S32>& indices,
                     const ffi::AnyBuffer& data, int64_t bandwidth, int64_t num_diagonals,
                     int64_t line_stride) {
  mf_operator_t op{};
  op.csr_bandwidth = bandwidth;
  op.csr_num_diagonals = (int32_t)num_diagonals;
  op.csr_line_stride = line_stride;
  op.kind = MF_OP_CSR;
  op.dtype = mf_dtype_of(data.element_type());
  op.n = (int64_t)indptr.element_count() - 1;
  op.nnz = (int64_t)data.element_count();
  op.values = data.untyped_data();
  op.indptr = indptr.typed_data();
  op.indices = indices.typed_data();
  return op;
}
mf_operator_t dense_op(int32_t kind, const ffi::AnyBuffer& A, const ffi::AnyBuffer& planes) {
  mf_operator_t op{};
  auto dims = A.dimensions();
  op.kind = kind;
  op.dtype = mf_dtype_of(A.element_type());
  op.m = (int64_t)dims[0];
  op.n = (int64_t)dims[1];
  op.lda = op.n;
  op.values = A.untyped_data();
  op.split_planes = planes.element_count() ? planes.untyped_data() : nullptr;
  return op;
}

struct Scratch {
  ffi::ScratchAllocator* alloc;
  bool ok = true;
  void* take(int64_t bytes) {
    auto p = alloc->Allocate((size_t)(bytes > 0 ? bytes : 16));
    if (!p.has_value()) {
      ok = false;
      return nullptr;
    }
    return *p;
  }
};

// jax.random.rademacher / normal for the (num, n) sample array of one key
// (matfree/stochtrace.py:957-964 -> backend/prng.py:14-29).
ffi::Error ProbeGenImpl(cudaStream_t stream, int64_t n, int64_t num, int32_t sampler,
                        int32_t x64_bits, uint32_t key0, uint32_t key1,
                        ffi::Result<ffi::AnyBuffer> out) {
  return to_error(mf_probe_gen(out->untyped_data(), mf_dtype_of(out->element_type()),
                               MF_LAYOUT_PROBE_MAJOR, n, n, 0, num, key0, key1, sampler,
                               x64_bits ? MF_PRNG_X64_BITS : 0, nullptr, stream));
}

// One-time TF32 hi/lo planes of an fp32 dense / Gram matrix (mf_operator_split).
ffi::Error OperatorSplitImpl(cudaStream_t stream, ffi::AnyBuffer A, int32_t kind,
                             ffi::Result<ffi::AnyBuffer> planes) {
  mf_operator_t op{};
  auto dims = A.dimensions();
  op.kind = kind;
  op.dtype = mf_dtype_of(A.element_type());
  op.m = (int64_t)dims[0];
  op.n = (int64_t)dims[1];
  op.lda = op.n;
  op.values = A.untyped_data();
  return to_error(mf_operator_split(&op, planes->untyped_data(), stream));
}

// decomp.tridiag_sym(num_matvecs, reortho=...)(matvec, vec) for one registered operator on a
// batch of start vectors (matfree/decomp.py:125-145,155-292,426-477).  Outputs in the
// reference's layouts: Q (k, B, n), diags (B, k), betas (B, k) [offdiags = betas[:, :k-1],
// betas[:, k-1] = residual norm], residual (B, n), init_length (B,).
ffi::Error LanczosCommon(cudaStream_t stream, ffi::ScratchAllocator& scratch, const mf_operator_t& op,
                         const ffi::AnyBuffer& vec, int64_t k, int32_t reortho,
                         ffi::Result<ffi::AnyBuffer>& Q, ffi::Result<ffi::AnyBuffer>& alphas,
                         ffi::Result<ffi::AnyBuffer>& betas, ffi::Result<ffi::AnyBuffer>& residual,
                         ffi::Result<ffi::AnyBuffer>& init_len) {
  int64_t B, n;
  batch_shape(vec, &B, &n);
  if (B < 1 || B > 256 || n != op.n)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "lanczos: vec must be (B <= 256, n)");
  const int64_t ld = tile_for(B), es = elem_size(op.dtype);
  const int64_t ws = mf_lanczos_workspace_bytes(&op, ld, k, reortho, 1);
  if (ws < 0) return ffi::Error(ffi::ErrorCode::kInvalidArgument, mf_last_error());
  Scratch s{&scratch};
  void* v0b = s.take(n * ld * es);
  void* ab = s.take((k + 1) * ld * es);
  void* bb = s.take((k + 1) * ld * es);
  void* lb = s.take(ld * es);
  void* qb = s.take((k > 0 ? k : 1) * n * ld * es);
  void* rb = s.take(n * ld * es);
  void* wsp = s.take(ws);
  if (!s.ok) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "mf_lanczos scratch");
  if (ld > B && cudaMemsetAsync(v0b, 0, (size_t)(n * ld * es), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "memset failed");
  MF_FFI_TRY(mf_to_blocked(vec.untyped_data(), v0b, op.dtype, n, B, ld, stream));
  MF_FFI_TRY(mf_lanczos(&op, v0b, ld, k, reortho, ab, bb, lb, qb, rb, wsp, ws, stream));
  for (int64_t j = 0; j < k; ++j)
    MF_FFI_TRY(mf_from_blocked((const char*)qb + j * n * ld * es,
                               (char*)Q->untyped_data() + j * B * n * es, op.dtype, n, B, ld, stream));
  if (k > 0) {
    MF_FFI_TRY(mf_from_blocked(ab, alphas->untyped_data(), op.dtype, k, B, ld, stream));
    MF_FFI_TRY(mf_from_blocked(bb, betas->untyped_data(), op.dtype, k, B, ld, stream));
  }
  MF_FFI_TRY(mf_from_blocked(rb, residual->untyped_data(), op.dtype, n, B, ld, stream));
  if (cudaMemcpyAsync(init_len->untyped_data(), lb, (size_t)(B * es), cudaMemcpyDeviceToDevice,
                      stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "copy failed");
  return ffi::Error::Success();
}

ffi::Error LanczosCsrImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                          ffi::Buffer<ffi::S32> indptr, ffi::Buffer<ffi::S32> indices,
                          ffi::AnyBuffer data, ffi::AnyBuffer vec, int64_t num_matvecs,
                          int32_t reortho, int64_t csr_bandwidth, int64_t csr_num_diagonals,
                          int64_t csr_line_stride, ffi::Result<ffi::AnyBuffer> Q,
                          ffi::Result<ffi::AnyBuffer> alphas, ffi::Result<ffi::AnyBuffer> betas,
                          ffi::Result<ffi::AnyBuffer> residual, ffi::Result<ffi::AnyBuffer> init_len) {
  const mf_operator_t op = csr_op(indptr, indices, data, csr_bandwidth, csr_num_diagonals, csr_line_stride);
  return LanczosCommon(stream, scratch, op, vec, num_matvecs, reortho, Q, alphas, betas, residual,
                       init_len);
}
ffi::Error LanczosDenseImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer A,
                            ffi::AnyBuffer planes, ffi::AnyBuffer vec, int64_t num_matvecs,
                            int32_t reortho, ffi::Result<ffi::AnyBuffer> Q,
                            ffi::Result<ffi::AnyBuffer> alphas, ffi::Result<ffi::AnyBuffer> betas,
                            ffi::Result<ffi::AnyBuffer> residual, ffi::Result<ffi::AnyBuffer> init_len) {
  const mf_operator_t op = dense_op(MF_OP_DENSE, A, planes);
  return LanczosCommon(stream, scratch, op, vec, num_matvecs, reortho, Q, alphas, betas, residual,
                       init_len);
}
ffi::Error LanczosGramImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer A,
                           ffi::AnyBuffer planes, ffi::AnyBuffer vec, int64_t num_matvecs,
                           int32_t reortho, ffi::Result<ffi::AnyBuffer> Q,
                           ffi::Result<ffi::AnyBuffer> alphas, ffi::Result<ffi::AnyBuffer> betas,
                           ffi::Result<ffi::AnyBuffer> residual, ffi::Result<ffi::AnyBuffer> init_len) {
  const mf_operator_t op = dense_op(MF_OP_GRAM, A, planes);
  return LanczosCommon(stream, scratch, op, vec, num_matvecs, reortho, Q, alphas, betas, residual,
                       init_len);
}

// dense_funm_sym_eigh + e1^T f(T) e1 (matfree/funm.py:239-241,322-335) for a batch of tridiagonal
// matrices: alphas (B, k), betas (B, k) as returned above, init_length (B,) -> quad (B,).
ffi::Error TridiagQuadImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer alphas,
                           ffi::AnyBuffer betas, ffi::AnyBuffer init_len, int32_t fn, double fn_param,
                           ffi::Result<ffi::AnyBuffer> quad) {
  int64_t B, k;
  batch_shape(alphas, &B, &k);
  if (B < 1 || B > 256 || k < 1)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "tridiag_quad: alphas must be (B <= 256, k >= 1)");
  const int32_t dt = mf_dtype_of(alphas.element_type());
  const int64_t ld = tile_for(B), es = elem_size(dt);
  const int64_t ws = mf_tridiag_quad_workspace_bytes(ld, k);
  Scratch s{&scratch};
  void* ab = s.take(k * ld * es);
  void* bb = s.take(k * ld * es);
  void* lb = s.take(ld * es);
  void* qb = s.take(ld * es);
  void* wsp = s.take(ws);
  if (!s.ok) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "mf_tridiag_quad scratch");
  MF_FFI_TRY(mf_to_blocked(alphas.untyped_data(), ab, dt, k, B, ld, stream));
  MF_FFI_TRY(mf_to_blocked(betas.untyped_data(), bb, dt, k, B, ld, stream));
  if (cudaMemcpyAsync(lb, init_len.untyped_data(), (size_t)(B * es), cudaMemcpyDeviceToDevice,
                      stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "copy failed");
  MF_FFI_TRY(mf_tridiag_quad(ab, bb, lb, dt, ld, B, k, fn, fn_param, qb, nullptr, nullptr, wsp, ws,
                             stream));
  if (cudaMemcpyAsync(quad->untyped_data(), qb, (size_t)(B * es), cudaMemcpyDeviceToDevice,
                      stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "copy failed");
  return ffi::Error::Success();
}

// estimate(matvec, key) of matfree/stochtrace.py:47-50 with the SLQ integrand: one value per
// probe of the range [p0, p0 + num_probes); the caller takes jnp.mean / jnp.std (or mf_mc_reduce).
ffi::Error EstimateCommon(cudaStream_t stream, ffi::ScratchAllocator& scratch, const mf_operator_t& op,
                          int64_t num_probes, int64_t p0, int64_t num_matvecs, int32_t reortho,
                          int32_t sampler, int32_t x64_bits, int32_t fn, double fn_param, int64_t tile,
                          uint32_t key0, uint32_t key1, ffi::Result<ffi::AnyBuffer>& quad) {
  const int64_t ws = mf_estimate_workspace_bytes(&op, tile, num_matvecs, reortho, MF_INTEGRAND_SLQ);
  if (ws < 0) return ffi::Error(ffi::ErrorCode::kInvalidArgument, mf_last_error());
  Scratch s{&scratch};
  void* wsp = s.take(ws);
  if (!s.ok) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "mf_estimate workspace");
  return to_error(mf_estimate(&op, MF_INTEGRAND_SLQ, sampler, x64_bits ? MF_PRNG_X64_BITS : 0, key0,
                              key1, p0, num_probes, tile, num_matvecs, reortho, fn, fn_param,
                              quad->untyped_data(), nullptr, nullptr, nullptr, wsp, ws, stream));
}

ffi::Error SlqEstimateCsrImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                              ffi::Buffer<ffi::S32> indptr, ffi::Buffer<ffi::S32> indices,
                              ffi::AnyBuffer data, int64_t num_probes, int64_t p0,
                              int64_t num_matvecs, int32_t reortho, int32_t sampler, int32_t x64_bits,
                              int32_t fn, double fn_param, int64_t tile, uint32_t key0, uint32_t key1,
                              int64_t csr_bandwidth, int64_t csr_num_diagonals, int64_t csr_line_stride,
                              ffi::Result<ffi::AnyBuffer> quad) {
  const mf_operator_t op = csr_op(indptr, indices, data, csr_bandwidth, csr_num_diagonals, csr_line_stride);
  return EstimateCommon(stream, scratch, op, num_probes, p0, num_matvecs, reortho, sampler, x64_bits,
                        fn, fn_param, tile, key0, key1, quad);
}
ffi::Error SlqEstimateDenseImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer A,
                                ffi::AnyBuffer planes, int64_t num_probes, int64_t p0,
                                int64_t num_matvecs, int32_t reortho, int32_t sampler, int32_t x64_bits,
                                int32_t fn, double fn_param, int64_t tile, uint32_t key0, uint32_t key1,
                                ffi::Result<ffi::AnyBuffer> quad) {
  const mf_operator_t op = dense_op(MF_OP_DENSE, A, planes);
  return EstimateCommon(stream, scratch, op, num_probes, p0, num_matvecs, reortho, sampler, x64_bits,
                        fn, fn_param, tile, key0, key1, quad);
}
ffi::Error SlqEstimateGramImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer A,
                               ffi::AnyBuffer planes, int64_t num_probes, int64_t p0,
                               int64_t num_matvecs, int32_t reortho, int32_t sampler, int32_t x64_bits,
                               int32_t fn, double fn_param, int64_t tile, uint32_t key0, uint32_t key1,
                               ffi::Result<ffi::AnyBuffer> quad) {
  const mf_operator_t op = dense_op(MF_OP_GRAM, A, planes);
  return EstimateCommon(stream, scratch, op, num_probes, p0, num_matvecs, reortho, sampler, x64_bits,
                        fn, fn_param, tile, key0, key1, quad);
}

// np.mean / np.std over the probe axis (matfree/stochtrace.py:50,85-86):
// stats = float64[4] {mean, std(ddof=0), sem, P}.
ffi::Error McReduceImpl(cudaStream_t stream, ffi::AnyBuffer values,
                        ffi::Result<ffi::Buffer<ffi::F64>> stats) {
  return to_error(mf_mc_reduce(values.untyped_data(), mf_dtype_of(values.element_type()),
                               (int64_t)values.element_count(), stats->typed_data(), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_probe_gen_ffi, ProbeGenImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int64_t>("n")
        .Attr<int64_t>("num")
        .Attr<int32_t>("sampler")
        .Attr<int32_t>("x64_bits")
        .Attr<uint32_t>("key0")
        .Attr<uint32_t>("key1")
        .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_operator_split_ffi, OperatorSplitImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::AnyBuffer>()
        .Attr<int32_t>("kind")
        .Ret<ffi::AnyBuffer>());

#define MF_CSR_HINTS                         \
  .Attr<int64_t>("csr_bandwidth")            \
      .Attr<int64_t>("csr_num_diagonals")    \
      .Attr<int64_t>("csr_line_stride")
#define MF_LANCZOS_ATTRS               \
  .Arg<ffi::AnyBuffer>()               \
      .Attr<int64_t>("num_matvecs")    \
      .Attr<int32_t>("reortho")
#define MF_LANCZOS_TAIL MF_LANCZOS_ATTRS MF_LANCZOS_RETS
#define MF_LANCZOS_RETS                \
  .Ret<ffi::AnyBuffer>()               \
      .Ret<ffi::AnyBuffer>()           \
      .Ret<ffi::AnyBuffer>()           \
      .Ret<ffi::AnyBuffer>()           \
      .Ret<ffi::AnyBuffer>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_lanczos_csr_ffi, LanczosCsrImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::AnyBuffer>() MF_LANCZOS_ATTRS MF_CSR_HINTS MF_LANCZOS_RETS);

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_lanczos_dense_ffi, LanczosDenseImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>() MF_LANCZOS_TAIL);

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_lanczos_gram_ffi, LanczosGramImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>() MF_LANCZOS_TAIL);

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_tridiag_quad_ffi, TridiagQuadImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Attr<int32_t>("fn")
        .Attr<double>("fn_param")
        .Ret<ffi::AnyBuffer>());

#define MF_ESTIMATE_TAIL MF_ESTIMATE_ATTRS.Ret<ffi::AnyBuffer>()
#define MF_ESTIMATE_ATTRS              \
  .Attr<int64_t>("num_probes")         \
      .Attr<int64_t>("p0")             \
      .Attr<int64_t>("num_matvecs")    \
      .Attr<int32_t>("reortho")        \
      .Attr<int32_t>("sampler")        \
      .Attr<int32_t>("x64_bits")       \
      .Attr<int32_t>("fn")             \
      .Attr<double>("fn_param")        \
      .Attr<int64_t>("tile")           \
      .Attr<uint32_t>("key0")          \
      .Attr<uint32_t>("key1")

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_slq_estimate_csr_ffi, SlqEstimateCsrImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::AnyBuffer>() MF_ESTIMATE_ATTRS MF_CSR_HINTS.Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_slq_estimate_dense_ffi, SlqEstimateDenseImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>() MF_ESTIMATE_TAIL);

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_slq_estimate_gram_ffi, SlqEstimateGramImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>() MF_ESTIMATE_TAIL);

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_mc_reduce_ffi, McReduceImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::AnyBuffer>()
        .Ret<ffi::Buffer<ffi::F64>>());

#endif  // MF_HAVE_XLA_FFI
