// XLA FFI handler shim: the symbols `jax.ffi.register_ffi_target` binds.
//
// NOT COMPILED IN THIS IMAGE: the XLA FFI headers (xla/ffi/api/ffi.h) ship inside jaxlib, which is
// neither installed nor installable here (no network, no wheel), so everything below is guarded
// by __has_include and the Makefile adds this file only when MF_XLA_INCLUDE points at the header
// tree.  The handlers are thin: they unpack XLA buffers into the plain C ABI of
// include/matfree_b200.h, take their scratch from XLA's ScratchAllocator and enqueue on XLA's
// stream -- no synchronisation, no allocation of their own, no retained pointers.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define MF_HAVE_XLA_FFI 1
#endif
#endif

#ifdef MF_HAVE_XLA_FFI
#include <cuda_runtime.h>

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

int32_t mf_dtype_of(ffi::DataType t) { return t == ffi::F64 ? MF_F64 : MF_F32; }

// jax.random.rademacher / normal for the (num, n) sample array of one key
// (matfree/stochtrace.py:957-964): key words are attributes (host values at trace time).
ffi::Error ProbeGenImpl(cudaStream_t stream, int64_t n, int64_t num, int32_t sampler,
                        int32_t x64_bits, uint32_t key0, uint32_t key1,
                        ffi::Result<ffi::AnyBuffer> out) {
  return to_error(mf_probe_gen(out->untyped_data(), mf_dtype_of(out->element_type()),
                               MF_LAYOUT_PROBE_MAJOR, n, n, 0, num, key0, key1, sampler,
                               x64_bits ? MF_PRNG_X64_BITS : 0, nullptr, stream));
}

// estimate(matvec, key) of matfree/stochtrace.py:47-50 with the SLQ integrand, CSR operator:
// one value per probe; the caller takes jnp.mean / jnp.std of the result.
ffi::Error SlqEstimateCsrImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                              ffi::Buffer<ffi::S32> indptr, ffi::Buffer<ffi::S32> indices,
                              ffi::AnyBuffer data, int64_t num_probes, int64_t p0,
                              int64_t num_matvecs, int32_t reortho, int32_t sampler, int32_t fn,
                              double fn_param, int64_t tile, uint32_t key0, uint32_t key1,
                              ffi::Result<ffi::AnyBuffer> quad) {
  mf_operator_t op{};
  op.kind = MF_OP_CSR;
  op.dtype = mf_dtype_of(data.element_type());
  op.n = (int64_t)indptr.element_count() - 1;
  op.nnz = (int64_t)data.element_count();
  op.values = data.untyped_data();
  op.indptr = indptr.typed_data();
  op.indices = indices.typed_data();
  const int64_t ws = mf_estimate_workspace_bytes(&op, tile, num_matvecs, reortho, MF_INTEGRAND_SLQ);
  if (ws < 0) return ffi::Error(ffi::ErrorCode::kInvalidArgument, mf_last_error());
  auto buf = scratch.Allocate((size_t)ws);
  if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "mf_estimate workspace");
  return to_error(mf_estimate(&op, MF_INTEGRAND_SLQ, sampler, 0, key0, key1, p0, num_probes, tile,
                              num_matvecs, reortho, fn, fn_param, quad->untyped_data(), nullptr,
                              nullptr, nullptr, *buf, ws, stream));
}

// decomp.tridiag_sym(...)(matvec, vec) for a CSR operator on a blocked start block [n][ld]
// (vmap_method="expand_dims" hands the whole probe block to one call).
ffi::Error LanczosCsrImpl(cudaStream_t stream, ffi::ScratchAllocator scratch,
                          ffi::Buffer<ffi::S32> indptr, ffi::Buffer<ffi::S32> indices,
                          ffi::AnyBuffer data, ffi::AnyBuffer v0_blocked, int64_t ld,
                          int64_t num_matvecs, int32_t reortho, ffi::Result<ffi::AnyBuffer> alphas,
                          ffi::Result<ffi::AnyBuffer> betas, ffi::Result<ffi::AnyBuffer> init_len,
                          ffi::Result<ffi::AnyBuffer> basis, ffi::Result<ffi::AnyBuffer> residual) {
  mf_operator_t op{};
  op.kind = MF_OP_CSR;
  op.dtype = mf_dtype_of(data.element_type());
  op.n = (int64_t)indptr.element_count() - 1;
  op.nnz = (int64_t)data.element_count();
  op.values = data.untyped_data();
  op.indptr = indptr.typed_data();
  op.indices = indices.typed_data();
  const int64_t ws = mf_lanczos_workspace_bytes(&op, ld, num_matvecs, reortho, 1);
  if (ws < 0) return ffi::Error(ffi::ErrorCode::kInvalidArgument, mf_last_error());
  auto buf = scratch.Allocate((size_t)ws);
  if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "mf_lanczos workspace");
  return to_error(mf_lanczos(&op, v0_blocked.untyped_data(), ld, num_matvecs, reortho,
                             alphas->untyped_data(), betas->untyped_data(),
                             init_len->untyped_data(), basis->untyped_data(),
                             residual->untyped_data(), *buf, ws, stream));
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
    mf_slq_estimate_csr_ffi, SlqEstimateCsrImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::AnyBuffer>()
        .Attr<int64_t>("num_probes")
        .Attr<int64_t>("p0")
        .Attr<int64_t>("num_matvecs")
        .Attr<int32_t>("reortho")
        .Attr<int32_t>("sampler")
        .Attr<int32_t>("fn")
        .Attr<double>("fn_param")
        .Attr<int64_t>("tile")
        .Attr<uint32_t>("key0")
        .Attr<uint32_t>("key1")
        .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    mf_lanczos_csr_ffi, LanczosCsrImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Ctx<ffi::ScratchAllocator>()
        .Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::AnyBuffer>()
        .Arg<ffi::AnyBuffer>()
        .Attr<int64_t>("ld")
        .Attr<int64_t>("num_matvecs")
        .Attr<int32_t>("reortho")
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>()
        .Ret<ffi::AnyBuffer>());

#endif  // MF_HAVE_XLA_FFI
