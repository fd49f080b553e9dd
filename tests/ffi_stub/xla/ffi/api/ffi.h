// TEST INFRASTRUCTURE -- a minimal stand-in for the part of XLA's FFI C++ API
// (xla/ffi/api/ffi.h, shipped inside jaxlib, which is not installable in this image) that
// matfree_b200/csrc/ffi_xla.cc uses.  It exists so that the CPU test suite can COMPILE the shim:
// `Bind()...To(impl)` checks at compile time that every handler is invocable with exactly the
// context / argument / attribute / result types its binding declares, and every mf_* call is
// checked against include/matfree_b200.h.  It executes nothing.
#pragma once

#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <type_traits>
#include <vector>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla {
namespace ffi {

enum class DataType { F32, F64, S32, U32 };
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType F64 = DataType::F64;
inline constexpr DataType S32 = DataType::S32;
inline constexpr DataType U32 = DataType::U32;

enum class ErrorCode { kInvalidArgument, kInternal, kResourceExhausted };

class Error {
 public:
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};

template <typename T>
struct Span {
  const T* ptr = nullptr;
  size_t len = 0;
  size_t size() const { return len; }
  const T& operator[](size_t i) const { return ptr[i]; }
};

class AnyBuffer {
 public:
  void* untyped_data() const { return nullptr; }
  DataType element_type() const { return DataType::F32; }
  size_t element_count() const { return 0; }
  Span<int64_t> dimensions() const { return {}; }
};

template <DataType dtype>
struct NativeOf;
template <> struct NativeOf<DataType::F32> { using type = float; };
template <> struct NativeOf<DataType::F64> { using type = double; };
template <> struct NativeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeOf<DataType::U32> { using type = uint32_t; };

template <DataType dtype>
class Buffer {
 public:
  using T = typename NativeOf<dtype>::type;
  T* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  size_t element_count() const { return 0; }
  Span<int64_t> dimensions() const { return {}; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};

class ScratchAllocator {
 public:
  std::optional<void*> Allocate(size_t) { return std::nullopt; }
};

template <typename T>
struct PlatformStream {
  using Type = T;
};

namespace internal {
template <typename T>
struct CtxType {
  using type = T;
};
template <typename T>
struct CtxType<PlatformStream<T>> {
  using type = T;
};

template <typename... Ts>
struct Binding {
  template <typename T>
  Binding<Ts..., typename CtxType<T>::type> Ctx() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Arg() const { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename F>
  int To(F) const {
    static_assert(std::is_invocable_r_v<Error, F, Ts...>,
                  "FFI handler signature does not match its binding");
    return 0;
  }
};
}  // namespace internal

struct Ffi {
  static internal::Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)      \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) {          \
    static const int bound = (binding).To(impl);                \
    (void)bound;                                                \
    return nullptr;                                             \
  }
