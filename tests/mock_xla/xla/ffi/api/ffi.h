/* Minimal stand-in for jaxlib's xla/ffi/api/ffi.h (the typed XLA FFI binding API, jaxlib >= 0.4.31), restricted to the
 * declarations enf_pde_b200/csrc/enf_xla_ffi.cc uses, with the real API's names, template shapes and member signatures:
 *   ffi::Error / ffi::ErrorCode, ffi::DataType constants F32 / U8, ffi::Buffer<dtype>, ffi::Result<T> / ffi::ResultBuffer,
 *   ffi::RemainingArgs / ffi::RemainingRets (get<T>(i) -> ErrorOr<...>), ffi::PlatformStream<T>, ffi::Ffi::Bind() with
 *   .Ctx / .Arg / .Ret / .Attr / .RemainingArgs / .RemainingRets, XLA_FFI_DEFINE_HANDLER_SYMBOL.
 * It exists so that the shim is type-checked (g++ -fsyntax-only) on every test run in an image that cannot install jaxlib;
 * it is NOT a substitute for compiling against the real headers. */
#ifndef ENF_MOCK_XLA_FFI_H_
#define ENF_MOCK_XLA_FFI_H_
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "xla/ffi/api/c_api.h"

namespace xla::ffi {

enum class ErrorCode { kOk, kInvalidArgument, kUnimplemented, kInternal };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool failure() const { return code_ != ErrorCode::kOk; }
  bool success() const { return code_ == ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

enum class DataType { U8, F32 };
inline constexpr DataType U8 = DataType::U8;
inline constexpr DataType F32 = DataType::F32;
template <DataType> struct NativeOf;
template <> struct NativeOf<DataType::U8> { using type = uint8_t; };
template <> struct NativeOf<DataType::F32> { using type = float; };

template <typename T> class Span {
 public:
  Span(const T* p, size_t n) : p_(p), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
 private:
  const T* p_; size_t n_;
};

template <DataType dtype> class Buffer {
 public:
  using T = typename NativeOf<dtype>::type;
  T* typed_data() const { return data_; }
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_.data(), dims_.size()); }
  size_t element_count() const { size_t n = 1; for (auto d : dims_) n *= (size_t)d; return n; }
 private:
  T* data_ = nullptr;
  std::vector<int64_t> dims_;
};

template <typename T> class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }
 private:
  T v_;
};
template <DataType dtype> using ResultBuffer = Result<Buffer<dtype>>;

template <typename T> class ErrorOr {
 public:
  bool has_value() const { return v_.has_value(); }
  T& operator*() { return *v_; }
  T* operator->() { return &*v_; }
  T& value() { return *v_; }
 private:
  std::optional<T> v_;
};

class RemainingArgs {
 public:
  size_t size() const { return 0; }
  template <typename T> ErrorOr<T> get(size_t) const { return ErrorOr<T>(); }
};
class RemainingRets {
 public:
  size_t size() const { return 0; }
  template <typename T> ErrorOr<Result<T>> get(size_t) const { return ErrorOr<Result<T>>(); }
};

template <typename T> struct PlatformStream {};

// the binding builder: every step returns the builder; To() type-checks nothing about arity here (the real one does)
class Binding {
 public:
  template <typename T> Binding& Ctx() { return *this; }
  template <typename T> Binding& Arg() { return *this; }
  template <typename T> Binding& Ret() { return *this; }
  template <typename T> Binding& Attr(const char*) { return *this; }
  Binding& RemainingArgs() { return *this; }
  Binding& RemainingRets() { return *this; }
};
struct Ffi { static Binding Bind() { return Binding(); } };

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) { (void)&impl; (void)(binding); return nullptr; }

#endif  // ENF_MOCK_XLA_FFI_H_
