// Minimal stand-in for "xla/ffi/api/ffi.h", used ONLY to syntax- and type-check
// ffi/precond_ffi.cc in images without JAX (g++ -fsyntax-only -DPC_FFI_SYNTAX_CHECK).  It
// mirrors the subset of the public XLA FFI C++ API the adapter uses: typed buffers, result
// buffers, the platform stream context, attributes, the Bind() builder (arguments are passed
// to the handler in declaration order) and XLA_FFI_DEFINE_HANDLER_SYMBOL.  Nothing here is
// linked into a product.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { F32, S32, S16, S8, U8 };
template <DataType> struct NativeType;
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<S16> { using type = int16_t; };
template <> struct NativeType<S8> { using type = int8_t; };
template <> struct NativeType<U8> { using type = uint8_t; };

class Error {
 public:
  static Error Success() { return Error(); }
  static Error Internal(std::string) { return Error(); }
  static Error InvalidArgument(std::string) { return Error(); }
};

template <DataType dtype>
class Buffer {
 public:
  using T = typename NativeType<dtype>::type;
  T* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  const std::vector<int64_t>& dimensions() const { return dims_; }
  size_t element_count() const { return 0; }
 private:
  std::vector<int64_t> dims_;
};

template <typename B>
class Result {
 public:
  B* operator->() { return &b_; }
 private:
  B b_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

template <typename T> struct PlatformStream {};

template <typename... Ts>
struct Binding {
  template <typename C> struct CtxOf;
  template <typename T> struct CtxOf<PlatformStream<T>> { using type = T; };
  template <typename C> Binding<Ts..., typename CtxOf<C>::type> Ctx() const { return {}; }
  template <typename B> Binding<Ts..., B> Arg() const { return {}; }
  template <typename B> Binding<Ts..., Result<B>> Ret() const { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename F>
  static constexpr bool Accepts() { return std::is_invocable_r<Error, F, Ts...>::value; }
};
struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, fn, binding)                                       \
  static_assert(decltype(binding)::Accepts<decltype(&fn)>(),                                   \
                #name ": handler signature does not match its binding");                       \
  extern "C" { void* name = nullptr; }                                                          \
  static_assert(true, "")
