// MINIMAL STAND-IN for XLA's header-only FFI API ("xla/ffi/api/ffi.h", shipped by jaxlib under jax.ffi.include_dir()),
// which is absent from this image (no jax).  It declares just the names csrc/xla_ffi_shim.cc uses, with the same
// shapes, so that the shim is compiled -- and its handler signatures are type-checked against their bindings -- by
// tests/test_host_cpu.py::test_xla_ffi_shim_compiles.  It is NOT the real API and nothing links against it; with
// jaxlib present build the shim with -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") instead.
#pragma once
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>

#define NSB200_XLA_FFI_STUB 1

namespace xla::ffi {

enum DataType { PRED, U8, U32, U64, S32, S64, F64 };
template <DataType dt> struct NativeOf;
template <> struct NativeOf<PRED> { using type = bool; };
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<U32> { using type = uint32_t; };
template <> struct NativeOf<U64> { using type = uint64_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<S64> { using type = int64_t; };
template <> struct NativeOf<F64> { using type = double; };

template <typename T>
struct Span {
    const T *ptr = nullptr;
    size_t n = 0;
    size_t size() const { return n; }
    const T &operator[](size_t i) const { return ptr[i]; }
    const T *begin() const { return ptr; }
    const T *end() const { return ptr + n; }
};

template <DataType dt>
struct Buffer {
    using T = typename NativeOf<dt>::type;
    T *data = nullptr;
    Span<int64_t> dims;
    T *typed_data() const { return data; }
    Span<int64_t> dimensions() const { return dims; }
    size_t element_count() const {
        size_t c = 1;
        for (int64_t d : dims) c *= (size_t) d;
        return c;
    }
};

template <typename T>
struct Result {
    T value;
    T *operator->() { return &value; }
    T &operator*() { return value; }
};
template <DataType dt>
using ResultBuffer = Result<Buffer<dt>>;

template <typename StreamT>
struct PlatformStream {};

enum class ErrorCode { kOk, kInvalidArgument, kInternal };
class Error {
  public:
    Error() = default;
    Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
    static Error Success() { return Error(); }
    static Error Internal(std::string message) { return Error(ErrorCode::kInternal, std::move(message)); }
    static Error InvalidArgument(std::string message) { return Error(ErrorCode::kInvalidArgument, std::move(message)); }
    bool success() const { return code_ == ErrorCode::kOk; }
  private:
    ErrorCode code_ = ErrorCode::kOk;
    std::string message_;
};

// Binding<Args...>: the argument list a handler must accept, accumulated by Ctx / Arg / Ret / Attr.
template <typename... Ts>
struct Binding {
    template <typename StreamT>
    Binding<Ts..., StreamT> CtxStream() const { return {}; }
    template <typename C>
    auto Ctx() const { return CtxImpl(static_cast<C *>(nullptr)); }
    template <typename StreamT>
    Binding<Ts..., StreamT> CtxImpl(PlatformStream<StreamT> *) const { return {}; }
    template <typename A>
    Binding<Ts..., A> Arg() const { return {}; }
    template <typename R>
    Binding<Ts..., Result<R>> Ret() const { return {}; }
    template <typename A>
    Binding<Ts..., A> Attr(const char *) const { return {}; }
    // what the real API does when the handler is registered: the callable has to be invocable with exactly these types
    template <typename Fn>
    static constexpr bool Accepts() { return std::is_invocable_r_v<Error, Fn, Ts...>; }
};

struct Ffi {
    static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                      \
    static_assert(decltype(binding)::template Accepts<decltype(&impl)>(), #impl " does not match its FFI binding"); \
    extern "C" XLA_FFI_Error *name(XLA_FFI_CallFrame *) { return nullptr; }
