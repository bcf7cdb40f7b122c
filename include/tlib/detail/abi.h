// tlib/detail/abi.h -- glue between the header-style C++17 interface and the C-ABI shim (include/ttv_b200.h).
// Not part of the reference; everything the templates need to reach libttv_b200.so lives here.
#pragma once

#include <complex>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../ttv_b200.h"

namespace tlib::ttv::detail::abi {

// element type -> enum ttv_b200_dtype.  Unsigned integers share the signed entry: wrap-around arithmetic is the same bits.
template<class value_t, class = void> struct dtype_of { static constexpr int value = -1; };
template<> struct dtype_of<float>                { static constexpr int value = TTV_B200_F32; };
template<> struct dtype_of<double>               { static constexpr int value = TTV_B200_F64; };
template<> struct dtype_of<std::complex<float>>  { static constexpr int value = TTV_B200_C64; };
template<> struct dtype_of<std::complex<double>> { static constexpr int value = TTV_B200_C128; };
template<class value_t>
struct dtype_of<value_t, std::enable_if_t<std::is_integral_v<value_t> && sizeof(value_t) == 4>> { static constexpr int value = TTV_B200_I32; };
template<class value_t>
struct dtype_of<value_t, std::enable_if_t<std::is_integral_v<value_t> && sizeof(value_t) == 8>> { static constexpr int value = TTV_B200_I64; };

template<class value_t>
constexpr int dtype_v = dtype_of<std::remove_cv_t<value_t>>::value;

template<class value_t>
constexpr void require_supported()
{
  static_assert(dtype_v<value_t> >= 0,
                "tlib::ttv (B200): value_t must be float, double, std::complex<float>, std::complex<double> or a 32/64-bit integer");
}

// any integral tuple type -> the uint64_t tuples of the C-ABI (null stays null so that the null checks still fire)
template<class size_t_>
struct tuple64 {
  std::vector<std::uint64_t> storage;
  bool null;
  tuple64(size_t_ const* v, std::size_t len) : storage(v ? len : 0), null(v == nullptr)
  {
    for (std::size_t r = 0; r < storage.size(); ++r) storage[r] = static_cast<std::uint64_t>(v[r]);
    if (!null && storage.empty()) storage.resize(1);
  }
  std::uint64_t const* get() const { return null ? nullptr : storage.data(); }
};

[[noreturn]] inline void raise(int status)
{
  char const* text = ttv_b200_last_error();
  throw std::runtime_error((text && *text) ? std::string(text) : std::string(ttv_b200_strerror(status)));
}

template<class execution_t, class slicing_t, class fusion_t>
inline ttv_b200_opts make_opts(unsigned flags = 0u)
{
  ttv_b200_opts o{};
  o.device = -1;
  o.execution = execution_t::code;
  o.slicing   = slicing_t::code;
  o.fusion    = fusion_t::code;
  o.flags     = flags;
  return o;
}

// C[outer][inner] (=|+=) sum_k A[outer][k][inner] b[k] on packed host or device buffers (the canonical view)
template<class value_t>
inline void view(std::uint64_t outer, std::uint64_t nq, std::uint64_t inner, value_t const* a, value_t const* b, value_t* c,
                 bool accumulate)
{
  require_supported<value_t>();
  ttv_b200_opts o{};
  o.device = -1;
  o.flags  = accumulate ? static_cast<unsigned>(TTV_B200_FLAG_ACCUMULATE) : 0u;
  if (int st = ttv_b200_view(dtype_v<value_t>, outer, nq, inner, a, b, c, &o)) raise(st);
}

} // namespace tlib::ttv::detail::abi
