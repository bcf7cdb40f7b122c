// tlib/detail/shape.h -- shape predicates and the output shape of a mode-q product.
// Restates the behaviour of bassoy/ttv detail/shape.h (cited per function); host-only integer code.
#pragma once

#include <algorithm>
#include <cassert>
#include <array>
#include <cstddef>
#include <iterator>
#include <stdexcept>
#include <vector>

namespace tlib::ttv::detail {

// a shape is valid when it has at least one mode and no extent is zero           (reference shape.h:30-34)
template<class InputIt>
inline bool is_valid_shape(InputIt first, InputIt last)
{
  if (first == last) return false;
  for (; first != last; ++first)
    if (*first == 0u) return false;
  return true;
}

namespace shape_impl {
template<class InputIt>
inline std::size_t count_greater_one(InputIt first, InputIt last)
{
  return static_cast<std::size_t>(std::count_if(first, last, [](auto const& n) { return n > 1u; }));
}
} // namespace shape_impl

// all extents equal one                                                          (reference shape.h:38-45)
template<class InputIt>
inline bool is_scalar(InputIt first, InputIt last)
{
  return is_valid_shape(first, last) && shape_impl::count_greater_one(first, last) == 0;
}

// exactly one of the first two extents exceeds one, every later extent is one     (reference shape.h:48-63)
template<class InputIt>
inline bool is_vector(InputIt first, InputIt last)
{
  if (!is_valid_shape(first, last)) return false;
  auto const p = std::distance(first, last);
  if (p == 1) return *first > 1u;
  return shape_impl::count_greater_one(first, first + 2) == 1 && shape_impl::count_greater_one(first + 2, last) == 0;
}

// the first two extents exceed one, every later extent is one                    (reference shape.h:65-76)
template<class InputIt>
inline bool is_matrix(InputIt first, InputIt last)
{
  if (!is_valid_shape(first, last) || std::distance(first, last) < 2) return false;
  return shape_impl::count_greater_one(first, first + 2) == 2 && shape_impl::count_greater_one(first + 2, last) == 0;
}

// order of at least three with some extent beyond the second exceeding one       (reference shape.h:79-89)
template<class InputIt>
inline bool is_tensor(InputIt first, InputIt last)
{
  if (!is_valid_shape(first, last) || std::distance(first, last) < 3) return false;
  return shape_impl::count_greater_one(first + 2, last) > 0;
}

// output shape of A x_q b: the input shape with entry q (one-based) erased       (reference shape.h:103-123)
template<class InputIt, class OutputIt, class SizeType>
inline void compute_output_shape(InputIt first, InputIt last, OutputIt out, SizeType q)
{
  if (!is_valid_shape(first, last))
    throw std::runtime_error("Error in tlib::detail::compute_output_shape(): input shape is not valid.");
  auto const p = static_cast<SizeType>(std::distance(first, last));
  if (q == 0u || q > p)
    throw std::runtime_error("Error in tlib::detail::compute_output_shape(): constraction mode q should be greater than 0 and less than or equal to the tensor order.");
  SizeType r = 1u;
  for (; first != last; ++first, ++r)
    if (r != q) *out++ = *first;
}

template<class SizeType, class ModeType>
inline auto generate_output_shape(std::vector<SizeType> const& input_shape, ModeType q)   // reference shape.h:126-141
{
  if (!is_valid_shape(input_shape.begin(), input_shape.end()))
    throw std::runtime_error("Error in tlib::detail::generate_output_shape(): input shape is not valid.");
  if (q == 0 || q > input_shape.size())
    throw std::runtime_error("Error in tlib::detail::generate_output_shape(): constraction mode q should be greater than 0 and less than or equal to the tensor order.");
  std::vector<SizeType> output_shape(input_shape.size() - 1);
  compute_output_shape(input_shape.begin(), input_shape.end(), output_shape.begin(), static_cast<std::size_t>(q));
  return output_shape;
}

template<class SizeType, class ModeType, std::size_t N>
inline auto generate_output_shape(std::array<SizeType, N> const& input_shape, ModeType q)   // reference shape.h:143-158
{
  static_assert(N > 0, "tlib::detail::generate_output_shape(): the order must be greater than zero.");
  if (!is_valid_shape(input_shape.begin(), input_shape.end()))
    throw std::runtime_error("Error in tlib::detail::generate_output_shape(): input shape is not valid.");
  if (q == 0 || q > N)
    throw std::runtime_error("Error in tlib::detail::generate_output_shape(): constraction mode q should be greater than 0 and less than or equal to the tensor order.");
  std::array<SizeType, N - 1> output_shape{};
  compute_output_shape(input_shape.begin(), input_shape.end(), output_shape.begin(), static_cast<std::size_t>(q));
  return output_shape;
}

} // namespace tlib::ttv::detail
