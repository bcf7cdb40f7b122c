// tlib/detail/layout.h -- layout tuples (permutations of the modes, fastest mode first, one-based).
// Restates the behaviour of bassoy/ttv detail/layout.h (cited per function); host-only integer code.
#pragma once

#include <algorithm>
#include <cassert>
#include <array>
#include <cstddef>
#include <iterator>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "shape.h"

namespace tlib::ttv::detail {

// a layout is valid when it is a permutation of 1..p, p > 0                      (reference layout.h:29-55)
template<class InputIt>
inline bool is_valid_layout(InputIt first, InputIt last)
{
  auto const len = std::distance(first, last);
  if (len <= 0) return false;
  auto const p = static_cast<std::size_t>(len);
  for (auto it = first; it != last; ++it) {
    auto const mode = *it;
    if (mode == 0u || static_cast<std::size_t>(mode) > p) return false;
    if (std::find(std::next(it), last, mode) != last) return false;   // repeated mode
  }
  return true;
}

// k-order layout (k, k-1, ..., 1, k+1, ..., p); k == 0 or k > p gives the last-order layout (p, ..., 1)
// (reference layout.h:57-76)
template<class OutputIt, class size_t>
inline void compute_k_order_layout(OutputIt first, OutputIt last, size_t k)
{
  auto const len = std::distance(first, last);
  if (len <= 0)
    throw std::runtime_error("Error in tlib::detail::compute_k_order: range provided by begin and end not correct!");
  auto const p = static_cast<std::size_t>(len);
  std::size_t const turn = (k == 0u || static_cast<std::size_t>(k) > p) ? p : static_cast<std::size_t>(k);
  std::size_t r = 0;
  for (; r < turn; ++r, ++first) *first = turn - r;     // descending head
  for (; r < p;    ++r, ++first) *first = r + 1;        // ascending tail
}

template<class size_t>
inline auto generate_k_order_layout(size_t p, size_t k)                             // reference layout.h:78-84
{
  std::vector<size_t> layout(p);
  compute_k_order_layout(layout.begin(), layout.end(), k);
  return layout;
}

template<class OutputIt>
inline void compute_first_order_layout(OutputIt first, OutputIt last) { compute_k_order_layout(first, last, 1u); }   // reference layout.h:87-91

template<class OutputIt>
inline void compute_last_order_layout(OutputIt first, OutputIt last) { compute_k_order_layout(first, last, 0u); }    // reference layout.h:93-97

// inverse permutation: out[pi_r - 1] = r                                          (reference layout.h:99-109)
template<class InputIt, class OutputIt>
inline void compute_inverse_layout(InputIt first, InputIt last, OutputIt out)
{
  if (!is_valid_layout(first, last))
    throw std::runtime_error("Error in tlib::detail::compute_inverse_layout: input layout is not valid!");
  unsigned position = 1u;
  for (; first != last; ++first, ++position) out[*first - 1] = position;
}

// one-based position of `mode` inside the layout tuple, i.e. pi^-1(mode)           (reference layout.h:114-138)
template<class InputIt, class SizeType>
inline auto inverse_mode(InputIt first, InputIt last, SizeType mode)
{
  using value_type = typename std::iterator_traits<InputIt>::value_type;
  if (!is_valid_layout(first, last))
    throw std::runtime_error("Error in tlib::detail::inverse_mode(): input layout is not valid.");
  auto const p = static_cast<value_type>(std::distance(first, last));
  if (mode == 0u || mode > SizeType(p))
    throw std::runtime_error("Error in tlib::detail::inverse_mode(): mode should be one-based and equal to or less than layout size.");
  auto const hit = std::find(first, last, value_type(mode));
  return static_cast<value_type>(std::distance(first, hit)) + value_type(1);
}

// layout of C = A x_q b: q erased from the tuple, every mode above q renumbered     (reference layout.h:143-172)
template<class InputIt, class OutputIt, class ModeType>
inline void compute_output_layout(InputIt first, InputIt last, OutputIt out, ModeType q)
{
  using value_type = typename std::iterator_traits<InputIt>::value_type;
  if (!is_valid_layout(first, last))
    throw std::runtime_error("Error in tlib::detail::compute_inverse_layout: input layout is not valid!");
  auto const p = static_cast<value_type>(std::distance(first, last));
  if (1u > q || value_type(q) > p)
    throw std::runtime_error("Error in tlib::detail::compute_inverse_layout: mode must be greater zero and less than or equal to the order!");
  for (; first != last; ++first) {
    value_type const mode = *first;
    if (mode == value_type(q)) continue;
    *out++ = mode > value_type(q) ? mode - 1 : mode;
  }
}

template<class SizeType, class ModeType>
inline auto generate_output_layout(std::vector<SizeType> const& input_layout, ModeType q)   // reference layout.h:175-190
{
  if (!is_valid_layout(input_layout.begin(), input_layout.end()))
    throw std::runtime_error("Error in tlib::detail::generate_output_layout(): input layout is not valid.");
  if (q == 0 || q > input_layout.size())
    throw std::runtime_error("Error in tlib::detail::generate_output_layout(): constraction mode q should be greater than 0 and less than or equal to the tensor order.");
  std::vector<SizeType> output_layout(input_layout.size() - 1);
  compute_output_layout(input_layout.begin(), input_layout.end(), output_layout.begin(), q);
  return output_layout;
}

template<class SizeType, class ModeType, std::size_t N>
inline auto generate_output_layout(std::array<SizeType, N> const& input_layout, ModeType q)   // reference layout.h:192-207
{
  static_assert(N > 0, "tlib::detail::generate_output_layout(): the order must be greater than zero.");
  if (!is_valid_layout(input_layout.begin(), input_layout.end()))
    throw std::runtime_error("Error in tlib::detail::generate_output_layout(): input layout is not valid.");
  if (q == 0 || q > N)
    throw std::runtime_error("Error in tlib::detail::generate_output_layout(): constraction mode q should be greater than 0 and less than or equal to the tensor order.");
  std::array<SizeType, N - 1> output_layout{};
  compute_output_layout(input_layout.begin(), input_layout.end(), output_layout.begin(), q);
  return output_layout;
}

} // namespace tlib::ttv::detail
