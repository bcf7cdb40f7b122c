// tlib/detail/strides.h -- stride tuples of packed tensors with a non-hierarchical layout.
// Restates the behaviour of bassoy/ttv detail/strides.h (cited per function); host-only integer code.
#pragma once

#include <algorithm>
#include <cassert>
#include <array>
#include <cstddef>
#include <iterator>
#include <stdexcept>
#include <vector>

#include "layout.h"
#include "shape.h"

namespace tlib::ttv::detail {

// w[pi_1] = 1, w[pi_r] = w[pi_{r-1}] * n[pi_{r-1}].  Scalar- and vector-shaped tensors keep all-one strides
// (reference strides.h:31-57, the early return at :42-45).
template<class InputIt1, class InputIt2, class OutputIt>
inline void compute_strides(InputIt1 shape_begin, InputIt1 shape_end, InputIt2 layout_begin, OutputIt strides_begin)
{
  if (!is_valid_shape(shape_begin, shape_end))
    throw std::runtime_error("Error in tlib::detail::compute_strides(): input shape is not valid.");
  auto const p = std::distance(shape_begin, shape_end);
  if (!is_valid_layout(layout_begin, layout_begin + p))
    throw std::runtime_error("Error in tlib::detail::compute_strides(): input layout is not valid.");

  std::fill_n(strides_begin, p, 1u);
  if (!is_matrix(shape_begin, shape_end) && !is_tensor(shape_begin, shape_end)) return;   // scalar or vector

  for (auto r = decltype(p){1}; r < p; ++r) {
    auto const slower = layout_begin[r] - 1;
    auto const faster = layout_begin[r - 1] - 1;
    strides_begin[slower] = strides_begin[faster] * shape_begin[faster];
  }
}

template<class size_type>
inline auto generate_strides(std::vector<size_type> const& shape, std::vector<size_type> const& layout)   // reference strides.h:59-65
{
  std::vector<size_type> strides(shape.size());
  compute_strides(shape.begin(), shape.end(), layout.begin(), strides.begin());
  return strides;
}

template<class size_type, std::size_t N>
inline auto generate_strides(std::array<size_type, N> const& shape, std::array<size_type, N> const& layout)   // reference strides.h:67-74
{
  static_assert(N > 0, "Static error in tlib::detail::generate_strides(): N, i.e. length of array should be greater than zero.");
  std::array<size_type, N> strides{};
  compute_strides(shape.begin(), shape.end(), layout.begin(), strides.begin());
  return strides;
}

// strides are valid when they never decrease along the layout order              (reference strides.h:76-101)
template<class InputIt1, class InputIt2>
inline bool is_valid_strides(InputIt1 layout_begin, InputIt1 layout_end, InputIt2 stride_begin)
{
  if (!is_valid_layout(layout_begin, layout_end))
    throw std::runtime_error("Error in tlib::detail::is_valid_strides(): input layout is not valid.");
  return std::adjacent_find(layout_begin, layout_end, [stride_begin](auto faster, auto slower) {
           return stride_begin[faster - 1] > stride_begin[slower - 1];
         }) == layout_end;
}

} // namespace tlib::ttv::detail
