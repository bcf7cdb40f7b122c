// tlib/detail/tensor.h -- minimal owning host tensor and the (tensor, mode) view behind `A(q) * b`.
// Restates the container of bassoy/ttv detail/tensor.h:36-114: shape + layout + zero-initialised std::vector.
// The data lives in HOST memory; the product stages it through the device inside the C-ABI call.
#pragma once

#include <algorithm>
#include <cstddef>
#include <functional>
#include <numeric>
#include <stdexcept>
#include <vector>

#include "layout.h"
#include "shape.h"
#include "strides.h"

namespace tlib::ttv {

template<class value_t> class tensor;

// what operator()(q) returns: a tensor together with a contraction mode               (reference tensor.h:36-52)
template<class _value_t>
struct tensor_view {
  using value_t  = _value_t;
  using tensor_t = tensor<value_t>;

  tensor_view() = delete;
  tensor_view(tensor_view const&) = delete;
  tensor_view& operator=(tensor_view const&) = delete;

  tensor_t const& get_tensor() const { return _tensor; }
  std::size_t contraction_mode() const { return _q; }

private:
  friend class tensor<value_t>;
  tensor_view(tensor_t const& t, std::size_t q) : _tensor(t), _q(q) {}
  tensor_t const& _tensor;
  std::size_t _q;
};

template<class value_t>
class tensor {
public:
  using shape_t   = std::vector<std::size_t>;
  using layout_t  = std::vector<std::size_t>;
  using strides_t = std::vector<std::size_t>;
  using vector_t  = std::vector<value_t>;

  tensor() = delete;

  // shape n, layout pi; all elements value-initialised                               (reference tensor.h:66-78)
  tensor(shape_t const& n, layout_t const& pi)
    : _n(n), _pi(pi), _data(std::accumulate(n.begin(), n.end(), std::size_t{1}, std::multiplies<>()))
  {
    if (n.size() != pi.size())
      throw std::runtime_error("Error in tlib::tensor: shape vector and layout vector must have the same length.");
    if (!detail::is_valid_shape(n.begin(), n.end()))
      throw std::runtime_error("Error in tlib::tensor: shape vector of tensor is not valid.");
    if (!detail::is_valid_layout(pi.begin(), pi.end()))
      throw std::runtime_error("Error in tlib::tensor: layout vector of tensor is not valid.");
  }

  // first-order layout by default                                                    (reference tensor.h:80-83)
  tensor(shape_t const& n) : tensor(n, detail::generate_k_order_layout(n.size(), std::size_t{1})) {}

  // fill with one value (the reference's version, tensor.h:85-88, forgets its return statement)
  tensor& operator=(value_t v)
  {
    std::fill(_data.begin(), _data.end(), v);
    return *this;
  }

  // A(q): pairs the tensor with a contraction mode, 1 <= q <= order                  (reference tensor.h:90-95)
  tensor_view<value_t> operator()(std::size_t contraction_mode) const
  {
    if (contraction_mode < 1ul || contraction_mode > order())
      throw std::runtime_error("Error in tlib::tensor: specified contraction mode should be greater than one and equal to or less than the order.");
    return tensor_view<value_t>(*this, contraction_mode);
  }

  auto begin() const { return _data.begin(); }
  auto end()   const { return _data.end(); }
  auto begin()       { return _data.begin(); }
  auto end()         { return _data.end(); }

  vector_t const& data()   const { return _data; }
  vector_t&       data()         { return _data; }
  shape_t const&  shape()  const { return _n; }
  layout_t const& layout() const { return _pi; }
  strides_t       strides() const { return detail::generate_strides(_n, _pi); }   // computed on demand (reference tensor.h:107)
  std::size_t     order()  const { return _n.size(); }

private:
  shape_t  _n;
  layout_t _pi;
  vector_t _data;
};

} // namespace tlib::ttv
