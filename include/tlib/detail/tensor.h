// tlib/detail/tensor.h -- minimal owning host tensor and the (tensor, mode) view behind `A(q) * b`.
// Restates the container of bassoy/ttv detail/tensor.h:36-114: shape + layout + zero-initialised std::vector.
// The data lives in HOST memory; the product stages it through the device inside the C-ABI call.
//
// Not in the reference: A.keep_on_device(true).  The reference's benchmark protocol contracts every mode of one tensor
// (README.md:59-64); as a drop-in every `A(q) * b` would move all of A across PCIe again.  A tensor that was asked to keep a
// device copy uploads on its first product (streamed under that product's kernels) and serves the following ones from HBM.
// The copy is dropped whenever the tensor hands out MUTABLE access (non-const begin / end / data, operator=), so code that
// writes through the container's own accessors stays correct; code that keeps a raw pointer from an earlier data().data()
// and writes through it later must call A.host_data_changed() -- which is why this is opt-in and not the default
// (TTV_B200_RESIDENT=1 in the environment opts every tensor in).  For data that should LIVE on the device see
// device_tensor.h.
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <functional>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <vector>

#include "../../ttv_b200.h"
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

  // a copy owns its own data and therefore its own (not yet made) device copy
  tensor(tensor const& other) : _n(other._n), _pi(other._pi), _data(other._data), _keep(other._keep) {}
  tensor(tensor&& other) noexcept = default;
  tensor& operator=(tensor const& other)
  {
    if (this != &other) { _n = other._n; _pi = other._pi; _data = other._data; _keep = other._keep; _twin.reset(); }
    return *this;
  }
  tensor& operator=(tensor&& other) noexcept = default;

  // fill with one value (the reference's version, tensor.h:85-88, forgets its return statement)
  tensor& operator=(value_t v)
  {
    host_data_changed();
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
  auto begin()       { host_data_changed(); return _data.begin(); }
  auto end()         { host_data_changed(); return _data.end(); }

  vector_t const& data()   const { return _data; }
  vector_t&       data()         { host_data_changed(); return _data; }
  shape_t const&  shape()  const { return _n; }
  layout_t const& layout() const { return _pi; }
  strides_t       strides() const { return detail::generate_strides(_n, _pi); }   // computed on demand (reference tensor.h:107)
  std::size_t     order()  const { return _n.size(); }

  // ---- device residency (not in the reference; see the top of this file) -------------------------------------------
  tensor& keep_on_device(bool on = true)
  {
    _keep = on;
    if (!on) _twin.reset();
    return *this;
  }
  bool kept_on_device() const { return _keep || env_opt_in(); }
  // the host data was written behind the container's back: the next product uploads again
  void host_data_changed() const { if (_twin) ttv_b200_resident_invalidate(_twin.get()); }
  // the device twin, created on first use; nullptr when the tensor is not kept on the device
  ttv_b200_resident* device_twin() const
  {
    if (!kept_on_device()) return nullptr;
    if (!_twin) {
      ttv_b200_resident* r = nullptr;
      if (ttv_b200_resident_create(&r, -1) != TTV_B200_OK) return nullptr;      // (no device: the product itself reports it)
      _twin = std::shared_ptr<ttv_b200_resident>(r, [](ttv_b200_resident* x) { ttv_b200_resident_destroy(x); });
    }
    return _twin.get();
  }

private:
  static bool env_opt_in()
  {
    static bool const on = [] { char const* e = std::getenv("TTV_B200_RESIDENT"); return e && *e && *e != '0'; }();
    return on;
  }

  shape_t  _n;
  layout_t _pi;
  vector_t _data;
  bool     _keep = false;
  mutable std::shared_ptr<ttv_b200_resident> _twin;     // moves with the tensor (the vector's buffer moves too), never copied
};

} // namespace tlib::ttv
