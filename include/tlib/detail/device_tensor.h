// tlib/detail/device_tensor.h -- a tensor whose elements LIVE in HBM, with the interface of tlib::ttv::tensor.
//
// Not in the reference (a CPU library: its container is a std::vector, detail/tensor.h:56-114).  Interfaces (1) and (2) of
// tlib/ttv.h exist for this type too:
//
//     tlib::ttv::device_tensor<float> dA(A), db(b);        // upload once (pageable memory is pipelined through pinned buffers)
//     auto dC1 = dA(1) * db;                               // operands and result stay on the device
//     auto dC2 = tlib::ttv::ttv(2, dA, db, ep, sp, fp);
//     tlib::ttv::tensor<float> C1 = dC1.to_host();
//
// which is how the reference's benchmark protocol -- every mode q = 1..p of one tensor (README.md:59-64) -- crosses PCIe
// once instead of p times.  Shape / layout / stride semantics are the host tensor's (same checks, same messages); the
// memory comes from the C-ABI (ttv_b200_device_alloc / ttv_b200_copy), so this header needs no CUDA runtime headers.
#pragma once

#include <cstddef>
#include <functional>
#include <numeric>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../ttv_b200.h"
#include "layout.h"
#include "shape.h"
#include "strides.h"
#include "tensor.h"

namespace tlib::ttv {

template<class value_t> class device_tensor;

// what dA(q) returns
template<class _value_t>
struct device_tensor_view {
  using value_t  = _value_t;
  using tensor_t = device_tensor<value_t>;

  device_tensor_view() = delete;
  device_tensor_view(device_tensor_view const&) = delete;
  device_tensor_view& operator=(device_tensor_view const&) = delete;

  tensor_t const& get_tensor() const { return _tensor; }
  std::size_t contraction_mode() const { return _q; }

private:
  friend class device_tensor<value_t>;
  device_tensor_view(tensor_t const& t, std::size_t q) : _tensor(t), _q(q) {}
  tensor_t const& _tensor;
  std::size_t _q;
};

template<class value_t>
class device_tensor {
public:
  using shape_t   = std::vector<std::size_t>;
  using layout_t  = std::vector<std::size_t>;
  using strides_t = std::vector<std::size_t>;

  device_tensor() = delete;

  // shape n, layout pi on `device` (-1: the current one); zero-initialised like the host tensor unless zero = false
  device_tensor(shape_t const& n, layout_t const& pi, int device = -1, bool zero = true) : _n(n), _pi(pi), _device(device)
  {
    if (n.size() != pi.size())
      throw std::runtime_error("Error in tlib::tensor: shape vector and layout vector must have the same length.");
    if (!detail::is_valid_shape(n.begin(), n.end()))
      throw std::runtime_error("Error in tlib::tensor: shape vector of tensor is not valid.");
    if (!detail::is_valid_layout(pi.begin(), pi.end()))
      throw std::runtime_error("Error in tlib::tensor: layout vector of tensor is not valid.");
    _count = std::accumulate(n.begin(), n.end(), std::size_t{1}, std::multiplies<>());
    void* ptr = nullptr;
    check(ttv_b200_device_alloc(&ptr, static_cast<std::uint64_t>(_count) * sizeof(value_t), device, zero ? 1 : 0));
    _data = static_cast<value_t*>(ptr);
  }

  // first-order layout by default, like tensor(shape)
  explicit device_tensor(shape_t const& n) : device_tensor(n, detail::generate_k_order_layout(n.size(), std::size_t{1})) {}

  // upload of a host tensor
  explicit device_tensor(tensor<value_t> const& host, int device = -1) : device_tensor(host.shape(), host.layout(), device, false)
  {
    check(ttv_b200_copy(_data, host.data().data(), static_cast<std::uint64_t>(_count) * sizeof(value_t), nullptr));
  }

  device_tensor(device_tensor const& other) : device_tensor(other._n, other._pi, other._device, false)
  {
    check(ttv_b200_copy(_data, other._data, static_cast<std::uint64_t>(_count) * sizeof(value_t), nullptr));
  }
  device_tensor(device_tensor&& other) noexcept
    : _n(std::move(other._n)), _pi(std::move(other._pi)), _data(other._data), _count(other._count), _device(other._device)
  {
    other._data = nullptr;
    other._count = 0;
  }
  device_tensor& operator=(device_tensor other) noexcept
  {
    std::swap(_n, other._n); std::swap(_pi, other._pi); std::swap(_data, other._data);
    std::swap(_count, other._count); std::swap(_device, other._device);
    return *this;
  }
  ~device_tensor() { if (_data) ttv_b200_device_free(_data); }

  // overwrite from / copy to host memory
  void upload(tensor<value_t> const& host)
  {
    if (host.shape() != _n || host.layout() != _pi)
      throw std::runtime_error("Error in tlib::device_tensor: shape and layout of the host tensor differ from the device tensor's.");
    check(ttv_b200_copy(_data, host.data().data(), static_cast<std::uint64_t>(_count) * sizeof(value_t), nullptr));
  }
  tensor<value_t> to_host() const
  {
    tensor<value_t> host(_n, _pi);
    check(ttv_b200_copy(host.data().data(), _data, static_cast<std::uint64_t>(_count) * sizeof(value_t), nullptr));
    return host;
  }

  // dA(q): pairs the tensor with a contraction mode, 1 <= q <= order
  device_tensor_view<value_t> operator()(std::size_t contraction_mode) const
  {
    if (contraction_mode < 1ul || contraction_mode > order())
      throw std::runtime_error("Error in tlib::tensor: specified contraction mode should be greater than one and equal to or less than the order.");
    return device_tensor_view<value_t>(*this, contraction_mode);
  }

  value_t const*  data()   const { return _data; }      // DEVICE pointers
  value_t*        data()         { return _data; }
  std::size_t     size()   const { return _count; }
  shape_t const&  shape()  const { return _n; }
  layout_t const& layout() const { return _pi; }
  strides_t       strides() const { return detail::generate_strides(_n, _pi); }
  std::size_t     order()  const { return _n.size(); }
  int             device() const { return _device; }

private:
  static void check(int status)
  {
    if (status == TTV_B200_OK) return;
    char const* text = ttv_b200_last_error();
    throw std::runtime_error((text && *text) ? std::string(text) : std::string(ttv_b200_strerror(status)));
  }

  shape_t     _n;
  layout_t    _pi;
  value_t*    _data = nullptr;
  std::size_t _count = 0;
  int         _device = -1;
};

} // namespace tlib::ttv
