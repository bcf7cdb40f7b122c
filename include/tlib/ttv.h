// tlib/ttv.h -- mode-q tensor-times-vector product C = A x_q b, B200-native drop-in for bassoy/ttv's include/tlib/ttv.h.
//
// Three interfaces, same names / argument order / error texts as the reference:
//   (1) auto C = A(q) * b;                                               operator*            reference ttv.h:122-127
//   (2) auto C = tlib::ttv::ttv(q, A, b, ep, sp, fp);                    tensor-level         reference ttv.h:99-114
//   (3) tlib::ttv::ttv(ep, sp, fp, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);   C-like     reference ttv.h:54-92
// (1) and (2) also exist on tlib::ttv::device_tensor (detail/device_tensor.h): operands and result stay in HBM; and a host
// tensor can keep a device copy between products (tensor::keep_on_device, detail/tensor.h).
//
// The arithmetic runs on the GPU (hand-written sm_100a kernels behind the C-ABI of include/ttv_b200.h; link with
// -lttv_b200).  a, b, c may be host pointers (staged inside the call) or device pointers (used in place).  There is no
// CPU fallback.  C is overwritten.
#pragma once

#include <cstddef>
#include <stdexcept>

#include "detail/tags.h"
#include "detail/tensor.h"
#include "detail/tensor_times_vector.h"
#include "detail/device_tensor.h"

namespace tlib::ttv {

/** Mode-q tensor-times-vector product, C-like interface.
 *
 * @tparam value_t      float, double, std::complex<float|double>, or a 32/64-bit integer
 * @tparam size_t       integral type of the extents, strides and layout elements (usually std::size_t)
 * @tparam execution_t  execution_policy tag, slicing_t slicing_policy tag, fusion_t fusion_policy tag: accepted as hints
 *
 * @param q    contraction mode, 1 <= q <= p              @param p    order of A, p >= 2
 * @param a    A (packed)                                 @param na, wa, pia   extents, strides, layout of A (length p)
 * @param b    b (unit stride)                            @param nb   extent of b (length 1), nb[0] == na[q-1]
 * @param c    C (packed, overwritten)                    @param nc, wc, pic   extents, strides, layout of C (length p-1)
 *
 * Throws std::runtime_error with the reference's message for every invalid argument, checked in the reference's
 * order (reference ttv.h:64-89); the checks themselves live behind the C-ABI so that every binding shares them.
 */
template<class value_t, class size_t, class execution_t, class slicing_t, class fusion_t>
inline void ttv(execution_t ep, slicing_t sp, fusion_t fp,
                size_t const q, size_t const p,
                value_t const* const a, size_t const* const na, size_t const* const wa, size_t const* const pia,
                value_t const* const b, size_t const* const nb,
                value_t* const c, size_t const* const nc, size_t const* const wc, size_t const* const pic)
{
  detail::ttv(ep, sp, fp, static_cast<unsigned>(q), static_cast<unsigned>(p), a, na, wa, pia, b, nb, c, nc, wc, pic);
}

/** Tensor-level interface: allocates C with the output shape / layout of (A, q) and returns it.   reference ttv.h:99-114 */
template<class value_t, class execution_t, class slicing_t, class fusion_t>
inline auto ttv(std::size_t q, tensor<value_t> const& a, tensor<value_t> const& b, execution_t ep, slicing_t sp, fusion_t fp)
{
  auto c = tensor<value_t>(detail::generate_output_shape(a.shape(), q), detail::generate_output_layout(a.layout(), q));
  auto const wa = a.strides();
  auto const wc = c.strides();
  if (ttv_b200_resident* twin = a.device_twin()) {
    // A keeps its copy in HBM (tensor::keep_on_device): same thirteen arguments, same checks, A crosses PCIe once
    detail::ttv_resident(twin, ep, sp, fp, q, a.order(),
                         a.data().data(), a.shape().data(), wa.data(), a.layout().data(),
                         b.data().data(), b.shape().data(),
                         c.data().data(), c.shape().data(), wc.data(), c.layout().data());
    return c;
  }
  ttv(ep, sp, fp, q, a.order(),
      a.data().data(), a.shape().data(), wa.data(), a.layout().data(),
      b.data().data(), b.shape().data(),
      c.data().data(), c.shape().data(), wc.data(), c.layout().data());
  return c;
}

/** Tensor-level interface on DEVICE tensors (device_tensor.h): A, b and the result live in HBM, nothing crosses PCIe. */
template<class value_t, class execution_t, class slicing_t, class fusion_t>
inline auto ttv(std::size_t q, device_tensor<value_t> const& a, device_tensor<value_t> const& b, execution_t ep, slicing_t sp, fusion_t fp)
{
  auto c = device_tensor<value_t>(detail::generate_output_shape(a.shape(), q), detail::generate_output_layout(a.layout(), q),
                                  a.device(), /*zero=*/false);
  auto const wa = a.strides();
  auto const wc = c.strides();
  ttv(ep, sp, fp, q, a.order(),
      a.data(), a.shape().data(), wa.data(), a.layout().data(),
      b.data(), b.shape().data(),
      c.data(), c.shape().data(), wc.data(), c.layout().data());
  return c;
}

} // namespace tlib::ttv

/** auto C = A(q) * b;   uses (par_loop, subtensor, all) like the reference.                      reference ttv.h:122-127 */
template<class value_t>
inline auto operator*(tlib::ttv::tensor_view<value_t> const& a, tlib::ttv::tensor<value_t> const& b)
{
  return tlib::ttv::ttv(a.contraction_mode(), a.get_tensor(), b, tlib::ttv::execution_policy::par_loop,
                        tlib::ttv::slicing_policy::subtensor, tlib::ttv::fusion_policy::all);
}

/** auto C = A(q) * b;  on device tensors: the result is a device tensor too. */
template<class value_t>
inline auto operator*(tlib::ttv::device_tensor_view<value_t> const& a, tlib::ttv::device_tensor<value_t> const& b)
{
  return tlib::ttv::ttv(a.contraction_mode(), a.get_tensor(), b, tlib::ttv::execution_policy::par_loop,
                        tlib::ttv::slicing_policy::subtensor, tlib::ttv::fusion_policy::all);
}
