// tlib/detail/tensor_times_vector.h -- detail::ttv, the layer below the public interface.
//
// In the reference (bassoy/ttv detail/tensor_times_vector.h:416-1361) this is 19 overloads, one per
// (execution, slicing, fusion) combination, each a differently parallelised loops-over-GEMV nest.  Here it is ONE
// function template: every combination forwards its thirteen arguments to the C-ABI shim, where the layout folder
// collapses (n, pi, q) to the canonical [outer x n_q x inner] view and a single sm_100a kernel launch does what the
// loop nest, the OpenMP parallel-for and the BLAS calls did.  The tags travel along as hints.
#pragma once

#include <cstddef>
#include <stdexcept>
#include <thread>

#include "abi.h"
#include "cases.h"
#include "index.h"
#include "matrix_times_vector.h"
#include "strides.h"
#include "tags.h"
#include "workload_computation.h"

namespace tlib::ttv::detail {

// hardware threads of the host; the reference shells out to lscpu (tensor_times_vector.h:56-88).  Only used by callers
// that size their own thread pools -- the product itself runs on the GPU.
static inline unsigned get_number_cores()
{
  unsigned const n = std::thread::hardware_concurrency();
  return n ? n : 1u;
}

// k = pi^-1(m), and a check that pic is pia without m                             (reference tensor_times_vector.h:147-168)
template<class size_t>
inline unsigned compute_inverse_pia_m(size_t const* const pia, size_t const* const pic, unsigned const p, unsigned const m)
{
  unsigned k = 0u;
  while (k < p && pia[k] != size_t(m)) ++k;
  auto renumbered = [m](size_t mode) { return mode > size_t(m) ? mode - 1 : mode; };
  for (unsigned i = 0u; i < k; ++i)
    if (pic[i] != renumbered(pia[i]))
      throw std::runtime_error("Error in tlib::detail::compute_inverse_pia_m: beginning of layout tuples of both tensors are not correct.");
  for (unsigned i = k; i + 1u < p; ++i)
    if (pic[i] != renumbered(pia[i + 1]))
      throw std::runtime_error("Error in tlib::detail::compute_inverse_pia_m: end of layout tuples of both tensors are not correct.");
  return k + 1u;
}

// inner = product of the extents in front of position k of the layout             (reference tensor_times_vector.h:171-180)
template<class size_t>
inline auto compute_ninvpia(size_t const* const na, size_t const* const pia, unsigned invpia_m)
{
  size_t inner = 1ul;
  for (unsigned r = 0u; r + 1u < invpia_m; ++r) inner *= na[pia[r] - 1];
  return inner;
}

// every (execution, slicing, fusion) combination                                 (reference tensor_times_vector.h:416-1361)
template<class value_t, class size_t, class execution_t, class slicing_t, class fusion_t>
inline void ttv(execution_t, slicing_t, fusion_t,
                unsigned const m, unsigned const p,
                value_t const* const a, size_t const* const na, size_t const* const wa, size_t const* const pia,
                value_t const* const b, size_t const* const nb,
                value_t* const c, size_t const* const nc, size_t const* const wc, size_t const* const pic)
{
  abi::require_supported<value_t>();
  std::size_t const pc = p > 0u ? p - 1u : 0u;
  abi::tuple64<size_t> na_(na, p), wa_(wa, p), pia_(pia, p), nb_(nb, 1), nc_(nc, pc), wc_(wc, pc), pic_(pic, pc);
  ttv_b200_opts opts = abi::make_opts<execution_t, slicing_t, fusion_t>();
  int const status = ttv_b200_run(abi::dtype_v<value_t>, m, p, a, na_.get(), wa_.get(), pia_.get(), b, nb_.get(),
                                  c, nc_.get(), wc_.get(), pic_.get(), &opts);
  if (status != TTV_B200_OK) abi::raise(status);
}

// the same call for a host tensor that keeps its copy in HBM between products (tensor::keep_on_device, ttv_b200_run_resident)
template<class value_t, class size_t, class execution_t, class slicing_t, class fusion_t>
inline void ttv_resident(ttv_b200_resident* twin, execution_t, slicing_t, fusion_t,
                         std::size_t const m, std::size_t const p,
                         value_t const* const a, size_t const* const na, size_t const* const wa, size_t const* const pia,
                         value_t const* const b, size_t const* const nb,
                         value_t* const c, size_t const* const nc, size_t const* const wc, size_t const* const pic)
{
  abi::require_supported<value_t>();
  std::size_t const pc = p > 0u ? p - 1u : 0u;
  abi::tuple64<size_t> na_(na, p), wa_(wa, p), pia_(pia, p), nb_(nb, 1), nc_(nc, pc), wc_(wc, pc), pic_(pic, pc);
  ttv_b200_opts opts = abi::make_opts<execution_t, slicing_t, fusion_t>();
  int const status = ttv_b200_run_resident(twin, abi::dtype_v<value_t>, m, p, a, na_.get(), wa_.get(), pia_.get(), b, nb_.get(),
                                           c, nc_.get(), wc_.get(), pic_.get(), &opts);
  if (status != TTV_B200_OK) abi::raise(status);
}

} // namespace tlib::ttv::detail
