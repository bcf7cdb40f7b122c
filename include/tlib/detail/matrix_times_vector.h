// tlib/detail/matrix_times_vector.h -- the matrix-times-vector entry points of the reference
// (bassoy/ttv detail/matrix_times_vector.h:51-386), served by the B200 kernels through the C-ABI.
//
// On the CPU these are the arithmetic layer of the LOG design (hand-written OpenMP-SIMD loops or cblas_{s,d}gemv).  Here
// each of them is one launch on the canonical view:
//     gemv_row / dot   rows of a contiguous  -> view (outer = M, nq = N, inner = 1)   DOT kernel
//     gemv_col         columns contiguous    -> view (outer = 1, nq = N, inner = M)   column-GEMV kernel
// a, b, c may be host or device pointers.  lda must be the packed leading dimension (N for rows, M for columns): that is
// the only value the reference's mtv ever passes (:325-331); other values are rejected.
//
// Semantics kept from the reference: the row kernels OVERWRITE c (:67); the non-BLAS column kernels ACCUMULATE into c
// (:124, :162); the BLAS column kernel overwrites (beta = 0, :213-215).
#pragma once

#include <cstddef>
#include <functional>
#include <numeric>
#include <stdexcept>

#include "abi.h"
#include "cases.h"
#include "tags.h"

namespace tlib::ttv::detail {

namespace mtv_impl {
template<class value_t, class size_t>
inline void rows(value_t const* a, value_t const* b, value_t* c, size_t M, size_t N, size_t lda)
{
  if (lda != N) throw std::runtime_error("Error in tlib::detail::gemv_row (B200): lda must equal the row length N (packed rows).");
  abi::view<value_t>(M, N, 1, a, b, c, /*accumulate=*/false);
}
template<class value_t, class size_t>
inline void cols(value_t const* a, value_t const* b, value_t* c, size_t M, size_t N, size_t lda, bool accumulate)
{
  if (lda != M) throw std::runtime_error("Error in tlib::detail::gemv_col (B200): lda must equal the column length M (packed columns).");
  abi::view<value_t>(1, N, M, a, b, c, accumulate);
}
} // namespace mtv_impl

// c[i] = sum_k a[i*lda + k] b[k]                                                  (reference matrix_times_vector.h:51-69)
template<class value_t, class size_t>
inline void gemv_row(value_t const* const a, value_t const* const b, value_t* const c, size_t const M, size_t const N, size_t const lda)
{ mtv_impl::rows(a, b, c, M, N, lda); }

template<class value_t, class size_t>                                           // reference matrix_times_vector.h:72-91
inline void gemv_row_parallel(value_t const* const a, value_t const* const b, value_t* const c, size_t const M, size_t const N, size_t const lda)
{ mtv_impl::rows(a, b, c, M, N, lda); }

template<class value_t, class size_t>                                           // reference matrix_times_vector.h:231-262
inline void gemv_row_blas(value_t const* const a, value_t const* const b, value_t* const c, size_t const M, size_t const N, size_t const lda)
{ mtv_impl::rows(a, b, c, M, N, lda); }

// c[j] += sum_k a[k*lda + j] b[k]                                                 (reference matrix_times_vector.h:108-127)
template<class value_t, class size_t>
inline void gemv_col(value_t const* const a, value_t const* const b, value_t* const c, size_t const M, size_t const N, size_t const lda)
{ mtv_impl::cols(a, b, c, M, N, lda, /*accumulate=*/true); }

template<class value_t, class size_t>                                           // reference matrix_times_vector.h:132-179
inline void gemv_col_parallel(value_t const* const a, value_t const* const b, value_t* const c, size_t const M, size_t const N, size_t const lda)
{ mtv_impl::cols(a, b, c, M, N, lda, /*accumulate=*/true); }

// c[j] = sum_k a[k*lda + j] b[k]  (BLAS semantics, beta = 0)                      (reference matrix_times_vector.h:189-221)
template<class value_t, class size_t>
inline void gemv_col_blas(value_t const* const a, value_t const* const b, value_t* const c, size_t const M, size_t const N, size_t const lda)
{ mtv_impl::cols(a, b, c, M, N, lda, /*accumulate=*/false); }

// c[0] = sum_k a[k] b[k]                                                          (reference matrix_times_vector.h:264-295)
template<class value_t, class size_t>
inline void dot(value_t const* const a, value_t const* const b, value_t* const c, size_t const M)
{ abi::view<value_t>(1, M, 1, a, b, c, false); }

template<class value_t, class size_t>
inline void dot_parallel(value_t const* const a, value_t const* const b, value_t* const c, size_t const M)
{ abi::view<value_t>(1, M, 1, a, b, c, false); }

template<class size_t>
inline auto compute_nfull(size_t const* const na, unsigned p)                    // reference matrix_times_vector.h:298-302
{
  return std::accumulate(na, na + p, 1ul, std::multiplies<>());
}

// cases 1-7: one product on the whole packed tensor; wa, nb, nc, wc, pic are not looked at   (reference matrix_times_vector.h:314-386)
template<class value_t, class size_t, class execution_policy>
inline void mtv(execution_policy, unsigned const m, unsigned const p,
                value_t const* const a, size_t const* const na, size_t const* const /*wa*/, size_t const* const pia,
                value_t const* const b, size_t const* const /*nb*/,
                value_t* const c, size_t const* const /*nc*/, size_t const* const /*wc*/, size_t const* const /*pic*/)
{
  auto const nq   = na[m - 1];
  auto const rest = compute_nfull(na, p) / nq;
  constexpr bool overwrite_columns = std::is_same_v<execution_policy, execution_policy::parallel_blas_t> ||
                                     std::is_same_v<execution_policy, execution_policy::sequential_blas_t>;
  if (is_case<1>(p, m, pia))                                   abi::view<value_t>(1, na[0], 1, a, b, c, false);
  else if (is_case<2>(p, m, pia) || is_case<5>(p, m, pia) || is_case<6>(p, m, pia))
                                                               abi::view<value_t>(rest, nq, 1, a, b, c, false);
  else if (is_case<3>(p, m, pia) || is_case<4>(p, m, pia) || is_case<7>(p, m, pia))
                                                               abi::view<value_t>(1, nq, rest, a, b, c, !overwrite_columns);
}

} // namespace tlib::ttv::detail
