// tlib/detail/cases.h -- the reference's 8-way classification of (p, q, pi)  (bassoy/ttv detail/cases.h:24-36).
//
// Kept for API parity.  The GPU path does not branch on it: cases 2,5,6 are the canonical view with inner == 1,
// cases 3,4,7 the view with outer == 1, case 8 everything else (see ttv_b200/csrc/plan.h).
#pragma once

namespace tlib::ttv::detail {

template<unsigned case_nr, typename size_t>
inline constexpr bool is_case(unsigned p, unsigned q, size_t const* const pi)
{
  static_assert(1u <= case_nr && case_nr <= 8u, "tlib::ttv::detail::is_case: cases are numbered 1..8.");
  bool const first  = pi[0] == size_t(q);       // q is the fastest mode
  if constexpr (case_nr == 1u) return p == 1u;
  else if constexpr (case_nr <= 5u) {
    if (p != 2u) return false;
    bool const col_major = pi[0] == size_t(1);
    if constexpr (case_nr == 2u) return  col_major && q == 1u;
    if constexpr (case_nr == 3u) return  col_major && q == 2u;
    if constexpr (case_nr == 4u) return !col_major && pi[0] == size_t(2) && q == 1u;
    if constexpr (case_nr == 5u) return !col_major && pi[0] == size_t(2) && q == 2u;
  }
  else {
    if (p < 3u) return false;
    bool const last = pi[p - 1] == size_t(q);   // q is the slowest mode
    if constexpr (case_nr == 6u) return first;
    if constexpr (case_nr == 7u) return last;
    if constexpr (case_nr == 8u) return !first && !last;
  }
  return false;
}

} // namespace tlib::ttv::detail
