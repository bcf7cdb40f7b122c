// tlib/detail/index.h -- multi-index <-> memory-offset maps.
// Restates the behaviour of bassoy/ttv detail/index.h (cited per function).  The GPU path needs none of this after
// the layout has been folded; it is kept for callers (and the reference's own workload tests) that use it.
#pragma once

namespace tlib::ttv::detail {

// offset of a multi-index: j = sum_r i_r * w_r                                   (reference index.h:30-37)
template<class size_type>
constexpr auto at(unsigned p, size_type const* const i, size_type const* const w)
{
  size_type offset{0u};
  for (unsigned r = p; r-- > 0;) offset += w[r] * i[r];
  return offset;
}

template<class container_type>
constexpr auto at(container_type const& i, container_type const& w)            // reference index.h:44-48
{
  return at(i.size(), i.data(), w.data());
}

// multi-index of an offset: peel the modes off from the slowest (pi_p) to the fastest (pi_1)   (reference index.h:52-63)
template<class size_type>
constexpr void at_1(unsigned const p, size_type* const i, size_type const j, size_type const* const w, size_type const* const pi)
{
  size_type rest = j;
  for (unsigned r = p; r-- > 0;) {
    auto const mode = pi[r] - 1;
    i[mode] = rest / w[mode];
    rest   %= w[mode];
  }
}

template<class container_type, class size_type>
constexpr auto at_1(size_type const j, container_type const& w, container_type const& pi)   // reference index.h:66-72
{
  container_type i = w;
  at_1(i.size(), i.data(), j, w.data(), pi.data());
  return i;
}

// offset in one stride system -> offset in another; modes visited in storage order 1..p of the tuples
// (reference index.h:76-90)
template<class size_type>
constexpr auto at_at_1(unsigned const p, size_type const j_view, size_type const* const w_view, size_type const* const w_array)
{
  size_type rest = j_view, offset = 0;
  for (unsigned r = 0; r < p; ++r) {
    offset += (rest / w_view[r]) * w_array[r];
    rest   %= w_view[r];
  }
  return offset;
}

template<class container_type, class size_type>
constexpr auto at_at_1(size_type const j_view, container_type const& w_view, container_type const& w_array)   // reference index.h:93-97
{
  return at_at_1(w_view.size(), j_view, w_view.data(), w_array.data());
}

// ... the same with the modes visited from the slowest to the fastest of the layout pi   (reference index.h:102-117)
template<class size_type>
constexpr auto at_at_1(unsigned const p, size_type const j_view, size_type const* const w_view, size_type const* const w_array,
                       size_type const* const pi)
{
  size_type rest = j_view, offset = 0;
  for (unsigned r = p; r-- > 0;) {
    auto const mode = pi[r] - 1;
    offset += (rest / w_view[mode]) * w_array[mode];
    rest   %= w_view[mode];
  }
  return offset;
}

template<class container_type, class size_type>
constexpr auto at_at_1(size_type const j_view, container_type const& w_view, container_type const& w_array, container_type const& pi)   // reference index.h:120-124
{
  return at_at_1(w_view.size(), j_view, w_view.data(), w_array.data(), pi.data());
}

} // namespace tlib::ttv::detail
