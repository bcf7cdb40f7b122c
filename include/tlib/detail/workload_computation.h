// tlib/detail/workload_computation.h -- splits layout / shape / stride tuples into the two "slice" modes {pi_1, m} and
// the remaining p-2 modes.  Restates bassoy/ttv detail/workload_computation.h (cited per function).
//
// The reference uses this for its (par_loop, slice, all) loop nest (tensor_times_vector.h:706-770): one flat parallel
// loop over the p-2 free modes, each iteration a small n[pi_1] x n[m] GEMV.  The B200 path gets the same effect from the
// layout folder -- all free modes after q collapse into `outer`, all modes before q into `inner` -- so these functions
// are kept for API parity only.
#pragma once

#include <stdexcept>
#include <utility>
#include <vector>

namespace tlib::ttv::detail {

// (psi, tau): psi orders the pair {pi_1, m}; tau is pi without pi_1 and m, renumbered densely   (reference workload_computation.h:36-70)
template<class size_type>
auto divide_layout(size_type const* const pi, unsigned const p, unsigned const m)
{
  if (p < m)  throw std::runtime_error("Error in tlib::detail::divide_layout: contraction mode cannot be greater than the length of layout tuple.");
  if (m == 0) throw std::runtime_error("Error in tlib::detail::divide_layout: contraction mode cannot be zero.");
  if (p < 3)  throw std::runtime_error("Error in tlib::detail::divide_layout: length of layout tuple must be greater than 2.");

  size_type const lead = pi[0], contract = size_type(m);
  std::vector<size_type> tau;
  tau.reserve(p - 2);
  for (unsigned r = 0; r < p; ++r) {
    size_type const mode = pi[r];
    if (mode == lead || mode == contract) continue;
    tau.push_back(mode - (mode > lead ? 1 : 0) - (mode > contract ? 1 : 0));
  }
  std::vector<size_type> psi = lead < contract ? std::vector<size_type>{1, 2} : std::vector<size_type>{2, 1};
  return std::make_pair(psi, tau);
}

// (x, y): x = {v[pi_1], v[m]}, y = the other entries of v in mode order                          (reference workload_computation.h:77-99)
template<class size_type>
auto divide(size_type const* const v, size_type const* const pi, unsigned const p, unsigned const m)
{
  size_type const lead = pi[0], contract = size_type(m);
  std::vector<size_type> rest;
  rest.reserve(p - 2);
  for (unsigned mode = 1; mode <= p; ++mode)
    if (size_type(mode) != lead && size_type(mode) != contract) rest.push_back(v[mode - 1]);
  return std::make_pair(std::vector<size_type>{v[lead - 1], v[contract - 1]}, rest);
}

// (x, y): x = {v[pi_1]}, y = the other entries of v in mode order (for the output tensor)        (reference workload_computation.h:106-124)
template<class size_type>
auto divide(size_type const* const v, size_type const* const pi, unsigned const p)
{
  size_type const lead = pi[0];
  std::vector<size_type> rest;
  rest.reserve(p - 1);
  for (unsigned mode = 1; mode <= p; ++mode)
    if (size_type(mode) != lead) rest.push_back(v[mode - 1]);
  return std::make_pair(std::vector<size_type>{v[lead - 1]}, rest);
}

} // namespace tlib::ttv::detail
