/*
 * ref_driver.cpp -- C entry points around the UNMODIFIED reference headers.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Compiled by oracle/Makefile with -I/root/reference/include (the sources stay where they lie; nothing is copied)
 * into oracle/_ref/libttv_ref.so (non-BLAS, OpenMP) and oracle/_ref/libttv_ref_openblas.so (-DUSE_OPENBLAS against
 * scipy's bundled OpenBLAS through oracle/blas_shim/cblas.h).  Used
 *   - by tests/ and tests/golden/make_golden.py to validate oracle/ttv_oracle.c and to generate the golden fixtures,
 *   - by bench.py's cpu_baseline / --impl reference legs as the timed CPU implementation ("kind": "reference").
 * The product (ttv_b200/, include/) never loads it.
 *
 * Every (execution, slicing, fusion) combination the reference defines and that is reachable through its public
 * wrapper ttv.h:54-92 is instantiated (tensor_times_vector.h:430-1361 minus the two par_blas_loop overloads, which take
 * an extra `ratio` argument).
 */
#include <complex>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#if defined(_OPENMP)
/* std::complex has no built-in OpenMP reduction; the reference's `omp simd reduction(+:sum)` needs these to compile
 * (SURVEY 8c).  They change nothing for the other element types. */
#pragma omp declare reduction(+ : std::complex<float>  : omp_out += omp_in) initializer(omp_priv = std::complex<float>{})
#pragma omp declare reduction(+ : std::complex<double> : omp_out += omp_in) initializer(omp_priv = std::complex<double>{})
#endif

#include <tlib/ttv.h>

namespace {

thread_local std::string g_error;

namespace ex = tlib::ttv::execution_policy;
namespace sl = tlib::ttv::slicing_policy;
namespace fu = tlib::ttv::fusion_policy;

/* numbering shared with include/ttv_b200.h (enum ttv_b200_execution / slicing / fusion) */
enum { SEQ = 0, SEQ_BLAS, PAR, PAR_LOOP, PAR_TASKLOOP, PAR_TASK, PAR_BLAS, PAR_BLAS_LOOP };
enum { SLICE = 0, SUBTENSOR = 1 };
enum { NONE = 0, OUTER = 1, ALL = 2 };

template<class T>
int run(int ep, int sp, int fp, std::size_t q, std::size_t p,
        T const* a, std::size_t const* na, std::size_t const* wa, std::size_t const* pia,
        T const* b, std::size_t const* nb,
        T* c, std::size_t const* nc, std::size_t const* wc, std::size_t const* pic)
{
  using tlib::ttv::ttv;
#define COMBO(E, S, F, etag, stag, ftag) \
  if (ep == E && sp == S && fp == F) { ttv(etag, stag, ftag, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic); return 0; }
  COMBO(SEQ,          SLICE,     NONE,  ex::seq,          sl::slice,     fu::none)
  COMBO(SEQ_BLAS,     SLICE,     NONE,  ex::seq_blas,     sl::slice,     fu::none)
  COMBO(PAR_TASK,     SLICE,     NONE,  ex::par_task,     sl::slice,     fu::none)
  COMBO(PAR_TASKLOOP, SLICE,     NONE,  ex::par_taskloop, sl::slice,     fu::none)
  COMBO(PAR,          SLICE,     NONE,  ex::par,          sl::slice,     fu::none)
  COMBO(PAR_LOOP,     SLICE,     NONE,  ex::par_loop,     sl::slice,     fu::none)
  COMBO(PAR_LOOP,     SLICE,     OUTER, ex::par_loop,     sl::slice,     fu::outer)
  COMBO(PAR_LOOP,     SLICE,     ALL,   ex::par_loop,     sl::slice,     fu::all)
  COMBO(PAR_BLAS,     SLICE,     ALL,   ex::par_blas,     sl::slice,     fu::all)
  COMBO(SEQ,          SUBTENSOR, NONE,  ex::seq,          sl::subtensor, fu::none)
  COMBO(SEQ_BLAS,     SUBTENSOR, NONE,  ex::seq_blas,     sl::subtensor, fu::none)
  COMBO(PAR_TASK,     SUBTENSOR, NONE,  ex::par_task,     sl::subtensor, fu::none)
  COMBO(PAR_TASKLOOP, SUBTENSOR, NONE,  ex::par_taskloop, sl::subtensor, fu::none)
  COMBO(PAR,          SUBTENSOR, NONE,  ex::par,          sl::subtensor, fu::none)
  COMBO(PAR_LOOP,     SUBTENSOR, NONE,  ex::par_loop,     sl::subtensor, fu::none)
  COMBO(PAR_LOOP,     SUBTENSOR, ALL,   ex::par_loop,     sl::subtensor, fu::all)
  COMBO(PAR_BLAS,     SUBTENSOR, ALL,   ex::par_blas,     sl::subtensor, fu::all)
#undef COMBO
  g_error = "ref_driver: the reference defines no ttv overload for this (execution, slicing, fusion) combination";
  return -2;
}

template<class T>
int guarded(int ep, int sp, int fp, std::uint64_t q, std::uint64_t p,
            void const* a, std::uint64_t const* na, std::uint64_t const* wa, std::uint64_t const* pia,
            void const* b, std::uint64_t const* nb,
            void* c, std::uint64_t const* nc, std::uint64_t const* wc, std::uint64_t const* pic)
{
  static_assert(sizeof(std::size_t) == sizeof(std::uint64_t), "LP64 expected");
  auto sz = [](std::uint64_t const* v) { return reinterpret_cast<std::size_t const*>(v); };
  try {
    return run<T>(ep, sp, fp, q, p, static_cast<T const*>(a), sz(na), sz(wa), sz(pia), static_cast<T const*>(b), sz(nb),
                  static_cast<T*>(c), sz(nc), sz(wc), sz(pic));
  } catch (std::exception const& e) {
    g_error = e.what();
    return -1;
  }
}

/* interface 1 / 2: tensor class + operator*  (ttv.h:99-127, tensor.h:56-114) */
template<class T>
int tensor_iface(int use_operator, std::uint64_t q, std::uint64_t p, void const* a, std::uint64_t const* na,
                 std::uint64_t const* pia, void const* b, void* c, std::uint64_t* nc, std::uint64_t* pic, std::uint64_t* wc)
{
  try {
    using tensor_t = tlib::ttv::tensor<T>;
    std::vector<std::size_t> shape(na, na + p), layout(pia, pia + p);
    tensor_t A(shape, layout);
    std::memcpy(A.data().data(), a, A.data().size() * sizeof(T));
    tensor_t B(std::vector<std::size_t>{shape.at(q - 1), 1});
    std::memcpy(B.data().data(), b, shape.at(q - 1) * sizeof(T));
    auto C = use_operator ? (A(q) * B)
                          : tlib::ttv::ttv(q, A, B, ex::seq, sl::subtensor, fu::none);
    std::memcpy(c, C.data().data(), C.data().size() * sizeof(T));
    auto s = C.strides();
    for (std::size_t r = 0; r + 1 < p; ++r) { nc[r] = C.shape()[r]; pic[r] = C.layout()[r]; wc[r] = s[r]; }
    return 0;
  } catch (std::exception const& e) {
    g_error = e.what();
    return -1;
  }
}

} // namespace

extern "C" {

int ttv_ref_run(int dtype, int ep, int sp, int fp, std::uint64_t q, std::uint64_t p,
                void const* a, std::uint64_t const* na, std::uint64_t const* wa, std::uint64_t const* pia,
                void const* b, std::uint64_t const* nb,
                void* c, std::uint64_t const* nc, std::uint64_t const* wc, std::uint64_t const* pic)
{
  switch (dtype) {
    case 0: return guarded<float>               (ep, sp, fp, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    case 1: return guarded<double>              (ep, sp, fp, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    case 2: return guarded<std::complex<float>> (ep, sp, fp, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    case 3: return guarded<std::complex<double>>(ep, sp, fp, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    case 4: return guarded<std::int32_t>        (ep, sp, fp, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    case 5: return guarded<std::int64_t>        (ep, sp, fp, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    default: g_error = "ref_driver: unknown dtype"; return -3;
  }
}

int ttv_ref_tensor(int dtype, int use_operator, std::uint64_t q, std::uint64_t p, void const* a, std::uint64_t const* na,
                   std::uint64_t const* pia, void const* b, void* c, std::uint64_t* nc, std::uint64_t* pic, std::uint64_t* wc)
{
  switch (dtype) {
    case 0: return tensor_iface<float> (use_operator, q, p, a, na, pia, b, c, nc, pic, wc);
    case 1: return tensor_iface<double>(use_operator, q, p, a, na, pia, b, c, nc, pic, wc);
    case 4: return tensor_iface<std::int32_t>(use_operator, q, p, a, na, pia, b, c, nc, pic, wc);
    default: g_error = "ref_driver: dtype not wired for the tensor interface"; return -3;
  }
}

/* L0 helpers of the reference, for the helper parity tests */
int ttv_ref_compute_strides(std::uint64_t const* n, std::uint64_t const* pi, std::uint64_t p, std::uint64_t* w)
{
  try { tlib::ttv::detail::compute_strides(n, n + p, pi, w); return 0; }
  catch (std::exception const& e) { g_error = e.what(); return -1; }
}
int ttv_ref_is_valid_strides(std::uint64_t const* pi, std::uint64_t p, std::uint64_t const* w)
{
  try { return tlib::ttv::detail::is_valid_strides(pi, pi + p, w) ? 1 : 0; }
  catch (std::exception const& e) { g_error = e.what(); return -1; }
}
int ttv_ref_is_valid_layout(std::uint64_t const* pi, std::uint64_t p) { return tlib::ttv::detail::is_valid_layout(pi, pi + p) ? 1 : 0; }
int ttv_ref_is_valid_shape (std::uint64_t const* n,  std::uint64_t p) { return tlib::ttv::detail::is_valid_shape(n, n + p) ? 1 : 0; }
int ttv_ref_output_shape(std::uint64_t const* na, std::uint64_t p, std::uint64_t q, std::uint64_t* nc)
{
  try { tlib::ttv::detail::compute_output_shape(na, na + p, nc, q); return 0; }
  catch (std::exception const& e) { g_error = e.what(); return -1; }
}
int ttv_ref_output_layout(std::uint64_t const* pia, std::uint64_t p, std::uint64_t q, std::uint64_t* pic)
{
  try { tlib::ttv::detail::compute_output_layout(pia, pia + p, pic, q); return 0; }
  catch (std::exception const& e) { g_error = e.what(); return -1; }
}
int ttv_ref_k_order_layout(std::uint64_t p, std::uint64_t k, std::uint64_t* pi)
{
  try { tlib::ttv::detail::compute_k_order_layout(pi, pi + p, k); return 0; }
  catch (std::exception const& e) { g_error = e.what(); return -1; }
}
int ttv_ref_case(std::uint64_t p, std::uint64_t q, std::uint64_t const* pia)
{
  using namespace tlib::ttv::detail;
  unsigned pp = unsigned(p), qq = unsigned(q);
  if (is_case<1>(pp, qq, pia)) return 1;
  if (is_case<2>(pp, qq, pia)) return 2;
  if (is_case<3>(pp, qq, pia)) return 3;
  if (is_case<4>(pp, qq, pia)) return 4;
  if (is_case<5>(pp, qq, pia)) return 5;
  if (is_case<6>(pp, qq, pia)) return 6;
  if (is_case<7>(pp, qq, pia)) return 7;
  if (is_case<8>(pp, qq, pia)) return 8;
  return 0;
}

char const* ttv_ref_last_error(void) { return g_error.c_str(); }

int ttv_ref_has_blas(void)
{
#if defined(USE_OPENBLAS) || defined(USE_MKL) || defined(USE_BLIS)
  return 1;
#else
  return 0;
#endif
}

int ttv_ref_has_openmp(void)
{
#if defined(_OPENMP)
  return 1;
#else
  return 0;
#endif
}

/* sockets x cores/socket as the reference counts them (tensor_times_vector.h:56-88) */
unsigned ttv_ref_cores(void) { return tlib::ttv::detail::get_number_cores(); }

} // extern "C"
