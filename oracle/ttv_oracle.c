/*
 * ttv_oracle.c -- CPU ORACLE for the mode-q tensor-times-vector product.  TEST INFRASTRUCTURE ONLY.
 * See ttv_oracle.h for what this is, who may call it and how its parity is pinned.
 * Cited lines are relative to /root/reference/include/tlib/.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fPIC -shared ttv_oracle.c -o libttv_oracle.so -lm   (oracle/Makefile)
 */
#include "ttv_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------------------------
 * L0 helpers
 * ------------------------------------------------------------------------------------------------------------- */

/* non-empty and no zero extent.                                                             detail/shape.h:30-34 */
int ttv_oracle_is_valid_shape(const uint64_t* n, uint64_t p)
{
  if (p == 0) return 0;
  for (uint64_t r = 0; r < p; ++r) if (n[r] == 0) return 0;
  return 1;
}

/* a permutation of 1..p.                                                                   detail/layout.h:29-55 */
int ttv_oracle_is_valid_layout(const uint64_t* pi, uint64_t p)
{
  if (p == 0) return 0;
  for (uint64_t r = 0; r < p; ++r) {
    if (pi[r] == 0 || pi[r] > p) return 0;
    for (uint64_t s = r + 1; s < p; ++s) if (pi[s] == pi[r]) return 0;
  }
  return 1;
}

/* strides must not decrease along layout order; a single mode with w == 1 is fine (and, because the loop below is
 * then empty, so is any other single stride).                                             detail/strides.h:76-101 */
int ttv_oracle_is_valid_strides(const uint64_t* pi, uint64_t p, const uint64_t* w)
{
  for (uint64_t r = 1; r < p; ++r)
    if (w[pi[r - 1] - 1] > w[pi[r] - 1]) return 0;
  return 1;
}

static int all_ones(const uint64_t* n, uint64_t from, uint64_t p)
{
  for (uint64_t r = from; r < p; ++r) if (n[r] != 1) return 0;
  return 1;
}

/* scalar / vector shapes get all-one strides, everything else the packed strides of (n, pi).
 * detail/strides.h:31-57 with the predicates of detail/shape.h:38-89 */
int ttv_oracle_compute_strides(const uint64_t* n, const uint64_t* pi, uint64_t p, uint64_t* w)
{
  if (!ttv_oracle_is_valid_shape(n, p) || !ttv_oracle_is_valid_layout(pi, p)) return -1;
  for (uint64_t r = 0; r < p; ++r) w[r] = 1;
  if (all_ones(n, 0, p)) return 0;                                   /* is_scalar */
  if (p == 1) return 0;                                              /* is_vector, one mode */
  if ((n[0] == 1 || n[1] == 1) && all_ones(n, 2, p)) return 0;       /* is_vector, p >= 2 */
  for (uint64_t r = 1; r < p; ++r)
    w[pi[r] - 1] = w[pi[r - 1] - 1] * n[pi[r - 1] - 1];
  return 0;
}

/* nc = na with entry q erased.                                                           detail/shape.h:103-123 */
int ttv_oracle_output_shape(const uint64_t* na, uint64_t p, uint64_t q, uint64_t* nc)
{
  if (!ttv_oracle_is_valid_shape(na, p) || q == 0 || q > p) return -1;
  uint64_t j = 0;
  for (uint64_t r = 0; r < p; ++r) if (r != q - 1) nc[j++] = na[r];
  return 0;
}

/* pic = pia with q erased, larger modes decremented.                                    detail/layout.h:143-172 */
int ttv_oracle_output_layout(const uint64_t* pia, uint64_t p, uint64_t q, uint64_t* pic)
{
  if (!ttv_oracle_is_valid_layout(pia, p) || q == 0 || q > p) return -1;
  uint64_t j = 0;
  for (uint64_t r = 0; r < p; ++r) {
    if (pia[r] == q) continue;
    pic[j++] = pia[r] > q ? pia[r] - 1 : pia[r];
  }
  return 0;
}

/* (k, k-1, ..., 1, k+1, ..., p); k == 0 or k > p means last-order.                        detail/layout.h:57-76 */
int ttv_oracle_k_order_layout(uint64_t p, uint64_t k, uint64_t* pi)
{
  if (p == 0) return -1;
  if (k == 0 || k > p) k = p;
  for (uint64_t r = 0; r < k; ++r) pi[r] = k - r;
  for (uint64_t r = k; r < p; ++r) pi[r] = r + 1;
  return 0;
}

/* detail/cases.h:24-36 */
int ttv_oracle_case(uint64_t p, uint64_t q, const uint64_t* pia)
{
  if (p == 1) return 1;
  if (p == 2) {
    if (q == 1 && pia[0] == 1) return 2;
    if (q == 2 && pia[0] == 1) return 3;
    if (q == 1 && pia[0] == 2) return 4;
    if (q == 2 && pia[0] == 2) return 5;
    return 0;
  }
  if (pia[0] == q)     return 6;
  if (pia[p - 1] == q) return 7;
  return 8;
}

/* ---------------------------------------------------------------------------------------------------------------
 * argument checks
 * ------------------------------------------------------------------------------------------------------------- */

static const char* const k_messages[] = {
  /* 0*/ "ok",
  /* 1*/ "Error in tlib::tensor_times_vector: input tensor order should be greater zero.",
  /* 2*/ "Error in tlib::tensor_times_vector: contraction mode should be greater zero or less than or equal to p.",
  /* 3*/ "Error in tlib::tensor_times_vector: pointer to input tensor A should not be zero.",
  /* 4*/ "Error in tlib::tensor_times_vector: pointer to input vector B should not be zero.",
  /* 5*/ "Error in tlib::tensor_times_vector: pointer to output tensor C should not be zero.",
  /* 6*/ "Error in tlib::tensor_times_vector: pointer to input tensor shape vector na should not be zero.",
  /* 7*/ "Error in tlib::tensor_times_vector: pointer to input vector shape vector nb should not be zero.",
  /* 8*/ "Error in tlib::tensor_times_vector: pointer to output tensor shape vector nc should not be zero.",
  /* 9*/ "Error in tlib::tensor_times_vector: pointer to input tensor stride vector wa should not be zero.",
  /*10*/ "Error in tlib::tensor_times_vector: pointer to output tensor stride vector wc should not be zero.",
  /*11*/ "Error in tlib::tensor_times_vector: pointer to input tensor permutation vector pia should not be zero.",
  /*12*/ "Error in tlib::tensor_times_vector: pointer to output tensor permutation vector pic should not be zero.",
  /*13*/ "Error in tlib::tensor_times_vector: contraction dimension of A and B are not equal.",
  /*14*/ "Error in tlib::tensor_times_vector: shape vector of A is not valid.",
  /*15*/ "Error in tlib::tensor_times_vector: shape vector of C is not valid.",
  /*16*/ "Error in tlib::tensor_times_vector: layout vector of A is not valid.",
  /*17*/ "Error in tlib::tensor_times_vector: layout vector of C is not valid.",
  /*18*/ "Error in tlib::tensor_times_vector: stride vector of A is not valid.",
  /*19*/ "Error in tlib::tensor_times_vector: stride vector of C is not valid.",
  /*20*/ "Error in tlib::detail::compute_inverse_pia_m: beginning of layout tuples of both tensors are not correct.",
  /*21*/ "Error in tlib::detail::compute_inverse_pia_m: end of layout tuples of both tensors are not correct.",
};

const char* ttv_oracle_strerror(int status)
{
  if (status < 0 || status > 21) return "unknown oracle status";
  return k_messages[status];
}

/* the sixteen checks of the low-level interface, in its order.                                       ttv.h:64-89 */
static int oracle_check_args(uint64_t q, uint64_t p,
                             const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                             const void* b, const uint64_t* nb,
                             const void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic)
{
  if (p == 0)          return 1;
  if (q == 0 || q > p) return 2;
  if (!a)   return 3;
  if (!b)   return 4;
  if (!c)   return 5;
  if (!na)  return 6;
  if (!nb)  return 7;
  if (!nc)  return 8;
  if (!wa)  return 9;
  if (!wc)  return 10;
  if (!pia) return 11;
  if (!pic) return 12;
  if (na[q - 1] != nb[0]) return 13;
  if (!ttv_oracle_is_valid_shape(na, p))      return 14;
  if (!ttv_oracle_is_valid_shape(nc, p - 1))  return 15;     /* p == 1: empty range, always invalid */
  if (!ttv_oracle_is_valid_layout(pia, p))    return 16;
  if (!ttv_oracle_is_valid_layout(pic, p - 1)) return 17;
  if (!ttv_oracle_is_valid_strides(pia, p, wa))     return 18;
  if (!ttv_oracle_is_valid_strides(pic, p - 1, wc)) return 19;
  return 0;
}

/* k = position of q in pia; pic must be pia without q, larger modes decremented.
 * detail/tensor_times_vector.h:147-168 (case 8 only -- the reference does not look at pic in cases 1-7) */
static int oracle_match_layouts(const uint64_t* pia, const uint64_t* pic, uint64_t p, uint64_t q, uint64_t* k_out)
{
  uint64_t k = 0;
  while (k < p && pia[k] != q) ++k;
  for (uint64_t i = 0; i < k; ++i) {
    const uint64_t want = pia[i] > q ? pia[i] - 1 : pia[i];
    if (pic[i] != want) return 20;
  }
  for (uint64_t i = k; i + 1 < p; ++i) {
    const uint64_t want = pia[i + 1] > q ? pia[i + 1] - 1 : pia[i + 1];
    if (pic[i] != want) return 21;
  }
  *k_out = k + 1;
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * per-type bodies
 * ------------------------------------------------------------------------------------------------------------- */
typedef float  _Complex c64_t;
typedef double _Complex c128_t;

#define T float
#define SFX f32
#include "ttv_oracle_impl.inc"
#undef T
#undef SFX

#define T double
#define SFX f64
#include "ttv_oracle_impl.inc"
#undef T
#undef SFX

#define T c64_t
#define SFX c64
#include "ttv_oracle_impl.inc"
#undef T
#undef SFX

#define T c128_t
#define SFX c128
#include "ttv_oracle_impl.inc"
#undef T
#undef SFX

/* integer types use unsigned arithmetic so that overflow wraps instead of being undefined */
#define T uint32_t
#define SFX i32
#include "ttv_oracle_impl.inc"
#undef T
#undef SFX

#define T uint64_t
#define SFX i64
#include "ttv_oracle_impl.inc"
#undef T
#undef SFX

int ttv_oracle_run(int dtype, int slicing, uint64_t q, uint64_t p,
                   const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                   const void* b, const uint64_t* nb,
                   void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic)
{
  switch (dtype) {
    case TTV_ORACLE_F32:  return run_f32 (slicing, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    case TTV_ORACLE_F64:  return run_f64 (slicing, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    case TTV_ORACLE_C64:  return run_c64 (slicing, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    case TTV_ORACLE_C128: return run_c128(slicing, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    case TTV_ORACLE_I32:  return run_i32 (slicing, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    case TTV_ORACLE_I64:  return run_i64 (slicing, q, p, a, na, wa, pia, b, nb, c, nc, wc, pic);
    default: return -1;
  }
}

int ttv_oracle_gemv_row(int dtype, const void* a, const void* b, void* c, uint64_t M, uint64_t N, uint64_t lda)
{
  switch (dtype) {
    case TTV_ORACLE_F32:  row_gemv_f32 (a, b, c, M, N, lda); return 0;
    case TTV_ORACLE_F64:  row_gemv_f64 (a, b, c, M, N, lda); return 0;
    case TTV_ORACLE_C64:  row_gemv_c64 (a, b, c, M, N, lda); return 0;
    case TTV_ORACLE_C128: row_gemv_c128(a, b, c, M, N, lda); return 0;
    case TTV_ORACLE_I32:  row_gemv_i32 (a, b, c, M, N, lda); return 0;
    case TTV_ORACLE_I64:  row_gemv_i64 (a, b, c, M, N, lda); return 0;
    default: return -1;
  }
}

int ttv_oracle_gemv_col(int dtype, const void* a, const void* b, void* c, uint64_t M, uint64_t N, uint64_t lda)
{
  switch (dtype) {
    case TTV_ORACLE_F32:  col_gemv_f32 (a, b, c, M, N, lda); return 0;
    case TTV_ORACLE_F64:  col_gemv_f64 (a, b, c, M, N, lda); return 0;
    case TTV_ORACLE_C64:  col_gemv_c64 (a, b, c, M, N, lda); return 0;
    case TTV_ORACLE_C128: col_gemv_c128(a, b, c, M, N, lda); return 0;
    case TTV_ORACLE_I32:  col_gemv_i32 (a, b, c, M, N, lda); return 0;
    case TTV_ORACLE_I64:  col_gemv_i64 (a, b, c, M, N, lda); return 0;
    default: return -1;
  }
}

/* ---------------------------------------------------------------------------------------------------------------
 * naive checker: C(i_1..i_{q-1}, i_{q+1}..i_p) = sum_k A(i_1..k..i_p) b(k)   (README.md:13-18), independent of the
 * reference's loop structure.  Addresses come from the packed strides of (na, pia) and (nc, pic = output layout).
 * ------------------------------------------------------------------------------------------------------------- */
#define MAXP 32

int ttv_oracle_naive(int dtype, uint64_t q, uint64_t p,
                     const void* a, const uint64_t* na, const uint64_t* pia,
                     const void* b, void* c, double* abs_out)
{
  if (p < 2 || p > MAXP || q == 0 || q > p) return -1;
  if (!ttv_oracle_is_valid_shape(na, p) || !ttv_oracle_is_valid_layout(pia, p)) return -1;
  uint64_t wa[MAXP], nc[MAXP], pic[MAXP], wc[MAXP], idx[MAXP];
  /* always the packed strides, also for vector-shaped tensors where compute_strides returns ones */
  wa[pia[0] - 1] = 1;
  for (uint64_t r = 1; r < p; ++r) wa[pia[r] - 1] = wa[pia[r - 1] - 1] * na[pia[r - 1] - 1];
  ttv_oracle_output_shape(na, p, q, nc);
  ttv_oracle_output_layout(pia, p, q, pic);
  wc[pic[0] - 1] = 1;
  for (uint64_t r = 1; r + 1 < p; ++r) wc[pic[r] - 1] = wc[pic[r - 1] - 1] * nc[pic[r - 1] - 1];

  uint64_t count = 1;
  for (uint64_t r = 0; r + 1 < p; ++r) count *= nc[r];
  const uint64_t nq = na[q - 1], wq = wa[q - 1];
  memset(idx, 0, sizeof idx);

  for (uint64_t e = 0; e < count; ++e) {
    uint64_t oa = 0, oc = 0;
    for (uint64_t r = 0, j = 0; r < p; ++r) {
      if (r == q - 1) continue;
      oa += idx[j] * wa[r];
      oc += idx[j] * wc[j];
      ++j;
    }
    double mag = 0.0;
    switch (dtype) {
      case TTV_ORACLE_F32: {
        long double s = 0;
        for (uint64_t k = 0; k < nq; ++k) {
          long double x = ((const float*)a)[oa + k * wq], y = ((const float*)b)[k];
          s += x * y; mag += fabs((double)(x * y));
        }
        ((float*)c)[oc] = (float)s; break; }
      case TTV_ORACLE_F64: {
        long double s = 0;
        for (uint64_t k = 0; k < nq; ++k) {
          long double x = ((const double*)a)[oa + k * wq], y = ((const double*)b)[k];
          s += x * y; mag += fabs((double)(x * y));
        }
        ((double*)c)[oc] = (double)s; break; }
      case TTV_ORACLE_C64: {
        long double sr = 0, si = 0;
        for (uint64_t k = 0; k < nq; ++k) {
          const float* x = (const float*)a + 2 * (oa + k * wq); const float* y = (const float*)b + 2 * k;
          long double xr = x[0], xi = x[1], yr = y[0], yi = y[1];
          sr += xr * yr - xi * yi; si += xr * yi + xi * yr;
          mag += hypot((double)xr, (double)xi) * hypot((double)yr, (double)yi);
        }
        ((float*)c)[2 * oc] = (float)sr; ((float*)c)[2 * oc + 1] = (float)si; break; }
      case TTV_ORACLE_C128: {
        long double sr = 0, si = 0;
        for (uint64_t k = 0; k < nq; ++k) {
          const double* x = (const double*)a + 2 * (oa + k * wq); const double* y = (const double*)b + 2 * k;
          long double xr = x[0], xi = x[1], yr = y[0], yi = y[1];
          sr += xr * yr - xi * yi; si += xr * yi + xi * yr;
          mag += hypot((double)xr, (double)xi) * hypot((double)yr, (double)yi);
        }
        ((double*)c)[2 * oc] = (double)sr; ((double*)c)[2 * oc + 1] = (double)si; break; }
      case TTV_ORACLE_I32: {
        uint32_t s = 0;
        for (uint64_t k = 0; k < nq; ++k) s += ((const uint32_t*)a)[oa + k * wq] * ((const uint32_t*)b)[k];
        ((uint32_t*)c)[oc] = s; break; }
      case TTV_ORACLE_I64: {
        uint64_t s = 0;
        for (uint64_t k = 0; k < nq; ++k) s += ((const uint64_t*)a)[oa + k * wq] * ((const uint64_t*)b)[k];
        ((uint64_t*)c)[oc] = s; break; }
      default: return -1;
    }
    if (abs_out) abs_out[oc] = mag;
    /* next multi-index of C, mode 1 fastest */
    for (uint64_t j = 0; j + 1 < p; ++j) {
      if (++idx[j] < nc[j]) break;
      idx[j] = 0;
    }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * synthetic data.  The CUDA library carries its own copy of this generator for device-side fills.
 * ------------------------------------------------------------------------------------------------------------- */
static uint64_t splitmix64(uint64_t x)
{
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static double unit_pm1(uint64_t u) { return (double)(u >> 11) * (2.0 / 9007199254740992.0) - 1.0; }

void ttv_oracle_fill(int dtype, void* x, uint64_t first, uint64_t count, uint64_t seed)
{
  for (uint64_t i = 0; i < count; ++i) {
    const uint64_t j = first + i;
    switch (dtype) {
      case TTV_ORACLE_F32:  ((float*)x)[i]  = (float)unit_pm1(splitmix64(seed ^ j)); break;
      case TTV_ORACLE_F64:  ((double*)x)[i] = unit_pm1(splitmix64(seed ^ j)); break;
      case TTV_ORACLE_C64:  ((float*)x)[2 * i]      = (float)unit_pm1(splitmix64(seed ^ (2 * j)));
                            ((float*)x)[2 * i + 1]  = (float)unit_pm1(splitmix64(seed ^ (2 * j + 1))); break;
      case TTV_ORACLE_C128: ((double*)x)[2 * i]     = unit_pm1(splitmix64(seed ^ (2 * j)));
                            ((double*)x)[2 * i + 1] = unit_pm1(splitmix64(seed ^ (2 * j + 1))); break;
      case TTV_ORACLE_I32:  ((int32_t*)x)[i] = (int32_t)(splitmix64(seed ^ j) % 17u) - 8; break;
      case TTV_ORACLE_I64:  ((int64_t*)x)[i] = (int64_t)(splitmix64(seed ^ j) % 17u) - 8; break;
      default: return;
    }
  }
}
