/*
 * ttv_oracle.h -- CPU ORACLE for the mode-q tensor-times-vector product.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithm (bassoy/ttv, include/tlib): argument checks, the 8-case
 * classifier, the recursive loops-over-GEMV nest in both slicing variants and the sequential row/column GEMV
 * micro-kernels.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or the reported CPU baseline.  Nothing under ttv_b200/ or include/
 * links, loads or calls it; the product path has no CPU fallback.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against
 *   - the reference's own known answers: gtest_tlib_ttv.cpp:132 closed form on its full grid, gtest_tlib_mtv.cpp:70-76,
 *     example/interface{1,2,3}.cpp output {15,18,21,24,51,54,57,60}, ttvpy/tests/test.py (einsum) and ttvpy/README.md:61-66;
 *   - outputs of the unmodified reference headers compiled here into oracle/_ref/ (oracle/Makefile), both live (when
 *     /root/reference is present) and as committed fixtures under tests/golden/ (made by tests/golden/make_golden.py).
 *
 * Status codes are the same numbers as enum ttv_b200_status in include/ttv_b200.h (1..21 = the reference's throw
 * sites in evaluation order); the oracle keeps its own message table so that the two can be compared.
 */
#ifndef TTV_ORACLE_H
#define TTV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* slicing variants of the loop nest (tags.h:43-50) */
#define TTV_ORACLE_SLICE     0   /* leaf = [n_pi1 x n_q] column-major GEMV, lda = wa[q-1]   tensor_times_vector.h:189-216 */
#define TTV_ORACLE_SUBTENSOR 1   /* leaf = [inner x n_q] column-major GEMV, lda = inner     tensor_times_vector.h:296-324 */

/* dtype numbering as in ttv_b200.h */
#define TTV_ORACLE_F32  0
#define TTV_ORACLE_F64  1
#define TTV_ORACLE_C64  2
#define TTV_ORACLE_C128 3
#define TTV_ORACLE_I32  4
#define TTV_ORACLE_I64  5

/* The reference's sequential path: ttv(seq, slice|subtensor, none, ...), tensor_times_vector.h:430-447 / :934-965.
 * Like the reference's non-BLAS column kernel it ACCUMULATES into c in the column cases and OVERWRITES in the row
 * cases (matrix_times_vector.h:67 vs :124) -- zero c before calling, as every caller in the reference does. */
int ttv_oracle_run(int dtype, int slicing, uint64_t q, uint64_t p,
                   const void* a, const uint64_t* na, const uint64_t* wa, const uint64_t* pia,
                   const void* b, const uint64_t* nb,
                   void* c, const uint64_t* nc, const uint64_t* wc, const uint64_t* pic);

/* Second, order-independent checker: naive strided contraction with wide accumulation (long double for
 * float/double/complex, int64 wrap-around for ints), written against the definition C = A x_q b (README.md:13-18),
 * not against the reference's loop structure.  abs_out (optional, same shape as c, real-valued double) receives
 * sum_k |a_k||b_k| per output element for the tolerance  2 n_q eps sum|a||b|. */
int ttv_oracle_naive(int dtype, uint64_t q, uint64_t p,
                     const void* a, const uint64_t* na, const uint64_t* pia,
                     const void* b, void* c, double* abs_out);

/* micro-kernels, exported for the mtv known-answer tests (matrix_times_vector.h:51-69, :108-127) */
int ttv_oracle_gemv_row(int dtype, const void* a, const void* b, void* c, uint64_t M, uint64_t N, uint64_t lda);
int ttv_oracle_gemv_col(int dtype, const void* a, const void* b, void* c, uint64_t M, uint64_t N, uint64_t lda);

/* L0 helpers restated (shape.h, layout.h, strides.h, cases.h) */
int      ttv_oracle_is_valid_shape  (const uint64_t* n,  uint64_t p);
int      ttv_oracle_is_valid_layout (const uint64_t* pi, uint64_t p);
int      ttv_oracle_is_valid_strides(const uint64_t* pi, uint64_t p, const uint64_t* w);
int      ttv_oracle_compute_strides (const uint64_t* n,  const uint64_t* pi, uint64_t p, uint64_t* w);
int      ttv_oracle_output_shape    (const uint64_t* na, uint64_t p, uint64_t q, uint64_t* nc);
int      ttv_oracle_output_layout   (const uint64_t* pia, uint64_t p, uint64_t q, uint64_t* pic);
int      ttv_oracle_k_order_layout  (uint64_t p, uint64_t k, uint64_t* pi);
int      ttv_oracle_case            (uint64_t p, uint64_t q, const uint64_t* pia);   /* 1..8, cases.h:24-36 */

/* deterministic synthetic data: splitmix64(seed ^ j) -> value; identical to ttv_b200's device generator */
void     ttv_oracle_fill(int dtype, void* x, uint64_t first, uint64_t count, uint64_t seed);

const char* ttv_oracle_strerror(int status);

#ifdef __cplusplus
}
#endif
#endif
