/*
 * cblas.h -- declaration shim so that the UNMODIFIED reference headers, built with -DUSE_OPENBLAS, link against the
 * OpenBLAS that ships inside the image's scipy wheel (libscipy_openblas, LP64, symbols prefixed "scipy_").
 * TEST / BASELINE INFRASTRUCTURE ONLY (used by oracle/Makefile for oracle/_ref/libttv_ref_openblas.so).
 * Only the four entry points the reference calls are declared (matrix_times_vector.h:213,215,254,256;
 * tensor_times_vector.h:94,106).
 */
#ifndef TTV_ORACLE_CBLAS_SHIM_H
#define TTV_ORACLE_CBLAS_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int blasint;
enum CBLAS_ORDER     { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };

void scipy_cblas_sgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, blasint m, blasint n, float alpha,
                       const float* a, blasint lda, const float* x, blasint incx, float beta, float* y, blasint incy);
void scipy_cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, blasint m, blasint n, double alpha,
                       const double* a, blasint lda, const double* x, blasint incx, double beta, double* y, blasint incy);
void scipy_openblas_set_num_threads(int n);
int  scipy_openblas_get_num_threads(void);

#define cblas_sgemv              scipy_cblas_sgemv
#define cblas_dgemv              scipy_cblas_dgemv
#define openblas_set_num_threads scipy_openblas_set_num_threads
#define openblas_get_num_threads scipy_openblas_get_num_threads

#ifdef __cplusplus
}
#endif
#endif
