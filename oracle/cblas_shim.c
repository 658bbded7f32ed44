/* oracle/cblas_shim.c — TEST INFRASTRUCTURE. Minimal cblas used only to link the
 * reference (it calls dgemm/dgemv/daxpy on <= 17x9 operands, reference
 * src/cblas.cpp:33-62). Plain triple loops; not product code. */
#include "cblas.h"

static double at(const double* A, int lda, int r, int c, CBLAS_ORDER o, CBLAS_TRANSPOSE t) {
    if (t != CblasNoTrans) { int q = r; r = c; c = q; }
    return o == CblasRowMajor ? A[r * lda + c] : A[c * lda + r];
}

void cblas_dgemm(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int M, int N, int K,
                 double alpha, const double* A, int lda, const double* B, int ldb, double beta,
                 double* C, int ldc) {
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            for (int l = 0; l < K; ++l) s += at(A, lda, i, l, o, ta) * at(B, ldb, l, j, o, tb);
            double* c = o == CblasRowMajor ? &C[i * ldc + j] : &C[j * ldc + i];
            *c = alpha * s + (beta == 0.0 ? 0.0 : beta * *c);
        }
}

void cblas_dgemv(CBLAS_ORDER o, CBLAS_TRANSPOSE ta, int M, int N, double alpha, const double* A,
                 int lda, const double* X, int incX, double beta, double* Y, int incY) {
    int rows = ta == CblasNoTrans ? M : N, cols = ta == CblasNoTrans ? N : M;
    for (int i = 0; i < rows; ++i) {
        double s = 0.0;
        for (int l = 0; l < cols; ++l) s += at(A, lda, i, l, o, ta) * X[l * incX];
        Y[i * incY] = alpha * s + (beta == 0.0 ? 0.0 : beta * Y[i * incY]);
    }
}

void cblas_daxpy(int n, double alpha, const double* x, int incx, double* y, int incy) {
    for (int i = 0; i < n; ++i) y[i * incy] += alpha * x[i * incx];
}
