// kernels.cuh -- internal C++ interface between the translation units of libnalgebra_b200.
#pragma once
#include "common.cuh"

namespace nab {

// ---- dgemm.cu ------------------------------------------------------------------------------------
// C <- alpha*A*B + beta*C on device pointers with arbitrary element strides (matrixmultiply::dgemm
// semantics).  lower_only: SYRK-shaped update, only elements with row >= col of a column-major C
// are computed/stored.
int dgemm_device(cudaStream_t s, bool lower_only, size_t m, size_t k, size_t n, double alpha,
                 const double* a, ptrdiff_t rsa, ptrdiff_t csa, const double* b, ptrdiff_t rsb, ptrdiff_t csb,
                 double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc);
int pack_strided(cudaStream_t s, double* dst, size_t ldd, const double* src, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols);
int scatter_strided(cudaStream_t s, double* dst, ptrdiff_t rs, ptrdiff_t cs, const double* src, size_t lds, size_t rows, size_t cols);
int scale_strided(cudaStream_t s, double* c, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols, double beta);
int fill_uniform(cudaStream_t s, double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed);

}  // namespace nab
