// GEMM entry points used by the C ABI (row-major fp32, arbitrary leading dimensions).
#pragma once
#include <cuda_runtime.h>

namespace spk {
int gemm_nn_simt(const float* A, long lda, const float* B, long ldb, float* C, long ldc,
                 long M, int N, int K, int accumulate, cudaStream_t s);
int gemm_tn_splits(long M, int Ka, int Nb);
long gemm_tn_workspace_floats(long M, int Ka, int Nb);
int gemm_tn_simt(const float* A, long lda, const float* B, long ldb, float* C, long ldc,
                 long M, int Ka, int Nb, int accumulate, float* workspace, cudaStream_t s);
}  // namespace spk

namespace spk {
// tcgen05 3xTF32 path (spk_gemm_tc.cu)
int gemm_tc_ldt(int K);
int gemm_nn_tc_supported(const float* A, long lda, long M, int N, int K);
int gemm_nn_tc(const float* A, long lda, const float* B, long ldb, float* C, long ldc, long M, int N, int K,
               int accumulate, float* workspace, cudaStream_t s, int act = 0);
int gemm_tn_tc_supported(const float* A, long lda, const float* B, long ldb, long M, int Ka, int Nb);
long gemm_tn_tc_workspace_floats(long M, int Ka, int Nb);
int gemm_tn_tc(const float* A, long lda, const float* B, long ldb, float* C, long ldc, long M, int Ka, int Nb,
               int accumulate, float* workspace, cudaStream_t s);
}  // namespace spk
