// gemm.cu -- prompt mat-mat dispatch: Y[n][m] = sum_k W[m][k] * X[n][k] for N > b200_mmv_max_n().
#include "kernels.h"

void launch_gemm_simt(const WPlanes & W, const __half * X, int64_t x_stride, int N, float * Y, int64_t y_stride, int epi_gelu, cudaStream_t stream);

size_t mmq_gemm_workspace_bytes(const WPlanes &, int) { return 256; }

void launch_mmq_gemm(const WPlanes & W, const __half * X, int64_t x_stride, int N, float * Y, int64_t y_stride,
                     int epi_gelu, void *, size_t, cudaStream_t stream) {
    launch_gemm_simt(W, X, x_stride, N, Y, y_stride, epi_gelu, stream);
}
