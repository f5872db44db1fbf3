// Dense-layer entry points (forward, data gradient, weight gradient) -- see include/b200asr.h.
#include "../../include/b200asr.h"
#include "common.cuh"
#include "kernels.h"

using namespace b200asr;

static int gemm_dispatch(const float* A, bool ak, int lda, const float* B, bool bk, int ldb, float* C, int ldc, int M,
                         int N, int K, const float* bias, int relu, const float* mask, int accumulate, bool allow_split,
                         int precision, cudaStream_t st, const float* b_split = nullptr, int b_rows = 0, float* rowsum = nullptr) {
  if (precision == B200ASR_PREC_FP32)
    return gemm_simt(A, ak, lda, B, bk, ldb, C, ldc, M, N, K, bias, relu, mask, accumulate, allow_split, st);
  if (precision == B200ASR_PREC_TF32 || precision == B200ASR_PREC_TF32X3)
    return gemm_tc(A, ak, lda, B, bk, ldb, C, ldc, M, N, K, bias, relu, mask, accumulate, precision, st,
                   precision == B200ASR_PREC_TF32X3 ? b_split : nullptr, b_rows, rowsum);
  set_error("unknown precision %d", precision);
  return B200ASR_BAD_ARG;
}

extern "C" {

int b200asr_linear_fwd(const float* x, const float* w, const float* bias, float* y, int M, int N, int K, int relu,
                       int precision, const float* w_split, b200asr_stream_t stream) {
  B200_REQUIRE(x && w && y && M >= 0 && N > 0 && K > 0, B200ASR_BAD_ARG, "linear_fwd: bad arguments");
  return gemm_dispatch(x, true, K, w, true, K, y, N, M, N, K, bias, relu, nullptr, 0, false, precision,
                       (cudaStream_t)stream, w_split, N);
}

int b200asr_linear_bwd_data(const float* dy, const float* w, const float* relu_out, float* dx, int M, int N, int K,
                            int accumulate, int precision, const float* w_split, b200asr_stream_t stream) {
  B200_REQUIRE(dy && w && dx && M >= 0 && N > 0 && K > 0, B200ASR_BAD_ARG, "linear_bwd_data: bad arguments");
  // dx[m,k] = sum_n dy[m,n] w[n,k]: contraction over n; A = dy (n contiguous), B(n,k) = w[n*K + k] (k contiguous)
  return gemm_dispatch(dy, true, N, w, false, K, dx, K, M, K, N, nullptr, 0, relu_out, accumulate, false, precision,
                       (cudaStream_t)stream, w_split, N);
}

int b200asr_linear_bwd_weight(const float* dy, const float* x, float* dw, float* dbias, int M, int N, int K,
                              int accumulate, int precision, b200asr_stream_t stream) {
  B200_REQUIRE(dy && x && dw && M >= 0 && N > 0 && K > 0, B200ASR_BAD_ARG, "linear_bwd_weight: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  // dw[n,k] = sum_m dy[m,n] x[m,k]: contraction over m; A(n,m) = dy[m*N + n], B(m,k) = x[m*K + k]
  // db[n] = sum_m dy[m,n] = row sums of A: the 3xTF32 kernel produces them from the tiles it stages (one launch and one
  // HBM pass over dy less); other precisions run the column-sum kernel
  const bool fuse = dbias && precision == B200ASR_PREC_TF32X3 && gemm_tc_fuses_rowsum(false, false, 3) && M > 0;
  if (fuse && !accumulate) cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)N, st);
  int rc = gemm_dispatch(dy, false, N, x, false, K, dw, K, N, K, M, nullptr, 0, nullptr, accumulate, true, precision, st,
                         nullptr, 0, fuse ? dbias : nullptr);
  if (rc) return rc;
  if (dbias && !fuse) return launch_colsum(dy, dbias, M, N, accumulate, st);
  return B200ASR_OK;
}

__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = src[i];
  const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  dst[i] = hi;
  dst[n + i] = x - hi;
}

int b200asr_split_tf32(const float* src, float* dst_hi_lo, long long n, b200asr_stream_t stream) {
  B200_REQUIRE(src && dst_hi_lo && n >= 0, B200ASR_BAD_ARG, "split_tf32: bad arguments");
  if (n == 0) return B200ASR_OK;
  split_tf32_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(src, dst_hi_lo, n);
  return check_launch("split_tf32");
}

}  // extern "C"
