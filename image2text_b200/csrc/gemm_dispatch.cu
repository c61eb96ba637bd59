// i2t_gemm: one entry point for every dense contraction on the path.
//   fp32 operands  -> exact fp32 FMA tiles (gemm_simt.cu), the parity anchor;
//   bf16 operands  -> tcgen05/TMEM tiles fed by TMA (gemm_tc.cu) when the shape fits its tiling,
//                     otherwise the same SIMT kernel reading bf16.
#include "common.cuh"

namespace i2t {
int gemm_simt(const void* A, const void* B, const float* bias, const void* residual, void* C, int64_t M, int64_t N,
              int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_kmajor, int b_kmajor, int act, int accumulate,
              int ab_dtype, int res_dtype, int c_dtype, cudaStream_t st);
// returns 1 if it handled the problem, 0 if the shape is not eligible, <0 on error
int gemm_tc_try(const void* A, const void* B, const float* bias, const void* residual, void* C, int64_t M, int64_t N,
                int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_kmajor, int b_kmajor, int act, int accumulate,
                int res_dtype, int c_dtype, int flags, cudaStream_t st);
static std::atomic<int> g_tc_mode{1};  // 1: use tcgen05 when eligible, 0: never (debug / A-B testing)
}  // namespace i2t

using namespace i2t;

extern "C" void i2t_set_tensor_core_gemm(int enabled) { g_tc_mode.store(enabled ? 1 : 0); }

extern "C" int i2t_gemm_ex(const void* A, const void* B, const float* bias, const void* residual, void* C, int64_t M,
                           int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_kmajor, int b_kmajor, int act,
                           int accumulate, int ab_dtype, int res_dtype, int c_dtype, int flags, void* stream) {
  I2T_REQUIRE(A && B && C, "gemm: null pointer");
  I2T_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm: bad sizes M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  I2T_REQUIRE(valid_dtype(ab_dtype) && valid_dtype(res_dtype) && valid_dtype(c_dtype), "gemm: bad dtype");
  I2T_REQUIRE(act >= 0 && act <= 2, "gemm: bad activation code");
  I2T_REQUIRE(lda >= (a_kmajor ? K : M) && ldb >= (b_kmajor ? K : N) && ldc >= N, "gemm: leading dimension too small");
  I2T_REQUIRE(!accumulate || c_dtype == I2T_F32, "gemm: accumulate needs an fp32 C");
  if (M == 0) return I2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (ab_dtype == I2T_BF16 && g_tc_mode.load() == 1) {
    const int r = gemm_tc_try(A, B, bias, residual, C, M, N, K, lda, ldb, ldc, a_kmajor, b_kmajor, act, accumulate,
                              res_dtype, c_dtype, flags, st);
    if (r != 0) return r < 0 ? r : I2T_OK;
  }
  return gemm_simt(A, B, bias, residual, C, M, N, K, lda, ldb, ldc, a_kmajor, b_kmajor, act, accumulate, ab_dtype,
                   res_dtype, c_dtype, st);
}


extern "C" int i2t_gemm(const void* A, const void* B, const float* bias, const void* residual, void* C, int64_t M,
                        int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_kmajor, int b_kmajor, int act,
                        int accumulate, int ab_dtype, int res_dtype, int c_dtype, void* stream) {
  return i2t_gemm_ex(A, B, bias, residual, C, M, N, K, lda, ldb, ldc, a_kmajor, b_kmajor, act, accumulate, ab_dtype, res_dtype,
                     c_dtype, 0, stream);
}
