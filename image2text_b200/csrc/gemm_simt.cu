// Exact-fp32 tiled GEMM on the FMA pipe: the parity anchor for every dense contraction on the path
// (1e-4 relative logits parity in fp32 rules out TF32/BF16 tensor-core rounding), and the
// fallback for shapes the tcgen05 kernel does not take.  128x128x16 CTA tile, 256 threads, 8x8 outputs per
// thread as 2x2 blocks of 4x4 (conflict-free float4 shared-memory reads), register-staged double buffering.
//   C[M,N] = act(op(A) op(B) + bias) + residual         (all four operand layouts: fwd / dgrad / wgrad / Conv1D)
#include "common.cuh"

namespace i2t {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, GEMM_THREADS = 256;

// Fetch 4 elements that are consecutive along the CONTIGUOUS dimension of an operand.
//   kmajor:  element (r, k) at p[r*ld + k]  -> the 4 run along k
//   !kmajor: element (r, k) at p[k*ld + r]  -> the 4 run along r
template <typename T>
__device__ __forceinline__ float4 fetch4(const T* __restrict__ p, int64_t ld, int64_t r, int64_t k, int64_t R,
                                         int64_t K, bool kmajor, bool vec) {
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kmajor) {
    if (r >= R) return o;
    const T* q = p + r * ld + k;
    if (vec && k + 3 < K) return load4(q);
    if (k + 0 < K) o.x = to_f32(q[0]);
    if (k + 1 < K) o.y = to_f32(q[1]);
    if (k + 2 < K) o.z = to_f32(q[2]);
    if (k + 3 < K) o.w = to_f32(q[3]);
  } else {
    if (k >= K) return o;
    const T* q = p + k * ld + r;
    if (vec && r + 3 < R) return load4(q);
    if (r + 0 < R) o.x = to_f32(q[0]);
    if (r + 1 < R) o.y = to_f32(q[1]);
    if (r + 2 < R) o.z = to_f32(q[2]);
    if (r + 3 < R) o.w = to_f32(q[3]);
  }
  return o;
}

template <typename TIN, typename TRES, typename TOUT, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_simt_kernel(const TIN* __restrict__ A, const TIN* __restrict__ B, const float* __restrict__ bias,
                 const TRES* __restrict__ residual, TOUT* __restrict__ C, int64_t M, int64_t N, int64_t K, int64_t lda,
                 int64_t ldb, int64_t ldc, int act, int accumulate, bool vecA, bool vecB) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int ty = t >> 4, tx = t & 15;

  // global->register staging coordinates (2 float4 per operand per thread)
  float4 ra[2], rb[2];
  auto gload = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (A_KMAJOR) {
        const int r = (t >> 2) + 64 * i, kv = (t & 3) * 4;
        ra[i] = fetch4(A, lda, m0 + r, k0 + kv, M, K, true, vecA);
      } else {
        const int k = (t >> 5) + 8 * i, rv = (t & 31) * 4;
        ra[i] = fetch4(A, lda, m0 + rv, k0 + k, M, K, false, vecA);
      }
      if (B_KMAJOR) {
        const int r = (t >> 2) + 64 * i, kv = (t & 3) * 4;
        rb[i] = fetch4(B, ldb, n0 + r, k0 + kv, N, K, true, vecB);
      } else {
        const int k = (t >> 5) + 8 * i, rv = (t & 31) * 4;
        rb[i] = fetch4(B, ldb, n0 + rv, k0 + k, N, K, false, vecB);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (A_KMAJOR) {
        const int r = (t >> 2) + 64 * i, kv = (t & 3) * 4;
        As[buf][kv + 0][r] = ra[i].x; As[buf][kv + 1][r] = ra[i].y;
        As[buf][kv + 2][r] = ra[i].z; As[buf][kv + 3][r] = ra[i].w;
      } else {
        const int k = (t >> 5) + 8 * i, rv = (t & 31) * 4;
        *reinterpret_cast<float4*>(&As[buf][k][rv]) = ra[i];
      }
      if (B_KMAJOR) {
        const int r = (t >> 2) + 64 * i, kv = (t & 3) * 4;
        Bs[buf][kv + 0][r] = rb[i].x; Bs[buf][kv + 1][r] = rb[i].y;
        Bs[buf][kv + 2][r] = rb[i].z; Bs[buf][kv + 3][r] = rb[i].w;
      } else {
        const int k = (t >> 5) + 8 * i, rv = (t & 31) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][k][rv]) = rb[i];
      }
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int64_t ktiles = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int64_t kt = 0; kt < ktiles; ++kt) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < ktiles) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
  const bool vec_out = (N % 4 == 0) && (ldc % 4 == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jb = 0; jb < 2; ++jb) {
      const int64_t n = n0 + (jb == 0 ? tx * 4 : 64 + tx * 4);
      if (n >= N) continue;
      float v[4] = {acc[i][jb * 4 + 0], acc[i][jb * 4 + 1], acc[i][jb * 4 + 2], acc[i][jb * 4 + 3]};
      TOUT* cp = C + m * ldc + n;
      const TRES* rp = residual ? residual + m * ldc + n : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j < N) {
          float x = v[j];
          if (bias) x += bias[n + j];
          x = apply_act(x, act);
          if (rp) x += to_f32(rp[j]);
          if (accumulate) x += to_f32(cp[j]);
          v[j] = x;
        }
      }
      if (vec_out && n + 3 < N) {
        store4(cp, make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) cp[j] = from_f32<TOUT>(v[j]);
      }
    }
  }
}

template <typename TIN, typename TRES, typename TOUT>
static int launch_layout(const void* A, const void* B, const float* bias, const void* residual, void* C, int64_t M,
                         int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_kmajor, int b_kmajor, int act,
                         int accumulate, cudaStream_t st) {
  const int esz = (int)sizeof(TIN);
  const int64_t align = 16 / esz == 4 ? 16 : 8;  // float4 / 4 x bf16
  const bool vecA = (lda % 4 == 0) && ((uintptr_t)A % align == 0) && ((a_kmajor ? K : M) % 4 == 0);
  const bool vecB = (ldb % 4 == 0) && ((uintptr_t)B % align == 0) && ((b_kmajor ? K : N) % 4 == 0);
  dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM)), block(GEMM_THREADS);
#define I2T_GEMM_LAUNCH(AK, BKM)                                                                              \
  gemm_simt_kernel<TIN, TRES, TOUT, AK, BKM><<<grid, block, 0, st>>>((const TIN*)A, (const TIN*)B, bias,      \
                                                                     (const TRES*)residual, (TOUT*)C, M, N, K, lda, \
                                                                     ldb, ldc, act, accumulate, vecA, vecB)
  if (a_kmajor && b_kmajor) I2T_GEMM_LAUNCH(true, true);
  else if (a_kmajor && !b_kmajor) I2T_GEMM_LAUNCH(true, false);
  else if (!a_kmajor && b_kmajor) I2T_GEMM_LAUNCH(false, true);
  else I2T_GEMM_LAUNCH(false, false);
#undef I2T_GEMM_LAUNCH
  I2T_LAUNCHED();
  return I2T_OK;
}

int gemm_simt(const void* A, const void* B, const float* bias, const void* residual, void* C, int64_t M, int64_t N,
              int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_kmajor, int b_kmajor, int act, int accumulate,
              int ab_dtype, int res_dtype, int c_dtype, cudaStream_t st) {
  if (ceil_div(M, BM) > 65535) return fail(I2T_ERR_INVALID, "gemm: M too large for the grid");
  if (ab_dtype == I2T_F32) {
    if (c_dtype == I2T_F32 && res_dtype == I2T_F32)
      return launch_layout<float, float, float>(A, B, bias, residual, C, M, N, K, lda, ldb, ldc, a_kmajor, b_kmajor, act, accumulate, st);
    if (c_dtype == I2T_BF16 && res_dtype == I2T_F32)
      return launch_layout<float, float, __nv_bfloat16>(A, B, bias, residual, C, M, N, K, lda, ldb, ldc, a_kmajor, b_kmajor, act, accumulate, st);
  } else {
    if (c_dtype == I2T_F32 && res_dtype == I2T_F32)
      return launch_layout<__nv_bfloat16, float, float>(A, B, bias, residual, C, M, N, K, lda, ldb, ldc, a_kmajor, b_kmajor, act, accumulate, st);
    if (c_dtype == I2T_BF16 && res_dtype == I2T_F32)
      return launch_layout<__nv_bfloat16, float, __nv_bfloat16>(A, B, bias, residual, C, M, N, K, lda, ldb, ldc, a_kmajor, b_kmajor, act, accumulate, st);
    if (c_dtype == I2T_BF16 && res_dtype == I2T_BF16)
      return launch_layout<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(A, B, bias, residual, C, M, N, K, lda, ldb, ldc, a_kmajor, b_kmajor, act, accumulate, st);
  }
  return fail(I2T_ERR_INVALID, "gemm: dtype combination (ab=%d,res=%d,c=%d) not built", ab_dtype, res_dtype, c_dtype);
}

// out[n] += sum_m X[m,n]: each CTA reduces a 32-column strip over a slab of rows, one atomicAdd per column.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, float* __restrict__ out, int64_t M,
                                                     int64_t N, int64_t ldx, int64_t rows_per_cta) {
  __shared__ float red[8][33];
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * 32 + lane;
  const int64_t mbeg = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t mend = mbeg + rows_per_cta < M ? mbeg + rows_per_cta : M;
  float s = 0.f;
  if (n < N)
    for (int64_t m = mbeg + w; m < mend; m += 8) s += to_f32(X[m * ldx + n]);
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && n < N) {
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tsum += red[i][lane];
    atomicAdd(out + n, tsum);
  }
}

// The same with 16-byte loads: a lane owns VEC = 16 / sizeof(T) adjacent columns, a CTA a strip of 32 * VEC columns (the bias
// gradients of the training step: hundreds of launches over (B*T, 768 .. 3072) bf16 gradients; the scalar kernel read 64 bytes per
// warp and row).  Needs N % VEC == 0, ldx % VEC == 0 and a 16-byte aligned X.
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ X, float* __restrict__ out, int64_t M, int64_t N,
                                                         int64_t ldx, int64_t rows_per_cta) {
  constexpr int VEC = 16 / (int)sizeof(T);
  __shared__ float red[8][32 * VEC + 4];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t n0 = ((int64_t)blockIdx.x * 32 + lane) * VEC;
  const int64_t mbeg = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t mend = mbeg + rows_per_cta < M ? mbeg + rows_per_cta : M;
  float s[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) s[j] = 0.f;
  if (n0 < N) {
#pragma unroll 4
    for (int64_t m = mbeg + w; m < mend; m += 8) {
      const T* p = X + m * ldx + n0;
#pragma unroll
      for (int h = 0; h < VEC / 4; ++h) {
        const float4 v = load4(p + 4 * h);
        s[4 * h] += v.x; s[4 * h + 1] += v.y; s[4 * h + 2] += v.z; s[4 * h + 3] += v.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) red[w][lane * VEC + j] = s[j];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
    const int64_t n = (int64_t)blockIdx.x * 32 * VEC + c;
    if (n < N) {
      float tsum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) tsum += red[i][c];
      atomicAdd(out + n, tsum);
    }
  }
}

}  // namespace i2t

using namespace i2t;

extern "C" int i2t_colsum(const void* X, float* out, int64_t M, int64_t N, int64_t ldx, int x_dtype, void* stream) {
  I2T_REQUIRE(X && out && M >= 0 && N > 0 && valid_dtype(x_dtype), "colsum: bad arguments");
  if (M == 0) return I2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int vec = x_dtype == I2T_F32 ? 4 : 8;
  const bool vec_ok = N % vec == 0 && ldx % vec == 0 && aligned16(X);
  const int64_t strips = ceil_div(N, vec_ok ? 32 * vec : 32);
  int64_t slabs = ceil_div((int64_t)num_sms() * 4, strips);
  if (slabs < 1) slabs = 1;
  int64_t rows_per = ceil_div(M, slabs);
  if (rows_per < 64) rows_per = 64;
  slabs = ceil_div(M, rows_per);
  dim3 grid((unsigned)strips, (unsigned)slabs);
  if (vec_ok) {
    if (x_dtype == I2T_F32)
      I2T_CUDA(launch_pdl(colsum_vec_kernel<float>, grid, dim3(256), 0, st, (const float*)X, out, M, N, ldx, rows_per));
    else
      I2T_CUDA(launch_pdl(colsum_vec_kernel<__nv_bfloat16>, grid, dim3(256), 0, st, (const __nv_bfloat16*)X, out, M, N, ldx, rows_per));
  } else if (x_dtype == I2T_F32) {
    I2T_CUDA(launch_pdl(colsum_kernel<float>, grid, dim3(256), 0, st, (const float*)X, out, M, N, ldx, rows_per));
  } else {
    I2T_CUDA(launch_pdl(colsum_kernel<__nv_bfloat16>, grid, dim3(256), 0, st, (const __nv_bfloat16*)X, out, M, N, ldx, rows_per));
  }
  I2T_LAUNCHED();
  return I2T_OK;
}
