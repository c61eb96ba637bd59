// LayerNorm forward / backward.  HBM-bound: one warp per row, the row lives in registers
// (128-bit loads), statistics by warp shuffle.  Algorithmic bytes: fwd = read x + write y;
// bwd = read dy, x + write dx (+ a column reduction for dgamma/dbeta done in registers per CTA).
// Replaces F.layer_norm at reference models/layers.py:357-358 and torchvision's nn.LayerNorm(eps=1e-6).
#include "common.cuh"

namespace i2t {

template <typename TX, typename TY, int MAXV>
__global__ void __launch_bounds__(128) ln_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, TY* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd,
                                                     int64_t rows, int cols, int64_t xstride, float eps) {
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const TX* xr = x + row * xstride;
  const int nvec = cols >> 2;
  float4 v[MAXV], gm[MAXV], bt[MAXV];
  float s = 0.f;
  // x, gamma and beta are requested together: with a handful of rows (decode: one row per sequence) the kernel is one
  // chain of memory latencies, and the parameters would otherwise wait for both reductions
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      v[i] = load4(xr + c * 4);
      gm[i] = load4(gamma + c * 4);
      bt[i] = beta ? load4(beta + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mu = warp_sum(s) / (float)cols;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float a = v[i].x - mu, b = v[i].y - mu, cc = v[i].z - mu, d = v[i].w - mu;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rs = 1.0f / sqrtf(warp_sum(q) / (float)cols + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
  TY* yr = y + row * (int64_t)cols;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float4 o;
      o.x = (v[i].x - mu) * rs * gm[i].x + bt[i].x;
      o.y = (v[i].y - mu) * rs * gm[i].y + bt[i].y;
      o.z = (v[i].z - mu) * rs * gm[i].z + bt[i].z;
      o.w = (v[i].w - mu) * rs * gm[i].w + bt[i].w;
      store4(yr + c * 4, o);
    }
  }
}

// Backward.  Each CTA (4 warps) walks rows with a grid stride; every lane keeps the dgamma/dbeta partial
// sums of the columns it owns in registers, reduced across the CTA's warps in shared memory and
// flushed with one atomicAdd per column per CTA.
// WG: dgamma / dbeta wanted; without them (frozen LayerNorms, most of nano.yaml) the column accumulators go: 48 registers less,
// 6 CTAs per SM instead of 4.
template <typename TDY, typename TX, typename TDX, int MAXV, bool WG>
__global__ void __launch_bounds__(128, MAXV <= 6 ? (WG ? 4 : 6) : 1) ln_bwd_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const TDX* __restrict__ dx_add,
                                                     TDX* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, int64_t rows, int cols) {
  extern __shared__ float red[];  // [4 warps][cols] reused for dgamma then dbeta
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = cols >> 2;
  float4 ag[WG ? MAXV : 1], ab[WG ? MAXV : 1];
#pragma unroll
  for (int i = 0; i < (WG ? MAXV : 1); ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = (int64_t)blockIdx.x * 4 + warp; row < rows; row += (int64_t)gridDim.x * 4) {
    const TDY* dyr = dy + row * (int64_t)cols;
    const TX* xr = x + row * (int64_t)cols;
    const float mu = mean[row], rs = rstd[row];
    float4 g_[MAXV], xh[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 d = load4(dyr + c * 4), xv = load4(xr + c * 4), g = load4(gamma + c * 4);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        if (WG) {
          ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
          ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
        }
        g_[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        s1 += (g_[i].x + g_[i].y) + (g_[i].z + g_[i].w);
        s2 += (g_[i].x * xh[i].x + g_[i].y * xh[i].y) + (g_[i].z * xh[i].z + g_[i].w * xh[i].w);
      }
    }
    s1 = warp_sum(s1) / (float)cols;
    s2 = warp_sum(s2) / (float)cols;
    TDX* dxr = dx + row * (int64_t)cols;
    const TDX* addr = dx_add ? dx_add + row * (int64_t)cols : nullptr;   // gradient arriving over the residual connection
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        float4 o;
        o.x = rs * (g_[i].x - s1 - xh[i].x * s2);
        o.y = rs * (g_[i].y - s1 - xh[i].y * s2);
        o.z = rs * (g_[i].z - s1 - xh[i].z * s2);
        o.w = rs * (g_[i].w - s1 - xh[i].w * s2);
        if (addr) {
          const float4 a = load4(addr + c * 4);
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        store4(dxr + c * 4, o);
      }
    }
  }
  if (!WG || (dgamma == nullptr && dbeta == nullptr)) return;
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? dgamma : dbeta;
    __syncthreads();
    if (dst != nullptr) {
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) store4(red + warp * cols + c * 4, pass == 0 ? ag[WG ? i : 0] : ab[WG ? i : 0]);
      }
    }
    __syncthreads();
    if (dst != nullptr) {
      for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        const float t = (red[c] + red[cols + c]) + (red[2 * cols + c] + red[3 * cols + c]);
        atomicAdd(dst + c, t);
      }
    }
  }
}

template <typename TX, typename TY>
static int launch_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      int64_t rows, int64_t cols, int64_t xstride, float eps, cudaStream_t st) {
  const int warps = 4;
  dim3 grid((unsigned)ceil_div(rows, warps)), block(warps * 32);
  if (cols <= 1024) {
    I2T_CUDA(launch_pdl(ln_fwd_kernel<TX, TY, 8>, grid, block, 0, st, (const TX*)x, gamma, beta, (TY*)y, mean, rstd, rows, (int)cols,
                        xstride, eps));
  } else {
    I2T_CUDA(launch_pdl(ln_fwd_kernel<TX, TY, 16>, grid, block, 0, st, (const TX*)x, gamma, beta, (TY*)y, mean, rstd, rows, (int)cols,
                        xstride, eps));
  }
  I2T_LAUNCHED();
  return I2T_OK;
}

template <typename TDY, typename TX, typename TDX>
static int launch_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const void* dx_add,
                      void* dx, float* dgamma, float* dbeta, int64_t rows, int64_t cols, cudaStream_t st) {
  int64_t ctas = ceil_div(rows, 4);
  const int64_t cap = (int64_t)num_sms() * ((dgamma != nullptr || dbeta != nullptr) ? 4 : 6);
  if (ctas > cap) ctas = cap;
  const bool wg = dgamma != nullptr || dbeta != nullptr;
  const size_t smem = wg ? (size_t)4 * cols * sizeof(float) : 0;
#define I2T_LNB(MV, WGV)                                                                                                              \
  I2T_CUDA(launch_pdl(ln_bwd_kernel<TDY, TX, TDX, MV, WGV>, dim3((unsigned)ctas), dim3(128), smem, st, (const TDY*)dy, (const TX*)x, gamma, \
                      mean, rstd, (const TDX*)dx_add, (TDX*)dx, dgamma, dbeta, rows, (int)cols))
  if (cols <= 768) {      // the model width: exactly 6 vectors per lane
    if (wg) I2T_LNB(6, true); else I2T_LNB(6, false);
  } else if (cols <= 1024) {
    if (wg) I2T_LNB(8, true); else I2T_LNB(8, false);
  } else {
    if (wg) I2T_LNB(16, true); else I2T_LNB(16, false);
  }
#undef I2T_LNB
  I2T_LAUNCHED();
  return I2T_OK;
}

}  // namespace i2t

using namespace i2t;

extern "C" int i2t_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                 int64_t rows, int64_t cols, int64_t x_row_stride, float eps, int x_dtype, int y_dtype,
                                 void* stream) {
  I2T_REQUIRE(x && gamma && y, "layernorm_fwd: null pointer");
  I2T_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0 && cols <= 2048, "layernorm_fwd: cols=%lld must be a multiple of 4, <= 2048",
              (long long)cols);
  I2T_REQUIRE(valid_dtype(x_dtype) && valid_dtype(y_dtype), "layernorm_fwd: bad dtype");
  I2T_REQUIRE(x_row_stride % 4 == 0 && aligned16(gamma) && (beta == nullptr || aligned16(beta)), "layernorm_fwd: alignment");
  I2T_REQUIRE(((uintptr_t)x % (x_dtype == I2T_F32 ? 16 : 8)) == 0 && ((uintptr_t)y % (y_dtype == I2T_F32 ? 16 : 8)) == 0,
              "layernorm_fwd: x/y alignment");
  if (rows == 0) return I2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == I2T_F32 && y_dtype == I2T_F32) return launch_fwd<float, float>(x, gamma, beta, y, mean, rstd, rows, cols, x_row_stride, eps, st);
  if (x_dtype == I2T_F32 && y_dtype == I2T_BF16) return launch_fwd<float, __nv_bfloat16>(x, gamma, beta, y, mean, rstd, rows, cols, x_row_stride, eps, st);
  if (x_dtype == I2T_BF16 && y_dtype == I2T_F32) return launch_fwd<__nv_bfloat16, float>(x, gamma, beta, y, mean, rstd, rows, cols, x_row_stride, eps, st);
  return launch_fwd<__nv_bfloat16, __nv_bfloat16>(x, gamma, beta, y, mean, rstd, rows, cols, x_row_stride, eps, st);
}

static int layernorm_bwd_any(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                             const void* dx_add, void* dx, float* dgamma, float* dbeta, int64_t rows, int64_t cols, int dy_dtype,
                             int x_dtype, int dx_dtype, void* stream) {
  I2T_REQUIRE(dy && x && gamma && mean && rstd && dx, "layernorm_bwd: null pointer");
  I2T_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0 && cols <= 2048, "layernorm_bwd: cols=%lld unsupported", (long long)cols);
  I2T_REQUIRE(valid_dtype(dy_dtype) && valid_dtype(x_dtype) && valid_dtype(dx_dtype), "layernorm_bwd: bad dtype");
  if (rows == 0) return I2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int key = dy_dtype * 4 + x_dtype * 2 + dx_dtype;
  switch (key) {
    case 0: return launch_bwd<float, float, float>(dy, x, gamma, mean, rstd, dx_add, dx, dgamma, dbeta, rows, cols, st);
    case 1: return launch_bwd<float, float, __nv_bfloat16>(dy, x, gamma, mean, rstd, dx_add, dx, dgamma, dbeta, rows, cols, st);
    case 5: return launch_bwd<__nv_bfloat16, float, __nv_bfloat16>(dy, x, gamma, mean, rstd, dx_add, dx, dgamma, dbeta, rows, cols, st);
    case 4: return launch_bwd<__nv_bfloat16, float, float>(dy, x, gamma, mean, rstd, dx_add, dx, dgamma, dbeta, rows, cols, st);
    default: return fail(I2T_ERR_INVALID, "layernorm_bwd: dtype combination (%d,%d,%d) not built", dy_dtype, x_dtype, dx_dtype);
  }
}

extern "C" int i2t_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                 void* dx, float* dgamma, float* dbeta, int64_t rows, int64_t cols, int dy_dtype,
                                 int x_dtype, int dx_dtype, void* stream) {
  return layernorm_bwd_any(dy, x, gamma, mean, rstd, nullptr, dx, dgamma, dbeta, rows, cols, dy_dtype, x_dtype, dx_dtype, stream);
}

extern "C" int i2t_layernorm_bwd_add(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                     const void* dx_add, void* dx, float* dgamma, float* dbeta, int64_t rows, int64_t cols,
                                     int dy_dtype, int x_dtype, int dx_dtype, void* stream) {
  I2T_REQUIRE(dx_add, "layernorm_bwd_add: null dx_add");
  return layernorm_bwd_any(dy, x, gamma, mean, rstd, dx_add, dx, dgamma, dbeta, rows, cols, dy_dtype, x_dtype, dx_dtype, stream);
}
