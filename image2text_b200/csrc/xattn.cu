// Cross attention over a SHORT key/value set (the encoder's n_cls summary tokens, S = 8 / 16 / <= 64).
// Replaces nn.MultiheadAttention's attention core at reference models/layers.py:537-542,600-605 and HF GPT-2's
// cross-attention core.  A tcgen05 flash tile is the wrong tool for S <= 64: K and V of one (batch, head) are 4-32 KB,
// so they sit in shared memory and every warp streams query rows against them with shuffle reductions.
#include "common.cuh"

namespace i2t {

constexpr int XA_WARPS = 4, XA_ROWS_PER_WARP = 8, XA_MAXS = 64;

template <typename TIN, typename TOUT, int HS>
__global__ void __launch_bounds__(XA_WARPS * 32)
xattn_fwd_kernel(const TIN* __restrict__ q, const TIN* __restrict__ k, const TIN* __restrict__ v, TOUT* __restrict__ out,
                 float* __restrict__ probs_out, int H, int T, int S, int64_t q_rs, int64_t kv_bs, int64_t kv_rs,
                 float scale) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;                       // [S][HS]
  float* Vs = Ks + S * HS;              // [S][HS]
  float* sc = Vs + S * HS;              // [XA_WARPS][XA_MAXS]
  constexpr int EPL = HS / 32;          // elements per lane
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const TIN* kb = k + b * kv_bs + (int64_t)h * HS;
  const TIN* vb = v + b * kv_bs + (int64_t)h * HS;
  for (int i = t; i < S * HS / 4; i += XA_WARPS * 32) {
    const int s = i / (HS / 4), c = i % (HS / 4);
    *reinterpret_cast<float4*>(&Ks[s * HS + c * 4]) = load4(kb + (int64_t)s * kv_rs + c * 4);
    *reinterpret_cast<float4*>(&Vs[s * HS + c * 4]) = load4(vb + (int64_t)s * kv_rs + c * 4);
  }
  __syncthreads();
  const int row0 = (blockIdx.x * XA_WARPS + w) * XA_ROWS_PER_WARP;
  for (int r = 0; r < XA_ROWS_PER_WARP; ++r) {
    const int ti = row0 + r;
    if (ti >= T) break;
    const TIN* qp = q + (b * T + ti) * q_rs + (int64_t)h * HS;
    float qv[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) qv[e] = to_f32(qp[lane + 32 * e]) * scale;
    float mx = -INFINITY;
    for (int s = 0; s < S; ++s) {
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) d = fmaf(qv[e], Ks[s * HS + lane + 32 * e], d);
      d = warp_sum(d);
      if (lane == 0) sc[w * XA_MAXS + s] = d;
      mx = fmaxf(mx, d);
    }
    __syncwarp();
    float den = 0.f;
    for (int s = 0; s < S; ++s) den += expf(sc[w * XA_MAXS + s] - mx);
    const float inv = 1.0f / den;
    float acc[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
    for (int s = 0; s < S; ++s) {
      const float p = expf(sc[w * XA_MAXS + s] - mx) * inv;
      if (probs_out != nullptr && lane == 0) probs_out[(((int64_t)b * H + h) * T + ti) * S + s] = p;
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[e] = fmaf(p, Vs[s * HS + lane + 32 * e], acc[e]);
    }
    TOUT* op = out + (b * T + ti) * ((int64_t)H * HS) + (int64_t)h * HS;
#pragma unroll
    for (int e = 0; e < EPL; ++e) op[lane + 32 * e] = from_f32<TOUT>(acc[e]);
    __syncwarp();
  }
}

// Backward: grid (T chunks, H, B).  Each warp walks its query rows; dq is written directly, dK/dV partial sums
// are accumulated in shared memory per CTA and flushed with atomicAdd into fp32 buffers the caller zeroed.
template <typename TIN, int HS>
__global__ void __launch_bounds__(XA_WARPS * 32)
xattn_bwd_kernel(const TIN* __restrict__ q, const TIN* __restrict__ k, const TIN* __restrict__ v,
                 const TIN* __restrict__ dout, TIN* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv,
                 int H, int T, int S, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, int64_t dkv_bs, int64_t dkv_rs,
                 float scale) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;
  float* Vs = Ks + S * HS;
  float* dKs = Vs + S * HS;
  float* dVs = dKs + S * HS;
  float* sc = dVs + S * HS;             // [XA_WARPS][2][XA_MAXS]  (p, dp)
  constexpr int EPL = HS / 32;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const TIN* kb = k + b * kv_bs + (int64_t)h * HS;
  const TIN* vb = v + b * kv_bs + (int64_t)h * HS;
  for (int i = t; i < S * HS; i += XA_WARPS * 32) {
    const int s = i / HS, e = i % HS;
    Ks[i] = to_f32(kb[(int64_t)s * kv_rs + e]);
    Vs[i] = to_f32(vb[(int64_t)s * kv_rs + e]);
    dKs[i] = 0.f;
    dVs[i] = 0.f;
  }
  __syncthreads();
  float* ps = sc + w * 2 * XA_MAXS;
  float* dps = ps + XA_MAXS;
  const int row0 = (blockIdx.x * XA_WARPS + w) * XA_ROWS_PER_WARP;
  for (int r = 0; r < XA_ROWS_PER_WARP; ++r) {
    const int ti = row0 + r;
    if (ti >= T) break;
    const TIN* qp = q + (b * T + ti) * q_rs + (int64_t)h * HS;
    const TIN* gp = dout + (b * T + ti) * ((int64_t)H * HS) + (int64_t)h * HS;
    float qv[EPL], gv[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      qv[e] = to_f32(qp[lane + 32 * e]);
      gv[e] = to_f32(gp[lane + 32 * e]);
    }
    float mx = -INFINITY;
    for (int s = 0; s < S; ++s) {
      float d = 0.f, g = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        d = fmaf(qv[e] * scale, Ks[s * HS + lane + 32 * e], d);
        g = fmaf(gv[e], Vs[s * HS + lane + 32 * e], g);
      }
      d = warp_sum(d);
      g = warp_sum(g);
      if (lane == 0) { ps[s] = d; dps[s] = g; }
      mx = fmaxf(mx, d);
    }
    __syncwarp();
    float den = 0.f;
    for (int s = 0; s < S; ++s) den += expf(ps[s] - mx);
    const float inv = 1.0f / den;
    float delta = 0.f;
    for (int s = 0; s < S; ++s) delta = fmaf(expf(ps[s] - mx) * inv, dps[s], delta);
    float aq[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) aq[e] = 0.f;
    for (int s = 0; s < S; ++s) {
      const float p = expf(ps[s] - mx) * inv;
      const float ds = p * (dps[s] - delta) * scale;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int ee = lane + 32 * e;
        aq[e] = fmaf(ds, Ks[s * HS + ee], aq[e]);
        atomicAdd(&dKs[s * HS + ee], ds * qv[e]);
        atomicAdd(&dVs[s * HS + ee], p * gv[e]);
      }
    }
    TIN* dqp = dq + (b * T + ti) * q_rs + (int64_t)h * HS;
#pragma unroll
    for (int e = 0; e < EPL; ++e) dqp[lane + 32 * e] = from_f32<TIN>(aq[e]);
    __syncwarp();
  }
  __syncthreads();
  for (int i = t; i < S * HS; i += XA_WARPS * 32) {
    const int s = i / HS, e = i % HS;
    atomicAdd(dk + b * dkv_bs + (int64_t)s * dkv_rs + (int64_t)h * HS + e, dKs[i]);
    atomicAdd(dv + b * dkv_bs + (int64_t)s * dkv_rs + (int64_t)h * HS + e, dVs[i]);
  }
}

}  // namespace i2t

using namespace i2t;

extern "C" int i2t_xattn_fwd(const void* q, const void* k, const void* v, void* out, int64_t B, int64_t H, int64_t T,
                             int64_t S, int64_t head_dim, int64_t q_row_stride, int64_t kv_batch_stride,
                             int64_t kv_row_stride, int in_dtype, int out_dtype, void* stream) {
  I2T_REQUIRE(q && k && v && out, "xattn_fwd: null pointer");
  I2T_REQUIRE(B > 0 && H > 0 && T > 0 && S > 0 && S <= XA_MAXS && B <= 65535 && H <= 65535, "xattn_fwd: bad sizes (S <= 64)");
  I2T_REQUIRE(head_dim == 64 || head_dim == 32, "xattn_fwd: head_dim %lld not built (32, 64)", (long long)head_dim);
  I2T_REQUIRE(kv_row_stride % 4 == 0 && kv_batch_stride % 4 == 0, "xattn_fwd: kv strides must be multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div(T, XA_WARPS * XA_ROWS_PER_WARP), (unsigned)H, (unsigned)B);
  const size_t smem = (size_t)(2 * S * head_dim + XA_WARPS * XA_MAXS) * sizeof(float);
  const float scale = 1.0f / sqrtf((float)head_dim);
#define I2T_XA(TI, TO, HSV)                                                                                          \
  do {                                                                                                               \
    xattn_fwd_kernel<TI, TO, HSV><<<grid, XA_WARPS * 32, smem, st>>>((const TI*)q, (const TI*)k, (const TI*)v, (TO*)out, \
                                                                     nullptr, (int)H, (int)T, (int)S, q_row_stride,  \
                                                                     kv_batch_stride, kv_row_stride, scale);         \
    I2T_LAUNCHED();                                                                                                  \
    return I2T_OK;                                                                                                   \
  } while (0)
  if (head_dim == 64) {
    if (in_dtype == I2T_F32 && out_dtype == I2T_F32) I2T_XA(float, float, 64);
    if (in_dtype == I2T_BF16 && out_dtype == I2T_BF16) I2T_XA(__nv_bfloat16, __nv_bfloat16, 64);
  } else {
    if (in_dtype == I2T_F32 && out_dtype == I2T_F32) I2T_XA(float, float, 32);
    if (in_dtype == I2T_BF16 && out_dtype == I2T_BF16) I2T_XA(__nv_bfloat16, __nv_bfloat16, 32);
  }
#undef I2T_XA
  return fail(I2T_ERR_INVALID, "xattn_fwd: dtype combination (%d,%d) not built", in_dtype, out_dtype);
}

extern "C" int i2t_xattn_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, float* dk, float* dv,
                             int64_t B, int64_t H, int64_t T, int64_t S, int64_t head_dim, int64_t q_row_stride,
                             int64_t kv_batch_stride, int64_t kv_row_stride, int64_t dkv_batch_stride,
                             int64_t dkv_row_stride, int dtype, void* stream) {
  I2T_REQUIRE(q && k && v && dout && dq && dk && dv, "xattn_bwd: null pointer");
  I2T_REQUIRE(B > 0 && H > 0 && T > 0 && S > 0 && S <= XA_MAXS && B <= 65535 && H <= 65535, "xattn_bwd: bad sizes (S <= 64)");
  I2T_REQUIRE(head_dim == 64 || head_dim == 32, "xattn_bwd: head_dim %lld not built (32, 64)", (long long)head_dim);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div(T, XA_WARPS * XA_ROWS_PER_WARP), (unsigned)H, (unsigned)B);
  const size_t smem = (size_t)(4 * S * head_dim + XA_WARPS * 2 * XA_MAXS) * sizeof(float);
  const float scale = 1.0f / sqrtf((float)head_dim);
#define I2T_XAB(TI, HSV)                                                                                             \
  do {                                                                                                               \
    auto kern = xattn_bwd_kernel<TI, HSV>;                                                                           \
    I2T_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
    kern<<<grid, XA_WARPS * 32, smem, st>>>((const TI*)q, (const TI*)k, (const TI*)v, (const TI*)dout, (TI*)dq, dk, dv, \
                                            (int)H, (int)T, (int)S, q_row_stride, kv_batch_stride, kv_row_stride,    \
                                            dkv_batch_stride, dkv_row_stride, scale);                                \
    I2T_LAUNCHED();                                                                                                  \
    return I2T_OK;                                                                                                   \
  } while (0)
  if (dtype == I2T_F32) {
    if (head_dim == 64) I2T_XAB(float, 64);
    I2T_XAB(float, 32);
  } else if (dtype == I2T_BF16) {
    if (head_dim == 64) I2T_XAB(__nv_bfloat16, 64);
    I2T_XAB(__nv_bfloat16, 32);
  }
#undef I2T_XAB
  return fail(I2T_ERR_INVALID, "xattn_bwd: bad dtype %d", dtype);
}
