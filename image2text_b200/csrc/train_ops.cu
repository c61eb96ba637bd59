// Memory-bound training-side kernels:
//   activation forward/backward  (nn.GELU('tanh') reference models/layers.py:477,483; torchvision nn.GELU(); HF gelu_new)
//   embedding backward           (scatter-add into the tied wte, reference models/decoder.py:234 autograd)
//   normalize_gradients          (reference models/functions.py:19-24: g / (||g||_2 + 1e-6) over the WHOLE tensor)
//   weighted / distilled LM loss (reference training/wrapper.py:80-96,120-151) forward + dlogits in ONE pass over V
#include "common.cuh"

namespace i2t {

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) act_fwd_kernel(const TI* __restrict__ z, TO* __restrict__ h, int64_t n4, int act) {
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = load4(z + i * 4);
    v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act);
    store4(h + i * 4, v);
  }
}

template <typename TZ, typename TG>
__global__ void __launch_bounds__(256) act_bwd_kernel(const TZ* __restrict__ z, const TG* __restrict__ dh,
                                                      TG* __restrict__ dz, int64_t n4, int act) {
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = load4(z + i * 4), g = load4(dh + i * 4);
    store4(dz + i * 4, make_float4(g.x * act_grad(v.x, act), g.y * act_grad(v.y, act), g.z * act_grad(v.z, act),
                                   g.w * act_grad(v.w, act)));
  }
}

// bf16 -> bf16 variants (the training path's MLP activations): 16-byte vectors (8 elements per thread and step) and, for the tanh
// GELU, the hardware tanh (tanh.approx.f32, relative error ~2^-11: below bf16 resolution) -- the precise tanhf made these kernels
// ALU-bound (53 / 70 us for 16384 x 3072 at ~60 % of the HBM rate).  fp32 inputs keep the precise functions (1e-4 parity anchor).
__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float act_fast(float x, int act) {
  if (act == I2T_ACT_GELU_TANH) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    return 0.5f * x * (1.0f + fast_tanh(k0 * (x + k1 * x * x * x)));
  }
  return apply_act(x, act);
}
__device__ __forceinline__ float act_grad_fast(float x, int act) {
  if (act == I2T_ACT_GELU_TANH) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float t = fast_tanh(k0 * (x + k1 * x * x * x));
    return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * x * x);
  }
  return act_grad(x, act);
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return r;
}
__global__ void __launch_bounds__(256) act_fwd_bf16_kernel(const uint4* __restrict__ z, uint4* __restrict__ h, int64_t n8, int act) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(z[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = act_fast(f[j], act);
    h[i] = pack8(f);
  }
}
__global__ void __launch_bounds__(256) act_bwd_bf16_kernel(const uint4* __restrict__ z, const uint4* __restrict__ dh,
                                                           uint4* __restrict__ dz, int64_t n8, int act) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8], g[8];
    unpack8(z[i], f);
    unpack8(dh[i], g);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = g[j] * act_grad_fast(f[j], act);
    dz[i] = pack8(f);
  }
}

// dwte[ids[b,s], :] += dx[b, n_prompt + s, :]   for n_prompt + s < T
__global__ void __launch_bounds__(256) embed_bwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ dx,
                                                        float* __restrict__ dwte, int64_t B, int T, int n_prompt, int S,
                                                        int C, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = C / 4;
  const int cv = (int)(i % c4);
  const int64_t r = i / c4;
  const int tt = (int)(r % T);
  const int64_t b = r / T;
  if (tt < n_prompt) return;
  const int64_t tok = ids[b * S + (tt - n_prompt)];
  const float4 g = load4(dx + r * C + cv * 4);
  float* dst = dwte + tok * C + cv * 4;
  atomicAdd(dst + 0, g.x); atomicAdd(dst + 1, g.y); atomicAdd(dst + 2, g.z); atomicAdd(dst + 3, g.w);
}

// sum of squares -> *acc (fp64 accumulation across CTAs keeps the norm reproducible to fp32 rounding)
template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ g, double* __restrict__ acc, int64_t n4) {
  __shared__ double red[8];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = load4(g + i * 4);
    s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(acc, t);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gradnorm_scale_kernel(const T* __restrict__ g, T* __restrict__ out,
                                                             const double* __restrict__ acc, int64_t n4) {
  const float inv = 1.0f / ((float)sqrt(*acc) + 1e-6f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = load4(g + i * 4);
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    store4(out + i * 4, v);
  }
}

// ---- loss weights: training/wrapper.py:80-96.  One CTA per batch row (sequence length <= a few thousand). ----
__global__ void __launch_bounds__(256) loss_weights_kernel(const int64_t* __restrict__ labels, float* __restrict__ w,
                                                           int Tl, int64_t ld_labels, int inv_sqrt_pos,
                                                           float eos_weight, int use_eos_weight, int64_t eos_id,
                                                           int64_t ignore_index, int B) {
  __shared__ float red[8];
  __shared__ float s_tot;
  const int b = blockIdx.x;
  float part = 0.f;
  for (int t = threadIdx.x; t < Tl; t += blockDim.x) {
    const int64_t y = labels[(int64_t)b * ld_labels + t];
    float v = inv_sqrt_pos ? 1.0f / sqrtf((float)(t + 1)) : 1.0f;
    if (use_eos_weight && y == eos_id) v = eos_weight;
    if (y == ignore_index) v = 0.f;
    w[(int64_t)b * Tl + t] = v;
    part += v;
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    s_tot = t;
  }
  __syncthreads();
  const float inv = 1.0f / (1e-3f + s_tot) / (float)B;
  for (int t = threadIdx.x; t < Tl; t += blockDim.x) w[(int64_t)b * Tl + t] *= inv;
}

// One CTA per (b, t) row of the vocabulary.  s = z / tau.
//   plain   : L = -w * (s_y - lse)                                      dz_v = w/tau * (p_v - 1[v=y])
//   distill : L = -w * [alpha * (sum_v q_v s_v - lse) + (1-alpha) * valid * (s_y - lse)]
//             dz_v = -w/tau * [alpha * (q_v - p_v) + (1-alpha) * valid * (1[v=y] - p_v)],   q = softmax(z_m / tau)
// Rows with w == 0 write zero gradients and add nothing.  loss_rows[row] receives the row's loss (summed on host side
// by a second tiny kernel to stay deterministic).
template <typename TL>
__global__ void __launch_bounds__(512)
lm_loss_kernel(const TL* __restrict__ logits, const TL* __restrict__ teacher, const int64_t* __restrict__ labels,
               const float* __restrict__ w, float* __restrict__ loss_rows, TL* __restrict__ dlogits, int V, int T_logits,
               int Tl, int64_t ld_labels, float inv_tau, float alpha, int64_t ignore_index, int64_t ld_logits,
               int64_t ld_teacher, float grad_scale) {
  __shared__ float redf[16];
  __shared__ float s_a, s_b;
  const int row = blockIdx.x;                 // row = b * Tl + t  (only the first Tl positions of each sequence)
  const int b = row / Tl, t = row % Tl;
  const TL* z = logits + ((int64_t)b * T_logits + t) * ld_logits;
  const TL* zm = teacher ? teacher + ((int64_t)b * T_logits + t) * ld_teacher : nullptr;
  TL* dz = dlogits ? dlogits + ((int64_t)b * T_logits + t) * ld_logits : nullptr;
  const float wt = w[row];
  const int64_t y = labels[(int64_t)b * ld_labels + t];
  const bool valid = y != ignore_index;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  if (wt == 0.f) {
    if (dz)
      for (int v = tid; v < V; v += blockDim.x) dz[v] = from_f32<TL>(0.f);
    if (tid == 0) loss_rows[row] = 0.f;
    return;
  }
  auto block_max = [&](float v) {
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) redf[wid] = v;
    __syncthreads();
    float r = redf[0];
    for (int i = 1; i < nw; ++i) r = fmaxf(r, redf[i]);
    return r;
  };
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) redf[wid] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < nw; ++i) r += redf[i];
    return r;
  };
  float mx = -INFINITY, mxm = -INFINITY;
  for (int v = tid; v < V; v += blockDim.x) {
    mx = fmaxf(mx, to_f32(z[v]) * inv_tau);
    if (zm) mxm = fmaxf(mxm, to_f32(zm[v]) * inv_tau);
  }
  mx = block_max(mx);
  if (zm) mxm = block_max(mxm);
  float se = 0.f, sem = 0.f, qs = 0.f;
  for (int v = tid; v < V; v += blockDim.x) {
    const float s = to_f32(z[v]) * inv_tau;
    se += expf(s - mx);
    if (zm) {
      const float e = expf(to_f32(zm[v]) * inv_tau - mxm);
      sem += e;
      qs = fmaf(e, s, qs);
    }
  }
  se = block_sum(se);
  if (zm) { sem = block_sum(sem); qs = block_sum(qs); }
  const float lse = mx + logf(se);
  const float sy = valid ? to_f32(z[y]) * inv_tau : 0.f;
  float loss;
  if (zm) loss = -wt * (alpha * (qs / sem - lse) + (1.f - alpha) * (valid ? (sy - lse) : 0.f));
  else loss = valid ? -wt * (sy - lse) : 0.f;
  if (tid == 0) loss_rows[row] = loss;
  if (dz) {
    const float scale = wt * inv_tau * grad_scale;
    const float hard = zm ? (1.f - alpha) * (valid ? 1.f : 0.f) : (valid ? 1.f : 0.f);
    const float soft = zm ? alpha : 0.f;
    const float inv_sem = zm ? 1.0f / sem : 0.f;
    for (int v = tid; v < V; v += blockDim.x) {
      const float p = expf(to_f32(z[v]) * inv_tau - lse);
      float g = (soft + hard) * p;
      if (zm) g -= soft * expf(to_f32(zm[v]) * inv_tau - mxm) * inv_sem;
      if (v == y) g -= hard;
      dz[v] = from_f32<TL>(scale * g);
    }
  }
}

__global__ void __launch_bounds__(256) sum_rows_kernel(const float* __restrict__ x, float* __restrict__ out, int n) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    *out = (float)t;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scale_inplace_kernel(T* __restrict__ x, const float* __restrict__ scale_ptr,
                                                            int64_t n) {
  const float s = *scale_ptr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = from_f32<T>(to_f32(x[i]) * s);
}

static unsigned grid_for(int64_t n_items, int per_block = 256) {
  int64_t g = ceil_div(n_items, per_block);
  const int64_t cap = (int64_t)num_sms() * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace i2t

using namespace i2t;

extern "C" int i2t_act_fwd(const void* z, void* h, int64_t n, int act, int z_dtype, int h_dtype, void* stream) {
  I2T_REQUIRE(z && h && n >= 0 && n % 4 == 0 && valid_dtype(z_dtype) && valid_dtype(h_dtype), "act_fwd: bad arguments");
  if (n == 0) return I2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(n / 4);
  const dim3 gd(g), bd(256);
  const int64_t n4 = n / 4;
  if (z_dtype == I2T_F32 && h_dtype == I2T_F32) I2T_CUDA(launch_pdl(act_fwd_kernel<float, float>, gd, bd, 0, st, (const float*)z, (float*)h, n4, act));
  else if (z_dtype == I2T_F32) I2T_CUDA(launch_pdl(act_fwd_kernel<float, __nv_bfloat16>, gd, bd, 0, st, (const float*)z, (__nv_bfloat16*)h, n4, act));
  else if (h_dtype == I2T_BF16 && n % 8 == 0 && aligned16(z) && aligned16(h))
    I2T_CUDA(launch_pdl(act_fwd_bf16_kernel, dim3(grid_for(n / 8)), bd, 0, st, (const uint4*)z, (uint4*)h, n / 8, act));
  else if (h_dtype == I2T_BF16) I2T_CUDA(launch_pdl(act_fwd_kernel<__nv_bfloat16, __nv_bfloat16>, gd, bd, 0, st, (const __nv_bfloat16*)z, (__nv_bfloat16*)h, n4, act));
  else I2T_CUDA(launch_pdl(act_fwd_kernel<__nv_bfloat16, float>, gd, bd, 0, st, (const __nv_bfloat16*)z, (float*)h, n4, act));
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_act_bwd(const void* z, const void* dh, void* dz, int64_t n, int act, int z_dtype, int g_dtype,
                           void* stream) {
  I2T_REQUIRE(z && dh && dz && n >= 0 && n % 4 == 0 && valid_dtype(z_dtype) && valid_dtype(g_dtype), "act_bwd: bad arguments");
  if (n == 0) return I2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(n / 4);
  const dim3 gd(g), bd(256);
  const int64_t n4 = n / 4;
  if (z_dtype == I2T_F32 && g_dtype == I2T_F32) I2T_CUDA(launch_pdl(act_bwd_kernel<float, float>, gd, bd, 0, st, (const float*)z, (const float*)dh, (float*)dz, n4, act));
  else if (z_dtype == I2T_BF16 && g_dtype == I2T_BF16 && n % 8 == 0 && aligned16(z) && aligned16(dh) && aligned16(dz))
    I2T_CUDA(launch_pdl(act_bwd_bf16_kernel, dim3(grid_for(n / 8)), bd, 0, st, (const uint4*)z, (const uint4*)dh, (uint4*)dz, n / 8, act));
  else if (z_dtype == I2T_BF16 && g_dtype == I2T_BF16) I2T_CUDA(launch_pdl(act_bwd_kernel<__nv_bfloat16, __nv_bfloat16>, gd, bd, 0, st, (const __nv_bfloat16*)z, (const __nv_bfloat16*)dh, (__nv_bfloat16*)dz, n4, act));
  else return fail(I2T_ERR_INVALID, "act_bwd: dtype combination not built");
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_embed_bwd(const int64_t* ids, const float* dx, float* dwte, int64_t B, int64_t T, int64_t n_prompt,
                             int64_t S, int64_t C, void* stream) {
  I2T_REQUIRE(ids && dx && dwte && B > 0 && T > 0 && C % 4 == 0, "embed_bwd: bad arguments");
  const int64_t total4 = B * T * C / 4;
  embed_bwd_kernel<<<(unsigned)ceil_div(total4, 256), 256, 0, (cudaStream_t)stream>>>(ids, dx, dwte, B, (int)T, (int)n_prompt,
                                                                                   (int)S, (int)C, total4);
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_gradnorm_scale(const void* g, void* out, double* acc, int64_t n, int dtype, void* stream) {
  I2T_REQUIRE(g && out && acc && n > 0 && n % 4 == 0 && valid_dtype(dtype), "gradnorm_scale: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  I2T_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
  const unsigned gr = grid_for(n / 4);
  if (dtype == I2T_F32) {
    sumsq_kernel<float><<<gr, 256, 0, st>>>((const float*)g, acc, n / 4);
    I2T_LAUNCHED();
    gradnorm_scale_kernel<float><<<gr, 256, 0, st>>>((const float*)g, (float*)out, acc, n / 4);
  } else {
    sumsq_kernel<__nv_bfloat16><<<gr, 256, 0, st>>>((const __nv_bfloat16*)g, acc, n / 4);
    I2T_LAUNCHED();
    gradnorm_scale_kernel<__nv_bfloat16><<<gr, 256, 0, st>>>((const __nv_bfloat16*)g, (__nv_bfloat16*)out, acc, n / 4);
  }
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_lm_loss(const void* logits, const void* teacher_logits, const int64_t* labels, float* weights,
                           float* loss_rows, float* loss_out, void* dlogits, int64_t B, int64_t T_logits, int64_t Tl,
                           int64_t V, int64_t ld_labels, float temperature, float alpha, int inv_sqrt_position,
                           int use_eos_weight, float eos_weight, int64_t eos_id, int64_t ignore_index, int64_t ld_logits,
                           int64_t ld_teacher, float grad_scale, int dtype, void* stream) {
  I2T_REQUIRE(logits && labels && weights && loss_rows && loss_out, "lm_loss: null pointer");
  I2T_REQUIRE(B > 0 && Tl > 0 && Tl <= T_logits && Tl <= ld_labels && V > 1 && temperature > 0.f && valid_dtype(dtype),
              "lm_loss: bad sizes");
  I2T_REQUIRE(ld_logits >= V && (teacher_logits == nullptr || ld_teacher >= V), "lm_loss: row pitch smaller than V");
  cudaStream_t st = (cudaStream_t)stream;
  loss_weights_kernel<<<(unsigned)B, 256, 0, st>>>(labels, weights, (int)Tl, ld_labels, inv_sqrt_position, eos_weight,
                                                  use_eos_weight, eos_id, ignore_index, (int)B);
  I2T_LAUNCHED();
  if (dtype == I2T_F32)
    lm_loss_kernel<float><<<(unsigned)(B * Tl), 512, 0, st>>>((const float*)logits, (const float*)teacher_logits, labels,
                                                              weights, loss_rows, (float*)dlogits, (int)V, (int)T_logits,
                                                              (int)Tl, ld_labels, 1.0f / temperature, alpha, ignore_index, ld_logits, ld_teacher, grad_scale);
  else
    lm_loss_kernel<__nv_bfloat16><<<(unsigned)(B * Tl), 512, 0, st>>>(
        (const __nv_bfloat16*)logits, (const __nv_bfloat16*)teacher_logits, labels, weights, loss_rows,
        (__nv_bfloat16*)dlogits, (int)V, (int)T_logits, (int)Tl, ld_labels, 1.0f / temperature, alpha, ignore_index, ld_logits, ld_teacher, grad_scale);
  I2T_LAUNCHED();
  sum_rows_kernel<<<1, 256, 0, st>>>(loss_rows, loss_out, (int)(B * Tl));
  I2T_LAUNCHED();
  return I2T_OK;
}

// ---- contrastive auxiliary loss: training/wrapper.py:98-118 after its (B*L, C) x (C, B*L) similarity GEMM -------------
// One CTA per row r of pred (R x R fp32, R = B * L): cross entropy of pred[r, :] / tau against target r over the VALID columns
// (labels != ignore_index; the reference writes -inf into the others), times the loss weight of row r; an infinite loss (the
// target column itself is masked) counts as 0, like the reference's isinf filter.  dpred (optional) = d loss / d pred.
namespace i2t {
__global__ void __launch_bounds__(256) contrastive_ce_kernel(const float* __restrict__ pred, const int64_t* __restrict__ labels,
                                                             const float* __restrict__ weights, float* __restrict__ loss_rows,
                                                             float* __restrict__ dpred, int R, int L, int64_t ld_labels,
                                                             float inv_tau, int64_t ignore_index, float grad_scale) {
  __shared__ float red[8];
  __shared__ float s_bcast;
  const int r = blockIdx.x, t = threadIdx.x;
  const float* row = pred + (int64_t)r * R;
  auto valid = [&](int c) { return labels[(int64_t)(c / L) * ld_labels + (c % L)] != ignore_index; };
  float mx = -INFINITY;
  for (int c = t; c < R; c += 256)
    if (valid(c)) mx = fmaxf(mx, row[c] * inv_tau);
  mx = warp_max(mx);
  if ((t & 31) == 0) red[t >> 5] = mx;
  __syncthreads();
  if (t == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    s_bcast = m;
  }
  __syncthreads();
  mx = s_bcast;
  __syncthreads();
  float sum = 0.f;
  for (int c = t; c < R; c += 256)
    if (valid(c)) sum += expf(row[c] * inv_tau - mx);
  sum = warp_sum(sum);
  if ((t & 31) == 0) red[t >> 5] = sum;
  __syncthreads();
  if (t == 0) {
    float s2 = 0.f;
    for (int i = 0; i < 8; ++i) s2 += red[i];
    s_bcast = s2;
  }
  __syncthreads();
  sum = s_bcast;
  const bool finite = valid(r) && sum > 0.f;          // target column masked (or nothing valid): the loss is inf -> dropped
  const float w = weights[r];
  const float lse = mx + logf(sum);
  if (t == 0) loss_rows[r] = finite ? w * (lse - row[r] * inv_tau) : 0.f;
  if (dpred != nullptr) {
    float* drow = dpred + (int64_t)r * R;
    const float k = finite ? w * inv_tau * grad_scale : 0.f;
    for (int c = t; c < R; c += 256) {
      float g = 0.f;
      if (finite && valid(c)) g = k * (expf(row[c] * inv_tau - lse) - (c == r ? 1.f : 0.f));
      drow[c] = g;
    }
  }
}
}  // namespace i2t

// pred (R, R) fp32 with R = B * L; labels (B, ld_labels) int64; weights (R) / loss_rows (R) fp32 scratch; loss_out: device scalar
extern "C" int i2t_contrastive_loss(const float* pred, const int64_t* labels, float* weights, float* loss_rows, float* loss_out,
                                    float* dpred, int64_t B, int64_t L, int64_t ld_labels, float temperature,
                                    int inv_sqrt_position, int use_eos_weight, float eos_weight, int64_t eos_id,
                                    int64_t ignore_index, float grad_scale, void* stream) {
  I2T_REQUIRE(pred && labels && weights && loss_rows && loss_out, "contrastive_loss: null pointer");
  I2T_REQUIRE(B > 0 && L > 0 && L <= ld_labels && temperature > 0.f && B * L < (1ll << 30), "contrastive_loss: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  loss_weights_kernel<<<(unsigned)B, 256, 0, st>>>(labels, weights, (int)L, ld_labels, inv_sqrt_position, eos_weight, use_eos_weight,
                                                  eos_id, ignore_index, (int)B);
  I2T_LAUNCHED();
  contrastive_ce_kernel<<<(unsigned)(B * L), 256, 0, st>>>(pred, labels, weights, loss_rows, dpred, (int)(B * L), (int)L, ld_labels,
                                                          1.0f / temperature, ignore_index, grad_scale);
  I2T_LAUNCHED();
  sum_rows_kernel<<<1, 256, 0, st>>>(loss_rows, loss_out, (int)(B * L));
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_scale_inplace(void* x, const float* scale_ptr, int64_t n, int dtype, void* stream) {
  I2T_REQUIRE(x && scale_ptr && n >= 0 && valid_dtype(dtype), "scale_inplace: bad arguments");
  if (n == 0) return I2T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == I2T_F32) scale_inplace_kernel<float><<<grid_for(n), 256, 0, st>>>((float*)x, scale_ptr, n);
  else scale_inplace_kernel<__nv_bfloat16><<<grid_for(n), 256, 0, st>>>((__nv_bfloat16*)x, scale_ptr, n);
  I2T_LAUNCHED();
  return I2T_OK;
}

// ---- row-wise L2 normalisation: F.normalize(x, p=2, dim=-1) at reference models/encoder.py:118-119 ------------------
namespace i2t {
__global__ void __launch_bounds__(128) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows,
                                                         int cols, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 4 + warp;
  if (r >= rows) return;
  const float* xr = x + r * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s = fmaf(xr[c], xr[c], s);
  const float n = fmaxf(sqrtf(warp_sum(s)), eps);
  for (int c = lane; c < cols; c += 32) y[r * cols + c] = xr[c] / n;
}

// y = x / n, n = max(||x||, eps):  dx = dy / n - x * (x . dy) / n^3   (the second term vanishes when the clamp is active)
__global__ void __launch_bounds__(128) l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dx, int64_t rows, int cols, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 4 + warp;
  if (r >= rows) return;
  const float* xr = x + r * cols;
  const float* gr = dy + r * cols;
  float s = 0.f, d = 0.f;
  for (int c = lane; c < cols; c += 32) {
    s = fmaf(xr[c], xr[c], s);
    d = fmaf(xr[c], gr[c], d);
  }
  s = warp_sum(s);
  d = warp_sum(d);
  const float nrm = sqrtf(s);
  const float n = fmaxf(nrm, eps);
  const float k = nrm >= eps ? d / (n * n * n) : 0.f;
  for (int c = lane; c < cols; c += 32) dx[r * cols + c] = gr[c] / n - xr[c] * k;
}
}  // namespace i2t

extern "C" int i2t_l2norm_fwd(const float* x, float* y, int64_t rows, int64_t cols, float eps, void* stream) {
  I2T_REQUIRE(x && y && rows >= 0 && cols > 0, "l2norm_fwd: bad arguments");
  if (rows == 0) return I2T_OK;
  i2t::l2norm_fwd_kernel<<<(unsigned)i2t::ceil_div(rows, 4), 128, 0, (cudaStream_t)stream>>>(x, y, rows, (int)cols, eps);
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_l2norm_bwd(const float* x, const float* dy, float* dx, int64_t rows, int64_t cols, float eps, void* stream) {
  I2T_REQUIRE(x && dy && dx && rows >= 0 && cols > 0, "l2norm_bwd: bad arguments");
  if (rows == 0) return I2T_OK;
  i2t::l2norm_bwd_kernel<<<(unsigned)i2t::ceil_div(rows, 4), 128, 0, (cudaStream_t)stream>>>(x, dy, dx, rows, (int)cols, eps);
  I2T_LAUNCHED();
  return I2T_OK;
}
