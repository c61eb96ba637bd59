// Fused multi-tensor optimiser kernels: one launch per parameter group instead of one python loop trip (SNRAdam,
// reference models/optimizer.py:66-111) or one foreach dispatch chain (torch.optim.AdamW, reference trainer.py:169-172)
// per tensor.  HBM-bound: AdamW / SNRAdam read p,g,m,v and write p,m,v (28 B per fp32 parameter); the EMA teacher
// update (reference training/wrapper.py:53-60) reads p_m,p and writes p_m (12 B per parameter, in place instead of the
// reference's allocate-and-rebind).
//
// Layout: `table` is a device array of n_tensors x 4 pointers [p, g, m, v] (EMA: [p_m, p, bf16 shadow of p_m or 0, -];
// cast: [src fp32, dst bf16, -, -]); work is cut into
// fixed-size chunks, chunk c covers elements [chunk_off[c], chunk_off[c] + chunk_len[c]) of tensor chunk_tensor[c].
#include "common.cuh"

namespace i2t {

constexpr int OPT_THREADS = 256;

// Scalars are derived on the host in double precision exactly the way the python optimisers derive them, then
// rounded once to fp32 (the dtype torch uses for a python scalar meeting an fp32 tensor).
struct AdamArgs {
  float lr, beta1, beta2, one_minus_b1, one_minus_b2, eps, weight_decay;
  float decay;           // 1 - lr * wd
  float step_size;       // lr / (1 - beta1^t)                     (AdamW)
  float bc2_sqrt;        // sqrt(1 - beta2^t)                      (AdamW)
  float inv_bc1, inv_bc2;  // 1 / (1 - beta^t)                     (SNRAdam)
  float prev_scale;      // 1 / (1 - beta1^(t-1)), 1 when t == 1   (SNRAdam)
  float grad_scale;      // multiplies g on the fly (1/world_size, 1/accum) -- 1.0 for the reference semantics
  int step;
};

__device__ __forceinline__ void chunk_range(const int32_t* chunk_tensor, const int64_t* chunk_off, const int32_t* chunk_len,
                                            int& tensor, int64_t& off, int& len) {
  tensor = chunk_tensor[blockIdx.x];
  off = chunk_off[blockIdx.x];
  len = chunk_len[blockIdx.x];
}

// torch.optim.AdamW (single-tensor formula):  p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(OPT_THREADS)
adamw_multi_kernel(const int64_t* __restrict__ table, const int32_t* __restrict__ chunk_tensor,
                   const int64_t* __restrict__ chunk_off, const int32_t* __restrict__ chunk_len, AdamArgs a) {
  int ti, len;
  int64_t off;
  chunk_range(chunk_tensor, chunk_off, chunk_len, ti, off, len);
  float* p = reinterpret_cast<float*>(table[ti * 4 + 0]) + off;
  const float* g = reinterpret_cast<const float*>(table[ti * 4 + 1]) + off;
  float* m = reinterpret_cast<float*>(table[ti * 4 + 2]) + off;
  float* v = reinterpret_cast<float*>(table[ti * 4 + 3]) + off;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= a.grad_scale;
    pp *= a.decay;
    mm = mm + (gg - mm) * a.one_minus_b1;            // lerp form used by torch (exp_avg.lerp_)
    vv = vv * a.beta2 + a.one_minus_b2 * gg * gg;
    const float denom = sqrtf(vv) / a.bc2_sqrt + a.eps;
    pp -= a.step_size * (mm / denom);
  };
  if (vec) {
    const int n4 = len >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 P = load4(p + i * 4), G = load4(g + i * 4), M = load4(m + i * 4), V = load4(v + i * 4);
      upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
      store4(p + i * 4, P); store4(m + i * 4, M); store4(v + i * 4, V);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < len; i += OPT_THREADS) upd(p[i], g[i], m[i], v[i]);
  } else {
    for (int i = threadIdx.x; i < len; i += OPT_THREADS) upd(p[i], g[i], m[i], v[i]);
  }
}

// SNRAdam (reference models/optimizer.py:85-111):
//   if wd: p *= 1 - lr*wd;  d = g - (t == 1 ? m : m / (1 - b1^(t-1)));  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*d*d;
//   p -= lr * (m / bc1) / (sqrt(v / bc2) + eps)
__global__ void __launch_bounds__(OPT_THREADS)
snradam_multi_kernel(const int64_t* __restrict__ table, const int32_t* __restrict__ chunk_tensor,
                     const int64_t* __restrict__ chunk_off, const int32_t* __restrict__ chunk_len, AdamArgs a) {
  int ti, len;
  int64_t off;
  chunk_range(chunk_tensor, chunk_off, chunk_len, ti, off, len);
  float* p = reinterpret_cast<float*>(table[ti * 4 + 0]) + off;
  const float* g = reinterpret_cast<const float*>(table[ti * 4 + 1]) + off;
  float* m = reinterpret_cast<float*>(table[ti * 4 + 2]) + off;
  float* v = reinterpret_cast<float*>(table[ti * 4 + 3]) + off;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= a.grad_scale;
    if (a.weight_decay != 0.f) pp *= a.decay;
    float d = gg - mm * a.prev_scale;
    d *= d;
    mm = mm * a.beta1 + a.one_minus_b1 * gg;
    vv = vv * a.beta2 + a.one_minus_b2 * d;
    pp -= a.lr * ((mm * a.inv_bc1) / (sqrtf(vv * a.inv_bc2) + a.eps));
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  if (vec) {
    const int n4 = len >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 P = load4(p + i * 4), G = load4(g + i * 4), M = load4(m + i * 4), V = load4(v + i * 4);
      upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
      store4(p + i * 4, P); store4(m + i * 4, M); store4(v + i * 4, V);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < len; i += OPT_THREADS) upd(p[i], g[i], m[i], v[i]);
  } else {
    for (int i = threadIdx.x; i < len; i += OPT_THREADS) upd(p[i], g[i], m[i], v[i]);
  }
}

// p_m = p_m * momentum + p * (1 - momentum)
__global__ void __launch_bounds__(OPT_THREADS)
ema_multi_kernel(const int64_t* __restrict__ table, const int32_t* __restrict__ chunk_tensor,
                 const int64_t* __restrict__ chunk_off, const int32_t* __restrict__ chunk_len, float momentum, float om) {
  int ti, len;
  int64_t off;
  chunk_range(chunk_tensor, chunk_off, chunk_len, ti, off, len);
  float* pm = reinterpret_cast<float*>(table[ti * 4 + 0]) + off;
  const float* p = reinterpret_cast<const float*>(table[ti * 4 + 1]) + off;
  // optional compute-dtype shadow of the teacher weight, refreshed in the same pass (bf16 forward passes read it)
  __nv_bfloat16* sh = table[ti * 4 + 2] != 0 ? reinterpret_cast<__nv_bfloat16*>(table[ti * 4 + 2]) + off : nullptr;
  const bool vec = ((reinterpret_cast<uintptr_t>(pm) | reinterpret_cast<uintptr_t>(p)) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(sh) & 7u) == 0;
  if (vec) {
    const int n4 = len >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 A = load4(pm + i * 4);
      const float4 Bv = load4(p + i * 4);
      A.x = A.x * momentum + Bv.x * om; A.y = A.y * momentum + Bv.y * om;
      A.z = A.z * momentum + Bv.z * om; A.w = A.w * momentum + Bv.w * om;
      store4(pm + i * 4, A);
      if (sh != nullptr) store4(sh + i * 4, A);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < len; i += OPT_THREADS) {
      pm[i] = pm[i] * momentum + p[i] * om;
      if (sh != nullptr) sh[i] = __float2bfloat16_rn(pm[i]);
    }
  } else {
    for (int i = threadIdx.x; i < len; i += OPT_THREADS) {
      pm[i] = pm[i] * momentum + p[i] * om;
      if (sh != nullptr) sh[i] = __float2bfloat16_rn(pm[i]);
    }
  }
}

// dst (bf16) = src (fp32) over a tensor list: refreshes the compute-dtype shadows of the parameters an optimiser step wrote
__global__ void __launch_bounds__(OPT_THREADS)
cast_bf16_multi_kernel(const int64_t* __restrict__ table, const int32_t* __restrict__ chunk_tensor,
                       const int64_t* __restrict__ chunk_off, const int32_t* __restrict__ chunk_len) {
  int ti, len;
  int64_t off;
  chunk_range(chunk_tensor, chunk_off, chunk_len, ti, off, len);
  const float* src = reinterpret_cast<const float*>(table[ti * 4 + 0]) + off;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(table[ti * 4 + 1]) + off;
  const bool vec = (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0;
  if (vec) {
    const int n4 = len >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) store4(dst + i * 4, load4(src + i * 4));
    for (int i = (n4 << 2) + threadIdx.x; i < len; i += OPT_THREADS) dst[i] = __float2bfloat16_rn(src[i]);
  } else {
    for (int i = threadIdx.x; i < len; i += OPT_THREADS) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

static AdamArgs make_args(double lr, double beta1, double beta2, double eps, double wd, int64_t step, double grad_scale) {
  AdamArgs a;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  a.lr = (float)lr; a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.weight_decay = (float)wd;
  a.one_minus_b1 = (float)(1.0 - beta1);
  a.one_minus_b2 = (float)(1.0 - beta2);
  a.decay = (float)(1.0 - lr * wd);
  a.step_size = (float)(lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.inv_bc1 = (float)(1.0 / bc1);
  a.inv_bc2 = (float)(1.0 / bc2);
  a.prev_scale = step > 1 ? (float)(1.0 / (1.0 - pow(beta1, (double)(step - 1)))) : 1.0f;
  a.grad_scale = (float)grad_scale;
  a.step = (int)step;
  return a;
}

}  // namespace i2t

using namespace i2t;

extern "C" int i2t_adamw_multi(const int64_t* table, const int32_t* chunk_tensor, const int64_t* chunk_off,
                               const int32_t* chunk_len, int64_t n_chunks, double lr, double beta1, double beta2, double eps,
                               double weight_decay, int64_t step, double grad_scale, void* stream) {
  I2T_REQUIRE(table && chunk_tensor && chunk_off && chunk_len && n_chunks >= 0 && step >= 1, "adamw_multi: bad arguments");
  if (n_chunks == 0) return I2T_OK;
  adamw_multi_kernel<<<(unsigned)n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(
      table, chunk_tensor, chunk_off, chunk_len, make_args(lr, beta1, beta2, eps, weight_decay, step, grad_scale));
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_snradam_multi(const int64_t* table, const int32_t* chunk_tensor, const int64_t* chunk_off,
                                 const int32_t* chunk_len, int64_t n_chunks, double lr, double beta1, double beta2, double eps,
                                 double weight_decay, int64_t step, double grad_scale, void* stream) {
  I2T_REQUIRE(table && chunk_tensor && chunk_off && chunk_len && n_chunks >= 0 && step >= 1, "snradam_multi: bad arguments");
  if (n_chunks == 0) return I2T_OK;
  snradam_multi_kernel<<<(unsigned)n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(
      table, chunk_tensor, chunk_off, chunk_len, make_args(lr, beta1, beta2, eps, weight_decay, step, grad_scale));
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_ema_multi(const int64_t* table, const int32_t* chunk_tensor, const int64_t* chunk_off,
                             const int32_t* chunk_len, int64_t n_chunks, double momentum, void* stream) {
  I2T_REQUIRE(table && chunk_tensor && chunk_off && chunk_len && n_chunks >= 0, "ema_multi: bad arguments");
  if (n_chunks == 0) return I2T_OK;
  ema_multi_kernel<<<(unsigned)n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(table, chunk_tensor, chunk_off, chunk_len,
                                                                               (float)momentum, (float)(1.0 - momentum));
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_cast_bf16_multi(const int64_t* table, const int32_t* chunk_tensor, const int64_t* chunk_off,
                                   const int32_t* chunk_len, int64_t n_chunks, void* stream) {
  I2T_REQUIRE(table && chunk_tensor && chunk_off && chunk_len && n_chunks >= 0, "cast_bf16_multi: bad arguments");
  if (n_chunks == 0) return I2T_OK;
  cast_bf16_multi_kernel<<<(unsigned)n_chunks, OPT_THREADS, 0, (cudaStream_t)stream>>>(table, chunk_tensor, chunk_off, chunk_len);
  I2T_LAUNCHED();
  return I2T_OK;
}
