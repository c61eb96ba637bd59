// KV-cached autoregressive decode step.  The reference has no KV cache: VisionEncoderDecoder.generate
// (reference models/vision_encoder_decoder.py:136-182) re-runs the whole decoder over the prefix for every token.
// These kernels compute the SAME last-position logits incrementally (DESIGN.md "decode algebra"):
//   * text rows never attend to the soft-prompt rows (reference :93-99 zeroes query rows), so the prompt
//     rows are dropped entirely and text position i just uses wpe[n_prompt + i];
//   * self K/V of earlier tokens and the cross K/V projections of the encoder output are cached.
// Everything here is a batch-of-B (B <= 16) weight-streaming problem: HBM-bound on the weight bytes.
// All kernels read the current position from DEVICE memory so that one captured CUDA graph replays every step.
#include "common.cuh"

namespace i2t {

// x[b,:] = wte[ids[b, pos]] + wpe[n_prompt + pos]          (models/decoder.py:234-243)
__global__ void __launch_bounds__(256) dec_embed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ wte,
                                                        const float* __restrict__ wpe, float* __restrict__ x,
                                                        const int32_t* __restrict__ pos_ptr, int B, int C, int64_t ids_ld,
                                                        int n_prompt) {
  const int pos = *pos_ptr;
  const int c4 = C / 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * c4; i += gridDim.x * blockDim.x) {
    const int b = i / c4, cv = i % c4;
    const int64_t tok = ids[(int64_t)b * ids_ld + pos];
    float4 v = load4(wte + tok * C + cv * 4);
    const float4 pe = load4(wpe + (int64_t)(n_prompt + pos) * C + cv * 4);
    v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
    store4(x + (int64_t)b * C + cv * 4, v);
  }
}

__global__ void dec_advance_kernel(int32_t* pos_ptr) { *pos_ptr += 1; }

// one 16-byte global load of weights -> fp32 registers (4 x fp32 or 8 x bf16), streaming (read once)
__device__ __forceinline__ void load_w16(const float* p, float (&o)[4]) {
  const float4 a = __ldcs(reinterpret_cast<const float4*>(p));
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
}
__device__ __forceinline__ void load_w16(const __nv_bfloat16* p, float (&o)[8]) {
  const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(p));
  const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __uint_as_float(u[i] << 16);
    o[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Skinny linear: out[b, n] = epi( sum_k LN?(x)[b,k] * W[n,k] + bias[n] ),  b < B <= MAXB.
// CTA = 4 warps, each warp owns R = 4 consecutive output rows and the whole K; the (optionally layer-normed)
// activations live in shared memory as fp32; weights stream from HBM with 128-bit loads, one pass, no reuse.
// Epilogues: bias, activation, residual add (in place on the fp32 residual stream) and the fused KV-cache append.
// ---------------------------------------------------------------------------------------------------------
struct DecLinearEpi {
  int mode;               // 0: out[b*ldo + n];  1: qkv split (q -> out, k/v -> caches at *pos_ptr)
  float* out;             // fp32 (B, ldo)
  int64_t ldo;
  const float* residual;  // optional fp32 (B, ldo) added after bias/act (may alias out)
  void* kcache;           // (B, Tmax, C) cache dtype
  void* vcache;
  int64_t cache_bs;       // Tmax * C
  int C;
  int cache_dtype;
  const int32_t* pos_ptr;
};

template <typename TW, int MAXB>
__global__ void __launch_bounds__(128)
dec_linear_kernel(const float* __restrict__ x, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float ln_eps,
                  const TW* __restrict__ W, const float* __restrict__ bias, int B, int N, int K, int act, DecLinearEpi epi) {
  extern __shared__ __align__(16) float xs[];  // [MAXB][K]
  constexpr int R = 4;
  constexpr int VEC = sizeof(TW) == 4 ? 4 : 8;  // elements per 16-byte load
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  // stage activations (+ LayerNorm prologue: one warp per batch row, two-pass statistics)
  for (int b = w; b < MAXB; b += 4) {
    float* xr = xs + b * K;
    if (b < B) {
      const float* src = x + (int64_t)b * K;
      float s = 0.f;
      for (int k = lane * 4; k < K; k += 128) {
        const float4 v = load4(src + k);
        *reinterpret_cast<float4*>(xr + k) = v;
        s += (v.x + v.y) + (v.z + v.w);
      }
      if (ln_g != nullptr) {
        const float mu = warp_sum(s) / (float)K;
        float q = 0.f;
        for (int k = lane * 4; k < K; k += 128) {
          const float4 v = *reinterpret_cast<const float4*>(xr + k);
          const float a = v.x - mu, bb = v.y - mu, c = v.z - mu, d = v.w - mu;
          q += (a * a + bb * bb) + (c * c + d * d);
        }
        const float rs = 1.0f / sqrtf(warp_sum(q) / (float)K + ln_eps);
        for (int k = lane * 4; k < K; k += 128) {
          float4 v = *reinterpret_cast<const float4*>(xr + k);
          const float4 g = load4(ln_g + k);
          v.x = (v.x - mu) * rs * g.x; v.y = (v.y - mu) * rs * g.y;
          v.z = (v.z - mu) * rs * g.z; v.w = (v.w - mu) * rs * g.w;
          if (ln_b != nullptr) {
            const float4 be = load4(ln_b + k);
            v.x += be.x; v.y += be.y; v.z += be.z; v.w += be.w;
          }
          if (sizeof(TW) == 2) {  // autocast semantics: the Linear sees bf16 activations
            v.x = __bfloat162float(__float2bfloat16_rn(v.x)); v.y = __bfloat162float(__float2bfloat16_rn(v.y));
            v.z = __bfloat162float(__float2bfloat16_rn(v.z)); v.w = __bfloat162float(__float2bfloat16_rn(v.w));
          }
          *reinterpret_cast<float4*>(xr + k) = v;
        }
      } else if (sizeof(TW) == 2) {
        for (int k = lane * 4; k < K; k += 128) {
          float4 v = *reinterpret_cast<const float4*>(xr + k);
          v.x = __bfloat162float(__float2bfloat16_rn(v.x)); v.y = __bfloat162float(__float2bfloat16_rn(v.y));
          v.z = __bfloat162float(__float2bfloat16_rn(v.z)); v.w = __bfloat162float(__float2bfloat16_rn(v.w));
          *reinterpret_cast<float4*>(xr + k) = v;
        }
      }
    } else {
      for (int k = lane * 4; k < K; k += 128) *reinterpret_cast<float4*>(xr + k) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();

  const int n0 = (blockIdx.x * 4 + w) * R;
  if (n0 >= N) return;
  float acc[R][MAXB];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int b = 0; b < MAXB; ++b) acc[r][b] = 0.f;
  const TW* wrow[R];
#pragma unroll
  for (int r = 0; r < R; ++r) wrow[r] = W + (int64_t)min(n0 + r, N - 1) * K;

  for (int k = lane * VEC; k < K; k += 32 * VEC) {
    float wv[R][VEC];
#pragma unroll
    for (int r = 0; r < R; ++r) load_w16(wrow[r] + k, wv[r]);
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      float xv[VEC];
#pragma unroll
      for (int j = 0; j < VEC; j += 4) {
        const float4 a = *reinterpret_cast<const float4*>(xs + b * K + k + j);
        xv[j] = a.x; xv[j + 1] = a.y; xv[j + 2] = a.z; xv[j + 3] = a.w;
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[r][b] = fmaf(wv[r][j], xv[j], acc[r][b]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int b = 0; b < MAXB; ++b) acc[r][b] = warp_sum(acc[r][b]);

  // lane (r, b) = (lane / MAXB', lane % ...) writes one output: spread the R*MAXB results over the lanes
  const int pos = epi.mode == 1 ? *epi.pos_ptr : 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      if (lane == ((r * MAXB + b) & 31)) {
        const int n = n0 + r;
        if (n < N && b < B) {
          float v = acc[r][b];
          if (bias != nullptr) v += bias[n];
          v = apply_act(v, act);
          if (epi.mode == 0) {
            if (epi.residual != nullptr) v += epi.residual[(int64_t)b * epi.ldo + n];
            epi.out[(int64_t)b * epi.ldo + n] = v;
          } else {
            const int seg = n / epi.C, nl = n % epi.C;
            if (seg == 0) {
              epi.out[(int64_t)b * epi.ldo + nl] = v;
            } else {
              void* base = seg == 1 ? epi.kcache : epi.vcache;
              const int64_t off = (int64_t)b * epi.cache_bs + (int64_t)pos * epi.C + nl;
              if (epi.cache_dtype == I2T_F32) ((float*)base)[off] = v;
              else ((__nv_bfloat16*)base)[off] = __float2bfloat16_rn(v);
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Single-query attention over a (B, Tmax, C) cache: grid (H, B), 4 warps; each warp takes keys w, w+4, ...
// with its 32 lanes across the head dimension, keeps an online softmax, and the 4 partial states are merged
// in shared memory.  len = *len_ptr + len_add (self: pos + 1) or the constant S (cross attention).
// ---------------------------------------------------------------------------------------------------------
template <typename TC, int HS>
__global__ void __launch_bounds__(128)
dec_attn_kernel(const float* __restrict__ q, int64_t q_ld, const TC* __restrict__ kc, const TC* __restrict__ vc,
                int64_t cache_bs, int64_t cache_rs, float* __restrict__ out, int64_t out_ld,
                const int32_t* __restrict__ len_ptr, int len_add, int round_q_bf16) {
  constexpr int EPL = HS / 32;
  __shared__ float s_m[4], s_l[4], s_acc[4][HS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int h = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int len = (len_ptr != nullptr ? *len_ptr : 0) + len_add;
  const float scale = 1.0f / sqrtf((float)HS);
  float qv[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    float v = q[b * q_ld + (int64_t)h * HS + lane + 32 * e];
    if (round_q_bf16) v = __bfloat162float(__float2bfloat16_rn(v));
    qv[e] = v * scale;
  }
  const TC* kb = kc + b * cache_bs + (int64_t)h * HS;
  const TC* vb = vc + b * cache_bs + (int64_t)h * HS;
  float m = -INFINITY, l = 0.f, acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  for (int j = w; j < len; j += 4) {
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) d = fmaf(qv[e], to_f32(kb[(int64_t)j * cache_rs + lane + 32 * e]), d);
    d = warp_sum(d);
    const float m_new = fmaxf(m, d);
    const float corr = expf(m - m_new);
    const float p = expf(d - m_new);
    l = l * corr + p;
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = fmaf(p, to_f32(vb[(int64_t)j * cache_rs + lane + 32 * e]), acc[e] * corr);
    m = m_new;
  }
  if (lane == 0) { s_m[w] = m; s_l[w] = l; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) s_acc[w][lane + 32 * e] = acc[e];
  __syncthreads();
  if (w == 0) {
    float M = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
    float L = 0.f, o[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) o[e] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float c = (s_m[i] == -INFINITY) ? 0.f : expf(s_m[i] - M);
      L += s_l[i] * c;
#pragma unroll
      for (int e = 0; e < EPL; ++e) o[e] = fmaf(s_acc[i][lane + 32 * e], c, o[e]);
    }
    const float inv = L > 0.f ? 1.0f / L : 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) out[b * out_ld + (int64_t)h * HS + lane + 32 * e] = o[e] * inv;
  }
}

}  // namespace i2t

using namespace i2t;

extern "C" int i2t_dec_embed(const int64_t* ids, const float* wte, const float* wpe, float* x, const int32_t* pos_ptr,
                             int64_t B, int64_t C, int64_t ids_ld, int64_t n_prompt, void* stream) {
  I2T_REQUIRE(ids && wte && wpe && x && pos_ptr && B > 0 && C % 4 == 0, "dec_embed: bad arguments");
  const int64_t n = B * C / 4;
  dec_embed_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(ids, wte, wpe, x, pos_ptr, (int)B, (int)C,
                                                                              ids_ld, (int)n_prompt);
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_dec_advance(int32_t* pos_ptr, void* stream) {
  I2T_REQUIRE(pos_ptr, "dec_advance: null pointer");
  dec_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(pos_ptr);
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_dec_linear(const float* x, const float* ln_gamma, const float* ln_beta, float ln_eps, const void* W,
                              const float* bias, const float* residual, float* out, int64_t ldo, int64_t B, int64_t N,
                              int64_t K, int act, int w_dtype, int qkv_split, void* kcache, void* vcache,
                              int64_t cache_batch_stride, int64_t C, int cache_dtype, const int32_t* pos_ptr,
                              void* stream) {
  I2T_REQUIRE(x && W && out, "dec_linear: null pointer");
  I2T_REQUIRE(B > 0 && B <= 16, "dec_linear: batch %lld outside 1..16 (larger batches go through i2t_gemm)", (long long)B);
  I2T_REQUIRE(N > 0 && K > 0 && K % 8 == 0, "dec_linear: K=%lld must be a multiple of 8", (long long)K);
  I2T_REQUIRE(valid_dtype(w_dtype) && aligned16(W) && aligned16(x), "dec_linear: dtype/alignment");
  I2T_REQUIRE(!qkv_split || (kcache && vcache && pos_ptr && C > 0 && N == 3 * C && valid_dtype(cache_dtype)),
              "dec_linear: qkv split needs caches, pos_ptr and N == 3C");
  DecLinearEpi epi;
  epi.mode = qkv_split ? 1 : 0;
  epi.out = out;
  epi.ldo = ldo;
  epi.residual = residual;
  epi.kcache = kcache;
  epi.vcache = vcache;
  epi.cache_bs = cache_batch_stride;
  epi.C = (int)C;
  epi.cache_dtype = cache_dtype;
  epi.pos_ptr = pos_ptr;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div(N, 16);
  const int maxb = B <= 8 ? 8 : 16;
  const size_t smem = (size_t)maxb * K * sizeof(float);
  I2T_REQUIRE(smem <= 200 * 1024, "dec_linear: B*K too large for shared memory");
#define I2T_DL(TW, MB)                                                                                              \
  do {                                                                                                              \
    auto kern = dec_linear_kernel<TW, MB>;                                                                          \
    if (smem > 48 * 1024) I2T_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, 128, smem, st>>>(x, ln_gamma, ln_beta, ln_eps, (const TW*)W, bias, (int)B, (int)N, (int)K, act, epi); \
  } while (0)
  if (w_dtype == I2T_F32) {
    if (maxb == 8) I2T_DL(float, 8); else I2T_DL(float, 16);
  } else {
    if (maxb == 8) I2T_DL(__nv_bfloat16, 8); else I2T_DL(__nv_bfloat16, 16);
  }
#undef I2T_DL
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_dec_attn(const float* q, int64_t q_ld, const void* kcache, const void* vcache,
                            int64_t cache_batch_stride, int64_t cache_row_stride, float* out, int64_t out_ld,
                            const int32_t* len_ptr, int64_t len_add, int64_t B, int64_t H, int64_t head_dim,
                            int cache_dtype, void* stream) {
  I2T_REQUIRE(q && kcache && vcache && out, "dec_attn: null pointer");
  I2T_REQUIRE(B > 0 && B <= 65535 && H > 0, "dec_attn: bad sizes");
  I2T_REQUIRE(head_dim == 64 || head_dim == 32, "dec_attn: head_dim %lld not built (32, 64)", (long long)head_dim);
  I2T_REQUIRE(valid_dtype(cache_dtype), "dec_attn: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)H, (unsigned)B);
  const int rq = cache_dtype == I2T_BF16 ? 1 : 0;
#define I2T_DA(TC, HSV)                                                                                              \
  dec_attn_kernel<TC, HSV><<<grid, 128, 0, st>>>(q, q_ld, (const TC*)kcache, (const TC*)vcache, cache_batch_stride,  \
                                                 cache_row_stride, out, out_ld, len_ptr, (int)len_add, rq)
  if (cache_dtype == I2T_F32) {
    if (head_dim == 64) I2T_DA(float, 64); else I2T_DA(float, 32);
  } else {
    if (head_dim == 64) I2T_DA(__nv_bfloat16, 64); else I2T_DA(__nv_bfloat16, 32);
  }
#undef I2T_DA
  I2T_LAUNCHED();
  return I2T_OK;
}
