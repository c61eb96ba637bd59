// KV-cached autoregressive decode step.  The reference has no KV cache: VisionEncoderDecoder.generate
// (reference models/vision_encoder_decoder.py:136-182) re-runs the whole decoder over the prefix for every token.
// These kernels compute the SAME last-position logits incrementally (DESIGN.md "decode algebra"):
//   * text rows never attend to the soft-prompt rows (reference :93-99 zeroes query rows), so the prompt
//     rows are dropped entirely and text position i just uses wpe[n_prompt + i];
//   * self K/V of earlier tokens and the cross K/V projections of the encoder output are cached.
// Everything here is a batch-of-8 weight-streaming problem: HBM-bound on the weight bytes, so the design goal is
// bytes in flight, not FLOPs:
//   * every CTA of the skinny linear pulls its slab of weight rows with ONE bulk async copy per 8 rows
//     (cp.async.bulk, the 1-D TMA path: no registers, completion on an mbarrier) issued before anything else;
//   * programmatic dependent launch: the next kernel of the step starts while the previous one drains, issues its
//     weight copies (weights never depend on the previous kernel) and only then waits on griddepcontrol.wait --
//     the HBM stream continues across kernel boundaries;
//   * all kernels read the current position from DEVICE memory so one captured CUDA graph replays every step.
#include <algorithm>

#include "common.cuh"

namespace i2t {

template <typename... KArgs, typename... Args>
static cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  return launch_pdl(kern, grid, block, smem, st, args...);
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_addr(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D bulk async copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

// x[b,:] = wte[ids[b, pos]] + wpe[n_prompt + pos]          (models/decoder.py:234-243)
__global__ void __launch_bounds__(256) dec_embed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ wte,
                                                        const float* __restrict__ wpe, float* __restrict__ x,
                                                        const int32_t* __restrict__ pos_ptr, int B, int C, int64_t ids_ld,
                                                        int n_prompt) {
  pdl_launch_dependents();
  pdl_wait();
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

__global__ void dec_advance_kernel(int32_t* pos_ptr) {
  pdl_launch_dependents();
  pdl_wait();
  *pos_ptr += 1;
}

// ---------------------------------------------------------------------------------------------------------
// Skinny linear: out[b, n] = epi( sum_k LN?(x)[b,k] * W[n,k] + bias[n] ),  b < B <= 8, W (N,K) row-major.
// CTA = 256 threads, owns `rows_per_cta` (8, 16 or 32) consecutive weight rows = one contiguous slab of W.
//   1. thread 0 arms one mbarrier per 8-row pass and issues the bulk copies of the whole slab (HBM -> smem);
//   2. griddepcontrol.wait (the activations come from the previous kernel), then the 8 warps stage x into shared
//      memory with the fused LayerNorm prologue (one batch row per warp, two-pass statistics, fp32);
//   3. per pass: K is split over all 256 threads; each thread keeps its 16-byte chunk of x for all 8 batch rows in
//      registers and walks the 8 weight rows in shared memory (1 LDS.128 per 32..64 FMAs);
//   4. the 8x8 partial sums are reduced with a transposing shuffle tree (62 shuffles instead of 320), across warps
//      through shared memory, and 64 threads apply bias / activation / residual / KV-cache append.
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

constexpr int DL_THREADS = 256, DL_R = 8, DL_B = 8, DL_MAX_PASSES = 4;

__device__ __forceinline__ void unpack16(const float4& raw, float (&o)[4]) {
  o[0] = raw.x; o[1] = raw.y; o[2] = raw.z; o[3] = raw.w;
}
__device__ __forceinline__ void unpack16(const float4& raw, float (&o)[8]) {
  const uint32_t u[4] = {__float_as_uint(raw.x), __float_as_uint(raw.y), __float_as_uint(raw.z), __float_as_uint(raw.w)};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __uint_as_float(u[i] << 16);
    o[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}

template <typename TW>
__global__ void __launch_bounds__(DL_THREADS)
dec_linear_kernel(const float* __restrict__ x, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float ln_eps,
                  const TW* __restrict__ W, const float* __restrict__ bias, int B, int N, int K, int act, int rows_per_cta,
                  DecLinearEpi epi) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[DL_MAX_PASSES];
  constexpr int VEC = 16 / (int)sizeof(TW);        // weight elements per 16-byte chunk
  TW* wslab = reinterpret_cast<TW*>(smem_raw);                                   // [rows_per_cta][K]
  float* xs = reinterpret_cast<float*>(smem_raw + (size_t)rows_per_cta * K * sizeof(TW));   // [DL_B][K]
  float* part = xs + DL_B * K;                                                   // [8 warps][64]
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int n0 = blockIdx.x * rows_per_cta;
  const int nrows = min(rows_per_cta, N - n0);
  const int passes = (nrows + DL_R - 1) / DL_R;

  if (t == 0) {
    for (int p = 0; p < passes; ++p) mbar_init(&bars[p], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int p = 0; p < passes; ++p) {
      const int rp = min(DL_R, nrows - p * DL_R);
      const uint32_t bytes = (uint32_t)rp * (uint32_t)K * (uint32_t)sizeof(TW);
      mbar_expect_tx(&bars[p], bytes);
      bulk_g2s(wslab + (size_t)p * DL_R * K, W + ((size_t)n0 + (size_t)p * DL_R) * K, bytes, &bars[p]);
    }
  }
  pdl_launch_dependents();
  pdl_wait();                       // activations (x, residual, position) are produced by the previous kernels

  // ---- stage activations with the LayerNorm prologue: warp w handles batch row w ----
  {
    float* xr = xs + w * K;
    if (w < B) {
      const float* src = x + (int64_t)w * K;
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
          *reinterpret_cast<float4*>(xr + k) = v;
        }
      }
      if (sizeof(TW) == 2) {  // autocast semantics: the Linear sees bf16 activations
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

  const int nchunks = K / VEC;
  const int pos = epi.mode == 1 ? *epi.pos_ptr : 0;
  for (int p = 0; p < passes; ++p) {
    float acc[DL_R * DL_B];
#pragma unroll
    for (int i = 0; i < DL_R * DL_B; ++i) acc[i] = 0.f;
    mbar_wait(&bars[p], 0u);
    const TW* wp = wslab + (size_t)p * DL_R * K;
    const int rp = min(DL_R, nrows - p * DL_R);
    for (int c = t; c < nchunks; c += DL_THREADS) {
      float xv[DL_B][VEC];
#pragma unroll
      for (int b = 0; b < DL_B; ++b)
#pragma unroll
        for (int j = 0; j < VEC; j += 4) {
          const float4 a = *reinterpret_cast<const float4*>(xs + b * K + c * VEC + j);
          xv[b][j] = a.x; xv[b][j + 1] = a.y; xv[b][j + 2] = a.z; xv[b][j + 3] = a.w;
        }
#pragma unroll
      for (int r = 0; r < DL_R; ++r) {
        if (r < rp) {
          float wv[VEC];
          unpack16(*reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(wp + (size_t)r * K) + (size_t)c * 16), wv);
#pragma unroll
          for (int b = 0; b < DL_B; ++b)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[r * DL_B + b] = fmaf(wv[j], xv[b][j], acc[r * DL_B + b]);
        }
      }
    }
    // transposing shuffle reduction: afterwards lane l holds the warp totals of flat indices 2l and 2l+1
#pragma unroll
    for (int off = 16, n = DL_R * DL_B; off >= 1; off >>= 1, n >>= 1) {
      const int half = n >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float send = upper ? acc[i] : acc[i + half];
        const float keep = upper ? acc[i + half] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    part[w * 64 + 2 * lane] = acc[0];
    part[w * 64 + 2 * lane + 1] = acc[1];
    __syncthreads();
    if (t < DL_R * DL_B) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < DL_THREADS / 32; ++i) v += part[i * 64 + t];
      const int r = t / DL_B, b = t % DL_B;
      const int n = n0 + p * DL_R + r;
      if (r < rp && b < B) {
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
    __syncthreads();   // `part` is reused by the next pass
  }
}

// ---------------------------------------------------------------------------------------------------------
// Single-query attention over a (B, Tmax, C) cache: grid (H, B), 4 warps; each warp takes keys in groups of 4
// (4 independent row loads in flight) with its 32 lanes across the head dimension, keeps an online softmax, and
// the 4 partial states are merged in shared memory.  len = *len_ptr + len_add (self: pos + 1) or the constant S
// (cross attention over the cached K/V projections of the encoder output).
// ---------------------------------------------------------------------------------------------------------
template <typename TC, int HS>
__global__ void __launch_bounds__(128)
dec_attn_kernel(const float* __restrict__ q, int64_t q_ld, const TC* __restrict__ kc, const TC* __restrict__ vc,
                int64_t cache_bs, int64_t cache_rs, float* __restrict__ out, int64_t out_ld,
                const int32_t* __restrict__ len_ptr, int len_add, int round_q_bf16,
                const float* __restrict__ knew, const float* __restrict__ vnew, int64_t new_ld, TC* __restrict__ kc_w,
                TC* __restrict__ vc_w, __nv_bfloat16* __restrict__ out16) {
  // knew / vnew (optional): the K / V row of the token being decoded, still in the projection's fp32 output.  It is key
  // len - 1: rounded to the cache dtype, used from registers and stored into the cache by this CTA (the separate append
  // launch disappears).  out16 (optional): the result is written as bf16, ready to be the next GEMM's A operand.
  constexpr int EPL = HS / 32;
  constexpr int G = 4;
  __shared__ float s_m[4], s_l[4], s_acc[4][HS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int h = blockIdx.x;
  const int64_t b = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
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
  const int len_c = knew != nullptr ? len - 1 : len;      // keys read from the cache (the new key comes from knew / vnew)
  for (int j0 = w * G; j0 < len_c; j0 += 4 * G) {
    float kk[G][EPL], vv[G][EPL], d[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int j = min(j0 + g, len_c - 1);
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        kk[g][e] = to_f32(kb[(int64_t)j * cache_rs + lane + 32 * e]);
        vv[g][e] = to_f32(vb[(int64_t)j * cache_rs + lane + 32 * e]);
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) s = fmaf(qv[e], kk[g][e], s);
      d[g] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int g = 0; g < G; ++g) d[g] += __shfl_xor_sync(0xffffffffu, d[g], o);
    float m_new = m;
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (j0 + g < len_c) m_new = fmaxf(m_new, d[g]);
    const float corr = expf(m - m_new);
    l *= corr;
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] *= corr;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (j0 + g < len_c) {
        const float p = expf(d[g] - m_new);
        l += p;
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[e] = fmaf(p, vv[g][e], acc[e]);
      }
    }
    m = m_new;
  }
  if (knew != nullptr && w == 0 && len > 0) {
    // key len - 1 = the token being decoded: rounded to the cache dtype, folded into warp 0's running softmax, appended
    float k1[EPL], v1[EPL], d = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const TC kq = from_f32<TC>(knew[b * new_ld + (int64_t)h * HS + lane + 32 * e]);
      const TC vq = from_f32<TC>(vnew[b * new_ld + (int64_t)h * HS + lane + 32 * e]);
      kc_w[b * cache_bs + (int64_t)(len - 1) * cache_rs + (int64_t)h * HS + lane + 32 * e] = kq;
      vc_w[b * cache_bs + (int64_t)(len - 1) * cache_rs + (int64_t)h * HS + lane + 32 * e] = vq;
      k1[e] = to_f32(kq);
      v1[e] = to_f32(vq);
      d = fmaf(qv[e], k1[e], d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    const float m_new = fmaxf(m, d);
    const float corr = expf(m - m_new), p = expf(d - m_new);
    l = l * corr + p;
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = fmaf(p, v1[e], acc[e] * corr);
    m = m_new;
  }
  if (lane == 0) { s_m[w] = m; s_l[w] = l; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) s_acc[w][lane + 32 * e] = acc[e];
  __syncthreads();
  if (w == 0) {
    const float M = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
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
    for (int e = 0; e < EPL; ++e) {
      if (out16 != nullptr) out16[b * out_ld + (int64_t)h * HS + lane + 32 * e] = __float2bfloat16_rn(o[e] * inv);
      else out[b * out_ld + (int64_t)h * HS + lane + 32 * e] = o[e] * inv;
    }
  }
}

}  // namespace i2t

using namespace i2t;

#define I2T_LAUNCH_CHECK(expr)                                                                              \
  do {                                                                                                      \
    cudaError_t e__ = (expr);                                                                               \
    ::i2t::g_launches.fetch_add(1, std::memory_order_relaxed);                                              \
    if (e__ != cudaSuccess) return ::i2t::fail(I2T_ERR_CUDA, "launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)


extern "C" int i2t_dec_embed(const int64_t* ids, const float* wte, const float* wpe, float* x, const int32_t* pos_ptr,
                             int64_t B, int64_t C, int64_t ids_ld, int64_t n_prompt, void* stream) {
  I2T_REQUIRE(ids && wte && wpe && x && pos_ptr && B > 0 && C % 4 == 0, "dec_embed: bad arguments");
  const int64_t n = B * C / 4;
  I2T_LAUNCH_CHECK(launch(dec_embed_kernel, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, ids, wte, wpe,
                          x, pos_ptr, (int)B, (int)C, ids_ld, (int)n_prompt));
  return I2T_OK;
}

namespace i2t {
// K / V rows of the token at *pos_ptr: columns [C,2C) and [2C,3C) of a packed (B, ld) fp32 qkv buffer -> row *pos_ptr of the
// (B, Tmax, C) caches (large-batch decode: the projections run as GEMMs, this is the cache append of dec_linear's epilogue)
template <typename TC>
__global__ void __launch_bounds__(256) dec_kv_append_kernel(const float* __restrict__ qkv, int64_t ld, TC* __restrict__ kcache,
                                                            TC* __restrict__ vcache, int64_t cache_bs, int C, int B,
                                                            const int32_t* __restrict__ pos_ptr) {
  const int pos = *pos_ptr;
  const int c4 = C / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)B * c4; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / c4), cv = (int)(i % c4);
    const float4 kx = load4(qkv + (int64_t)b * ld + C + cv * 4);
    const float4 vx = load4(qkv + (int64_t)b * ld + 2 * C + cv * 4);
    store4(kcache + (int64_t)b * cache_bs + (int64_t)pos * C + cv * 4, kx);
    store4(vcache + (int64_t)b * cache_bs + (int64_t)pos * C + cv * 4, vx);
  }
}
}  // namespace i2t

namespace i2t {
// x[b,:] = rows[b, *pos_ptr, :] + wpe[*pos_ptr, :]  -- a soft-prompt row as the decoder input of step *pos_ptr (HF decoders
// see the prompt rows as ordinary positions, reference models/decoder.py:343-360)
__global__ void __launch_bounds__(256) dec_embed_rows_kernel(const float* __restrict__ rows, int64_t batch_stride,
                                                             const float* __restrict__ wpe, float* __restrict__ x,
                                                             const int32_t* __restrict__ pos_ptr, int B, int C) {
  const int pos = *pos_ptr;
  const int c4 = C / 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * c4; i += gridDim.x * blockDim.x) {
    const int b = i / c4, cv = i % c4;
    float4 v = load4(rows + (int64_t)b * batch_stride + (int64_t)pos * C + cv * 4);
    const float4 pe = load4(wpe + (int64_t)pos * C + cv * 4);
    v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
    store4(x + (int64_t)b * C + cv * 4, v);
  }
}
}  // namespace i2t

extern "C" int i2t_dec_embed_rows(const float* rows, int64_t batch_stride, const float* wpe, float* x, const int32_t* pos_ptr,
                                  int64_t B, int64_t C, void* stream) {
  I2T_REQUIRE(rows && wpe && x && pos_ptr && B > 0 && C > 0 && C % 4 == 0, "dec_embed_rows: bad arguments");
  const int64_t n = B * (C / 4);
  dec_embed_rows_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 4096), 256, 0, (cudaStream_t)stream>>>(
      rows, batch_stride, wpe, x, pos_ptr, (int)B, (int)C);
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_dec_kv_append(const float* qkv, int64_t ld, void* kcache, void* vcache, int64_t cache_batch_stride, int64_t C,
                                 int64_t B, int cache_dtype, const int32_t* pos_ptr, void* stream) {
  I2T_REQUIRE(qkv && kcache && vcache && pos_ptr, "dec_kv_append: null pointer");
  I2T_REQUIRE(B > 0 && C > 0 && C % 4 == 0 && ld >= 3 * C && ld % 4 == 0 && valid_dtype(cache_dtype), "dec_kv_append: bad sizes");
  const int64_t n = B * (C / 4);
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), 4096);
  cudaStream_t st = (cudaStream_t)stream;
  if (cache_dtype == I2T_F32)
    dec_kv_append_kernel<float><<<grid, 256, 0, st>>>(qkv, ld, (float*)kcache, (float*)vcache, cache_batch_stride, (int)C, (int)B, pos_ptr);
  else
    dec_kv_append_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(qkv, ld, (__nv_bfloat16*)kcache, (__nv_bfloat16*)vcache,
                                                              cache_batch_stride, (int)C, (int)B, pos_ptr);
  I2T_LAUNCHED();
  return I2T_OK;
}

namespace i2t {
// LayerNorm of the (B, C) residual stream for the batched decode step, PDL-aware: gamma / beta (weights) are fetched before
// the dependency wait; optionally zero-fills `zero_ptr` (the fp32 output of the split-K projection that consumes this
// LayerNorm: its partial tiles are ADDED, and a memset node would break the programmatic launch chain).
template <typename TY, int MAXV>
__global__ void __launch_bounds__(128) dec_ln_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, TY* __restrict__ y, int rows, int cols,
                                                     float eps, float* __restrict__ zero_ptr, int64_t zero_n4) {
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;
  const int nvec = cols >> 2;
  float4 v[MAXV], gm[MAXV], bt[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      gm[i] = load4(gamma + c * 4);
      bt[i] = beta ? load4(beta + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  pdl_wait();                 // x was written by the previous kernel; zero_ptr may still be read by an earlier one
  if (zero_ptr != nullptr) {
    float4* z4 = reinterpret_cast<float4*>(zero_ptr);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero_n4; i += (int64_t)gridDim.x * blockDim.x)
      z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * cols;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      v[i] = load4(xr + c * 4);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
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
  TY* yr = y + (int64_t)row * cols;
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

// h = act(z) with a cast, PDL-aware (the GELU between the split-K FC and the MLP down projection of the batched decode step)
template <typename TO>
__global__ void __launch_bounds__(256) dec_act_kernel(const float* __restrict__ z, TO* __restrict__ h, int64_t n4, int act) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = load4(z + i * 4);
    v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act);
    store4(h + i * 4, v);
  }
}
}  // namespace i2t

extern "C" int i2t_dec_layernorm(const float* x, const float* gamma, const float* beta, void* y, int64_t rows, int64_t cols,
                                 float eps, int y_dtype, float* zero_ptr, int64_t zero_count, void* stream) {
  I2T_REQUIRE(x && gamma && y && rows > 0, "dec_layernorm: bad arguments");
  I2T_REQUIRE(cols > 0 && cols % 4 == 0 && cols <= 2048, "dec_layernorm: cols=%lld must be a multiple of 4, <= 2048", (long long)cols);
  I2T_REQUIRE(valid_dtype(y_dtype) && aligned16(x) && aligned16(gamma) && (beta == nullptr || aligned16(beta)) && aligned16(y),
              "dec_layernorm: alignment / dtype");
  I2T_REQUIRE(zero_ptr == nullptr || (zero_count % 4 == 0 && aligned16(zero_ptr)), "dec_layernorm: zero-fill range must be 16-byte granular");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)ceil_div(rows, 4)), block(128);
  const int64_t z4 = zero_ptr ? zero_count / 4 : 0;
#define I2T_DLN(TY, MV) I2T_LAUNCH_CHECK(launch(dec_ln_kernel<TY, MV>, grid, block, 0, st, x, gamma, beta, (TY*)y, (int)rows, (int)cols, eps, zero_ptr, z4))
  if (y_dtype == I2T_F32) {
    if (cols <= 1024) I2T_DLN(float, 8); else I2T_DLN(float, 16);
  } else {
    if (cols <= 1024) I2T_DLN(__nv_bfloat16, 8); else I2T_DLN(__nv_bfloat16, 16);
  }
#undef I2T_DLN
  return I2T_OK;
}

extern "C" int i2t_dec_act(const float* z, void* h, int64_t n, int act, int h_dtype, void* stream) {
  I2T_REQUIRE(z && h && n > 0 && n % 4 == 0 && valid_dtype(h_dtype), "dec_act: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t blocks = std::min<int64_t>(ceil_div(n / 4, 256), (int64_t)num_sms() * 8);
  if (h_dtype == I2T_F32)
    I2T_LAUNCH_CHECK(launch(dec_act_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, st, z, (float*)h, n / 4, act));
  else
    I2T_LAUNCH_CHECK(launch(dec_act_kernel<__nv_bfloat16>, dim3((unsigned)blocks), dim3(256), 0, st, z, (__nv_bfloat16*)h, n / 4, act));
  return I2T_OK;
}

extern "C" int i2t_dec_advance(int32_t* pos_ptr, void* stream) {
  I2T_REQUIRE(pos_ptr, "dec_advance: null pointer");
  I2T_LAUNCH_CHECK(launch(dec_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, pos_ptr));
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
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t esz = w_dtype == I2T_F32 ? 4 : 2;
  // slab height: as tall as 48 KB allows, but keep >= 2 CTAs per SM when the matrix is small
  int rows = DL_R * DL_MAX_PASSES;
  while (rows > DL_R && (rows * K * esz > 48 * 1024 || ceil_div(N, rows) < 2 * (int64_t)num_sms())) rows >>= 1;
  const size_t smem = (size_t)rows * K * esz + (size_t)DL_B * K * sizeof(float) + (size_t)(DL_THREADS / 32) * 64 * sizeof(float);
  // 227 KB per CTA is the opt-in ceiling for static + dynamic shared memory together
  I2T_REQUIRE(smem <= 226 * 1024, "dec_linear: K=%lld too large for shared memory", (long long)K);
  static std::atomic<size_t> attr_f32{0}, attr_bf16{0};
  if (w_dtype == I2T_F32 && smem > attr_f32.load()) {
    I2T_CUDA(cudaFuncSetAttribute(dec_linear_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_f32.store(smem);
  }
  if (w_dtype == I2T_BF16 && smem > attr_bf16.load()) {
    I2T_CUDA(cudaFuncSetAttribute(dec_linear_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_bf16.store(smem);
  }
  for (int64_t b0 = 0; b0 < B; b0 += DL_B) {     // batches above 8 re-stream the weights (second pass hits L2)
    const int bn = (int)(B - b0 < DL_B ? B - b0 : DL_B);
    DecLinearEpi epi;
    epi.mode = qkv_split ? 1 : 0;
    epi.out = out + b0 * ldo;
    epi.ldo = ldo;
    epi.residual = residual ? residual + b0 * ldo : nullptr;
    const int64_t cesz = cache_dtype == I2T_F32 ? 4 : 2;
    epi.kcache = kcache ? (void*)((uint8_t*)kcache + b0 * cache_batch_stride * cesz) : nullptr;
    epi.vcache = vcache ? (void*)((uint8_t*)vcache + b0 * cache_batch_stride * cesz) : nullptr;
    epi.cache_bs = cache_batch_stride;
    epi.C = (int)C;
    epi.cache_dtype = cache_dtype;
    epi.pos_ptr = pos_ptr;
    const dim3 grid((unsigned)ceil_div(N, rows));
    if (w_dtype == I2T_F32)
      I2T_LAUNCH_CHECK(launch(dec_linear_kernel<float>, grid, dim3(DL_THREADS), smem, st, x + b0 * K, ln_gamma, ln_beta, ln_eps,
                              (const float*)W, bias, bn, (int)N, (int)K, act, rows, epi));
    else
      I2T_LAUNCH_CHECK(launch(dec_linear_kernel<__nv_bfloat16>, grid, dim3(DL_THREADS), smem, st, x + b0 * K, ln_gamma, ln_beta,
                              ln_eps, (const __nv_bfloat16*)W, bias, bn, (int)N, (int)K, act, rows, epi));
  }
  return I2T_OK;
}

static int dec_attn_impl(const float* q, int64_t q_ld, const void* kcache, const void* vcache, int64_t cache_batch_stride,
                         int64_t cache_row_stride, void* out, int64_t out_ld, const int32_t* len_ptr, int64_t len_add, int64_t B,
                         int64_t H, int64_t head_dim, int cache_dtype, const float* knew, const float* vnew, int64_t new_ld,
                         int out_dtype, void* stream) {
  I2T_REQUIRE(q && kcache && vcache && out, "dec_attn: null pointer");
  I2T_REQUIRE(B > 0 && B <= 65535 && H > 0, "dec_attn: bad sizes");
  I2T_REQUIRE(head_dim == 64 || head_dim == 32, "dec_attn: head_dim %lld not built (32, 64)", (long long)head_dim);
  I2T_REQUIRE(valid_dtype(cache_dtype) && valid_dtype(out_dtype), "dec_attn: bad dtype");
  I2T_REQUIRE((knew == nullptr) == (vnew == nullptr), "dec_attn: knew and vnew go together");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)H, (unsigned)B);
  const int rq = cache_dtype == I2T_BF16 ? 1 : 0;
  float* o32 = out_dtype == I2T_F32 ? (float*)out : nullptr;
  __nv_bfloat16* o16 = out_dtype == I2T_BF16 ? (__nv_bfloat16*)out : nullptr;
#define I2T_DA(TC, HSV)                                                                                              \
  I2T_LAUNCH_CHECK(launch(dec_attn_kernel<TC, HSV>, grid, dim3(128), 0, st, q, q_ld, (const TC*)kcache, (const TC*)vcache, \
                          cache_batch_stride, cache_row_stride, o32, out_ld, len_ptr, (int)len_add, rq, knew, vnew, new_ld,  \
                          (TC*)const_cast<void*>(kcache), (TC*)const_cast<void*>(vcache), o16))
  if (cache_dtype == I2T_F32) {
    if (head_dim == 64) I2T_DA(float, 64); else I2T_DA(float, 32);
  } else {
    if (head_dim == 64) I2T_DA(__nv_bfloat16, 64); else I2T_DA(__nv_bfloat16, 32);
  }
#undef I2T_DA
  return I2T_OK;
}

extern "C" int i2t_dec_attn(const float* q, int64_t q_ld, const void* kcache, const void* vcache,
                            int64_t cache_batch_stride, int64_t cache_row_stride, float* out, int64_t out_ld,
                            const int32_t* len_ptr, int64_t len_add, int64_t B, int64_t H, int64_t head_dim,
                            int cache_dtype, void* stream) {
  return dec_attn_impl(q, q_ld, kcache, vcache, cache_batch_stride, cache_row_stride, out, out_ld, len_ptr, len_add, B, H,
                       head_dim, cache_dtype, nullptr, nullptr, 0, I2T_F32, stream);
}

extern "C" int i2t_dec_attn_append(const float* q, int64_t q_ld, void* kcache, void* vcache, int64_t cache_batch_stride,
                                   int64_t cache_row_stride, void* out, int64_t out_ld, const int32_t* len_ptr, int64_t len_add,
                                   int64_t B, int64_t H, int64_t head_dim, int cache_dtype, const float* knew,
                                   const float* vnew, int64_t new_ld, int out_dtype, void* stream) {
  return dec_attn_impl(q, q_ld, kcache, vcache, cache_batch_stride, cache_row_stride, out, out_ld, len_ptr, len_add, B, H,
                       head_dim, cache_dtype, knew, vnew, new_ld, out_dtype, stream);
}
