// bf16 flash attention forward on the tensor cores (mma.sync m16n8k16, fp32 accumulate): same contract as
// attn_fwd_kernel in attention.cu (packed strided q/k/v read in place, closed-form masks, row log-sum-exp out),
// used for bf16 activations.  64-query x 64-key tiles, 4 warps (16 query rows each); Q fragments live in registers,
// K / V tiles in padded shared memory (144-byte rows: conflict-free ldmatrix); S = Q K^T and O += P V are both
// m16n8k16 MMAs, the S accumulator layout is re-used directly as the A fragment of P V (no shared-memory round trip).
// Replaces F.scaled_dot_product_attention at reference models/layers.py:465 and torchvision's MHA core (:113) under
// bf16 autocast.  (The sequence lengths here -- 197..272 -- make this kernel latency- not throughput-bound; a tcgen05
// version would not change its duration, so the legacy MMA path is the pragmatic choice for these shapes.)
#include "common.cuh"

namespace i2t {

constexpr int ATC_BQ = 64, ATC_BK = 64, ATC_THREADS = 128;

__device__ __forceinline__ bool atc_visible(int mode, int n_prompt, int qi, int kj) {
  if (mode == I2T_MASK_NONE) return true;
  if (kj > qi) return false;
  if (mode == I2T_MASK_CAUSAL) return true;
  return qi < n_prompt ? true : kj >= n_prompt;
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int HS>
__global__ void __launch_bounds__(ATC_THREADS)
attn_fwd_tc_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int H, int Tq, int Tk, int64_t q_bs, int64_t q_rs,
                   int64_t kv_bs, int64_t kv_rs, int mode, int n_prompt, float scale_log2) {
  constexpr int PITCH = HS + 8;               // bf16 elements per shared-memory row (16 bytes of padding)
  constexpr int KS = HS / 16;                 // k-steps over the head dimension
  constexpr int NT_O = HS / 8;                // 8-wide output column tiles
  __shared__ __align__(16) __nv_bfloat16 Qs[ATC_BQ][PITCH];
  __shared__ __align__(16) __nv_bfloat16 Ks[ATC_BK][PITCH];
  __shared__ __align__(16) __nv_bfloat16 Vs[ATC_BK][PITCH];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;    // ldmatrix: matrix index / row inside the 8x8 matrix
  const int q0 = blockIdx.x * ATC_BQ, h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* qb = q + (int64_t)b * q_bs + (int64_t)h * HS;
  const __nv_bfloat16* kb = k + (int64_t)b * kv_bs + (int64_t)h * HS;
  const __nv_bfloat16* vb = v + (int64_t)b * kv_bs + (int64_t)h * HS;

  constexpr int CH = HS / 8;                  // 16-byte chunks per row
  for (int i = t; i < ATC_BQ * CH; i += ATC_THREADS) {
    const int r = i / CH, c = i % CH;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (q0 + r < Tq) val = *reinterpret_cast<const uint4*>(qb + (int64_t)(q0 + r) * q_rs + c * 8);
    *reinterpret_cast<uint4*>(&Qs[r][c * 8]) = val;
  }
  __syncthreads();
  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) ldsm_x4(qf[ks], &Qs[w * 16 + (mi & 1) * 8 + r8][ks * 16 + (mi >> 1) * 8]);

  float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  int kend = Tk;
  if (mode != I2T_MASK_NONE) kend = min(Tk, min(q0 + ATC_BQ, Tq));
  for (int k0 = 0; k0 < kend; k0 += ATC_BK) {
    __syncthreads();
    for (int i = t; i < ATC_BK * CH; i += ATC_THREADS) {
      const int r = i / CH, c = i % CH;
      uint4 kv4 = make_uint4(0u, 0u, 0u, 0u), vv4 = kv4;
      if (k0 + r < Tk) {
        kv4 = *reinterpret_cast<const uint4*>(kb + (int64_t)(k0 + r) * kv_rs + c * 8);
        vv4 = *reinterpret_cast<const uint4*>(vb + (int64_t)(k0 + r) * kv_rs + c * 8);
      }
      *reinterpret_cast<uint4*>(&Ks[r][c * 8]) = kv4;
      *reinterpret_cast<uint4*>(&Vs[r][c * 8]) = vv4;
    }
    __syncthreads();

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bf[4];
        ldsm_x4(bf, &Ks[(2 * np + (mi >> 1)) * 8 + r8][ks * 16 + (mi & 1) * 8]);
        mma_bf16(s[2 * np], qf[ks], bf[0], bf[1]);
        mma_bf16(s[2 * np + 1], qf[ks], bf[2], bf[3]);
      }
    }
    // scale (log2 domain), mask, online softmax; rows g and g+8 of this warp's 16-row slab
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int qi = q0 + w * 16 + g + half * 8;
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kj = k0 + nt * 8 + tq * 2 + e;
          float x = s[nt][half * 2 + e] * scale_log2;
          if (kj >= Tk || !atc_visible(mode, n_prompt, qi, kj)) x = -INFINITY;
          s[nt][half * 2 + e] = x;
          mx = fmaxf(mx, x);
        }
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_i[half], mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f(m_i[half] - m_use);
      float rs = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float p = exp2f(s[nt][half * 2 + e] - m_use);
          s[nt][half * 2 + e] = p;
          rs += p;
        }
      }
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);
      l_i[half] = l_i[half] * corr + rs;
      m_i[half] = m_new;
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt) {
        o[nt][half * 2] *= corr;
        o[nt][half * 2 + 1] *= corr;
      }
    }
    // O += P V : the S accumulators of key tiles (2kk, 2kk+1) are exactly the A fragment of key k-step kk
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pf[4];
      pf[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < NT_O / 2; ++np) {
        uint32_t bf[4];
        ldsm_x4_t(bf, &Vs[kk * 16 + (mi & 1) * 8 + r8][(2 * np + (mi >> 1)) * 8]);
        mma_bf16(o[2 * np], pf, bf[0], bf[1]);
        mma_bf16(o[2 * np + 1], pf, bf[2], bf[3]);
      }
    }
  }

#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int qi = q0 + w * 16 + g + half * 8;
    if (qi >= Tq) continue;
    const float inv = l_i[half] > 0.f ? 1.0f / l_i[half] : 0.f;
    __nv_bfloat16* op = out + ((int64_t)b * Tq + qi) * ((int64_t)H * HS) + (int64_t)h * HS;
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt)
      *reinterpret_cast<uint32_t*>(op + nt * 8 + tq * 2) = pack_bf16(o[nt][half * 2] * inv, o[nt][half * 2 + 1] * inv);
    if (lse != nullptr && tq == 0)
      lse[((int64_t)b * H + h) * Tq + qi] = l_i[half] > 0.f ? (m_i[half] + log2f(l_i[half])) * 0.6931471805599453f : -INFINITY;
  }
}

int attn_fwd_tc(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H, int64_t Tq, int64_t Tk,
                int64_t head_dim, int64_t q_bs, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, int mode, int64_t n_prompt,
                cudaStream_t st) {
  // 16-byte vector loads: strides and head offsets must be multiples of 8 bf16 elements
  if ((q_rs | q_bs | kv_rs | kv_bs) % 8 != 0 || !aligned16(q) || !aligned16(k) || !aligned16(v)) return 0;
  dim3 grid((unsigned)ceil_div(Tq, ATC_BQ), (unsigned)H, (unsigned)B);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
  if (head_dim == 64)
    attn_fwd_tc_kernel<64><<<grid, ATC_THREADS, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                                                       (__nv_bfloat16*)out, lse, (int)H, (int)Tq, (int)Tk, q_bs, q_rs, kv_bs, kv_rs,
                                                       mode, (int)n_prompt, scale_log2);
  else if (head_dim == 32)
    attn_fwd_tc_kernel<32><<<grid, ATC_THREADS, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                                                       (__nv_bfloat16*)out, lse, (int)H, (int)Tq, (int)Tk, q_bs, q_rs, kv_bs, kv_rs,
                                                       mode, (int)n_prompt, scale_log2);
  else
    return 0;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "attn_fwd_tc launch failed: %s", cudaGetErrorString(e));
  return 1;
}

}  // namespace i2t
