// bf16 flash attention forward on the tensor cores (mma.sync m16n8k16, fp32 accumulate): same contract as
// attn_fwd_kernel in attention.cu (packed strided q/k/v read in place, closed-form masks, row log-sum-exp out),
// used for bf16 activations.  64-query x 64-key tiles, 4 warps (16 query rows each); Q fragments live in registers,
// K / V tiles in padded shared memory (144-byte rows: conflict-free ldmatrix); S = Q K^T and O += P V are both
// m16n8k16 MMAs, the S accumulator layout is re-used directly as the A fragment of P V (no shared-memory round trip).
// Replaces F.scaled_dot_product_attention at reference models/layers.py:465 and torchvision's MHA core (:113) under
// bf16 autocast.  (The sequence lengths here -- 197..272 -- make this kernel latency- not throughput-bound; a tcgen05
// version would not change its duration, so the legacy MMA path is the pragmatic choice for these shapes.)
#include "common.cuh"
#include "rng.cuh"

namespace i2t {

// launch with programmatic stream serialization; a failed launch is reported by the cudaGetLastError() check that follows
#define I2T_PDL_LAUNCH(kern, grid, block, smem, st, ...) (void)launch_pdl(kern, grid, block, smem, st, __VA_ARGS__)

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
__device__ __forceinline__ float atc_ex2(float x) {     // one MUFU op (exp2f() adds range fix-ups the softmax does not need)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void atc_cp_async16(void* dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(valid ? 16 : 0)
               : "memory");
}
__device__ __forceinline__ void atc_cp_async4(void* dst, const void* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(valid ? 4 : 0)
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int HS>
__global__ void __launch_bounds__(ATC_THREADS)
attn_fwd_tc_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int H, int Tq, int Tk, int64_t q_bs, int64_t q_rs,
                   int64_t kv_bs, int64_t kv_rs, int mode, int n_prompt, float scale_log2, DropArgs drop) {
  constexpr int PITCH = HS + 8;               // bf16 elements per shared-memory row (16 bytes of padding)
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
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
  DropKey dkey{0u, 0u, 0u};
  if (drop.thr != 0u) dkey = drop_key(drop);

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
      if (drop.thr != 0u) {   // dropout on the probabilities (the row sum above keeps every key); one Philox call per
                              // 16-key block = this thread's four columns {2tq, 2tq+1, 8+2tq, 9+2tq} of the block
        const uint32_t row = (uint32_t)(((int64_t)b * H + h) * Tq + qi);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const Philox4 r = drop_attn8(drop, dkey, row, (uint32_t)(k0 >> 4) + kk, (uint32_t)tq >> 1);
          const int hb = (tq & 1) * 4;                     // this thread's four half-words of the call
          if (philox_half(r, hb) < drop.thr16) s[2 * kk][half * 2] = 0.f;
          if (philox_half(r, hb + 1) < drop.thr16) s[2 * kk][half * 2 + 1] = 0.f;
          if (philox_half(r, hb + 2) < drop.thr16) s[2 * kk + 1][half * 2] = 0.f;
          if (philox_half(r, hb + 3) < drop.thr16) s[2 * kk + 1][half * 2 + 1] = 0.f;
        }
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
    const float inv = l_i[half] > 0.f ? drop.inv_keep / l_i[half] : 0.f;
    __nv_bfloat16* op = out + ((int64_t)b * Tq + qi) * ((int64_t)H * HS) + (int64_t)h * HS;
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt)
      *reinterpret_cast<uint32_t*>(op + nt * 8 + tq * 2) = pack_bf16(o[nt][half * 2] * inv, o[nt][half * 2 + 1] * inv);
    if (lse != nullptr && tq == 0)
      lse[((int64_t)b * H + h) * Tq + qi] = l_i[half] > 0.f ? (m_i[half] + log2f(l_i[half])) * 0.6931471805599453f : -INFINITY;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward on the tensor cores.  CTA = (64-key tile, head, batch), 4 warps, each warp owns 16 keys; loop over
// 64-query tiles.  Everything is computed TRANSPOSED (rows = keys) so that dK / dV accumulate in registers:
//   S^T = K Q^T,  P^T = exp(S^T - lse[q]),  dP^T = V dO^T,  dS^T = P^T * (dP^T - delta[q]) * scale,
//   dV += P^T dO,  dK += dS^T Q            (the S^T / dS^T accumulator fragments are re-used as A fragments)
//   dQ[q,:] += dS K  over this CTA's 64 keys: dS^T goes through shared memory once, each warp takes 16 queries and
//   adds its 16 x HS result to the fp32 dQ accumulator with atomics (other key tiles add to the same rows).
// delta[q] = sum_e dO[q,e] O[q,e] comes from attn_delta_kernel (attention.cu).
// ---------------------------------------------------------------------------------------------------------
template <int HS>
__global__ void __launch_bounds__(ATC_THREADS)
attn_bwd_tc_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                   const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ delta,
                   float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int H, int Tq,
                   int Tk, int64_t q_bs, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, int mode, int n_prompt, float scale,
                   DropArgs drop) {
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  constexpr int PITCH = HS + 8;
  constexpr int KS = HS / 16, NT_O = HS / 8;
  // dynamic shared memory (56 KB at head_dim 64): K | V (dead after the fragment loads: becomes Q buffer 1) | Q buffer 0 |
  // dO buffer 0 | dO buffer 1 | dS^T | lse, delta (2 buffers each).  Q / dO tiles are double buffered and filled with
  // cp.async: the loads of query tile i+1 fly while tile i is in the MMAs (ncu on the single-buffered loop: long-scoreboard
  // stalls 4.2 per issue at 8 warps / SM -- the kernel waited on its own synchronous tile loads)
  extern __shared__ __align__(16) uint8_t atc_smem[];
  typedef __nv_bfloat16 (*Tile)[PITCH];
  constexpr int TILE_BYTES = ATC_BK * PITCH * 2;
  Tile Ks = reinterpret_cast<Tile>(atc_smem);
  Tile Vs = reinterpret_cast<Tile>(atc_smem + TILE_BYTES);
  Tile Qb[2] = {reinterpret_cast<Tile>(atc_smem + 2 * TILE_BYTES), reinterpret_cast<Tile>(atc_smem + TILE_BYTES)};
  Tile dOb[2] = {reinterpret_cast<Tile>(atc_smem + 3 * TILE_BYTES), reinterpret_cast<Tile>(atc_smem + 4 * TILE_BYTES)};
  __nv_bfloat16 (*dSs)[ATC_BQ + 8] = reinterpret_cast<__nv_bfloat16 (*)[ATC_BQ + 8]>(atc_smem + 5 * TILE_BYTES);   // dS^T: [key][query]
  float* lse_b = reinterpret_cast<float*>(atc_smem + 5 * TILE_BYTES + ATC_BK * (ATC_BQ + 8) * 2);                  // [2][64]
  float* delta_b = lse_b + 2 * ATC_BQ;                                                                             // [2][64]
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;
  const int k0 = blockIdx.x * ATC_BK, h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* qb = q + (int64_t)b * q_bs + (int64_t)h * HS;
  const __nv_bfloat16* kb = k + (int64_t)b * kv_bs + (int64_t)h * HS;
  const __nv_bfloat16* vb = v + (int64_t)b * kv_bs + (int64_t)h * HS;
  const int64_t do_rs = (int64_t)H * HS;
  const __nv_bfloat16* dob = dout + (int64_t)b * Tq * do_rs + (int64_t)h * HS;
  const float LOG2E = 1.4426950408889634f;
  const float scale_log2 = scale * LOG2E;
  DropKey dkey{0u, 0u, 0u};
  if (drop.thr != 0u) dkey = drop_key(drop);

  constexpr int CH = HS / 8;
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
  // A fragments of this warp's 16 keys (rows of K and V), kept in registers for the whole kernel
  uint32_t kf[KS][4], vf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    ldsm_x4(kf[ks], &Ks[w * 16 + (mi & 1) * 8 + r8][ks * 16 + (mi >> 1) * 8]);
    ldsm_x4(vf[ks], &Vs[w * 16 + (mi & 1) * 8 + r8][ks * 16 + (mi >> 1) * 8]);
  }
  float adk[NT_O][4], adv[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) adk[i][0] = adk[i][1] = adk[i][2] = adk[i][3] = adv[i][0] = adv[i][1] = adv[i][2] = adv[i][3] = 0.f;

  int qstart = 0;
  if (mode != I2T_MASK_NONE) qstart = (k0 / ATC_BQ) * ATC_BQ;       // queries before the key tile never see it
  // asynchronous fill of one (Q, dO, lse, delta) tile; rows past Tq are zero-filled (src-size 0)
  auto issue_tile = [&](int q0, int buf) {
    for (int i = t; i < ATC_BQ * CH; i += ATC_THREADS) {
      const int r = i / CH, c = i % CH;
      const bool ok = q0 + r < Tq;
      const int64_t row = ok ? q0 + r : 0;
      atc_cp_async16(&Qb[buf][r][c * 8], qb + row * q_rs + c * 8, ok);
      atc_cp_async16(&dOb[buf][r][c * 8], dob + row * do_rs + c * 8, ok);
    }
    {
      const int r = t & (ATC_BQ - 1);
      const bool ok = q0 + r < Tq;
      const int64_t li = ((int64_t)b * H + h) * Tq + (ok ? q0 + r : 0);
      if (t < ATC_BQ) atc_cp_async4(lse_b + buf * ATC_BQ + r, lse + li, ok);
      else atc_cp_async4(delta_b + buf * ATC_BQ + r, delta + li, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (qstart < Tq) issue_tile(qstart, 0);
  int it = 0;
  for (int q0 = qstart; q0 < Tq; q0 += ATC_BQ, ++it) {
    const int buf = it & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // tile `it` is visible; every warp has left iteration it-1 (and, for
                                                       // it = 0, has loaded its V fragments: buffer 1 of Q aliases V)
    if (q0 + ATC_BQ < Tq) issue_tile(q0 + ATC_BQ, buf ^ 1);
    const Tile Qs = Qb[buf], dOs = dOb[buf];
    const float* lse_s = lse_b + buf * ATC_BQ;
    const float* delta_s = delta_b + buf * ATC_BQ;

    bool tile_full = k0 + ATC_BK <= Tk && q0 + ATC_BQ <= Tq;
    if (mode != I2T_MASK_NONE) {
      tile_full = tile_full && (k0 + ATC_BK - 1 <= q0);                                   // every key <= every query
      if (mode == I2T_MASK_PROMPT)
        tile_full = tile_full && ((q0 + ATC_BQ <= n_prompt) || (q0 >= n_prompt && k0 >= n_prompt));
    }
    // S^T = K Q^T and dP^T = V dO^T : rows = this warp's 16 keys, columns = the tile's 64 queries
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bq[4], bd[4];
        ldsm_x4(bq, &Qs[(2 * np + (mi >> 1)) * 8 + r8][ks * 16 + (mi & 1) * 8]);
        ldsm_x4(bd, &dOs[(2 * np + (mi >> 1)) * 8 + r8][ks * 16 + (mi & 1) * 8]);
        mma_bf16(s[2 * np], kf[ks], bq[0], bq[1]);
        mma_bf16(s[2 * np + 1], kf[ks], bq[2], bq[3]);
        mma_bf16(dp[2 * np], vf[ks], bd[0], bd[1]);
        mma_bf16(dp[2 * np + 1], vf[ks], bd[2], bd[3]);
      }
    }
    // P^T and dS^T in place: s <- P^T (with the dropout multiplier: it feeds dV), dp <- dS^T (scaled)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float mk[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop.thr != 0u) {
        // the forward's mask: for query qi the half-words of block (keys k0 + w*16 ..+15), pair (g >> 1) cover keys g and g + 8
        // of this thread at half-words 4 (pair & 1) + (g & 1) and + 2 (rng.cuh: drop_attn8)
#pragma unroll
        for (int e1 = 0; e1 < 2; ++e1) {
          const int qi = q0 + nt * 8 + tq * 2 + e1;
          const uint32_t pr = (uint32_t)(g >> 1) & 3u;
          const Philox4 r = drop_attn8(drop, dkey, (uint32_t)(((int64_t)b * H + h) * Tq + qi), (uint32_t)((k0 + w * 16) >> 4), pr >> 1);
          const int hb = (int)(pr & 1u) * 4;
          mk[e1] = philox_half(r, hb + (g & 1)) >= drop.thr16 ? drop.inv_keep : 0.f;
          mk[2 + e1] = philox_half(r, hb + 2 + (g & 1)) >= drop.thr16 ? drop.inv_keep : 0.f;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kj = k0 + w * 16 + g + (e >> 1) * 8;
        const int ql = nt * 8 + tq * 2 + (e & 1);
        const int qi = q0 + ql;
        const float l = lse_s[ql];
        float p = 0.f;
        // tile_full (uniform over the CTA): every (query, key) pair of this 64 x 64 tile is in range and visible -- the
        // off-diagonal tiles of a causal mask -- so the per-element mask algebra is skipped
        if (tile_full ? (l != -INFINITY) : (kj < Tk && qi < Tq && l != -INFINITY && atc_visible(mode, n_prompt, qi, kj)))
          p = atc_ex2(fmaf(s[nt][e], scale_log2, -l * LOG2E));
        s[nt][e] = p * mk[e];
        dp[nt][e] = p * (dp[nt][e] * mk[e] - delta_s[ql]) * scale;
      }
    }
    // dS^T -> shared memory [key][query] (for dQ), packed pairs along the query index
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<uint32_t*>(&dSs[w * 16 + g][nt * 8 + tq * 2]) = pack_bf16(dp[nt][0], dp[nt][1]);
      *reinterpret_cast<uint32_t*>(&dSs[w * 16 + g + 8][nt * 8 + tq * 2]) = pack_bf16(dp[nt][2], dp[nt][3]);
    }
    // dV += P^T dO ; dK += dS^T Q   (k dimension = the 64 queries: 4 k-steps of 16)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pf[4], df[4];
      pf[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      df[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);
      df[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
      df[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      df[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < NT_O / 2; ++np) {
        uint32_t bo[4], bqq[4];
        ldsm_x4_t(bo, &dOs[kk * 16 + (mi & 1) * 8 + r8][(2 * np + (mi >> 1)) * 8]);
        ldsm_x4_t(bqq, &Qs[kk * 16 + (mi & 1) * 8 + r8][(2 * np + (mi >> 1)) * 8]);
        mma_bf16(adv[2 * np], pf, bo[0], bo[1]);
        mma_bf16(adv[2 * np + 1], pf, bo[2], bo[3]);
        mma_bf16(adk[2 * np], df, bqq[0], bqq[1]);
        mma_bf16(adk[2 * np + 1], df, bqq[2], bqq[3]);
      }
    }
    __syncthreads();                                                  // dS^T of all four warps is in shared memory
    // dQ[q0 + w*16 .. +15][:] += dS[q][keys] K[keys][:]   (A = dS from the transposed tile, B = K)
    {
      float adq[NT_O][4];
#pragma unroll
      for (int i = 0; i < NT_O; ++i) adq[i][0] = adq[i][1] = adq[i][2] = adq[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < ATC_BK / 16; ++kk) {
        uint32_t af[4];
        // A[m = query][k = key] read from dSs[key][query]: matrices (q 0-7,k 0-7), (q 8-15,k 0-7), (q 0-7,k 8-15), (q 8-15,k 8-15)
        ldsm_x4_t(af, &dSs[kk * 16 + (mi >> 1) * 8 + r8][w * 16 + (mi & 1) * 8]);
#pragma unroll
        for (int np = 0; np < NT_O / 2; ++np) {
          uint32_t bk[4];
          ldsm_x4_t(bk, &Ks[kk * 16 + (mi & 1) * 8 + r8][(2 * np + (mi >> 1)) * 8]);
          mma_bf16(adq[2 * np], af, bk[0], bk[1]);
          mma_bf16(adq[2 * np + 1], af, bk[2], bk[3]);
        }
      }
      // 16-byte reductions (red.global.add.v4.f32): lanes (tq, tq^1) swap halves so that the even lane owns four adjacent
      // columns of row g and the odd lane the same four columns of row g + 8 -- a quarter of the atomic operations of the
      // scalar form (the L2 atomic rate, not the MMAs, bounded this kernel)
      {
        const bool odd = (tq & 1) != 0;
        const int qi = q0 + w * 16 + g + (odd ? 8 : 0);
        float* dst = dq_acc + (((int64_t)b * H + h) * Tq + qi) * HS + (tq & ~1) * 2;
#pragma unroll
        for (int nt = 0; nt < NT_O; ++nt) {
          const float sx = odd ? adq[nt][0] : adq[nt][2], sy = odd ? adq[nt][1] : adq[nt][3];
          const float rx = __shfl_xor_sync(0xffffffffu, sx, 1), ry = __shfl_xor_sync(0xffffffffu, sy, 1);
          const float4 v4 = odd ? make_float4(rx, ry, adq[nt][2], adq[nt][3]) : make_float4(adq[nt][0], adq[nt][1], rx, ry);
          if (qi < Tq) atomicAdd(reinterpret_cast<float4*>(dst + nt * 8), v4);
        }
      }
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int kj = k0 + w * 16 + g + half * 8;
    if (kj >= Tk) continue;
    __nv_bfloat16* dkp = dk + (int64_t)b * kv_bs + (int64_t)kj * kv_rs + (int64_t)h * HS + tq * 2;
    __nv_bfloat16* dvp = dv + (int64_t)b * kv_bs + (int64_t)kj * kv_rs + (int64_t)h * HS + tq * 2;
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt) {
      *reinterpret_cast<uint32_t*>(dkp + nt * 8) = pack_bf16(adk[nt][half * 2], adk[nt][half * 2 + 1]);
      *reinterpret_cast<uint32_t*>(dvp + nt * 8) = pack_bf16(adv[nt][half * 2], adv[nt][half * 2 + 1]);
    }
  }
}

// returns 1 when it handled the call, 0 when the shape is not eligible (the caller then runs the fp32-math kernel)
int attn_bwd_tc(const void* q, const void* k, const void* v, const void* dout, const float* lse, const float* delta, float* dq_acc,
                void* dk, void* dv, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_bs, int64_t q_rs,
                int64_t kv_bs, int64_t kv_rs, int mode, int64_t n_prompt, DropArgs drop, cudaStream_t st) {
  if ((q_rs | q_bs | kv_rs | kv_bs) % 8 != 0 || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(dout) ||
      (H * head_dim) % 8 != 0)
    return 0;
  if (((uintptr_t)dk & 3u) != 0 || ((uintptr_t)dv & 3u) != 0) return 0;
  dim3 grid((unsigned)ceil_div(Tk, ATC_BK), (unsigned)H, (unsigned)B);
  const float scale = 1.0f / sqrtf((float)head_dim);
  const size_t smem = (size_t)5 * ATC_BK * (head_dim + 8) * 2 + (size_t)ATC_BK * (ATC_BQ + 8) * 2 + 4 * ATC_BQ * sizeof(float);
  static bool attr64 = false, attr32 = false;
  if (head_dim == 64) {
    if (!attr64) {
      I2T_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr64 = true;
    }
    I2T_PDL_LAUNCH(attn_bwd_tc_kernel<64>, grid, dim3(ATC_THREADS), smem, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                   (const __nv_bfloat16*)v, (const __nv_bfloat16*)dout, lse, delta, dq_acc, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv,
                   (int)H, (int)Tq, (int)Tk, q_bs, q_rs, kv_bs, kv_rs, mode, (int)n_prompt, scale, drop);
  } else if (head_dim == 32) {
    if (!attr32) {
      I2T_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr32 = true;
    }
    I2T_PDL_LAUNCH(attn_bwd_tc_kernel<32>, grid, dim3(ATC_THREADS), smem, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                   (const __nv_bfloat16*)v, (const __nv_bfloat16*)dout, lse, delta, dq_acc, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv,
                   (int)H, (int)Tq, (int)Tk, q_bs, q_rs, kv_bs, kv_rs, mode, (int)n_prompt, scale, drop);
  }
  else
    return 0;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "attn_bwd_tc launch failed: %s", cudaGetErrorString(e));
  return 1;
}

int attn_fwd_tc(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H, int64_t Tq, int64_t Tk,
                int64_t head_dim, int64_t q_bs, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, int mode, int64_t n_prompt,
                DropArgs drop, cudaStream_t st) {
  // 16-byte vector loads: strides and head offsets must be multiples of 8 bf16 elements
  if ((q_rs | q_bs | kv_rs | kv_bs) % 8 != 0 || !aligned16(q) || !aligned16(k) || !aligned16(v)) return 0;
  dim3 grid((unsigned)ceil_div(Tq, ATC_BQ), (unsigned)H, (unsigned)B);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
  if (head_dim == 64)
    I2T_PDL_LAUNCH(attn_fwd_tc_kernel<64>, grid, dim3(ATC_THREADS), 0, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                                                       (__nv_bfloat16*)out, lse, (int)H, (int)Tq, (int)Tk, q_bs, q_rs, kv_bs, kv_rs,
                                                       mode, (int)n_prompt, scale_log2, drop);
  else if (head_dim == 32)
    I2T_PDL_LAUNCH(attn_fwd_tc_kernel<32>, grid, dim3(ATC_THREADS), 0, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                                                       (__nv_bfloat16*)out, lse, (int)H, (int)Tq, (int)Tk, q_bs, q_rs, kv_bs, kv_rs,
                                                       mode, (int)n_prompt, scale_log2, drop);
  else
    return 0;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "attn_fwd_tc launch failed: %s", cudaGetErrorString(e));
  return 1;
}

}  // namespace i2t
