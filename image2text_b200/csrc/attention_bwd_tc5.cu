// bf16 attention BACKWARD on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), head_dim 64, any sequence length.
// Same contract as attn_bwd_tc (attention_tc.cu): packed strided q/k/v read in place, closed-form masks, the forward's
// log-sum-exp, dropout masks regenerated from the counter RNG, dQ accumulated in an fp32 workspace.  Replaces the backward
// of F.scaled_dot_product_attention at reference models/layers.py:465 (and torchvision's MHA core) under bf16 autocast.
//
// One CTA = one 128-key block j of one (batch, head); it walks the 128-query blocks i that can see those keys:
//   warp 0   : TMA -- K_j, V_j once; Q_i, dO_i per query block into a 2-deep ring (128-byte swizzle, one mbarrier per stage);
//   warp 1   : one thread issues, per query block,
//                S  = Q_i K_j^T , dP = dO_i V_j^T        (M128 N128 K16 x 4 each, all operands K-major)  -> TMEM columns [0,256)
//                dV += P^T dO_i , dK += dS^T Q_i         (M128 N64 K16 x 8: A = P / dS read MN-major (keys contiguous) from the
//                                                         [query][key] tiles the softmax threads wrote, B = dO_i / Q_i MN-major)
//                dQ  = dS K_j                            (M128 N64 K16 x 8: A = dS K-major, B = K_j MN-major)
//              dV, dK stay in TMEM columns [256,384) over the whole loop, dQ_i in [384,448) is drained per query block;
//   warps 4-11: two warpgroups, thread = (query row, 64-key half): reads its S / dP half rows from TMEM, forms
//              P = exp2(S * scale - lse) (masked, dropout multiplier) and dS = P (dP * mask - delta) * scale, writes both as bf16
//              into the swizzled shared-memory tiles; delta = sum_e dO[q,e] O[q,e] is computed here (no separate pass); drains
//              dQ_i with 16-byte fp32 reductions into the workspace, and dK_j / dV_j at the end.
// TMEM: 448 of 512 columns -> one CTA per SM; shared memory 160 KB.
#include "common.cuh"
#include "rng.cuh"
#include "tc_common.cuh"

namespace i2t {

constexpr int B5_BQ = 128, B5_BK = 128, B5_HS = 64, B5_THREADS = 384;
constexpr int B5_TILE = 128 * B5_HS * 2;                 // 16 KB: K, V, one Q / dO stage, one 64-key panel of P / dS
constexpr int B5_SLACK = 1024, B5_BAR_BYTES = 128;
constexpr int B5_SMEM = 10 * B5_TILE + B5_BAR_BYTES + B5_SLACK;
constexpr uint32_t B5_COL_S = 0, B5_COL_DP = 128, B5_COL_DV = 256, B5_COL_DK = 320, B5_COL_DQ = 384;

__device__ __forceinline__ float b5_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t b5_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void b5_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// log-sum-exp (base 2) of query row qi and delta = sum_e dO[qi,e] O[qi,e]
__device__ __forceinline__ void b5_row_scalars(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                               const float* __restrict__ lse, int b, int h, int H, int Tq, int qi, float& l2, float& delta) {
  l2 = -INFINITY;
  delta = 0.f;
  if (qi >= Tq) return;
  const int64_t row_stride = (int64_t)H * B5_HS;
  l2 = lse[((int64_t)b * H + h) * Tq + qi] * 1.4426950408889634f;
  const uint4* orow = reinterpret_cast<const uint4*>(out + ((int64_t)b * Tq + qi) * row_stride + (int64_t)h * B5_HS);
  const uint4* drow = reinterpret_cast<const uint4*>(dout + ((int64_t)b * Tq + qi) * row_stride + (int64_t)h * B5_HS);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 a = orow[c], d = drow[c];
    const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* dp2 = reinterpret_cast<const __nv_bfloat162*>(&d);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = __bfloat1622float2(ap[e]), y = __bfloat1622float2(dp2[e]);
      delta = fmaf(x.x, y.x, fmaf(x.y, y.y, delta));
    }
  }
}

// visible keys of query row qi: one interval [lo, hi) (empty for rows beyond Tq or without any visible key)
__device__ __forceinline__ void b5_visible(int mode, int n_prompt, int Tq, int Tk, int qi, float l2, int& lo, int& hi) {
  lo = 0;
  hi = Tk;
  if (mode != I2T_MASK_NONE) {
    hi = min(Tk, qi + 1);
    if (mode == I2T_MASK_PROMPT && qi >= n_prompt) lo = n_prompt;
  }
  if (qi >= Tq || l2 == -INFINITY) hi = lo;
}

// One thread's half row (64 keys starting at key kbase, TMEM columns col_s / col_dp + [0,64)) of a (query block, key block) pair:
// P = exp2(S * scale - lse) (masked, dropout multiplier) and dS = P (dP * mask - delta) * scale as bf16 into row r of the
// 64-key panels prow / drow (128-byte rows, 16-byte chunks XOR-swizzled by the row like the TMA tiles).
__device__ __forceinline__ void b5_softmax_half(uint32_t taddr_s, uint32_t taddr_dp, int kbase, int lo, int hi, float l2, float delta,
                                                float scale, const DropArgs& drop, const DropKey& dkey, uint32_t rng_row, uint8_t* prow,
                                                uint8_t* drow, int r) {
  const float scale_log2 = scale * 1.4426950408889634f;
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    const int c0 = kbase + c * 32;                                  // first key of this 32-key chunk
    uint32_t sv[32], dpv[32];
    tmem_ld32(taddr_s + (uint32_t)(c * 32), sv);
    tmem_ld32(taddr_dp + (uint32_t)(c * 32), dpv);
    float* p = reinterpret_cast<float*>(sv);                        // P and dS are formed in place (register budget: 168 / thread)
    float* ds = reinterpret_cast<float*>(dpv);
    if (c0 >= lo && c0 + 32 <= hi) {
#pragma unroll
      for (int i = 0; i < 32; ++i) p[i] = b5_ex2(fmaf(p[i], scale_log2, -l2));
    } else if (c0 + 32 > lo && c0 < hi) {
#pragma unroll
      for (int i = 0; i < 32; ++i) p[i] = (c0 + i >= lo && c0 + i < hi) ? b5_ex2(fmaf(p[i], scale_log2, -l2)) : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) p[i] = 0.f;
    }
    if (drop.thr != 0u && c0 + 32 > lo && c0 < hi) {     // the forward's masks: 4 Philox calls per 32 keys (rng.cuh: drop_attn8);
                                                         // a chunk without a visible key has P = 0 already
#pragma unroll
      for (int bl = 0; bl < 2; ++bl) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {                     // one call = 8 keys: pairs 2 hh and 2 hh + 1 of the 16-key block
          const Philox4 rr = drop_attn8(drop, dkey, rng_row, (uint32_t)(c0 >> 4) + bl, (uint32_t)hh);
#pragma unroll
          for (int sp = 0; sp < 2; ++sp) {
            const int e0 = bl * 16 + 2 * (2 * hh + sp);
            const uint32_t w0 = sp == 0 ? rr.x : rr.z, w1 = sp == 0 ? rr.y : rr.w;
            const float m0 = (w0 & 0xFFFFu) < drop.thr16 ? 0.f : drop.inv_keep, m1 = (w0 >> 16) < drop.thr16 ? 0.f : drop.inv_keep;
            const float m2 = (w1 & 0xFFFFu) < drop.thr16 ? 0.f : drop.inv_keep, m3 = (w1 >> 16) < drop.thr16 ? 0.f : drop.inv_keep;
            ds[e0] = p[e0] * (ds[e0] * m0 - delta) * scale;
            ds[e0 + 1] = p[e0 + 1] * (ds[e0 + 1] * m1 - delta) * scale;
            ds[e0 + 8] = p[e0 + 8] * (ds[e0 + 8] * m2 - delta) * scale;
            ds[e0 + 9] = p[e0 + 9] * (ds[e0 + 9] * m3 - delta) * scale;
            p[e0] *= m0;
            p[e0 + 1] *= m1;
            p[e0 + 8] *= m2;
            p[e0 + 9] *= m3;
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) ds[i] = p[i] * (ds[i] - delta) * scale;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      uint4 pk, dk4;
      pk.x = b5_pack(p[8 * t], p[8 * t + 1]);
      pk.y = b5_pack(p[8 * t + 2], p[8 * t + 3]);
      pk.z = b5_pack(p[8 * t + 4], p[8 * t + 5]);
      pk.w = b5_pack(p[8 * t + 6], p[8 * t + 7]);
      dk4.x = b5_pack(ds[8 * t], ds[8 * t + 1]);
      dk4.y = b5_pack(ds[8 * t + 2], ds[8 * t + 3]);
      dk4.z = b5_pack(ds[8 * t + 4], ds[8 * t + 5]);
      dk4.w = b5_pack(ds[8 * t + 6], ds[8 * t + 7]);
      const int c16 = c * 4 + t;
      *reinterpret_cast<uint4*>(prow + ((c16 ^ (r & 7)) << 4)) = pk;
      *reinterpret_cast<uint4*>(drow + ((c16 ^ (r & 7)) << 4)) = dk4;
    }
  }
}

// Coalesced drains.  A thread owns one accumulator ROW, so storing it straight to global memory touches 32 different cache lines
// per warp instruction (the LSU serialises them: measured 1.6 us per 128 x 64 tile pair, and the store storm also held up the MMA
// thread's next issue).  Instead the bf16 tile is staged in shared memory (128-byte rows, chunks XOR-swizzled by the row) and the
// 256 softmax threads write it out 8 lanes per row: 4 rows = 4 lines per warp instruction.
__device__ __forceinline__ void b5_stage32(const uint32_t (&v)[32], uint8_t* tile, int r, int wg, float sc = 1.f) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    uint4 pk;
    pk.x = b5_pack(__uint_as_float(v[8 * t]) * sc, __uint_as_float(v[8 * t + 1]) * sc);
    pk.y = b5_pack(__uint_as_float(v[8 * t + 2]) * sc, __uint_as_float(v[8 * t + 3]) * sc);
    pk.z = b5_pack(__uint_as_float(v[8 * t + 4]) * sc, __uint_as_float(v[8 * t + 5]) * sc);
    pk.w = b5_pack(__uint_as_float(v[8 * t + 6]) * sc, __uint_as_float(v[8 * t + 7]) * sc);
    *reinterpret_cast<uint4*>(tile + r * 128 + (((wg * 4 + t) ^ (r & 7)) << 4)) = pk;
  }
}
// tile row i -> gbase + i * row_stride (elements), rows [0, rows_valid); tid = 0..255
__device__ __forceinline__ void b5_store_tile(const uint8_t* tile, __nv_bfloat16* gbase, int64_t row_stride, int rows_valid, int tid) {
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int row = pass * 32 + (tid >> 3), c = tid & 7;
    if (row < rows_valid)
      *reinterpret_cast<uint4*>(gbase + (int64_t)row * row_stride + c * 8) =
          *reinterpret_cast<const uint4*>(tile + row * 128 + ((c ^ (row & 7)) << 4));
  }
}
__device__ __forceinline__ void b5_sync_softmax_warps() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 32 fp32 accumulator columns of one TMEM row -> 32 bf16 (64 bytes) at dst
__device__ __forceinline__ void b5_store32(const uint32_t (&v)[32], __nv_bfloat16* dst) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    uint4 pk;
    pk.x = b5_pack(__uint_as_float(v[8 * t]), __uint_as_float(v[8 * t + 1]));
    pk.y = b5_pack(__uint_as_float(v[8 * t + 2]), __uint_as_float(v[8 * t + 3]));
    pk.z = b5_pack(__uint_as_float(v[8 * t + 4]), __uint_as_float(v[8 * t + 5]));
    pk.w = b5_pack(__uint_as_float(v[8 * t + 6]), __uint_as_float(v[8 * t + 7]));
    *reinterpret_cast<uint4*>(dst + 8 * t) = pk;
  }
}

__global__ void __launch_bounds__(B5_THREADS, 1)
attn_bwd_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                    const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                    float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int H, int Tq, int Tk,
                    int64_t kv_bs, int64_t kv_rs, int mode, int n_prompt, float scale, DropArgs drop) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sK = smem;
  uint8_t* sV = sK + B5_TILE;
  uint8_t* sQ = sV + B5_TILE;                  // 2 stages
  uint8_t* sdO = sQ + 2 * B5_TILE;             // 2 stages
  uint8_t* sP = sdO + 2 * B5_TILE;             // [2 panels of 64 keys][128 queries][128 B]
  uint8_t* sdS = sP + 2 * B5_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * B5_TILE);
  uint64_t& bar_kv = bars[0];
  uint64_t* bar_q = bars + 1;                  // [2] Q_i / dO_i landed
  uint64_t* bar_qfree = bars + 3;              // [2] the MMAs that read the stage have finished
  uint64_t& bar_s = bars[5];                   // S / dP in TMEM
  uint64_t& bar_p = bars[6];                   // P / dS in shared memory (256 arrivals); S / dP have been read
  uint64_t& bar_mma2 = bars[7];                // dV / dK / dQ MMAs of the query block finished
  uint64_t& bar_dqfree = bars[8];              // dQ columns drained (256 arrivals)
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = j * B5_BK;
  const int i0 = mode != I2T_MASK_NONE ? j : 0;                 // under a causal-type mask earlier queries see none of these keys
  const int n_it = max(0, (Tq + B5_BQ - 1) / B5_BQ - i0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmdO) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bar_kv, 1);
    mbar_init(&bar_q[0], 1);
    mbar_init(&bar_q[1], 1);
    mbar_init(&bar_qfree[0], 1);
    mbar_init(&bar_qfree[1], 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_p, 256);
    mbar_init(&bar_mma2, 1);
    mbar_init(&bar_dqfree, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0 && n_it > 0) {
      const int col = h * B5_HS;
      mbar_expect_tx(&bar_kv, 2 * B5_TILE);
      tma_load_2d(sK, &tmK, col, b * Tk + k0, &bar_kv);
      tma_load_2d(sV, &tmV, col, b * Tk + k0, &bar_kv);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        if (it >= 2) mbar_wait(&bar_qfree[s], (uint32_t)((it - 2) >> 1) & 1u);
        mbar_expect_tx(&bar_q[s], 2 * B5_TILE);
        tma_load_2d(sQ + s * B5_TILE, &tmQ, col, b * Tq + (i0 + it) * B5_BQ, &bar_q[s]);
        tma_load_2d(sdO + s * B5_TILE, &tmdO, col, b * Tq + (i0 + it) * B5_BQ, &bar_q[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10);                                       // D = f32, A = B = bf16
      const uint32_t idesc_s = base | ((uint32_t)(B5_BK >> 3) << 17) | ((uint32_t)(B5_BQ >> 4) << 24);                 // K-major x K-major
      const uint32_t idesc_kv = base | (1u << 15) | (1u << 16) | ((uint32_t)(B5_HS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // MN x MN
      const uint32_t idesc_q = base | (1u << 16) | ((uint32_t)(B5_HS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);      // K-major x MN
      const uint64_t kdesc = umma_desc_sw128(smem_u32(sK)), vdesc = umma_desc_sw128(smem_u32(sV));
      const uint64_t kdesc_mn = umma_desc_sw128_mn(smem_u32(sK));
      const uint64_t pdesc_mn = umma_desc_sw128_mn_lbo(smem_u32(sP), (uint32_t)B5_TILE);
      const uint64_t dsdesc_mn = umma_desc_sw128_mn_lbo(smem_u32(sdS), (uint32_t)B5_TILE);
      mbar_wait(&bar_kv, 0u);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        mbar_wait(&bar_q[s], (uint32_t)(it >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + s * B5_TILE)), dodesc = umma_desc_sw128(smem_u32(sdO + s * B5_TILE));
#pragma unroll
        for (int k = 0; k < B5_HS / 16; ++k)
          umma_bf16(tmem_base + B5_COL_S, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < B5_HS / 16; ++k)
          umma_bf16(tmem_base + B5_COL_DP, dodesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bar_s);
        mbar_wait(&bar_p, (uint32_t)it & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t qdesc_mn = umma_desc_sw128_mn(smem_u32(sQ + s * B5_TILE)), dodesc_mn = umma_desc_sw128_mn(smem_u32(sdO + s * B5_TILE));
#pragma unroll
        for (int kk = 0; kk < B5_BQ / 16; ++kk)          // reduction over the 128 queries, 16 (= 2048 B of every tile) per step
          umma_bf16(tmem_base + B5_COL_DV, pdesc_mn + (uint64_t)(128 * kk), dodesc_mn + (uint64_t)(128 * kk), idesc_kv, (it | kk) != 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < B5_BQ / 16; ++kk)
          umma_bf16(tmem_base + B5_COL_DK, dsdesc_mn + (uint64_t)(128 * kk), qdesc_mn + (uint64_t)(128 * kk), idesc_kv, (it | kk) != 0 ? 1u : 0u);
        if (it > 0) {
          mbar_wait(&bar_dqfree, (uint32_t)(it - 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
#pragma unroll
        for (int kk = 0; kk < B5_BK / 16; ++kk) {        // reduction over the 128 keys: K-major dS panels of 64 keys
          const uint64_t dsdesc = umma_desc_sw128(smem_u32(sdS + (kk >> 2) * B5_TILE)) + (uint64_t)(2 * (kk & 3));
          umma_bf16(tmem_base + B5_COL_DQ, dsdesc, kdesc_mn + (uint64_t)(128 * kk), idesc_q, kk != 0 ? 1u : 0u);
        }
        umma_commit(&bar_mma2);
        umma_commit(&bar_qfree[s]);
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3, wg = (warp - 4) >> 2;
    const int r = wq * 32 + lane;                                   // row of the tile = TMEM lane
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
    DropKey dkey{0u, 0u, 0u};
    if (drop.thr != 0u) dkey = drop_key(drop);
    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i0 + it) * B5_BQ, qi = q0 + r;
      // ---- per-row scalars: log-sum-exp (base 2) and delta = dO . O, read while the S / dP MMAs run ----
      float l2, delta;
      b5_row_scalars(out, dout, lse, b, h, H, Tq, qi, l2, delta);
      int lo, hi;
      b5_visible(mode, n_prompt, Tq, Tk, qi, l2, lo, hi);
      // ---- dQ of the previous query block: TMEM -> fp32 reductions ----
      if (it > 0) {
        mbar_wait(&bar_mma2, (uint32_t)(it - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[32];
        tmem_ld32(trow + B5_COL_DQ + (uint32_t)(wg * 32), v);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        b5_arrive(&bar_dqfree);
        const int qp = q0 - B5_BQ + r;
        if (qp < Tq) {
          float* dst = dq_acc + (((int64_t)b * H + h) * Tq + qp) * B5_HS + wg * 32;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            atomicAdd(reinterpret_cast<float4*>(dst + 4 * t), make_float4(__uint_as_float(v[4 * t]), __uint_as_float(v[4 * t + 1]),
                                                                         __uint_as_float(v[4 * t + 2]), __uint_as_float(v[4 * t + 3])));
        }
      }
      mbar_wait(&bar_s, (uint32_t)it & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // the MMAs of the previous query block have read P / dS (bar_mma2 was awaited above when it > 0)
      b5_softmax_half(trow + B5_COL_S + (uint32_t)(wg * 64), trow + B5_COL_DP + (uint32_t)(wg * 64), k0 + wg * 64, lo, hi, l2, delta,
                      scale, drop, dkey, (uint32_t)(((int64_t)b * H + h) * Tq + qi), sP + wg * B5_TILE + r * 128,
                      sdS + wg * B5_TILE + r * 128, r);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      b5_arrive(&bar_p);
    }
    if (n_it > 0) {
      // ---- last dQ block, then dK_j / dV_j ----
      mbar_wait(&bar_mma2, (uint32_t)(n_it - 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        uint32_t v[32];
        tmem_ld32(trow + B5_COL_DQ + (uint32_t)(wg * 32), v);
        const int qp = (i0 + n_it - 1) * B5_BQ + r;
        if (qp < Tq) {
          float* dst = dq_acc + (((int64_t)b * H + h) * Tq + qp) * B5_HS + wg * 32;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            atomicAdd(reinterpret_cast<float4*>(dst + 4 * t), make_float4(__uint_as_float(v[4 * t]), __uint_as_float(v[4 * t + 1]),
                                                                         __uint_as_float(v[4 * t + 2]), __uint_as_float(v[4 * t + 3])));
        }
      }
    }
    const int kj = k0 + r;                                         // TMEM lane = key row for dK / dV
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      uint32_t v[32];
      if (n_it > 0) {                                              // CTA-uniform: tcgen05.ld is a warp-collective, so the row guard
        tmem_ld32(trow + (which == 0 ? B5_COL_DV : B5_COL_DK) + (uint32_t)(wg * 32), v);   // below must not enclose it
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (kj < Tk) {
        b5_store32(v, (which == 0 ? dv : dk) + (int64_t)b * kv_bs + (int64_t)kj * kv_rs + (int64_t)h * B5_HS + wg * 32);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Resident variant for sequences of at most 256 rows (the decoder's block_size, the ViT's 197 tokens): ONE CTA per (batch, head)
// holds both key blocks, both query blocks (every tile is loaded exactly once) and, next to dK_j / dV_j of the current key block,
// BOTH dQ accumulators in TMEM (columns [384,512)): dQ needs no atomics, no zeroed fp32 workspace and no conversion pass -- the
// kernel writes dq / dk / dv in bf16 straight into the packed gradient buffer.  Pair order: key block j outer, query block i inner.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int R5_SMEM = 12 * B5_TILE + B5_BAR_BYTES + 2 * 128 * 4 + B5_SLACK;      // tiles, barriers, delta[2][128]
// timeline of CTA (0,0), SM clock ticks relative to its first stamp (i2t_attn_bwd_trace): [0..15] MMA thread, [16..31] one softmax thread
__device__ long long g_r5_trace[32];
#define R5_STAMP(slot) do { if (trace_on) g_r5_trace[slot] = clock64(); } while (0)

__global__ void __launch_bounds__(B5_THREADS, 1)
attn_bwd_tc5r_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                     const __grid_constant__ CUtensorMap tmO, const float* __restrict__ lse,
                     __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int H, int Tq, int Tk,
                     int64_t q_bs, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, int mode, int n_prompt, float scale, DropArgs drop,
                     DropArgs tok, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sK = smem;                          // [2 key blocks]
  uint8_t* sV = sK + 2 * B5_TILE;
  uint8_t* sQ = sV + 2 * B5_TILE;              // [2 query blocks]
  uint8_t* sdO = sQ + 2 * B5_TILE;
  uint8_t* sP = sdO + 2 * B5_TILE;             // [2 panels of 64 keys][128 queries][128 B]
  uint8_t* sdS = sP + 2 * B5_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 2 * B5_TILE);
  uint64_t* bar_kv = bars;                     // [2]
  uint64_t* bar_q = bars + 2;                  // [2]
  uint64_t& bar_s = bars[4];
  uint64_t& bar_p = bars[5];
  uint64_t& bar_mma2 = bars[6];
  uint64_t& bar_kvfree = bars[7];              // dK / dV columns of the finished key block drained (256 arrivals)
  uint64_t& bar_o = bars[8];                   // the forward's output rows (for delta), parked in the P panels until the first P
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 9);
  float* s_delta = reinterpret_cast<float*>(bars + 10);      // [2][128] behind the barriers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int nq = (Tq + B5_BQ - 1) / B5_BQ, nk = (Tk + B5_BK - 1) / B5_BK;      // 1 or 2 each (host-checked)
  const bool causal = mode != I2T_MASK_NONE;
  const bool trace_on = dbg < 0 && blockIdx.x == 0 && blockIdx.y == 0;
  if (trace_on && threadIdx.x == 32) g_r5_trace[0] = clock64();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmdO) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(&bar_q[0], 1);
    mbar_init(&bar_q[1], 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_p, 256);
    mbar_init(&bar_mma2, 1);
    mbar_init(&bar_kvfree, 256);
    mbar_init(&bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  if (trace_on && threadIdx.x == 32) g_r5_trace[13] = clock64();     // barriers + TMEM ready
  pdl_wait();
  if (trace_on && threadIdx.x == 32) g_r5_trace[14] = clock64();     // predecessor grid complete

  if (warp == 0) {
    if (lane == 0) {
      const int col = h * B5_HS;
      mbar_expect_tx(&bar_kv[0], 2 * B5_TILE);
      tma_load_2d(sK, &tmK, col, b * Tk, &bar_kv[0]);
      tma_load_2d(sV, &tmV, col, b * Tk, &bar_kv[0]);
      mbar_expect_tx(&bar_q[0], 2 * B5_TILE);
      tma_load_2d(sQ, &tmQ, col, b * Tq, &bar_q[0]);
      tma_load_2d(sdO, &tmdO, col, b * Tq, &bar_q[0]);
      mbar_expect_tx(&bar_o, nq * B5_TILE);
      tma_load_2d(sP, &tmO, col, b * Tq, &bar_o);
      if (nq > 1) tma_load_2d(sP + B5_TILE, &tmO, col, b * Tq + B5_BQ, &bar_o);
      if (nq > 1) {
        mbar_expect_tx(&bar_q[1], 2 * B5_TILE);
        tma_load_2d(sQ + B5_TILE, &tmQ, col, b * Tq + B5_BQ, &bar_q[1]);
        tma_load_2d(sdO + B5_TILE, &tmdO, col, b * Tq + B5_BQ, &bar_q[1]);
      }
      if (nk > 1) {
        mbar_expect_tx(&bar_kv[1], 2 * B5_TILE);
        tma_load_2d(sK + B5_TILE, &tmK, col, b * Tk + B5_BK, &bar_kv[1]);
        tma_load_2d(sV + B5_TILE, &tmV, col, b * Tk + B5_BK, &bar_kv[1]);
      }
      if (trace_on) g_r5_trace[15] = clock64();                        // all loads issued
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10);
      const uint32_t idesc_s = base | ((uint32_t)(B5_BK >> 3) << 17) | ((uint32_t)(B5_BQ >> 4) << 24);
      const uint32_t idesc_kv = base | (1u << 15) | (1u << 16) | ((uint32_t)(B5_HS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_q = base | (1u << 16) | ((uint32_t)(B5_HS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t pdesc_mn = umma_desc_sw128_mn_lbo(smem_u32(sP), (uint32_t)B5_TILE);
      const uint64_t dsdesc_mn = umma_desc_sw128_mn_lbo(smem_u32(sdS), (uint32_t)B5_TILE);
      int t = 0;
      for (int j = 0; j < nk; ++j) {
        const int ifirst = causal ? j : 0;
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + j * B5_TILE)), vdesc = umma_desc_sw128(smem_u32(sV + j * B5_TILE));
        const uint64_t kdesc_mn = umma_desc_sw128_mn(smem_u32(sK + j * B5_TILE));
        for (int i = ifirst; i < nq; ++i, ++t) {
          mbar_wait(&bar_kv[j], 0u);
          mbar_wait(&bar_q[i], 0u);
          R5_STAMP(1 + 4 * t);           // operands in shared memory
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + i * B5_TILE)), dodesc = umma_desc_sw128(smem_u32(sdO + i * B5_TILE));
#pragma unroll
          for (int k = 0; k < B5_HS / 16; ++k)
            umma_bf16(tmem_base + B5_COL_S, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < B5_HS / 16; ++k)
            umma_bf16(tmem_base + B5_COL_DP, dodesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&bar_s);
          R5_STAMP(2 + 4 * t);           // S / dP issued
          mbar_wait(&bar_p, (uint32_t)t & 1u);
          if (j > 0 && i == ifirst) mbar_wait(&bar_kvfree, (uint32_t)(j - 1) & 1u);   // dK / dV of the previous key block are out
          R5_STAMP(3 + 4 * t);           // P / dS arrived
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t qdesc_mn = umma_desc_sw128_mn(smem_u32(sQ + i * B5_TILE)), dodesc_mn = umma_desc_sw128_mn(smem_u32(sdO + i * B5_TILE));
          const uint32_t acc_kv = i != ifirst ? 1u : 0u;
#pragma unroll
          for (int kk = 0; kk < B5_BQ / 16; ++kk)
            umma_bf16(tmem_base + B5_COL_DV, pdesc_mn + (uint64_t)(128 * kk), dodesc_mn + (uint64_t)(128 * kk), idesc_kv, acc_kv | (kk != 0 ? 1u : 0u));
#pragma unroll
          for (int kk = 0; kk < B5_BQ / 16; ++kk)
            umma_bf16(tmem_base + B5_COL_DK, dsdesc_mn + (uint64_t)(128 * kk), qdesc_mn + (uint64_t)(128 * kk), idesc_kv, acc_kv | (kk != 0 ? 1u : 0u));
          const uint32_t acc_q = j != 0 ? 1u : 0u;                       // key block 0 reaches every query block first
#pragma unroll
          for (int kk = 0; kk < B5_BK / 16; ++kk) {
            const uint64_t dsdesc = umma_desc_sw128(smem_u32(sdS + (kk >> 2) * B5_TILE)) + (uint64_t)(2 * (kk & 3));
            umma_bf16(tmem_base + B5_COL_DQ + (uint32_t)(64 * i), dsdesc, kdesc_mn + (uint64_t)(128 * kk), idesc_q, acc_q | (kk != 0 ? 1u : 0u));
          }
          umma_commit(&bar_mma2);
          R5_STAMP(4 + 4 * t);           // dV / dK / dQ issued
        }
      }
    }
  } else if (warp >= 4) {
    const int wq = warp & 3, wg = (warp - 4) >> 2;
    const int r = wq * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
    const bool trace_c = trace_on && threadIdx.x == 128;
    DropKey dkey{0u, 0u, 0u}, tkey{0u, 0u, 0u};
    if (drop.thr != 0u) dkey = drop_key(drop);
    if (tok.thr != 0u) tkey = drop_key(tok);
    // ---- per-row scalars.  delta[i][r] = dO_i[r,:] . O_i[r,:]: warpgroup i takes query block i, both rows come from the swizzled
    // shared-memory tiles (O_i is parked in P panel i, whose row r only this thread writes later), exchanged through s_delta ----
    float l2[2], delta[2];
    l2[0] = r < Tq ? lse[((int64_t)b * H + h) * Tq + r] * 1.4426950408889634f : -INFINITY;
    l2[1] = B5_BQ + r < Tq ? lse[((int64_t)b * H + h) * Tq + B5_BQ + r] * 1.4426950408889634f : -INFINITY;
    if (wg < nq) {
      mbar_wait(&bar_o, 0u);
      mbar_wait(&bar_q[wg], 0u);
      const uint8_t* orow = sP + wg * B5_TILE + r * 128;
      const uint8_t* drow = sdO + wg * B5_TILE + r * 128;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 a = *reinterpret_cast<const uint4*>(orow + ((c ^ (r & 7)) << 4));
        const uint4 d = *reinterpret_cast<const uint4*>(drow + ((c ^ (r & 7)) << 4));
        const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* dp2 = reinterpret_cast<const __nv_bfloat162*>(&d);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = __bfloat1622float2(ap[e]), y = __bfloat1622float2(dp2[e]);
          acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
        }
      }
      s_delta[wg * 128 + r] = acc;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");            // the eight softmax warps only
    delta[0] = s_delta[r];
    delta[1] = nq > 1 ? s_delta[128 + r] : 0.f;
    if (trace_c) g_r5_trace[16] = clock64();        // row scalars read
    int t = 0;
    for (int j = 0; j < nk; ++j) {
      const int ifirst = causal ? j : 0;
      const int kj = j * B5_BK + r;
      for (int i = ifirst; i < nq; ++i, ++t) {
        const int qi = i * B5_BQ + r;
        const float l2i = i == 0 ? l2[0] : l2[1], di = i == 0 ? delta[0] : delta[1];
        int lo, hi;
        b5_visible(mode, n_prompt, Tq, Tk, qi, l2i, lo, hi);
        if (t > 0) mbar_wait(&bar_mma2, (uint32_t)(t - 1) & 1u);          // the previous pair's MMAs have read P / dS
        mbar_wait(&bar_s, (uint32_t)t & 1u);
        if (trace_c) g_r5_trace[17 + 4 * t] = clock64();   // S / dP complete
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        b5_softmax_half(trow + B5_COL_S + (uint32_t)(wg * 64), trow + B5_COL_DP + (uint32_t)(wg * 64), j * B5_BK + wg * 64, lo, hi, l2i,
                        di, scale, drop, dkey, (uint32_t)(((int64_t)b * H + h) * Tq + qi), sP + wg * B5_TILE + r * 128,
                        sdS + wg * B5_TILE + r * 128, r);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        b5_arrive(&bar_p);
        if (trace_c) g_r5_trace[18 + 4 * t] = clock64();   // P / dS written
        if (i == nq - 1) {                                               // last pair of key block j: dK_j / dV_j are complete
          mbar_wait(&bar_mma2, (uint32_t)t & 1u);
          if (trace_c) g_r5_trace[19 + 4 * t] = clock64();   // dV / dK / dQ MMAs complete
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t v0[32], v1[32];
          tmem_ld32(trow + B5_COL_DV + (uint32_t)(wg * 32), v0);
          tmem_ld32(trow + B5_COL_DK + (uint32_t)(wg * 32), v1);
          if (trace_c) g_r5_trace[29 + (j + 1 < nk ? 0 : 1)] = clock64();   // dK / dV read from TMEM
          if (j + 1 < nk) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            b5_arrive(&bar_kvfree);
          }
          // the pair's MMAs are complete: the P / dS panels are free until the next softmax
          // token-level q / k / v dropout of the reference's attention (models/layers.py:454-461), backward half: the gradient of
          // the packed qkv buffer is scaled per (row, segment) on its way out instead of by a separate pass over the buffer
          float kd = 1.f, vd = 1.f;
          if (tok.thr != 0u) {
            const Philox4 tr = drop_elem4(tok, tkey, (uint64_t)((int64_t)b * Tk + kj));
            kd = tr.y >= tok.thr ? tok.inv_keep : 0.f;
            vd = tr.z >= tok.thr ? tok.inv_keep : 0.f;
          }
          b5_stage32(v0, sP, r, wg, vd);
          b5_stage32(v1, sP + B5_TILE, r, wg, kd);
          if (j + 1 == nk) {                                             // last key block: dQ_0 / dQ_1 are complete as well
            for (int i2 = 0; i2 < nq; ++i2) {
              tmem_ld32(trow + B5_COL_DQ + (uint32_t)(64 * i2 + wg * 32), v0);
              float qd = 1.f;
              if (tok.thr != 0u) {
                const Philox4 tr = drop_elem4(tok, tkey, (uint64_t)((int64_t)b * Tq + i2 * B5_BQ + r));
                qd = tr.x >= tok.thr ? tok.inv_keep : 0.f;
              }
              b5_stage32(v0, sdS + i2 * B5_TILE, r, wg, qd);
            }
          }
          b5_sync_softmax_warps();
          const int tid = threadIdx.x - 128;
          b5_store_tile(sP, dv + (int64_t)b * kv_bs + (int64_t)(j * B5_BK) * kv_rs + (int64_t)h * B5_HS, kv_rs, Tk - j * B5_BK, tid);
          b5_store_tile(sP + B5_TILE, dk + (int64_t)b * kv_bs + (int64_t)(j * B5_BK) * kv_rs + (int64_t)h * B5_HS, kv_rs, Tk - j * B5_BK, tid);
          if (j + 1 == nk) {
            for (int i2 = 0; i2 < nq; ++i2)
              b5_store_tile(sdS + i2 * B5_TILE, dq + (int64_t)b * q_bs + (int64_t)(i2 * B5_BQ) * q_rs + (int64_t)h * B5_HS, q_rs,
                            Tq - i2 * B5_BQ, tid);
          } else {
            b5_sync_softmax_warps();                                     // the panels are rewritten by the next softmax
          }
          if (trace_c) g_r5_trace[20 + 4 * t] = clock64();   // dK / dV stored
        }
      }
      if (ifirst >= nq && kj < Tk) {                                     // keys no query can see (Tk > Tq under a causal mask)
        uint32_t z[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) z[e] = 0u;
        b5_store32(z, dv + (int64_t)b * kv_bs + (int64_t)kj * kv_rs + (int64_t)h * B5_HS + wg * 32);
        b5_store32(z, dk + (int64_t)b * kv_bs + (int64_t)kj * kv_rs + (int64_t)h * B5_HS + wg * 32);
      }
    }
    if (trace_c) g_r5_trace[31] = clock64();          // dQ stored
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace i2t

// Debugging aid: with I2T_ATTN_BWD_DEBUG=-1 in the environment CTA (0,0) of every resident-kernel launch stamps its timeline; this
// copies the 32 stamps of the most recent launch to `host_out` (synchronises the device).
extern "C" int i2t_attn_bwd_trace(long long* host_out) {
  I2T_REQUIRE(host_out, "attn_bwd_trace: null pointer");
  I2T_CUDA(cudaDeviceSynchronize());
  I2T_CUDA(cudaMemcpyFromSymbol(host_out, i2t::g_r5_trace, sizeof(long long) * 32));
  return I2T_OK;
}

namespace i2t {
// returns 1 when it handled the call (dq, dk, dv written; no workspace used), 0 when the shape is not eligible
int attn_bwd_tc5r(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, void* dq, void* dk,
                  void* dv, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_bs, int64_t q_rs, int64_t kv_bs,
                  int64_t kv_rs, int mode, int64_t n_prompt, DropArgs drop, DropArgs tok, cudaStream_t st) {
  if (head_dim != B5_HS || Tq > 2 * B5_BQ || Tk > 2 * B5_BK) return 0;
  if (mode != I2T_MASK_NONE && Tk > Tq && (Tk + B5_BK - 1) / B5_BK > (Tq + B5_BQ - 1) / B5_BQ) return 0;
  if (q_bs != Tq * q_rs || kv_bs != Tk * kv_rs) return 0;
  if (q_rs % 8 != 0 || kv_rs % 8 != 0 || (H * B5_HS) % 8 != 0) return 0;
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out) || !aligned16(dout) || !aligned16(dq) || !aligned16(dk) ||
      !aligned16(dv))
    return 0;
  if (B > 65535 || H > 65535) return 0;
  CUtensorMap mq, mk, mv, mdo;
  int rc = tc_make_map(q, B * Tq, H * B5_HS, q_rs, 128, &mq);
  if (rc != I2T_OK) return rc;
  rc = tc_make_map(k, B * Tk, H * B5_HS, kv_rs, 128, &mk);
  if (rc != I2T_OK) return rc;
  rc = tc_make_map(v, B * Tk, H * B5_HS, kv_rs, 128, &mv);
  if (rc != I2T_OK) return rc;
  rc = tc_make_map(dout, B * Tq, H * B5_HS, H * B5_HS, 128, &mdo);
  if (rc != I2T_OK) return rc;
  CUtensorMap mo;
  rc = tc_make_map(out, B * Tq, H * B5_HS, H * B5_HS, 128, &mo);
  if (rc != I2T_OK) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc5r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, R5_SMEM);
    if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "cudaFuncSetAttribute(attn_bwd_tc5r_kernel): %s", cudaGetErrorString(e));
    attr = true;
  }
  static const int dbg = (getenv("I2T_ATTN_BWD_DEBUG") && atoi(getenv("I2T_ATTN_BWD_DEBUG")) < 0) ? -1 : 0;   // -1: CTA (0,0) stamps its timeline
  dim3 grid((unsigned)H, (unsigned)B);
  (void)launch_pdl(attn_bwd_tc5r_kernel, grid, dim3(B5_THREADS), (size_t)R5_SMEM, st, mq, mk, mv, mdo, mo, lse, (__nv_bfloat16*)dq, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, (int)H, (int)Tq, (int)Tk, q_bs,
                   q_rs, kv_bs, kv_rs, mode, (int)n_prompt, 1.0f / sqrtf((float)head_dim), drop, tok, dbg);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "attn_bwd_tc5r launch failed: %s", cudaGetErrorString(e));
  return 1;
}

// returns 1 when it handled the call, 0 when the shape is not eligible (the caller then runs the mma.sync kernel)
int attn_bwd_tc5(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, float* dq_acc,
                 void* dk, void* dv, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_bs, int64_t q_rs,
                 int64_t kv_bs, int64_t kv_rs, int mode, int64_t n_prompt, DropArgs drop, cudaStream_t st) {
  if (head_dim != B5_HS) return 0;
  if (q_bs != Tq * q_rs || kv_bs != Tk * kv_rs) return 0;                // batches must be row-contiguous for one 2-D tensor map
  if (q_rs % 8 != 0 || kv_rs % 8 != 0 || (H * B5_HS) % 8 != 0) return 0;
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out) || !aligned16(dout) || !aligned16(dk) || !aligned16(dv) ||
      !aligned16(dq_acc))
    return 0;
  if (B > 65535 || H > 65535) return 0;
  CUtensorMap mq, mk, mv, mdo;
  int rc = tc_make_map(q, B * Tq, H * B5_HS, q_rs, 128, &mq);
  if (rc != I2T_OK) return rc;
  rc = tc_make_map(k, B * Tk, H * B5_HS, kv_rs, 128, &mk);
  if (rc != I2T_OK) return rc;
  rc = tc_make_map(v, B * Tk, H * B5_HS, kv_rs, 128, &mv);
  if (rc != I2T_OK) return rc;
  rc = tc_make_map(dout, B * Tq, H * B5_HS, H * B5_HS, 128, &mdo);
  if (rc != I2T_OK) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B5_SMEM);
    if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "cudaFuncSetAttribute(attn_bwd_tc5_kernel): %s", cudaGetErrorString(e));
    attr = true;
  }
  dim3 grid((unsigned)ceil_div(Tk, B5_BK), (unsigned)H, (unsigned)B);
  (void)launch_pdl(attn_bwd_tc5_kernel, grid, dim3(B5_THREADS), (size_t)B5_SMEM, st, mq, mk, mv, mdo, (const __nv_bfloat16*)out,
                   (const __nv_bfloat16*)dout, lse, dq_acc, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, (int)H, (int)Tq, (int)Tk, kv_bs, kv_rs,
                   mode, (int)n_prompt, 1.0f / sqrtf((float)head_dim), drop);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "attn_bwd_tc5 launch failed: %s", cudaGetErrorString(e));
  return 1;
}

}  // namespace i2t
