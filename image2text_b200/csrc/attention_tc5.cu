// bf16 attention forward on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), head_dim 64, up to 384 keys.
// Same contract as attn_fwd_kernel (attention.cu): packed strided q/k/v read in place, closed-form masks, row
// log-sum-exp out.  Replaces F.scaled_dot_product_attention at reference models/layers.py:465 and torchvision's MHA core
// (:113) under bf16 autocast.
//
// One CTA = 128 queries of one (batch, head).  The sequences on this path are short (197 ViT tokens, 256 / 272 decoder
// rows), so the WHOLE score row block lives in TMEM and the softmax is exact, not online:
//   warp 0  : TMA -- Q tile, then every K block and V block (128 keys x 64, 128-byte swizzle), one mbarrier each;
//   warp 1  : one thread issues S_j = Q K_j^T (tcgen05.mma M128 N128 K16 x 4) for every key block j into TMEM columns
//             [128 j, 128 j + 128); later O += P_j V_j (M128 N64 K16 x 8, V as the MN-major B operand straight from its
//             [key][dim] rows) into TMEM columns [0, 64) -- the columns of S_0, dead once P_0 exists;
//   warps 4-7: one thread per query row: pass A reads the row from TMEM (tcgen05.ld) for the masked maximum, pass B
//             re-reads it, exponentiates, accumulates the row sum and writes P_j as bf16 into shared memory in the
//             K-major 128-byte-swizzled layout the MMA's A operand expects (fence.proxy.async before signalling);
//             epilogue: O from TMEM, scaled by 1 / sum, bf16 rows out, log-sum-exp out.
// Keys beyond the causal diagonal of the tile are never loaded.  2 CTAs per SM when the sequence has <= 256 keys
// (113 KB of shared memory, 256 TMEM columns each).
#include "common.cuh"
#include "rng.cuh"
#include "tc_common.cuh"

namespace i2t {

constexpr int A5_BQ = 128, A5_BK = 128, A5_HS = 64, A5_THREADS = 256;
constexpr int A5_TILE_BYTES = 128 * A5_HS * 2;            // 16 KB: Q tile, one K block, one V block
constexpr int A5_P_BYTES = A5_BQ * A5_BK * 2;             // 32 KB: one P block (two 64-key K-major sub-blocks)
constexpr int A5_SLACK = 512, A5_BAR_BYTES = 128;
constexpr int a5_smem(int nb) { return (1 + 2 * nb) * A5_TILE_BYTES + A5_P_BYTES + A5_BAR_BYTES + A5_SLACK; }

__device__ __forceinline__ bool a5_visible(int mode, int n_prompt, int qi, int kj) {
  if (mode == I2T_MASK_NONE) return true;
  if (kj > qi) return false;
  if (mode == I2T_MASK_CAUSAL) return true;
  return qi < n_prompt ? true : kj >= n_prompt;
}
__device__ __forceinline__ float a5_ex2(float x) {      // one MUFU op; exp2f() adds range fix-ups the softmax does not need
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t a5_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int NB>     // key blocks the kernel can hold (2 or 3)
__global__ void __launch_bounds__(A5_THREADS, NB == 2 ? 2 : 1)
attn_fwd_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int H,
                    int Tq, int Tk, int mode, int n_prompt, float scale_log2, DropArgs drop) {
  constexpr uint32_t TMEM_COLS = NB == 2 ? 256 : 512;
  // no static shared memory: two CTAs of the 2-block variant must fit one SM, so the barriers live behind the tiles and
  // the 1024-byte alignment of the swizzled tiles may cost at most A5_SLACK bytes
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  if (pad > (uint32_t)A5_SLACK) asm volatile("trap;");
  uint8_t* smem = smem_raw + pad;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + A5_TILE_BYTES;
  uint8_t* sV = sK + NB * A5_TILE_BYTES;
  uint8_t* sP = sV + NB * A5_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + A5_P_BYTES);
  uint64_t* bar_k = bars;
  uint64_t* bar_v = bars + 3;
  uint64_t* bar_s = bars + 6;
  uint64_t& bar_p = bars[9];
  uint64_t& bar_pfree = bars[10];
  uint64_t& bar_o = bars[11];
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * A5_BQ, h = blockIdx.y, b = blockIdx.z;
  int kend = Tk;
  if (mode != I2T_MASK_NONE) kend = min(Tk, q0 + A5_BQ);        // keys beyond the tile's last query are never visible
  const int nb = (kend + A5_BK - 1) / A5_BK;                     // <= NB (host-checked)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int j = 0; j < NB; ++j) {
      mbar_init(&bar_k[j], 1);
      mbar_init(&bar_v[j], 1);
      mbar_init(&bar_s[j], 1);
    }
    mbar_init(&bar_p, 128);
    mbar_init(&bar_pfree, 1);
    mbar_init(&bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();     // programmatic dependent launch: barrier init / TMEM allocation above overlap the previous kernel
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                  // q / k / v are produced by the previous kernels

  if (warp == 0) {
    if (lane == 0) {
      const int col = h * A5_HS;
      for (int j = 0; j < nb; ++j) {
        mbar_expect_tx(&bar_k[j], (j == 0 ? 2 : 1) * A5_TILE_BYTES);
        if (j == 0) tma_load_2d(sQ, &tmQ, col, b * Tq + q0, &bar_k[0]);
        tma_load_2d(sK + j * A5_TILE_BYTES, &tmK, col, b * Tk + j * A5_BK, &bar_k[j]);
      }
      for (int j = 0; j < nb; ++j) {
        mbar_expect_tx(&bar_v[j], A5_TILE_BYTES);
        tma_load_2d(sV + j * A5_TILE_BYTES, &tmV, col, b * Tk + j * A5_BK, &bar_v[j]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // S_j = Q K_j^T : D=f32, A=B=bf16, both K-major, M=128, N=128
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(A5_BK >> 3) << 17) | ((uint32_t)(A5_BQ >> 4) << 24);
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      for (int j = 0; j < nb; ++j) {
        mbar_wait(&bar_k[j], 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + j * A5_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < A5_HS / 16; ++k)
          umma_bf16(tmem_base + (uint32_t)(j * A5_BK), qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bar_s[j]);
      }
      // O += P_j V_j : A = P (K-major, shared memory), B = V_j (MN-major: [key][dim] rows), M=128, N=64
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(A5_HS >> 3) << 17) |
                               ((uint32_t)(A5_BQ >> 4) << 24);
      for (int j = 0; j < nb; ++j) {
        mbar_wait(&bar_v[j], 0u);
        mbar_wait(&bar_p, (uint32_t)j & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t vdesc = umma_desc_sw128_mn(smem_u32(sV + j * A5_TILE_BYTES));
#pragma unroll
        for (int kk = 0; kk < A5_BK / 16; ++kk) {
          const uint64_t pdesc = umma_desc_sw128(smem_u32(sP + (kk >> 2) * (A5_P_BYTES / 2))) + (uint64_t)(2 * (kk & 3));
          umma_bf16(tmem_base, pdesc, vdesc + (uint64_t)(128 * kk), idesc_o, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&bar_pfree);       // P may be overwritten once these MMAs have read it
      }
      umma_commit(&bar_o);
    }
  } else if (warp >= 4) {
    const int wq = warp & 3;
    const int r = wq * 32 + lane;                                   // query row of the tile = TMEM lane
    const int qi = q0 + r;
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
    // The closed-form masks make the visible keys of a row ONE interval [lo, hi): a 32-key chunk is classified once
    // (all visible / none / straddling) and only straddling chunks -- the diagonal -- pay per-element tests.  The softmax
    // threads are instruction-bound (one row of up to 384 keys each), so this is what sets the kernel's latency.
    int lo = 0, hi = Tk;
    if (mode != I2T_MASK_NONE) {
      hi = min(Tk, qi + 1);
      if (mode == I2T_MASK_PROMPT && qi >= n_prompt) lo = n_prompt;
    }
    // ---- pass A: masked row maximum over every key block ----
    float mx = -INFINITY;
    for (int j = 0; j < nb; ++j) {
      mbar_wait(&bar_s[j], 0u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c = 0; c < A5_BK / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(trow + (uint32_t)(j * A5_BK + c * 32), v);
        const int c0 = j * A5_BK + c * 32;
        if (c0 >= lo && c0 + 32 <= hi) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else if (c0 + 32 > lo && c0 < hi) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i >= lo && c0 + i < hi) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
    }
    const float m_use = mx == -INFINITY ? 0.f : mx;
    const float m_scaled = m_use * scale_log2;
    // ---- pass B: P_j = exp2(S_j * scale - m * scale) as bf16 into the swizzled A-operand tile; row sum ----
    float l = 0.f;
    for (int j = 0; j < nb; ++j) {
      if (j > 0) mbar_wait(&bar_pfree, (uint32_t)(j - 1) & 1u);    // the MMAs that read P_{j-1} have finished
#pragma unroll 1
      for (int c = 0; c < A5_BK / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(trow + (uint32_t)(j * A5_BK + c * 32), v);
        float p[32];
        const int c0 = j * A5_BK + c * 32;
        if (c0 >= lo && c0 + 32 <= hi) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            p[i] = a5_ex2(fmaf(__uint_as_float(v[i]), scale_log2, -m_scaled));
            l += p[i];
          }
        } else if (c0 + 32 > lo && c0 < hi) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool vis = c0 + i >= lo && c0 + i < hi;
            p[i] = vis ? a5_ex2(fmaf(__uint_as_float(v[i]), scale_log2, -m_scaled)) : 0.f;
            l += p[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) p[i] = 0.f;
        }
        if (drop.thr != 0u && c0 + 32 > lo && c0 < hi) {   // dropout on the probabilities (l keeps every key): 4 Philox calls per
                                                           // 32 keys (16 bits per key); chunks without a visible key are zero already
          const DropKey dkey = drop_key(drop);
          const uint32_t row = (uint32_t)(((int64_t)b * H + h) * Tq + qi);
#pragma unroll
          for (int bl = 0; bl < 2; ++bl) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {                 // one call = 8 keys: pairs 2 hh and 2 hh + 1 of the 16-key block
              const Philox4 rr = drop_attn8(drop, dkey, row, (uint32_t)((j * A5_BK + c * 32) >> 4) + bl, (uint32_t)hh);
#pragma unroll
              for (int sp = 0; sp < 2; ++sp) {
                const int e0 = bl * 16 + 2 * (2 * hh + sp);
                const uint32_t w0 = sp == 0 ? rr.x : rr.z, w1 = sp == 0 ? rr.y : rr.w;
                if ((w0 & 0xFFFFu) < drop.thr16) p[e0] = 0.f;
                if ((w0 >> 16) < drop.thr16) p[e0 + 1] = 0.f;
                if ((w1 & 0xFFFFu) < drop.thr16) p[e0 + 8] = 0.f;
                if ((w1 >> 16) < drop.thr16) p[e0 + 9] = 0.f;
              }
            }
          }
        }
        // 32 keys = four 16-byte chunks of the (c / 2)-th 64-key sub-block, chunk index (c % 2) * 4 + t, XOR-swizzled by row
        uint8_t* sub = sP + (c >> 1) * (A5_P_BYTES / 2) + r * 128;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint4 pk;
          pk.x = a5_pack(p[8 * t], p[8 * t + 1]);
          pk.y = a5_pack(p[8 * t + 2], p[8 * t + 3]);
          pk.z = a5_pack(p[8 * t + 4], p[8 * t + 5]);
          pk.w = a5_pack(p[8 * t + 6], p[8 * t + 7]);
          const int c16 = (c & 1) * 4 + t;
          *reinterpret_cast<uint4*>(sub + ((c16 ^ (r & 7)) << 4)) = pk;
        }
      }
      // generic-proxy writes -> visible to the tensor core (async proxy), then signal
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_p)) : "memory");
    }
    // ---- epilogue: O / l ----
    mbar_wait(&bar_o, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = l > 0.f ? drop.inv_keep / l : 0.f;
    __nv_bfloat16* op = out + ((int64_t)b * Tq + qi) * ((int64_t)H * A5_HS) + (int64_t)h * A5_HS;
#pragma unroll 1
    for (int c = 0; c < A5_HS / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(trow + (uint32_t)(c * 32), v);
      if (qi < Tq) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint4 pk;
          pk.x = a5_pack(__uint_as_float(v[8 * t]) * inv, __uint_as_float(v[8 * t + 1]) * inv);
          pk.y = a5_pack(__uint_as_float(v[8 * t + 2]) * inv, __uint_as_float(v[8 * t + 3]) * inv);
          pk.z = a5_pack(__uint_as_float(v[8 * t + 4]) * inv, __uint_as_float(v[8 * t + 5]) * inv);
          pk.w = a5_pack(__uint_as_float(v[8 * t + 6]) * inv, __uint_as_float(v[8 * t + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c * 32 + t * 8) = pk;
        }
      }
    }
    if (lse != nullptr && qi < Tq)
      lse[((int64_t)b * H + h) * Tq + qi] = l > 0.f ? (m_use * scale_log2 + log2f(l)) * 0.6931471805599453f : -INFINITY;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// returns 1 when it handled the call, 0 when the shape is not eligible (the caller then runs the mma.sync kernel)
int attn_fwd_tc5(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H, int64_t Tq, int64_t Tk,
                 int64_t head_dim, int64_t q_bs, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, int mode, int64_t n_prompt,
                 DropArgs drop, cudaStream_t st) {
  if (head_dim != A5_HS || Tk > 3 * A5_BK) return 0;
  if (q_bs != Tq * q_rs || kv_bs != Tk * kv_rs) return 0;                // batches must be row-contiguous for one 2-D tensor map
  if (q_rs % 8 != 0 || kv_rs % 8 != 0 || !aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(out)) return 0;
  if (B > 65535 || H > 65535) return 0;
  CUtensorMap mq, mk, mv;
  int rc = tc_make_map(q, B * Tq, H * A5_HS, q_rs, 128, &mq);
  if (rc != I2T_OK) return rc;
  rc = tc_make_map(k, B * Tk, H * A5_HS, kv_rs, 128, &mk);
  if (rc != I2T_OK) return rc;
  rc = tc_make_map(v, B * Tk, H * A5_HS, kv_rs, 128, &mv);
  if (rc != I2T_OK) return rc;
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
  dim3 grid((unsigned)ceil_div(Tq, A5_BQ), (unsigned)H, (unsigned)B);
  const int nbmax = (int)ceil_div(Tk, A5_BK);
  static bool attr2 = false, attr3 = false;
  if (nbmax <= 2) {
    constexpr int SMEM = a5_smem(2);
    if (!attr2) {
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc5_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "cudaFuncSetAttribute(attn_fwd_tc5_kernel): %s", cudaGetErrorString(e));
      attr2 = true;
    }
    (void)launch_pdl(attn_fwd_tc5_kernel<2>, grid, dim3(A5_THREADS), (size_t)SMEM, st, mq, mk, mv, (__nv_bfloat16*)out, lse, (int)H, (int)Tq,
                     (int)Tk, mode, (int)n_prompt, scale_log2, drop);
  } else {
    constexpr int SMEM = a5_smem(3);
    if (!attr3) {
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc5_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "cudaFuncSetAttribute(attn_fwd_tc5_kernel): %s", cudaGetErrorString(e));
      attr3 = true;
    }
    (void)launch_pdl(attn_fwd_tc5_kernel<3>, grid, dim3(A5_THREADS), (size_t)SMEM, st, mq, mk, mv, (__nv_bfloat16*)out, lse, (int)H, (int)Tq,
                     (int)Tk, mode, (int)n_prompt, scale_log2, drop);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "attn_fwd_tc5 launch failed: %s", cudaGetErrorString(e));
  return 1;
}

}  // namespace i2t
