// bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (sm_100a): TMA (cp.async.bulk.tensor, 128-byte swizzle)
// feeds a 6-stage shared-memory ring, ONE elected thread issues tcgen05.mma (M=128, N=128, K=16) with the accumulator
// in TMEM, four epilogue warps read it back with tcgen05.ld and apply bias / GELU / residual / cast on the way out.
// PERSISTENT: one CTA per SM loops over output tiles with a double-buffered TMEM accumulator, so a tile's epilogue and
// the next tile's TMA prologue overlap the tensor-core main loop (the training shapes have K = 768: 12 k-blocks).
//   C[M,N] = act(A[M,K] * B[N,K]^T + bias) + residual      (both operands K-major: activations x nn.Linear weight)
// Warp roles (256 threads): 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue (TMEM lanes 32w..).
// M / N / K tails are handled by TMA zero fill on the way in and predicated stores on the way out.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_common.cuh"

namespace i2t {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 6, TC_THREADS = 256;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2, TC_B_BYTES = TC_BN * TC_BK * 2;
constexpr int TC_SMEM = TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + 2 * 128 * 128 + 1024;   // ring + 2 epilogue staging boxes

struct TcEpilogue {
  const float* bias;
  const void* residual;
  void* C;
  int64_t ldc;
  int M, N;
  int act, accumulate, res_dtype, c_dtype;
  int debug;
  int tma_store;      // 1: the output tile leaves through shared memory + cp.async.bulk.tensor stores (tmC is valid)
};

// Split-K epilogue chunk: the partial sums of one K slice are ADDED to the fp32 output with red.global (the host zeroed C,
// or C already holds the in-place residual / the running weight gradient); the bias rides on slice 0.
__device__ __forceinline__ void tc_epilogue_chunk_splitk(const TcEpilogue& epi, const uint32_t (&r)[32], int64_t row, int64_t n0,
                                                         const float* sb) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  const int nvalid = (int)min((int64_t)32, epi.N - n0);
  if (sb != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + j);
      v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
    }
  }
  float* cp = (float*)epi.C + row * epi.ldc + n0;
  if (nvalid == 32 && ((uintptr_t)cp & 15u) == 0) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) atomicAdd(reinterpret_cast<float4*>(cp + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nvalid) atomicAdd(cp + j, v[j]);
  }
}

// One epilogue chunk: 32 accumulator columns of this thread's row -> bias / activation / residual / cast -> global.
__device__ __forceinline__ void tc_epilogue_chunk(const TcEpilogue& epi, const uint32_t (&r)[32], int64_t row, int64_t n0,
                                                  const float* sb) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  const int nvalid = (int)min((int64_t)32, epi.N - n0);
  if (sb != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + j);
      v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
    }
  }
  if (epi.act != I2T_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], epi.act);
  }
  const int64_t off = row * epi.ldc + n0;
  if (epi.residual != nullptr) {
    if (epi.res_dtype == I2T_F32) {
      const float* rp = (const float*)epi.residual + off;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += rp[j];
    } else {
      const __nv_bfloat16* rp = (const __nv_bfloat16*)epi.residual + off;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += __bfloat162float(rp[j]);
    }
  }
  if (epi.c_dtype == I2T_F32) {
    float* cp = (float*)epi.C + off;
    if (epi.accumulate) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += cp[j];
    }
    if (nvalid == 32 && ((uintptr_t)cp & 15u) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) cp[j] = v[j];
    }
  } else {
    __nv_bfloat16* cp = (__nv_bfloat16*)epi.C + off;
    if (nvalid == 32 && ((uintptr_t)cp & 15u) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(cp + j) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) cp[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// ---- staged epilogue: the 128 x 32 accumulator chunk of the four epilogue warps goes through a 128-byte-swizzled
//      shared-memory box and leaves with ONE cp.async.bulk.tensor store per 128-byte-wide box (full lines), instead of
//      32 rows x 16 B scattered stores per warp instruction (measured: 62 % -> 83 % of peak at K = 768) ----
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1)
               : "memory");
}
// bias / activation / residual on a chunk of 32 columns held by one thread (one row)
__device__ __forceinline__ void tc_chunk_math(const TcEpilogue& epi, const uint32_t (&r)[32], float (&v)[32], int64_t row, int64_t n0,
                                              bool row_ok, const float* sb) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  const int nvalid = (int)max((int64_t)0, min((int64_t)32, epi.N - n0));
  if (sb != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + j);
      v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
    }
  }
  if (epi.act != I2T_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], epi.act);
  }
  if (epi.residual != nullptr && row_ok) {
    const int64_t off = row * epi.ldc + n0;
    if (epi.res_dtype == I2T_F32) {
      const float* rp = (const float*)epi.residual + off;
      if (nvalid == 32 && ((uintptr_t)rp & 15u) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 x = *reinterpret_cast<const float4*>(rp + j);
          v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) v[j] += rp[j];
      }
    } else {
      const __nv_bfloat16* rp = (const __nv_bfloat16*)epi.residual + off;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += __bfloat162float(rp[j]);
    }
  }
}
// write the chunk into the staging box: row r_in_tile (0..127), 128-byte rows, 16-byte pieces XOR-swizzled by (row & 7)
__device__ __forceinline__ void tc_stage_chunk(uint8_t* box, int r_in_tile, const float (&v)[32], int c_dtype, int half) {
  uint8_t* rowp = box + r_in_tile * 128;
  const int sw = r_in_tile & 7;
  if (c_dtype == I2T_F32) {          // 32 fp32 = the whole 128-byte row
#pragma unroll
    for (int t = 0; t < 8; ++t)
      *reinterpret_cast<float4*>(rowp + ((t ^ sw) << 4)) = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
  } else {                           // 32 bf16 = half a row (`half` selects which)
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      uint4 pk;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * t], v[8 * t + 1]), h1 = __floats2bfloat162_rn(v[8 * t + 2], v[8 * t + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * t + 4], v[8 * t + 5]), h3 = __floats2bfloat162_rn(v[8 * t + 6], v[8 * t + 7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(rowp + (((half * 4 + t) ^ sw) << 4)) = pk;
    }
  }
}
// 32 bf16 of row r_in_tile into a 64-byte-pitch box, 16-byte pieces XOR-swizzled by ((row >> 1) & 3) (CU_TENSOR_MAP_SWIZZLE_64B)
__device__ __forceinline__ void tc_stage_chunk64(uint8_t* box, int r_in_tile, const float (&v)[32]) {
  uint8_t* rowp = box + r_in_tile * 64;
  const int sw = (r_in_tile >> 1) & 3;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    uint4 pk;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * t], v[8 * t + 1]), h1 = __floats2bfloat162_rn(v[8 * t + 2], v[8 * t + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * t + 4], v[8 * t + 5]), h3 = __floats2bfloat162_rn(v[8 * t + 6], v[8 * t + 7]);
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(rowp + ((t ^ sw) << 4)) = pk;
  }
}
constexpr int TC_STORE_BOX_BYTES = 128 * 128;        // one staging box: 128 rows x 128 bytes
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps
// The tile's bias slice goes to shared memory BEFORE the epilogue waits for the accumulator, so its L2 latency hides behind
// the main loop instead of being paid once per 32-column chunk (measured: 8.7 -> 5.x us for a 64 x 768 x 768 projection).
// Called by all 128 epilogue threads (tid = 0..127); zero beyond N.  The two barriers order it against the readers of the
// previous tile's slice and against this tile's readers.
template <int BN>
__device__ __forceinline__ void tc_stage_bias(const TcEpilogue& epi, float* sbias, int64_t n_first, int tid) {
  epi_bar_sync();
#pragma unroll
  for (int j = tid; j < BN; j += 128) sbias[j] = (n_first + j < epi.N) ? epi.bias[n_first + j] : 0.f;
  epi_bar_sync();
}

// Persistent kernel: one CTA per SM walks the output tiles t = blockIdx.x, + gridDim.x, ... (m fastest, so concurrently
// running CTAs share the weight tile in L2).  The TMA producer and the MMA issuer run ahead across tile boundaries
// through the 6-stage ring; the accumulator is double buffered in TMEM (2 x 128 columns), so the four epilogue warps
// drain tile i while the tensor core already works on tile i+1.
constexpr int TC_EPI2_THREADS = 384;          // warps 0-3: TMA / MMA / TMEM allocator / idle; warps 4-11: two epilogue warpgroups
template <bool A_MN, bool B_MN, int BN>      // BN = 128, or 96 (N = 768 at M = 2048: 128 tiles instead of 96 on the 148 SMs)
__global__ void __launch_bounds__(TC_EPI2_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, int num_k_blocks, int m_tiles, int n_tiles, int splits, int b_is_weight,
               TcEpilogue epi) {
  // splits > 1 (split-K, for problems with few output tiles): work item w = (tile w % num_tiles, K slice w / num_tiles);
  // a slice covers k-blocks [slice * kpb, min(+kpb, num_k_blocks)) and its partial tile is added atomically (fp32 C only)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float sbias[TC_BN];
  constexpr int B_TX = B_MN ? TC_B_BYTES : BN * TC_BK * 2;          // bytes one stage of B brings (MN-major: two 64-wide boxes)
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + TC_STAGES * TC_A_BYTES;
  uint8_t* smemC = smem + TC_STAGES * (TC_A_BYTES + TC_B_BYTES);     // 2 staging boxes for the TMA-store epilogue
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = m_tiles * n_tiles;
  const int num_work = num_tiles * splits;
  const int kpb = (num_k_blocks + splits - 1) / splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 8);      // one arrival per epilogue warp (two warpgroups)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  pdl_launch_dependents();                                       // the next kernel may start staging ITS weights
  if (warp == 0) {
    if (lane == 0) {
      // K-major operand ([rows][K], K contiguous): one box {64 k, 128 rows}.
      // MN-major operand ([K][rows], rows contiguous): two boxes {64 rows, 64 k}, one per 64-wide half of the tile.
      auto load_a = [&](int s, int kb, int m_blk) {
        if (!A_MN) {
          tma_load_2d(smemA + s * TC_A_BYTES, &tmA, kb * TC_BK, m_blk * TC_BM, &full_bar[s]);
        } else {
          tma_load_2d(smemA + s * TC_A_BYTES, &tmA, m_blk * TC_BM, kb * TC_BK, &full_bar[s]);
          tma_load_2d(smemA + s * TC_A_BYTES + TC_A_BYTES / 2, &tmA, m_blk * TC_BM + 64, kb * TC_BK, &full_bar[s]);
        }
      };
      auto load_b = [&](int s, int kb, int n_blk) {
        if (!B_MN) {
          tma_load_2d(smemB + s * TC_B_BYTES, &tmB, kb * TC_BK, n_blk * BN, &full_bar[s]);       // box {64 k, BN rows}
        } else {
          tma_load_2d(smemB + s * TC_B_BYTES, &tmB, n_blk * BN, kb * TC_BK, &full_bar[s]);
          tma_load_2d(smemB + s * TC_B_BYTES + TC_B_BYTES / 2, &tmB, n_blk * BN + 64, kb * TC_BK, &full_bar[s]);
        }
      };
      int it = 0;                                            // running k-block counter across tiles
      bool first = true;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int t = w % num_tiles, kb0 = (w / num_tiles) * kpb, kb1 = min(kb0 + kpb, num_k_blocks);
        const int m_blk = t % m_tiles, n_blk = t / m_tiles;
        int kb = kb0;
        if (first) {
          // Programmatic dependent launch: B (`b_is_weight`: a weight matrix) does not depend on the previous kernel, so the
          // first ring-full of B tiles is requested BEFORE waiting for it; A (activations) follows after the wait.
          first = false;
          const int npre = b_is_weight ? min(TC_STAGES, kb1 - kb0) : 0;
          for (int i = 0; i < npre; ++i) {
            mbar_expect_tx(&full_bar[i], TC_A_BYTES + B_TX);           // all slots are free at kernel start
            load_b(i, kb0 + i, n_blk);
          }
          pdl_wait();
          for (int i = 0; i < npre; ++i) load_a(i, kb0 + i, m_blk);
          it = npre;
          kb = kb0 + npre;
        }
        for (; kb < kb1; ++kb, ++it) {
          const int s = it % TC_STAGES;
          const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], TC_A_BYTES + B_TX);
          load_a(s, kb, m_blk);
          load_b(s, kb, n_blk);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128 (cute::UMMA::InstrDescriptor)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(A_MN ? 1u : 0u) << 15) |
                             ((uint32_t)(B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int it = 0, i = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++i) {
        const int kb0 = (w / num_tiles) * kpb, kb1 = min(kb0 + kpb, num_k_blocks);
        const int b = i & 1;
        mbar_wait(&tmem_empty_bar[b], (((uint32_t)i >> 1) & 1u) ^ 1u);     // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * TC_BN);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % TC_STAGES;
          const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t adesc = A_MN ? umma_desc_sw128_mn(smem_u32(smemA + s * TC_A_BYTES)) : umma_desc_sw128(smem_u32(smemA + s * TC_A_BYTES));
          const uint64_t bdesc = B_MN ? umma_desc_sw128_mn(smem_u32(smemB + s * TC_B_BYTES)) : umma_desc_sw128(smem_u32(smemB + s * TC_B_BYTES));
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // one UMMA consumes 16 values of K.  K-major tile: 16 bf16 = 32 bytes inside the 128-byte swizzle span
            // (+2 in the addr>>4 field).  MN-major tile: 16 K-rows of 128 bytes = 2 swizzle atoms = 2048 bytes (+128).
            umma_bf16(tmem_d, adesc + (uint64_t)((A_MN ? 128 : 2) * k), bdesc + (uint64_t)((B_MN ? 128 : 2) * k), idesc,
                      ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
        }
        umma_commit(&tmem_full_bar[b]);   // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // TWO epilogue warpgroups (warps 4-7 and 8-11; a warp may touch the TMEM lanes 32 (warp % 4) ..): group g drains the store
    // boxes g, g + 2, ... of a tile through its own staging box with its own TMA-store issuer (see the pair kernel)
    const int wq = warp & 3, eg = (warp - 4) >> 2;
    const int r_in_tile = wq * 32 + lane;
    const bool issuer = wq == 0 && lane == 0;
    auto group_sync = [&]() {
      if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };
    if (issuer && epi.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    pdl_wait();               // C / residual belong to the previous kernels (the bias does not, but it is staged per tile)
    int i = 0;
    // bf16 output: two 32-column chunks fill one 128-byte-wide box; the 96-column tile stores 64-byte-wide boxes instead
    // (one per chunk, SWIZZLE_64B tensor map: see gemm_tc_try)
    const bool narrow = BN == 96 && epi.c_dtype != I2T_F32;
    const int chunks_per_box = (epi.c_dtype == I2T_F32 || narrow) ? 1 : 2;
    constexpr int n_chunks = BN / 32;
    int last_c = -1;                                           // this group's last chunk of a tile (-1: none, BN = 32 ...)
    for (int c = 0; c < n_chunks; ++c)
      if (((c / chunks_per_box) & 1) == eg) last_c = c;
    uint8_t* box = smemC + eg * TC_STORE_BOX_BYTES;
    const int etid = (warp - 4) * 32 + lane;                   // 0..255 over both groups
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++i) {
      const int t = w % num_tiles, split = w / num_tiles;
      const int m_blk = t % m_tiles, n_blk = t / m_tiles;
      const int b = i & 1;
      const bool use_bias = epi.bias != nullptr && split == 0;
      if (use_bias) {                                          // the tile's bias slice, staged by all 256 epilogue threads
        asm volatile("bar.sync 3, 256;" ::: "memory");
        for (int j = etid; j < BN; j += 256) sbias[j] = ((int64_t)n_blk * BN + j < epi.N) ? epi.bias[(int64_t)n_blk * BN + j] : 0.f;
        asm volatile("bar.sync 3, 256;" ::: "memory");
      }
      mbar_wait(&tmem_full_bar[b], ((uint32_t)i >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t row0 = (int64_t)m_blk * TC_BM;
      const int64_t row = row0 + r_in_tile;
      const bool row_ok = row < epi.M;
      if (last_c < 0) {                                        // nothing to read: hand the accumulator straight back
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[b])) : "memory");
      }
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        if (((c / chunks_per_box) & 1) != eg) continue;        // the other group's box
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(b * TC_BN + c * 32), r);
        if (c == last_c) {
          // the accumulator is in registers: hand the TMEM buffer back before the (slow) global stores
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[b])) : "memory");
        }
        const int64_t n0 = (int64_t)n_blk * BN + c * 32;
        if (!epi.tma_store) {
          if (!row_ok || n0 >= epi.N) continue;
          if (splits > 1) tc_epilogue_chunk_splitk(epi, r, row, n0, use_bias ? sbias + c * 32 : nullptr);
          else tc_epilogue_chunk(epi, r, row, n0, use_bias ? sbias + c * 32 : nullptr);
          continue;
        }
        float v[32];
        tc_chunk_math(epi, r, v, row, n0, row_ok, use_bias ? sbias + c * 32 : nullptr);
        if (c % chunks_per_box == 0) {
          // first chunk of a box: the group's previous store must have finished READING the staging box
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          group_sync();
        }
        if (narrow) tc_stage_chunk64(box, r_in_tile, v);
        else tc_stage_chunk(box, r_in_tile, v, epi.c_dtype, c % chunks_per_box);
        if ((c + 1) % chunks_per_box == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          group_sync();
          if (issuer) {
            const int col0 = (int)((int64_t)n_blk * BN + (c + 1 - chunks_per_box) * 32);
            if (col0 < epi.N && row0 < epi.M) {
              tma_store_2d(&tmC, box, col0, (int)row0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
      }
    }
    if (issuer && epi.tma_store) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// CTA-PAIR kernel (tcgen05 cta_group::2): two CTAs of one cluster (one TPC) compute a 256 x BN tile together.
// Each CTA stages ITS 128 rows of A and ITS half (BN/2 rows) of B; one thread of the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' shared memory and writes each CTA's 128 accumulator rows
// into that CTA's TMEM.  Per MAC this moves half the bytes through L2 / shared memory of the 128 x 128 single-CTA tile
// (0.016 vs 0.031 B/MAC), which is what large GEMMs need to leave the L2-bandwidth bound.
//   full[s]   : leader's barrier; BOTH CTAs' TMA loads complete_tx on it (the peer addresses it through mapa)
//   empty[s], tmem_full[b] : per CTA; tcgen05.commit ... multicast::cluster arrives on both
//   tmem_empty[b] : leader's; 16 arrivals (8 epilogue warps x 2 CTAs)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_cta0(uint32_t saddr) {    // same offset in the shared memory of cluster CTA 0
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(0u));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int BN>
struct PairCfg {
  static constexpr int B_ROWS = BN / 2;                       // rows of B each CTA stages
  static constexpr int B_BYTES = B_ROWS * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;    // 6 (BN 256) / 8 (BN 128)
  static constexpr int SMEM = STAGES * STAGE_BYTES + 2 * 128 * 128 + 1024;   // ring + 2 epilogue staging boxes + alignment
  static constexpr int TMEM_COLS = 2 * BN;                     // double-buffered accumulator
};

constexpr int TC_PAIR_THREADS = 384;          // warps 0-3: TMA / MMA / TMEM allocator / idle; warps 4-11: two epilogue warpgroups
template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(TC_PAIR_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, int num_k_blocks, int m_tiles, int n_tiles, TcEpilogue epi) {
  using Cfg = PairCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float sbias[BN];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + STAGES * TC_A_BYTES;
  uint8_t* smemC = smem + STAGES * Cfg::STAGE_BYTES;     // 2 staging boxes for the TMA-store epilogue (1024-aligned)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int num_tiles = m_tiles * n_tiles;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 16);     // 8 epilogue warps of each CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();                     // barrier init / TMEM allocation above overlap the previous kernel's tail
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                          // the peer's barriers are initialised before anything remote happens
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                                  // operands, bias, residual and C belong to the previous kernels

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int m_blk = t % m_tiles, n_blk = t / m_tiles;
        const int m0 = m_blk * 256 + (int)rank * 128;            // this CTA's rows of A (and of the output)
        const int n0 = n_blk * BN + (int)rank * Cfg::B_ROWS;     // this CTA's half of the B tile
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
          const uint32_t fb = mapa_cta0(smem_u32(&full_bar[s]));
          uint8_t* a_dst = smemA + s * TC_A_BYTES;
          uint8_t* b_dst = smemB + s * Cfg::B_BYTES;
          if (!A_MN) {
            tma_load_2d_pair(a_dst, &tmA, kb * TC_BK, m0, fb);
          } else {
            tma_load_2d_pair(a_dst, &tmA, m0, kb * TC_BK, fb);
            tma_load_2d_pair(a_dst + TC_A_BYTES / 2, &tmA, m0 + 64, kb * TC_BK, fb);
          }
          if (!B_MN) {
            tma_load_2d_pair(b_dst, &tmB, kb * TC_BK, n0, fb);
          } else {
#pragma unroll
            for (int hb = 0; hb < Cfg::B_ROWS / 64; ++hb)
              tma_load_2d_pair(b_dst + hb * (64 * TC_BK * 2), &tmB, n0 + hb * 64, kb * TC_BK, fb);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, N = BN, M = 256 (the pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(A_MN ? 1u : 0u) << 15) |
                             ((uint32_t)(B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int it = 0, i = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++i) {
        const int b = i & 1;
        mbar_wait(&tmem_empty_bar[b], (((uint32_t)i >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * BN);
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(smemA + s * TC_A_BYTES), b_addr = smem_u32(smemB + s * Cfg::B_BYTES);
          const uint64_t adesc = A_MN ? umma_desc_sw128_mn(a_addr) : umma_desc_sw128(a_addr);
          const uint64_t bdesc = B_MN ? umma_desc_sw128_mn(b_addr) : umma_desc_sw128(b_addr);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16_pair(tmem_d, adesc + (uint64_t)((A_MN ? 128 : 2) * k), bdesc + (uint64_t)((B_MN ? 128 : 2) * k), idesc,
                           (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[s]);     // both CTAs' producers may refill the slot
        }
        umma_commit_pair(&tmem_full_bar[b]);   // both CTAs' epilogues may read their half of the accumulator
      }
    }
  } else if (warp >= 4) {
    // TWO epilogue warpgroups (warps 4-7 and 8-11; a warp may touch the TMEM lanes 32 (warp % 4) ..).  Group g drains the store
    // boxes g, g + 2, ... of a tile through its OWN staging box and issues its own TMA stores, so a 256-column accumulator leaves in
    // half the time: at K = 768 the pair kernel was paced by its epilogue (tmem load -> math -> staging -> store per 32 columns:
    // 71 % tensor pipe with one group).
    const int wq = warp & 3, eg = (warp - 4) >> 2;
    const int r_in_tile = wq * 32 + lane;
    const bool issuer = wq == 0 && lane == 0;
    auto group_sync = [&]() {
      if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };
    if (issuer && epi.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    int i = 0;
    const int chunks_per_box = epi.c_dtype == I2T_F32 ? 1 : 2;
    const int n_chunks = BN / 32;
    int last_c = 0;                                            // this group's last chunk of a tile
    for (int c = 0; c < n_chunks; ++c)
      if (((c / chunks_per_box) & 1) == eg) last_c = c;
    uint8_t* box = smemC + eg * TC_STORE_BOX_BYTES;
    const int etid = (warp - 4) * 32 + lane;                   // 0..255 over both groups
    for (int t = pair; t < num_tiles; t += num_pairs, ++i) {
      const int m_blk = t % m_tiles, n_blk = t / m_tiles;
      const int b = i & 1;
      const bool use_bias = epi.bias != nullptr;
      if (use_bias) {                                          // the tile's bias slice, staged by all 256 epilogue threads
        asm volatile("bar.sync 3, 256;" ::: "memory");
#pragma unroll
        for (int j = etid; j < BN; j += 256) sbias[j] = ((int64_t)n_blk * BN + j < epi.N) ? epi.bias[(int64_t)n_blk * BN + j] : 0.f;
        asm volatile("bar.sync 3, 256;" ::: "memory");
      }
      mbar_wait(&tmem_full_bar[b], ((uint32_t)i >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t row0 = (int64_t)m_blk * 256 + (int64_t)rank * 128;
      const int64_t row = row0 + r_in_tile;
      const bool row_ok = row < epi.M;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        if (((c / chunks_per_box) & 1) != eg) continue;        // the other group's box
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(b * BN + c * 32), r);
        if (c == last_c) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            const uint32_t eb = mapa_cta0(smem_u32(&tmem_empty_bar[b]));
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(eb) : "memory");
          }
        }
        const int64_t n0 = (int64_t)n_blk * BN + c * 32;
        if (!epi.tma_store) {
          if (!row_ok || n0 >= epi.N || epi.debug == 2) continue;
          if (epi.debug == 1 && r[0] != 0x7fc12345u) continue;
          tc_epilogue_chunk(epi, r, row, n0, use_bias ? sbias + c * 32 : nullptr);
          continue;
        }
        // staged path (uniform control flow for the 128 threads of the group: out-of-range rows / columns are clipped by TMA)
        float v[32];
        tc_chunk_math(epi, r, v, row, n0, row_ok, use_bias ? sbias + c * 32 : nullptr);
        if (c % chunks_per_box == 0) {
          // first chunk of a box: the group's previous store must have finished READING the staging box
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          group_sync();
        }
        tc_stage_chunk(box, r_in_tile, v, epi.c_dtype, c % chunks_per_box);
        if ((c + 1) % chunks_per_box == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          group_sync();
          if (issuer) {
            const int col0 = (int)((int64_t)n_blk * BN + (c + 1 - chunks_per_box) * 32);
            if (col0 < epi.N && row0 < epi.M) {
              tma_store_2d(&tmC, box, col0, (int)row0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
      }
    }
    if (issuer && epi.tma_store) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                          // nobody may still be reading our shared memory / signalling our barriers
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// ---- host side: tensor-map construction through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct MapKey {
  const void* p;
  int64_t rows, cols, ld;
  int box_rows;     // bit 16 set: fp32 elements (32-column boxes) instead of bf16 (64-column boxes)
  bool operator==(const MapKey& o) const {
    return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.p) ^ (std::hash<int64_t>()(k.rows) * 1315423911u) ^
           (std::hash<int64_t>()(k.cols) * 2654435761u) ^ (std::hash<int64_t>()(k.ld) << 7) ^ (size_t)k.box_rows;
  }
};

// 2-D bf16 [rows][cols] (cols contiguous, pitch ld elements), box {64 cols, box_rows}, 128-byte swizzle, zero fill out
// of bounds.  K-major operand: cols = K, box_rows = 128.  MN-major operand: cols = MN, rows = K, box_rows = 64.
static int make_map_dt(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool f32, CUtensorMap* out,
                       bool narrow = false);
int tc_make_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
  return make_map_dt(ptr, rows, cols, ld, box_rows, false, out);
}
// same for either element type: the box is always 128 bytes wide (64 bf16 / 32 fp32) and 128-byte swizzled
// (narrow: bf16 box of 32 columns = 64 bytes, 64-byte swizzle -- the store boxes of the 96-column tile)
static int make_map_dt(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool f32, CUtensorMap* out, bool narrow) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, rows, cols, ld, box_rows | (f32 ? 0x10000 : 0) | (narrow ? 0x20000 : 0)};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return I2T_OK;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(I2T_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {(f32 || narrow) ? 32u : 64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, narrow ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(I2T_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
  }
  return I2T_OK;
}

template <bool A_MN, bool B_MN, int BN>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, int kblocks, const TcEpilogue& epi, dim3 grid,
                     int splits, int b_is_weight, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "cudaFuncSetAttribute(gemm_tc_kernel): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int m_tiles = (int)grid.y, n_tiles = (int)grid.x;
  const int ctas = (int)std::min<int64_t>((int64_t)m_tiles * n_tiles * splits, (int64_t)num_sms());
  // programmatic stream serialization: the kernel calls griddepcontrol.wait before it touches A, C or the residual
  cudaError_t e = launch_pdl(gemm_tc_kernel<A_MN, B_MN, BN>, dim3(ctas), dim3(TC_EPI2_THREADS), TC_SMEM, st, ma, mb, mc, kblocks, m_tiles,
                             n_tiles, splits, b_is_weight, epi);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(I2T_ERR_CUDA, "gemm_tc launch failed: %s", cudaGetErrorString(e));
  }
  return 1;
}

static std::atomic<int> g_split_k{1};      // 1: split-K for few-tile problems (fp32 atomics); 0: never (A/B testing, determinism)

// K slices per output tile: minimise waves x (k-blocks per slice + fixed cost of a slice: pipeline fill, 64 KB of atomics),
// keep at least two k-blocks per slice, split only for a clear (>= 15 %) gain
static int pick_splits(int64_t tiles, int kb, int sms) {
  const int C0 = 8;
  auto cost = [&](int s) { return (double)ceil_div(tiles * s, sms) * (double)(ceil_div(kb, s) + C0); };
  int best = 1;
  double best_cost = cost(1);
  for (int s = 2; s <= 32 && s * 2 <= kb; ++s) {
    const int kpb = (int)ceil_div(kb, s);
    if ((int64_t)(s - 1) * kpb >= kb) continue;          // the last slice would be empty
    const double c = cost(s);
    if (c < best_cost) {
      best_cost = c;
      best = s;
    }
  }
  return best_cost <= 0.85 * cost(1) ? best : 1;
}

static std::atomic<int> g_tile96{getenv("I2T_GEMM_TILE96") ? atoi(getenv("I2T_GEMM_TILE96")) : 1};   // A/B switch for the 96-column tile
static std::atomic<int> g_tma_store{1};    // 1: CTA-pair epilogue through shared memory + TMA stores; 0: per-thread row stores
static std::atomic<int> g_pair_mode{1};   // 1: CTA-pair kernel where the problem has enough 256-row tiles; 0: never

template <bool A_MN, bool B_MN, int BN>
static int launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, int kblocks, const TcEpilogue& epi,
                       int m_tiles, int n_tiles, cudaStream_t st) {
  using Cfg = PairCfg<BN>;
  static bool attr_set = false;
  auto kern = gemm_tc_pair_kernel<A_MN, B_MN, BN>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return fail(I2T_ERR_CUDA, "cudaFuncSetAttribute(gemm_tc_pair_kernel): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int pairs = (int)std::min<int64_t>((int64_t)m_tiles * n_tiles, (int64_t)(num_sms() / 2));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(TC_PAIR_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // the kernel waits (griddepcontrol) after its setup
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl.load() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, kblocks, m_tiles, n_tiles, epi);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(I2T_ERR_CUDA, "gemm_tc_pair launch failed: %s", cudaGetErrorString(e));
  }
  return 1;
}

template <int BN>
static int launch_pair_layout(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, int kb, const TcEpilogue& epi,
                              int mt, int nt, int a_kmajor, int b_kmajor, cudaStream_t st) {
  if (a_kmajor && b_kmajor) return launch_pair<false, false, BN>(ma, mb, mc, kb, epi, mt, nt, st);
  if (a_kmajor && !b_kmajor) return launch_pair<false, true, BN>(ma, mb, mc, kb, epi, mt, nt, st);
  if (!a_kmajor && b_kmajor) return launch_pair<true, false, BN>(ma, mb, mc, kb, epi, mt, nt, st);
  return launch_pair<true, true, BN>(ma, mb, mc, kb, epi, mt, nt, st);
}

int gemm_tc_try(const void* A, const void* B, const float* bias, const void* residual, void* C, int64_t M, int64_t N,
                int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_kmajor, int b_kmajor, int act, int accumulate,
                int res_dtype, int c_dtype, int flags, cudaStream_t st) {
  // leading dimensions are row pitches in elements: 16-byte multiples for TMA
  if (lda % 8 != 0 || ldb % 8 != 0 || !aligned16(A) || !aligned16(B)) return 0;
  if (M > 128 * 65535LL || N > 128LL * 0x7fffffff) return 0;
  TcEpilogue epi;
  epi.bias = bias;
  epi.residual = residual;
  epi.C = C;
  epi.ldc = ldc;
  epi.M = (int)M;
  epi.N = (int)N;
  epi.act = act;
  epi.accumulate = accumulate;
  epi.res_dtype = res_dtype;
  epi.c_dtype = c_dtype;
  epi.debug = getenv("I2T_GEMM_DEBUG") ? atoi(getenv("I2T_GEMM_DEBUG")) : 0;
  epi.tma_store = 0;
  const int kb = (int)ceil_div(K, TC_BK);
  CUtensorMap ma, mb;
  int rc = a_kmajor ? tc_make_map(A, M, K, lda, TC_BM, &ma) : tc_make_map(A, K, M, lda, 64, &ma);
  if (rc != I2T_OK) return rc;
  // CTA-pair tiles (256 x 256, else 256 x 128) when they fill at least ~3/4 of the 74 SM pairs; otherwise 128 x 128 tiles
  if (g_pair_mode.load() == 1) {
    static const int64_t want_env = getenv("I2T_GEMM_PAIR_MIN") ? atoi(getenv("I2T_GEMM_PAIR_MIN")) : 0;    // tuning experiments
    const int64_t want = want_env > 0 ? want_env : (int64_t)(num_sms() / 2) * 3 / 4;
    const int64_t mt = ceil_div(M, 256);
    int bn = 0;
    if (mt * ceil_div(N, 256) >= want) bn = 256;
    else if (mt * ceil_div(N, 128) >= want) bn = 128;
    if (bn != 0) {
      rc = b_kmajor ? tc_make_map(B, N, K, ldb, bn / 2, &mb) : tc_make_map(B, K, N, ldb, 64, &mb);
      if (rc != I2T_OK) return rc;
      const int nt = (int)ceil_div(N, bn);
      // the output leaves through TMA stores when the epilogue does not read C back and the rows are 16-byte pitched
      const int esz = c_dtype == I2T_F32 ? 4 : 2;
      CUtensorMap mc = ma;
      epi.tma_store = 0;
      if (!accumulate && aligned16(C) && (ldc * esz) % 16 == 0 && epi.debug == 0 && g_tma_store.load() == 1) {
        rc = make_map_dt(C, M, N, ldc, 128, c_dtype == I2T_F32, &mc);
        if (rc != I2T_OK) return rc;
        epi.tma_store = 1;
      }
      return bn == 256 ? launch_pair_layout<256>(ma, mb, mc, kb, epi, (int)mt, nt, a_kmajor, b_kmajor, st)
                       : launch_pair_layout<128>(ma, mb, mc, kb, epi, (int)mt, nt, a_kmajor, b_kmajor, st);
    }
  }
  {
    const int esz = c_dtype == I2T_F32 ? 4 : 2;
    dim3 grid((unsigned)ceil_div(N, TC_BN), (unsigned)ceil_div(M, TC_BM));
    // split-K: few output tiles and a long K (decode projections over a batch, weight gradients) leave most SMs idle and
    // each busy SM bound by its own TMA rate; K slices on the idle SMs add their partial tiles with fp32 atomics.
    // Eligible: fp32 C, no activation, and C either accumulates, holds the residual in place, or can be zeroed first.
    int splits = 1;
    if (g_split_k.load() == 1 && c_dtype == I2T_F32 && act == I2T_ACT_NONE && (residual == nullptr || residual == C) &&
        epi.debug == 0)
      splits = pick_splits((int64_t)grid.x * grid.y, kb, num_sms());
    // 96-column tiles: when the 128 x 128 tiling leaves SMs idle and 128 x 96 tiles still fit one wave (N = 768 at M = 2048:
    // 96 -> 128 tiles on 148 SMs), every CTA's serial work shrinks by a quarter.  Not with split-K (its own way to fill the SMs).
    int bn = TC_BN;
    if (g_tile96.load() == 1 && splits == 1 && epi.debug == 0) {
      const int64_t t128 = (int64_t)grid.x * grid.y, t96 = (int64_t)ceil_div(N, 96) * grid.y;
      if (t128 < num_sms() && t96 <= num_sms() && t96 > t128 && N >= 96) bn = 96;
    }
    rc = b_kmajor ? tc_make_map(B, N, K, ldb, bn, &mb) : tc_make_map(B, K, N, ldb, 64, &mb);
    if (rc != I2T_OK) return rc;
    CUtensorMap mc = ma;
    epi.tma_store = 0;
    if (!accumulate && aligned16(C) && (ldc * esz) % 16 == 0 && epi.debug == 0 && g_tma_store.load() == 1) {
      rc = make_map_dt(C, M, N, ldc, 128, c_dtype == I2T_F32, &mc, bn == 96 && c_dtype != I2T_F32);
      if (rc != I2T_OK) return rc;
      epi.tma_store = 1;
    }
    if (splits > 1) {
      epi.tma_store = 0;
      if (!accumulate && residual == nullptr)
        I2T_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
    }
    // B is fetched ahead of the dependency wait only when the caller promises (I2T_GEMM_B_STABLE) that no preceding work in
    // the stream writes it: the decode step's weights.  (Training cannot promise that: a bf16 weight shadow may have been
    // refreshed by the kernel right before this one.)
    const int b_is_weight = (flags & I2T_GEMM_B_STABLE) ? 1 : 0;
    if (bn == 96) {
      grid.x = (unsigned)ceil_div(N, 96);
      if (a_kmajor && b_kmajor) return launch_tc<false, false, 96>(ma, mb, mc, kb, epi, grid, splits, b_is_weight, st);
      if (a_kmajor && !b_kmajor) return launch_tc<false, true, 96>(ma, mb, mc, kb, epi, grid, splits, b_is_weight, st);
      if (!a_kmajor && b_kmajor) return launch_tc<true, false, 96>(ma, mb, mc, kb, epi, grid, splits, b_is_weight, st);
      return launch_tc<true, true, 96>(ma, mb, mc, kb, epi, grid, splits, b_is_weight, st);
    }
    if (a_kmajor && b_kmajor) return launch_tc<false, false, 128>(ma, mb, mc, kb, epi, grid, splits, b_is_weight, st);
    if (a_kmajor && !b_kmajor) return launch_tc<false, true, 128>(ma, mb, mc, kb, epi, grid, splits, b_is_weight, st);
    if (!a_kmajor && b_kmajor) return launch_tc<true, false, 128>(ma, mb, mc, kb, epi, grid, splits, b_is_weight, st);
    return launch_tc<true, true, 128>(ma, mb, mc, kb, epi, grid, splits, b_is_weight, st);
  }
}

}  // namespace i2t

extern "C" void i2t_set_gemm_cta_pair(int enabled) { i2t::g_pair_mode.store(enabled ? 1 : 0); }
extern "C" void i2t_set_gemm_tma_store(int enabled) { i2t::g_tma_store.store(enabled ? 1 : 0); }
extern "C" void i2t_set_gemm_split_k(int enabled) { i2t::g_split_k.store(enabled ? 1 : 0); }
