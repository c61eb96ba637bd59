// Decode megakernel v2 (bf16 weights): ONE cooperative launch runs a whole generate() loop -- every decode step of
// every layer, the LM head, and the greedy pick / sampler -- for up to 8 sequences.
//
// Why v2.  The per-stage trace of v1 (profiles/r01_trace_mega_v1_bf16.txt) showed a decode step is ~80 dependent stages
// of 24-32 KB of weights per SM, i.e. pure latency: 2.4 us activation staging + 3.5 us FMA/shuffle compute + 1.9 us grid
// barrier per stage while the bytes need < 1 us.  v2 attacks each term:
//   * the batch (<= 8 sequences) is the N=8 of mma.sync.m16n8k16 (bf16 x bf16 -> fp32): a tile of 16 weight rows x 16 k
//     per instruction; the 256 FMAs + 63 shuffles per thread of v1 become 6 MMAs.  A dot product is invariant under a
//     permutation of k, so each thread owns 16 CONTIGUOUS bytes of two weight rows per 32-k block and uses them as the
//     (a0,a1 | a4,a5) halves of two MMAs; the activation fragment is read with the same permutation.
//   * weights travel HBM -> shared memory with 16-byte cp.async in that per-thread fragment order (a thread reads back
//     only what it copied: no CTA barrier, no mbarrier, no producer warp).  The ring holds 4 tiles (96 KB) per SM; the
//     first tiles of a stage are issued BEFORE the grid barrier that precedes it (weights never depend on the previous
//     stage), LayerNorm parameters and the cached K/V rows of an attention stage likewise.
//   * one CTA of 8 warps per SM; the warps split the K of a tile (3 x 32 k each for K <= 768, 12 x 32 k each for the
//     MLP down projection) and reduce partial tiles through shared memory in a fixed order: deterministic results.
//   * every stage's metadata (schedule, op table) is copied to shared memory once, and the hot loop is kept small:
//     a 160 KB loop body (first cut) missed the 32 KB instruction cache on every stage.
//   * attention: one key per thread for q.k (K row in registers), V rows staged in shared memory for P.V.
//   * greedy decode fuses the no-repeat-n-gram ban and the arg-max into the LM-head epilogue (packed atomicMax keys --
//     an order-independent integer max), so the logits round trip and the sampler stage disappear; the next step's
//     embedding reads the key directly.
//   * the token loop lives inside the kernel: no per-token launch, graph replay or drain.
// Replaces, for KV-cached decode, reference models/vision_encoder_decoder.py:144-180 (generate loop),
// models/decoder.py:214-256 and models/layers.py:447-486,565-614 (one-token forward).
//
// Tables: same layout as decode_mega.cu (lin[op][20], att[a][8], sched[s][4]); lin[op][17] bit 0 = "arg-max epilogue".
#include "common.cuh"
#include "sampler.cuh"

namespace i2t {

constexpr int M2_WARPS = 8;
constexpr int M2_THREADS = M2_WARPS * 32;
constexpr int M2_B = 8;                       // batch rows = MMA N
constexpr int M2_ROWS = 16;                   // weight rows per tile = MMA M
constexpr int M2_TILE = M2_ROWS * M2_B;       // outputs per tile
constexpr int M2_BLK = 32;                    // k elements per block (one 16-byte copy per thread and row)
constexpr int M2_KA = 768;                    // mode A (K <= 768): 3 blocks per warp, one ring slot per tile
constexpr int M2_KB = 3072;                   // mode B (K <= 3072): 12 blocks per warp, the whole ring per tile
constexpr int M2_NBA = M2_KA / M2_BLK / M2_WARPS;     // 3
constexpr int M2_RING = 4;                    // ring slots
constexpr int M2_SLOT_VECS = 2 * M2_NBA;      // 16-byte vectors per thread and slot
constexpr int M2_RING_BYTES = M2_RING * M2_SLOT_VECS * M2_THREADS * 16;    // 96 KB
constexpr int M2_LIN_FIELDS = 20;
constexpr int M2_MAX_BANNED = 256;            // per sequence
constexpr int M2_MAX_KEYS = M2_THREADS;       // attention: one key per thread
constexpr int M2_TRACE = 8;                   // stamps per stage

struct M2Args {
  const int64_t* lin;
  const int64_t* att;
  const int32_t* sched_sample;
  const int32_t* sched_prefill;
  int n_sched_sample, n_sched_prefill, n_ops, n_att;
  int n_prefill, n_sample;
  int B, C, H, V, n_prompt, max_k;
  int64_t* ids;
  int64_t ids_ld;
  int32_t* pos;
  float* q;        // (B, C) query scratch
  float* y;        // (B, C) attention output scratch
  float* logits;   // (B, V)   (written only when top_k != 1)
  unsigned int* bar;
  int32_t* error_flag;
  unsigned long long* keys;   // [3][8] packed arg-max keys (zeroed by the host before the launch)
  float temperature;
  int top_k;
  const int32_t* ngrams;
  int n_ngrams;
  const uint64_t* seed_ptr;
  long long* trace;           // optional [n_sched_sample][8] clock64 stamps of CTA 0 for the LAST sampled step
};

struct __align__(16) M2Fixed {
  float red[2][M2_WARPS][2 * M2_TILE];           // per-warp partial tiles (a pair of them), double buffered (16 KB)
  int banned[M2_B][M2_MAX_BANNED];
  int nbanned[M2_B];
  int hist[M2_B][M2_MAX_KEYS + 8];               // token history of every sequence (n-gram ban)
  float ln_g[M2_KA];                             // LayerNorm gamma / beta of the coming stage (cp.async before the barrier)
  float ln_b[M2_KA];
  float att_q[64];
  float att_p[M2_MAX_KEYS];
  float att_red[M2_WARPS];
  float att_o[M2_WARPS][64];
  unsigned long long best[M2_WARPS][M2_B];
};

// Dynamic shared memory: [ring | M2Fixed | xs | lin | att | sched_prefill | sched_sample].  Every access goes through
// the extern symbol (not through pointers kept in a struct) so that the compiler emits LDS/STS, not generic LD/ST.
extern __shared__ __align__(128) uint8_t m2_smem[];
__host__ __device__ inline size_t m2_align16(size_t x) { return (x + 15) & ~(size_t)15; }
constexpr uint32_t M2_FIXED_OFF = M2_RING_BYTES;
constexpr uint32_t M2_XS_OFF = M2_FIXED_OFF + ((sizeof(M2Fixed) + 15) / 16) * 16;
struct M2Sm {
  int xpitch;                         // bf16 elements between the staged activation rows
  uint32_t lin_off, att_off, sched_off[2];     // byte offsets of the tables ([0] prefill, [1] sample)
};
__device__ __forceinline__ uint4* m2_ring() { return reinterpret_cast<uint4*>(m2_smem); }   // aliased by V rows in attention
__device__ __forceinline__ M2Fixed* m2_f() { return reinterpret_cast<M2Fixed*>(m2_smem + M2_FIXED_OFF); }
__device__ __forceinline__ __nv_bfloat16* m2_xs() { return reinterpret_cast<__nv_bfloat16*>(m2_smem + M2_XS_OFF); }
__device__ __forceinline__ const int64_t* m2_lin(const M2Sm& S, int op) {
  return reinterpret_cast<const int64_t*>(m2_smem + S.lin_off) + (size_t)op * M2_LIN_FIELDS;
}
__device__ __forceinline__ const int64_t* m2_att(const M2Sm& S, int ai) {
  return reinterpret_cast<const int64_t*>(m2_smem + S.att_off) + (size_t)ai * 8;
}
__device__ __forceinline__ const int32_t* m2_sched(const M2Sm& S, int which) {
  return reinterpret_cast<const int32_t*>(m2_smem + S.sched_off[which]);
}

__device__ __forceinline__ void m2_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void m2_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void m2_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void m2_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// grid-wide barrier (monotonic counter zeroed by the host per launch), split in two so that the next stage's HBM
// traffic can be issued between "arrive" and "wait".
__device__ __forceinline__ void m2_grid_arrive(unsigned int* bar, unsigned int& epoch, int tid) {
  __syncthreads();
  if (tid == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
  }
}
__device__ __forceinline__ void m2_grid_wait(unsigned int* bar, unsigned int epoch, int32_t* error_flag, int tid) {
  if (tid == 0) {
    unsigned int seen = 0;
    uint32_t spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if (++spins > (1u << 26)) {   // never hang the GPU on a protocol bug
        atomicExch(error_flag, 2);
        break;
      }
    } while ((int)(seen - epoch) < 0);
  }
  __syncthreads();
}

// ---- work decomposition of a linear op ----
struct M2Op {
  const __nv_bfloat16* W;
  int N, K, total, nb;           // total = tiles of 16 rows = work units; nb = 32-k blocks per warp (3 or 12)
  bool mode_b;
};
__device__ __forceinline__ M2Op m2_op(const int64_t* d) {
  M2Op o;
  o.W = reinterpret_cast<const __nv_bfloat16*>(d[0]);
  o.N = (int)d[7];
  o.K = (int)d[8];
  o.total = (o.N + M2_ROWS - 1) / M2_ROWS;
  o.mode_b = o.K > M2_KA;
  o.nb = o.mode_b ? M2_RING * M2_NBA : M2_NBA;
  return o;
}
__device__ __forceinline__ int m2_first_unit(int op) {
  return (int)((blockIdx.x + (unsigned)op * 37u) % gridDim.x);     // rotate the CTA <-> unit map from stage to stage
}
// number of ring slots the stage's first tiles occupy when issued ahead of time
__device__ __forceinline__ int m2_lead_tiles(const M2Op& o, int u0) {
  if (u0 >= o.total) return 0;
  if (o.mode_b) return 1;
  return min(M2_RING, (o.total - u0 + (int)gridDim.x - 1) / (int)gridDim.x);
}

// issue the copies of one tile: this warp's blocks (strided over the 8 warps) x 2 rows, 16 bytes per thread each,
// into ring vectors [vec0 .. vec0 + 2 * nb).  One out-of-line copy: the hot loop must stay inside the instruction cache.
__device__ __noinline__ void m2_issue_tile(const __nv_bfloat16* W, int N, int K, int nb, int tile, int vec0) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = K / M2_BLK;
  const int g = lane >> 2, qd = lane & 3;
  const int r0 = min(tile * M2_ROWS + g, N - 1), r1 = min(tile * M2_ROWS + g + 8, N - 1);
  const __nv_bfloat16* p0 = W + (size_t)r0 * K + qd * 8;
  const __nv_bfloat16* p1 = W + (size_t)r1 * K + qd * 8;
  uint4* dst = m2_ring() + (size_t)vec0 * M2_THREADS + tid;
#pragma unroll 1
  for (int i0 = 0; i0 < nb; i0 += M2_NBA) {
#pragma unroll
    for (int ii = 0; ii < M2_NBA; ++ii) {
      const int i = i0 + ii, blk = warp + M2_WARPS * i;
      if (blk < nblk) {
        m2_cp_async16(dst + (2 * i) * M2_THREADS, p0 + blk * M2_BLK);
        m2_cp_async16(dst + (2 * i + 1) * M2_THREADS, p1 + blk * M2_BLK);
      }
    }
  }
}
__device__ __forceinline__ void m2_issue(const M2Op& o, const M2Sm& S, int tile, int vec0, int warp, int lane, int tid) {
  m2_issue_tile(o.W, o.N, o.K, o.nb, tile, vec0);
}
// everything a linear stage can fetch before the grid barrier in front of it: LayerNorm parameters + its first tiles
__device__ __forceinline__ void m2_lead_issue(const M2Op& o, const int64_t* d, const M2Sm& S, int u0, int warp, int lane, int tid) {
  const float* ln_g = reinterpret_cast<const float*>(d[2]);
  const float* ln_b = reinterpret_cast<const float*>(d[3]);
  if (ln_g != nullptr) {
    const int half = M2_THREADS / 2;
    if (tid < half) {
      for (int k = tid * 4; k < o.K; k += half * 4) m2_cp_async16(m2_f()->ln_g + k, ln_g + k);
    } else if (ln_b != nullptr) {
      for (int k = (tid - half) * 4; k < o.K; k += half * 4) m2_cp_async16(m2_f()->ln_b + k, ln_b + k);
    }
  }
  const int lead = m2_lead_tiles(o, u0);
  for (int j = 0; j < lead; ++j) m2_issue(o, S, u0 + j * (int)gridDim.x, j * M2_SLOT_VECS, warp, lane, tid);
  m2_commit();
}

__device__ __forceinline__ int m2_token(const M2Args& a, int b, int pos, int keybuf) {
  if (keybuf >= 0) {
    const unsigned long long key = __ldcg(a.keys + keybuf * M2_B + b);
    return (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
  }
  return (int)__ldcg(a.ids + (int64_t)b * a.ids_ld + pos);
}

__device__ __forceinline__ uint32_t m2_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- stage the activations: warp b owns batch row b.  K <= 768: all loads in flight at once, LayerNorm on registers.
//      K > 768 (no LayerNorm): chunks of 768 ----
constexpr int M2_NVA = M2_KA / 128;      // float4 per lane and chunk
__device__ __forceinline__ void m2_stage_x(const M2Args& a, const int64_t* d, const M2Sm& S, int pos, int keybuf, int warp,
                                           int lane, long long* trace) {
  const int K = (int)d[8];
  const float* in = reinterpret_cast<const float*>(d[4]);
  const int in_mode = (int)d[13];
  const bool has_ln = d[2] != 0, has_beta = d[3] != 0;
  __nv_bfloat16* xr = m2_xs() + warp * S.xpitch;
  const float* src = in + (int64_t)warp * K;
  if (in_mode == 1 && warp < a.B) src = in + (int64_t)m2_token(a, warp, pos, keybuf) * K;     // token embedding row
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += M2_KA) {
    float4 v[M2_NVA];
#pragma unroll
    for (int i = 0; i < M2_NVA; ++i) {
      const int k = k0 + (lane + 32 * i) * 4;
      v[i] = (k < K && warp < a.B) ? __ldcg(reinterpret_cast<const float4*>(src + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (in_mode == 1 && warp < a.B) {   // x = wte[tok] + wpe[n_prompt + pos]; CTA 0 publishes it as the residual stream
      const float* wpe = reinterpret_cast<const float*>(d[14]) + (int64_t)(a.n_prompt + pos) * K;
      float* xout = reinterpret_cast<float*>(d[6]);
#pragma unroll
      for (int i = 0; i < M2_NVA; ++i) {
        const int k = (lane + 32 * i) * 4;
        if (k < K) {
          const float4 pe = __ldcg(reinterpret_cast<const float4*>(wpe + k));
          v[i].x += pe.x; v[i].y += pe.y; v[i].z += pe.z; v[i].w += pe.w;
          if (blockIdx.x == 0) store4(xout + (int64_t)warp * K + k, v[i]);
        }
      }
    }
    if (has_ln) {                 // (K <= 768: one chunk)  parameters arrive by cp.async, issued before the barrier
      m2_wait_group<0>();
      __syncthreads();
      if (trace != nullptr) trace[4] = clock64();
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < M2_NVA; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      const float mu = warp_sum(s) / (float)K;
      float qq = 0.f;
#pragma unroll
      for (int i = 0; i < M2_NVA; ++i)
        if ((lane + 32 * i) * 4 < K) {
          const float c0 = v[i].x - mu, c1 = v[i].y - mu, c2 = v[i].z - mu, c3 = v[i].w - mu;
          qq += (c0 * c0 + c1 * c1) + (c2 * c2 + c3 * c3);
        }
      const float rs = 1.0f / sqrtf(warp_sum(qq) / (float)K + 1e-5f);
#pragma unroll
      for (int i = 0; i < M2_NVA; ++i)
        if ((lane + 32 * i) * 4 < K) {
          const float4 gg = *reinterpret_cast<const float4*>(m2_f()->ln_g + (lane + 32 * i) * 4);
          v[i].x = (v[i].x - mu) * rs * gg.x; v[i].y = (v[i].y - mu) * rs * gg.y;
          v[i].z = (v[i].z - mu) * rs * gg.z; v[i].w = (v[i].w - mu) * rs * gg.w;
          if (has_beta) {
            const float4 bb = *reinterpret_cast<const float4*>(m2_f()->ln_b + (lane + 32 * i) * 4);
            v[i].x += bb.x; v[i].y += bb.y; v[i].z += bb.z; v[i].w += bb.w;
          }
        }
    }
#pragma unroll
    for (int i = 0; i < M2_NVA; ++i) {
      const int k = k0 + (lane + 32 * i) * 4;
      if (k < K) *reinterpret_cast<uint2*>(xr + k) = make_uint2(m2_pack(v[i].x, v[i].y), m2_pack(v[i].z, v[i].w));
    }
  }
}

// ---- banned next tokens of every sequence (transformers NoRepeatNGramLogitsProcessor), into shared memory ----
__device__ __noinline__ void m2_banned(const int64_t* ids, int64_t ids_ld, int B, const int32_t* ngrams, int n_ngrams,
                                       int32_t* error_flag, int cur_len, int tid) {
  M2Fixed* f = m2_f();
  if (tid < M2_B) f->nbanned[tid] = 0;
  for (int w = tid; w < cur_len * B; w += M2_THREADS) {
    const int b = w / cur_len, i = w - b * cur_len;
    f->hist[b][i] = (int)__ldcg(ids + (int64_t)b * ids_ld + i);
  }
  __syncthreads();
  for (int g = 0; g < n_ngrams; ++g) {
    const int n = ngrams[g];
    if (n <= 0 || cur_len + 1 < n) continue;
    const int tail = cur_len + 1 - n;
    const int span = cur_len - n + 1;                 // candidate start positions 0 .. cur_len - n
    for (int w = tid; w < span * B; w += M2_THREADS) {
      const int b = w / span, i = w - b * span;
      const int* idr = f->hist[b];
      bool same = true;
      for (int j = 0; j < n - 1; ++j) same = same && (idr[i + j] == idr[tail + j]);
      if (same) {
        const int slot = atomicAdd(&f->nbanned[b], 1);
        if (slot < M2_MAX_BANNED) f->banned[b][slot] = idr[i + n - 1];
        else atomicExch(error_flag, 5);
      }
    }
  }
  __syncthreads();
}

// bf16 path: tanh.approx (|err| ~ 5e-4 relative) is far below the bf16 rounding of the value it feeds
__device__ __forceinline__ float m2_act(float x, int act) {
  if (act == I2T_ACT_GELU_TANH) {
    float t;
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    return 0.5f * x * (1.0f + t);
  }
  if (act == I2T_ACT_GELU_ERF) return gelu_erf_f(x);
  return x;
}

// ---- one linear stage.  `lead` = its first tiles (and LayerNorm parameters) were issued before the barrier. ----
struct M2Epi {
  const float* bias;
  float* out;
  const float* residual;
  __nv_bfloat16* kcache;
  __nv_bfloat16* vcache;
  int64_t ldo, cache_bs;
  int act, mode, in_mode, N;
  bool argmax;
};
// epilogue of one output: v = reduced dot product of weight row n with batch row eb
__device__ __forceinline__ void m2_epilogue(const M2Args& a, const M2Epi& e, float v, int n0, int n, int eb, int pos, float e_bias,
                                            float e_res, float& best_v, int& best_n) {
  v = m2_act(v + e_bias, e.act);
  if (e.mode == 0) {
    if (e.residual != nullptr && e.in_mode == 0) v += e_res;
    if (e.argmax) {
      bool ban = false;
      const int nb = min(m2_f()->nbanned[eb], M2_MAX_BANNED);
      for (int i = 0; i < nb; ++i) ban = ban || (m2_f()->banned[eb][i] == n);
      if (!ban && (v > best_v || best_n == 0x7fffffff)) { best_v = v; best_n = n; }
    } else {
      e.out[(int64_t)eb * e.ldo + n] = v;
    }
  } else {   // packed q | k | v (a tile never straddles two of them): q to the scratch, k / v appended at `pos`
    const int seg = n0 / a.C, nl = n - seg * a.C;
    if (seg == 0) {
      e.out[(int64_t)eb * e.ldo + nl] = v;
    } else {
      __nv_bfloat16* base = seg == 1 ? e.kcache : e.vcache;
      base[(int64_t)eb * e.cache_bs + (int64_t)pos * a.C + nl] = __float2bfloat16_rn(v);
    }
  }
}

__device__ __forceinline__ void m2_linear(const M2Args& a, int op, const M2Sm& S, bool lead, int pos, int keybuf, int keyout,
                                          int tid, long long* trace) {
  const int lane = tid & 31, warp = tid >> 5;
  const int64_t* d = m2_lin(S, op);
  const M2Op o = m2_op(d);
  M2Epi e;
  e.bias = reinterpret_cast<const float*>(d[1]);
  e.out = reinterpret_cast<float*>(d[5]);
  e.residual = reinterpret_cast<const float*>(d[6]);
  e.kcache = reinterpret_cast<__nv_bfloat16*>(d[11]);
  e.vcache = reinterpret_cast<__nv_bfloat16*>(d[12]);
  e.ldo = d[15]; e.cache_bs = d[16];
  e.act = (int)d[9]; e.mode = (int)d[10]; e.in_mode = (int)d[13]; e.N = o.N;
  e.argmax = ((int)d[17] & 1) != 0 && a.top_k == 1;
  const int G = gridDim.x;
  const int u0 = m2_first_unit(op);
  const int g = lane >> 2, qd = lane & 3;
  float best_v = -INFINITY;
  int best_n = 0x7fffffff;
  if (u0 < o.total) {
    if (!lead) m2_lead_issue(o, d, S, u0, warp, lane, tid);
    m2_stage_x(a, d, S, pos, keybuf, warp, lane, trace);
    m2_wait_group<0>();                        // this thread's copies of the lead tiles have landed
    __syncthreads();                           // xs visible
    if (trace != nullptr) trace[1] = clock64();
    const int nblk = o.K / M2_BLK;
    const __nv_bfloat16* xp = m2_xs() + g * S.xpitch + qd * 8;
    // Tiles are consumed in PAIRS when K <= 768 (two ring slots, one CTA barrier, 256 epilogue threads; a CTA with a single
    // tile runs the same code with the second half switched off).  K > 768: one tile fills the whole ring.
    const int half = tid >> 7, et = tid & 127;        // epilogue: threads 0..127 -> first tile, 128..255 -> second tile
    const int er = et >> 3, eb = et & 7;
    const int ustep = o.mode_b ? G : 2 * G;
    int pidx = 0;
#pragma unroll 1
    for (int u = u0; u < o.total; u += ustep, ++pidx) {
      const bool two = !o.mode_b && u + G < o.total;
      int vec0 = 0;
      if (o.mode_b) {
        if (pidx > 0) { m2_issue(o, S, u, 0, warp, lane, tid); m2_commit(); m2_wait_group<0>(); }
      } else {
        if (pidx >= 2) m2_wait_group<1>();             // pair p was committed two iterations ago
        vec0 = (pidx & 1) * 2 * M2_SLOT_VECS;
      }
      const int n0 = (u + half * G) * M2_ROWS;
      const bool e_on = (half == 0 || two) && n0 + er < o.N && eb < a.B;
      float e_bias = 0.f, e_res = 0.f;                 // epilogue operands do not depend on the MMAs: fetch them now
      if (e_on) {
        if (e.bias != nullptr) e_bias = __ldcg(e.bias + n0 + er);
        if (e.mode == 0 && e.residual != nullptr && e.in_mode == 0) e_res = __ldcg(e.residual + (int64_t)eb * e.ldo + n0 + er);
      }
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      const uint4* wv = m2_ring() + (size_t)vec0 * M2_THREADS + tid;
#pragma unroll 1
      for (int i0 = 0; i0 < o.nb; i0 += M2_NBA) {
#pragma unroll
        for (int ii = 0; ii < M2_NBA; ++ii) {
          const int i = i0 + ii, blk = warp + M2_WARPS * i;
          if (blk < nblk) {
            const uint4 xf = *reinterpret_cast<const uint4*>(xp + blk * M2_BLK);
            const uint4 w0 = wv[(2 * i) * M2_THREADS], w1 = wv[(2 * i + 1) * M2_THREADS];
            m2_mma(acc[0], w0.x, w1.x, w0.y, w1.y, xf.x, xf.y);
            m2_mma(acc[0], w0.z, w1.z, w0.w, w1.w, xf.z, xf.w);
            if (two) {
              const uint4 v0 = wv[(M2_SLOT_VECS + 2 * i) * M2_THREADS], v1 = wv[(M2_SLOT_VECS + 2 * i + 1) * M2_THREADS];
              m2_mma(acc[1], v0.x, v1.x, v0.y, v1.y, xf.x, xf.y);
              m2_mma(acc[1], v0.z, v1.z, v0.w, v1.w, xf.z, xf.w);
            }
          }
        }
      }
      if (trace != nullptr && pidx == 0) trace[7] = clock64() + (long long)(acc[0][0] == 123.f);
      if (!o.mode_b) {   // refill both slots with the pair two iterations ahead (thread-private data: no barrier needed)
        if (u + 4 * G < o.total) m2_issue(o, S, u + 4 * G, vec0, warp, lane, tid);
        if (u + 5 * G < o.total) m2_issue(o, S, u + 5 * G, vec0 + M2_SLOT_VECS, warp, lane, tid);
        m2_commit();
      }
      // acc: D[g][2qd], D[g][2qd+1], D[g+8][2qd], D[g+8][2qd+1]  ->  flat (row * 8 + batch), second tile at +128
      float* red = &m2_f()->red[pidx & 1][warp][0];
      *reinterpret_cast<float2*>(red + lane * 2) = make_float2(acc[0][0], acc[0][1]);
      *reinterpret_cast<float2*>(red + 64 + lane * 2) = make_float2(acc[0][2], acc[0][3]);
      if (two) {
        *reinterpret_cast<float2*>(red + 128 + lane * 2) = make_float2(acc[1][0], acc[1][1]);
        *reinterpret_cast<float2*>(red + 192 + lane * 2) = make_float2(acc[1][2], acc[1][3]);
      }
      __syncthreads();
      if (e_on) {
        const float* rp = &m2_f()->red[pidx & 1][0][tid];
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < M2_WARPS; ++i) v += rp[i * 2 * M2_TILE];
        m2_epilogue(a, e, v, n0, n0 + er, eb, pos, e_bias, e_res, best_v, best_n);
      }
    }
  }
  if (e.argmax) {
    // CTA-level arg-max per sequence: rows of a warp (lane bits 3,4), then the 8 warps, then one atomicMax
    unsigned long long key = 0ull;
    if (best_n != 0x7fffffff)
      key = ((unsigned long long)float_key(best_v) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)best_n);
    const unsigned long long k1 = __shfl_xor_sync(0xffffffffu, key, 8);
    key = k1 > key ? k1 : key;
    const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, 16);
    key = k2 > key ? k2 : key;
    if (lane < M2_B) m2_f()->best[warp][lane] = key;
    __syncthreads();
    if (tid < a.B) {
      unsigned long long k = m2_f()->best[0][tid];
#pragma unroll
      for (int i = 1; i < M2_WARPS; ++i) k = m2_f()->best[i][tid] > k ? m2_f()->best[i][tid] : k;
      if (k != 0ull) atomicMax(a.keys + keyout * M2_B + tid, k);
    }
  }
}

// ---- single-query attention for one (batch, head) per CTA: one key per thread for q.k (K row in registers),
//      V rows staged in shared memory (aliasing the idle weight ring) for P.V ----
// m2_att_fetch(fresh = false) runs BEFORE the grid barrier that follows the QKV stage: every cached position except
// the one that stage appends (j == pos in the self cache) is already final; fresh = true fetches that row afterwards.
template <int HS>
__device__ __forceinline__ void m2_att_fetch(const M2Args& a, const M2Sm& S, int ai, uint4 (&kr)[HS / 8], int pos, int tid,
                                             bool fresh) {
  constexpr int NV = HS / 8;                 // 16-byte vectors per row
  const int unit = blockIdx.x;
  if (unit >= a.B * a.H) return;
  const int64_t* d = m2_att(S, ai);
  const int len = d[4] == 0 ? pos + 1 : (int)d[5];
  const bool is_fresh = d[4] == 0 && tid == pos;
  if (tid >= len || is_fresh != fresh) return;
  const int b = unit / a.H, h = unit - b * a.H;
  const __nv_bfloat16* kp = reinterpret_cast<const __nv_bfloat16*>(d[0]) + b * d[2] + (int64_t)tid * d[3] + h * HS;
  const __nv_bfloat16* vp = reinterpret_cast<const __nv_bfloat16*>(d[1]) + b * d[2] + (int64_t)tid * d[3] + h * HS;
  uint4* vdst = m2_ring() + (size_t)tid * NV;
#pragma unroll
  for (int i = 0; i < NV; ++i) kr[i] = __ldcg(reinterpret_cast<const uint4*>(kp) + i);
#pragma unroll
  for (int i = 0; i < NV; ++i) m2_cp_async16(vdst + i, reinterpret_cast<const uint4*>(vp) + i);
}

template <int HS>
__device__ __forceinline__ void m2_attention(const M2Args& a, int ai, const M2Sm& S, uint4 (&kr)[HS / 8], bool lead, int pos,
                                             int tid) {
  constexpr int NV = HS / 8, DPL = HS / 32;          // dims per lane in P.V
  const int lane = tid & 31, warp = tid >> 5;
  const int unit = blockIdx.x;               // B * H <= gridDim.x is checked by the host
  if (unit >= a.B * a.H) return;
  const int64_t* d = m2_att(S, ai);
  const int len = d[4] == 0 ? pos + 1 : (int)d[5];
  const float scale = 1.0f / sqrtf((float)HS);
  const int b = unit / a.H, h = unit - b * a.H;
  const bool on = tid < len;
  if (!lead) m2_att_fetch<HS>(a, S, ai, kr, pos, tid, false);
  m2_att_fetch<HS>(a, S, ai, kr, pos, tid, true);
  m2_commit();
  if (tid < HS) {
    const float v = __ldcg(a.q + (int64_t)b * a.C + h * HS + tid);
    m2_f()->att_q[tid] = __bfloat162float(__float2bfloat16_rn(v)) * scale;      // autocast: SDPA sees a bf16 query
  }
  m2_wait_group<0>();
  __syncthreads();                           // q and every V row visible
  float s = -INFINITY;
  if (on) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 q0 = *reinterpret_cast<const float4*>(&m2_f()->att_q[i * 8]);
      const float4 q1 = *reinterpret_cast<const float4*>(&m2_f()->att_q[i * 8 + 4]);
      acc = fmaf(__uint_as_float(kr[i].x << 16), q0.x, acc); acc = fmaf(__uint_as_float(kr[i].x & 0xffff0000u), q0.y, acc);
      acc = fmaf(__uint_as_float(kr[i].y << 16), q0.z, acc); acc = fmaf(__uint_as_float(kr[i].y & 0xffff0000u), q0.w, acc);
      acc = fmaf(__uint_as_float(kr[i].z << 16), q1.x, acc); acc = fmaf(__uint_as_float(kr[i].z & 0xffff0000u), q1.y, acc);
      acc = fmaf(__uint_as_float(kr[i].w << 16), q1.z, acc); acc = fmaf(__uint_as_float(kr[i].w & 0xffff0000u), q1.w, acc);
    }
    s = acc;
  }
  float mx = warp_max(s);
  if (lane == 0) m2_f()->att_red[warp] = mx;
  __syncthreads();
  mx = m2_f()->att_red[0];
#pragma unroll
  for (int i = 1; i < M2_WARPS; ++i) mx = fmaxf(mx, m2_f()->att_red[i]);
  m2_f()->att_p[tid] = on ? expf(s - mx) : 0.f;
  __syncthreads();
  // P.V: warp w takes keys w, w+8, ...; lane l owns dims [l * DPL, (l + 1) * DPL); the softmax sum rides along
  const __nv_bfloat16* vs = reinterpret_cast<const __nv_bfloat16*>(m2_ring());
  float o0 = 0.f, o1 = 0.f, psum = 0.f;
  for (int j = warp; j < len; j += M2_WARPS) {
    const float p = m2_f()->att_p[j];
    psum += p;
    if (DPL == 2) {
      const uint32_t r = *reinterpret_cast<const uint32_t*>(vs + (size_t)j * HS + lane * 2);
      o0 = fmaf(p, __uint_as_float(r << 16), o0);
      o1 = fmaf(p, __uint_as_float(r & 0xffff0000u), o1);
    } else {
      o0 = fmaf(p, __bfloat162float(vs[(size_t)j * HS + lane]), o0);
    }
  }
  if (DPL == 2) *reinterpret_cast<float2*>(&m2_f()->att_o[warp][lane * 2]) = make_float2(o0, o1);
  else m2_f()->att_o[warp][lane] = o0;
  if (lane == 0) m2_f()->att_red[warp] = psum;
  __syncthreads();
  if (tid < HS) {
    float tot = 0.f, ov = 0.f;
#pragma unroll
    for (int i = 0; i < M2_WARPS; ++i) { tot += m2_f()->att_red[i]; ov += m2_f()->att_o[i][tid]; }
    a.y[(int64_t)b * a.C + h * HS + tid] = tot > 0.f ? ov / tot : 0.f;
  }
}

// general sampler on B CTAs, vocabulary row processed in place (global / L2).  Not inlined: its register needs (double
// precision prefix sums) and code size must not shape the hot loop.
__device__ __noinline__ void m2_sample_stage(float* logits, int V, int B, int64_t* ids, int64_t ids_ld, float temperature,
                                             int top_k, const int32_t* ngrams, int n_ngrams, const uint64_t* seed_ptr, int pos,
                                             int tid) {
  if ((int)blockIdx.x >= B) return;
  const int b = blockIdx.x;
  float* row = logits + (int64_t)b * V;
  SampleScratch& samp = *reinterpret_cast<SampleScratch*>(m2_f());
  static_assert(sizeof(SampleScratch) <= sizeof(M2Fixed), "sampler scratch must fit the stage buffers");
  const int choice = sample_row_smem(row, samp, row, V, ids + (int64_t)b * ids_ld, pos + 1, temperature, top_k, ngrams,
                                     n_ngrams, *seed_ptr, b, nullptr, tid, M2_THREADS);
  if (tid == 0) ids[(int64_t)b * ids_ld + pos + 1] = (int64_t)choice;
}

template <int HS>
__global__ void __launch_bounds__(M2_THREADS, 1) decode_mega2_kernel(M2Args a) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- lay out dynamic shared memory and copy the tables (the schedule is read ~80 times per token) ----
  M2Sm S;
  {
    S.xpitch = a.max_k + 32;                 // bytes = 2 * max_k + 64 = 64 (mod 128): conflict-free LDS.128 of the B fragments
    uint32_t off = M2_XS_OFF + (uint32_t)m2_align16((size_t)M2_B * S.xpitch * 2);
    S.lin_off = off; off += (uint32_t)m2_align16((size_t)a.n_ops * M2_LIN_FIELDS * 8);
    S.att_off = off; off += (uint32_t)m2_align16((size_t)a.n_att * 8 * 8);
    S.sched_off[0] = off; off += (uint32_t)m2_align16((size_t)a.n_sched_prefill * 4 * 4);
    S.sched_off[1] = off;
    int64_t* lin = reinterpret_cast<int64_t*>(m2_smem + S.lin_off);
    int64_t* att = reinterpret_cast<int64_t*>(m2_smem + S.att_off);
    int32_t* sp = reinterpret_cast<int32_t*>(m2_smem + S.sched_off[0]);
    int32_t* ss = reinterpret_cast<int32_t*>(m2_smem + S.sched_off[1]);
    for (int i = tid; i < a.n_ops * M2_LIN_FIELDS; i += M2_THREADS) lin[i] = a.lin[i];
    for (int i = tid; i < a.n_att * 8; i += M2_THREADS) att[i] = a.att[i];
    for (int i = tid; i < a.n_sched_prefill * 4; i += M2_THREADS) sp[i] = a.sched_prefill[i];
    for (int i = tid; i < a.n_sched_sample * 4; i += M2_THREADS) ss[i] = a.sched_sample[i];
  }
  const int pos0 = *a.pos;
  __syncthreads();
  unsigned int epoch = 0;
  uint4 kr[HS / 8];         // K row of this thread's key in the coming attention stage
  int lead_s = -1;          // schedule index (within the step it belongs to) of the stage whose lead traffic is in flight
  const int n_steps = a.n_prefill + a.n_sample;
  const bool greedy = a.top_k == 1;
#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
    const bool sampling = step >= a.n_prefill;
    const int32_t* sched = m2_sched(S, sampling ? 1 : 0);
    const int n_sched = sampling ? a.n_sched_sample : a.n_sched_prefill;
    const int pos = pos0 + step;
    const int sstep = step - a.n_prefill;                       // index among the sampled steps
    // token source of this step's embedding: the previous sampled step's arg-max key (greedy), else the ids buffer
    const int keybuf = (greedy && sampling && sstep > 0) ? (sstep - 1) % 3 : -1;
    const int keyout = sstep >= 0 ? sstep % 3 : 0;
    const bool trace_step = a.trace != nullptr && step == n_steps - 1 && sampling && blockIdx.x == 0 && tid == 0;
    if (greedy && sampling && blockIdx.x == 0 && tid < a.B) {
      // publish the previous pick for the host and the n-gram ban; recycle the key slot two steps ahead
      if (sstep > 0) a.ids[(int64_t)tid * a.ids_ld + pos] = (int64_t)m2_token(a, tid, pos, keybuf);
      a.keys[((sstep + 1) % 3) * M2_B + tid] = 0ull;
    }
#pragma unroll 1
    for (int s = 0; s < n_sched; ++s) {
      const int kind = sched[s * 4], idx = sched[s * 4 + 1];
      long long* trace = trace_step ? a.trace + s * M2_TRACE : nullptr;
      if (trace_step) trace[0] = clock64();
      bool need_sync = true;
      const bool lead = lead_s == s && kind <= 1;
      if (kind <= 1) lead_s = -1;
      if (kind == 0) {
        m2_linear(a, idx, S, lead, pos, keybuf, keyout, tid, trace);
      } else if (kind == 1) {
        m2_attention<HS>(a, idx, S, kr, lead, pos, tid);
      } else if (kind == 2) {
        if (!greedy) m2_sample_stage(a.logits, a.V, a.B, a.ids, a.ids_ld, a.temperature, a.top_k, a.ngrams, a.n_ngrams, a.seed_ptr, pos, tid);
        else need_sync = false;               // the arg-max keys were completed by the LM-head stage's barrier
      } else {
        need_sync = false;                    // ADVANCE: the position is a kernel-local counter here
      }
      if (trace_step) trace[2] = clock64();
      if (need_sync) {
        m2_grid_arrive(a.bar, epoch, tid);
        if (trace_step) trace[5] = clock64();
        // what the next stage streams from HBM does not depend on this stage: put it in flight while waiting
        int s2 = s + 1, step2 = step, n2 = n_sched;
        const int32_t* sc2 = sched;
        bool found = false;
        for (int guard = 0; guard < 4 && !found; ++guard) {
          if (s2 >= n2) {
            if (++step2 >= n_steps) break;
            const bool samp2 = step2 >= a.n_prefill;
            sc2 = m2_sched(S, samp2 ? 1 : 0);
            n2 = samp2 ? a.n_sched_sample : a.n_sched_prefill;
            s2 = 0;
          }
          const int k2 = sc2[s2 * 4];
          if (k2 == 0 || k2 == 1) found = true; else ++s2;
        }
        if (found && (s2 == s + 1 || (greedy && kind == 0))) {     // (never across a sampler stage: it aliases the buffers)
          const int k2 = sc2[s2 * 4], i2 = sc2[s2 * 4 + 1];
          if (k2 == 0) {
            const int64_t* d2 = m2_lin(S, i2);
            const M2Op o2 = m2_op(d2);
            const int u2 = m2_first_unit(i2);
            if (u2 < o2.total) m2_lead_issue(o2, d2, S, u2, warp, lane, tid);
            if (((int)d2[17] & 1) != 0 && greedy) m2_banned(a.ids, a.ids_ld, a.B, a.ngrams, a.n_ngrams, a.error_flag, pos + 1, tid);   // n-gram ban list, off the critical path
            lead_s = s2;
          } else if (step2 == step) {
            m2_att_fetch<HS>(a, S, i2, kr, pos, tid, false);
            lead_s = s2;
          }
        }
        if (trace_step) trace[6] = clock64();
        m2_grid_wait(a.bar, epoch, a.error_flag, tid);
      }
      if (trace_step) trace[3] = clock64();
    }
  }
  // final bookkeeping: last pick -> ids, position counter
  const int pos_end = pos0 + n_steps;
  if (blockIdx.x == 0) {
    if (greedy && a.n_sample > 0 && tid < a.B)
      a.ids[(int64_t)tid * a.ids_ld + pos_end] = (int64_t)m2_token(a, tid, pos_end, (a.n_sample - 1) % 3);
    if (tid == 0) *a.pos = pos_end;
  }
}

}  // namespace i2t

using namespace i2t;

extern "C" int i2t_decode_mega2_max_keys(void) { return M2_MAX_KEYS; }

// Runs n_prefill steps of `sched_prefill` (no LM head) followed by n_sample steps of `sched_sample`, starting at the
// device-side position *pos.  bf16 weights only.  keys: device uint64[24]; bar: device uint32; both zeroed here.
extern "C" int i2t_decode_mega2(const int64_t* lin, const int64_t* att, const int32_t* sched_sample, int64_t n_sched_sample,
                                const int32_t* sched_prefill, int64_t n_sched_prefill, int64_t n_ops, int64_t n_att,
                                int64_t n_prefill, int64_t n_sample, int64_t B, int64_t C, int64_t H, int64_t V,
                                int64_t n_prompt, int64_t* ids, int64_t ids_ld, int32_t* pos, float* q, float* y,
                                float* logits, uint32_t* bar, int32_t* error_flag, uint64_t* keys, float temperature,
                                int64_t top_k, const int32_t* ngrams, int64_t n_ngrams, const uint64_t* seed_ptr,
                                int64_t max_k, int64_t max_len, int64_t* trace, void* stream) {
  I2T_REQUIRE(lin && att && sched_sample && sched_prefill && ids && pos && q && y && logits && bar && error_flag && keys &&
                  seed_ptr, "decode_mega2: null pointer");
  I2T_REQUIRE(B > 0 && B <= M2_B, "decode_mega2: batch %lld outside 1..8", (long long)B);
  I2T_REQUIRE(H > 0 && C % H == 0 && (C / H == 64 || C / H == 32), "decode_mega2: head_dim must be 32 or 64");
  I2T_REQUIRE(C % 64 == 0 && max_k % 64 == 0, "decode_mega2: widths must be multiples of 64");
  I2T_REQUIRE(C <= M2_KA && max_k <= M2_KB, "decode_mega2: n_embd=%lld / widest input %lld above %d / %d", (long long)C,
              (long long)max_k, M2_KA, M2_KB);
  I2T_REQUIRE(max_len <= M2_MAX_KEYS, "decode_mega2: %lld cached positions exceed the one-key-per-thread limit %d",
              (long long)max_len, M2_MAX_KEYS);
  I2T_REQUIRE(temperature > 0.f && n_prefill >= 0 && n_sample >= 0 && n_prefill + n_sample > 0, "decode_mega2: bad step counts / temperature");
  cudaStream_t st = (cudaStream_t)stream;
  M2Args a;
  a.lin = lin; a.att = att; a.sched_sample = sched_sample; a.sched_prefill = sched_prefill;
  a.n_sched_sample = (int)n_sched_sample; a.n_sched_prefill = (int)n_sched_prefill; a.n_ops = (int)n_ops; a.n_att = (int)n_att;
  a.n_prefill = (int)n_prefill; a.n_sample = (int)n_sample;
  a.B = (int)B; a.C = (int)C; a.H = (int)H; a.V = (int)V; a.n_prompt = (int)n_prompt; a.max_k = (int)max_k;
  a.ids = ids; a.ids_ld = ids_ld; a.pos = pos; a.q = q; a.y = y; a.logits = logits; a.bar = bar; a.error_flag = error_flag;
  a.keys = reinterpret_cast<unsigned long long*>(keys);
  a.temperature = temperature; a.top_k = (int)(top_k > 0 ? top_k : 0);
  a.ngrams = ngrams; a.n_ngrams = (int)n_ngrams; a.seed_ptr = seed_ptr;
  a.trace = reinterpret_cast<long long*>(trace);
  const void* kern = (C / H == 64) ? (const void*)decode_mega2_kernel<64> : (const void*)decode_mega2_kernel<32>;
  const size_t smem = (size_t)M2_RING_BYTES + m2_align16(sizeof(M2Fixed)) + m2_align16((size_t)M2_B * (max_k + 32) * 2) +
                      m2_align16((size_t)n_ops * M2_LIN_FIELDS * 8) + m2_align16((size_t)n_att * 64) +
                      m2_align16((size_t)n_sched_prefill * 16) + m2_align16((size_t)n_sched_sample * 16);
  I2T_REQUIRE(smem <= 227 * 1024, "decode_mega2: tables need %zu bytes of shared memory", smem);
  I2T_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  I2T_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, M2_THREADS, smem));
  I2T_REQUIRE(per_sm >= 1, "decode_mega2: kernel does not fit on an SM");
  const int grid = num_sms();
  I2T_REQUIRE(B * H <= grid, "decode_mega2: %lld (batch, head) pairs exceed the %d CTAs", (long long)(B * H), grid);
  I2T_CUDA(cudaMemsetAsync(bar, 0, sizeof(uint32_t), st));
  I2T_CUDA(cudaMemsetAsync(keys, 0, sizeof(uint64_t) * 3 * M2_B, st));
  void* params[] = {&a};
  I2T_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(M2_THREADS), params, smem, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return I2T_OK;
}
