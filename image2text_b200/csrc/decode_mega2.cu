// Decode megakernel v2 (bf16 weights): ONE cooperative launch runs a whole generate() loop -- every decode step of
// every layer, the LM head, and the greedy pick / sampler -- for up to 8 sequences.
//
// Why v2.  The per-stage trace of v1 (profiles/r01_trace_mega_bf16.txt) showed a decode step is ~80 dependent stages of
// 24-32 KB of weights per SM, i.e. pure latency: 2.4 us activation staging + 3.5 us FMA/shuffle compute + 1.9 us grid
// barrier per stage while the bytes need < 1 us.  v2 attacks each term:
//   * weights go HBM -> REGISTERS as the A fragments of mma.sync.m16n8k16 (bf16 x bf16 -> fp32), issued BEFORE the grid
//     barrier that precedes the stage (weights never depend on the previous stage), so they are in flight while the
//     CTA waits and stages its activations; no shared-memory ring, no mbarriers, no producer warp.  A dot product is
//     invariant under a permutation of k, so each thread loads 16 contiguous bytes of a weight row and uses them as the
//     (a0,a1 | a4,a5) halves of two MMAs; the activation fragment is read from shared memory with the same permutation.
//   * the batch (<= 8 sequences) is the MMA's N=8: one m16n8k16 per 16 weight rows x 16 k, fp32 accumulate -- the
//     256 FMAs + 63 shuffles per thread of v1 become 6 MMAs.
//   * work unit = 16 weight rows x (<= 768) k for one CTA of 8 warps (warps split k, partial tiles reduced through
//     shared memory); K > 768 is split over CTAs and accumulated with fp32 atomics into the residual stream; 2 CTAs/SM.
//   * attention: one key per thread (K and V rows of a head loaded up front in one round trip), softmax across the CTA,
//     P.V reduced with a transposing shuffle tree.
//   * greedy decode fuses the no-repeat-n-gram ban and the arg-max into the LM-head epilogue (packed atomicMax keys), so
//     the 1.6 MB logits round trip and the sampler stage disappear; the next step's embedding reads the key directly.
//   * the token loop lives inside the kernel: no per-token launch, graph replay or drain.
// Replaces, for KV-cached decode, reference models/vision_encoder_decoder.py:144-180 (generate loop),
// models/decoder.py:214-256 and models/layers.py:447-486,565-614 (one-token forward).
//
// Tables: same layout as decode_mega.cu (lin[op][20], att[a][8], sched[s][4]); lin[op][17] bit 0 = "arg-max epilogue".
#include <type_traits>

#include "common.cuh"
#include "sampler.cuh"

namespace i2t {

constexpr int M2_WARPS = 8;
constexpr int M2_THREADS = M2_WARPS * 32;
constexpr int M2_B = 8;                       // batch rows = MMA N
constexpr int M2_ROWS = 16;                   // weight rows per unit = MMA M
constexpr int M2_BLK = 32;                    // k elements per block (one 16-byte load per thread and row)
constexpr int M2_CHUNK_BLKS = 24;             // blocks per K chunk (768 elements): 3 blocks per warp
constexpr int M2_WB = M2_CHUNK_BLKS / M2_WARPS;
constexpr int M2_XPITCH = M2_CHUNK_BLKS * M2_BLK + 32;   // bf16 elements; pitch bytes = 1600 = 64 (mod 128): conflict-free LDS.128
constexpr int M2_LIN_FIELDS = 20;
constexpr int M2_MAX_BANNED = 256;            // per sequence
constexpr int M2_MAX_KEYS = 256;              // attention: one key per thread

struct M2Args {
  const int64_t* lin;
  const int64_t* att;
  const int32_t* sched_sample;
  const int32_t* sched_prefill;
  int n_sched_sample, n_sched_prefill;
  int n_prefill, n_sample;
  int B, C, H, V, n_prompt;
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
  int32_t* ticket;
  long long* trace;           // optional [n_sched_sample][4] clock64 stamps of CTA 0 for the LAST sampled step
};

struct __align__(16) M2Smem {
  __nv_bfloat16 xs[M2_B * M2_XPITCH];            // staged activations of the current K chunk (bf16, autocast semantics)
  float red[2][M2_WARPS][M2_ROWS * M2_B];        // per-warp partial tiles, double buffered
  int banned[M2_B][M2_MAX_BANNED];
  int nbanned[M2_B];
  float att_q[64];
  float att_red[M2_WARPS];
  float att_o[M2_WARPS][64];
  unsigned long long best[4][M2_B];
};

__device__ __forceinline__ uint4 m2_ldg(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void m2_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// grid-wide barrier (monotonic counter zeroed by the host per launch)
__device__ __forceinline__ void m2_grid_sync(unsigned int* bar, unsigned int& epoch, int32_t* error_flag, int tid) {
  __syncthreads();
  if (tid == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
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
  int N, K, rows16, nchunks, total;
};
__device__ __forceinline__ M2Op m2_op(const int64_t* d) {
  M2Op o;
  o.W = reinterpret_cast<const __nv_bfloat16*>(d[0]);
  o.N = (int)d[7];
  o.K = (int)d[8];
  o.rows16 = (o.N + M2_ROWS - 1) / M2_ROWS;
  const int nblk = o.K / M2_BLK;
  o.nchunks = (nblk + M2_CHUNK_BLKS - 1) / M2_CHUNK_BLKS;
  o.total = o.rows16 * o.nchunks;
  return o;
}
__device__ __forceinline__ int m2_first_unit(int op) {
  const int G = gridDim.x;
  return (int)(((int64_t)blockIdx.x + (int64_t)op * 37) % G);     // rotate the CTA <-> unit map from stage to stage
}

struct M2Pref {
  uint4 w[M2_WB][2];
  int op, unit;      // what the registers hold (op < 0: nothing)
};

// issue the weight loads of unit `u` (chunk-major numbering: u = chunk * rows16 + row_group)
__device__ __forceinline__ void m2_issue(const M2Op& o, int u, uint4 (&w)[M2_WB][2], int warp, int lane) {
  const int chunk = u / o.rows16, rg = u - chunk * o.rows16;
  const int blk0 = chunk * M2_CHUNK_BLKS;
  const int nblk = min(M2_CHUNK_BLKS, o.K / M2_BLK - blk0);
  const int g = lane >> 2, qd = lane & 3;
  const int r0 = min(rg * M2_ROWS + g, o.N - 1), r1 = min(rg * M2_ROWS + g + 8, o.N - 1);
  const __nv_bfloat16* p0 = o.W + (size_t)r0 * o.K + (size_t)blk0 * M2_BLK + qd * 8;
  const __nv_bfloat16* p1 = o.W + (size_t)r1 * o.K + (size_t)blk0 * M2_BLK + qd * 8;
#pragma unroll
  for (int i = 0; i < M2_WB; ++i) {
    const int blk = warp + M2_WARPS * i;
    if (blk < nblk) {
      w[i][0] = m2_ldg(p0 + blk * M2_BLK);
      w[i][1] = m2_ldg(p1 + blk * M2_BLK);
    } else {
      w[i][0] = make_uint4(0u, 0u, 0u, 0u);
      w[i][1] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
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

// ---- stage one K chunk of the activations: warp b owns batch row b (all loads in flight at once, LayerNorm on
//      registers, bf16 into shared memory) ----
__device__ void m2_stage_x(const M2Args& a, const int64_t* d, M2Smem& S, int chunk, int pos, int keybuf, int warp, int lane) {
  const int K = (int)d[8];
  const float* ln_g = reinterpret_cast<const float*>(d[2]);
  const float* ln_b = reinterpret_cast<const float*>(d[3]);
  const float* in = reinterpret_cast<const float*>(d[4]);
  const int in_mode = (int)d[13];
  const int k0 = chunk * M2_CHUNK_BLKS * M2_BLK;
  const int kc = min(M2_CHUNK_BLKS * M2_BLK, K - k0);
  constexpr int NV = M2_CHUNK_BLKS * M2_BLK / 128;      // float4 per lane
  __nv_bfloat16* xr = S.xs + warp * M2_XPITCH;
  if (warp >= a.B) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int k = (lane + 32 * i) * 4;
      if (k < kc) *reinterpret_cast<uint2*>(xr + k) = make_uint2(0u, 0u);
    }
    return;
  }
  float4 v[NV];
  const float* src = in + (int64_t)warp * K + k0;
  if (in_mode == 1) src = in + (int64_t)m2_token(a, warp, pos, keybuf) * K;     // token embedding row (K == C, one chunk)
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int k = (lane + 32 * i) * 4;
    if (k < kc) {
      v[i] = in_mode == 1 ? load4(src + k) : __ldcg(reinterpret_cast<const float4*>(src + k));
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (in_mode == 1) {   // x = wte[tok] + wpe[n_prompt + pos]; CTA 0 publishes it as the residual stream
    const float* wpe = reinterpret_cast<const float*>(d[14]) + (int64_t)(a.n_prompt + pos) * K;
    float* xout = reinterpret_cast<float*>(d[6]);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int k = (lane + 32 * i) * 4;
      if (k < kc) {
        const float4 pe = load4(wpe + k);
        v[i].x += pe.x; v[i].y += pe.y; v[i].z += pe.z; v[i].w += pe.w;
        if (blockIdx.x == 0) store4(xout + (int64_t)warp * K + k, v[i]);
      }
    }
  }
  if (ln_g != nullptr) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mu = warp_sum(s) / (float)K;
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if ((lane + 32 * i) * 4 < kc) {
        const float c0 = v[i].x - mu, c1 = v[i].y - mu, c2 = v[i].z - mu, c3 = v[i].w - mu;
        qq += (c0 * c0 + c1 * c1) + (c2 * c2 + c3 * c3);
      }
    const float rs = 1.0f / sqrtf(warp_sum(qq) / (float)K + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if ((lane + 32 * i) * 4 < kc) {
        // gamma / beta: the same 3 KB for every warp and stage of a layer -> L1 / L2 hits, not worth 48 registers
        const float4 gg = load4(ln_g + (lane + 32 * i) * 4);
        v[i].x = (v[i].x - mu) * rs * gg.x; v[i].y = (v[i].y - mu) * rs * gg.y;
        v[i].z = (v[i].z - mu) * rs * gg.z; v[i].w = (v[i].w - mu) * rs * gg.w;
        if (ln_b != nullptr) {
          const float4 bb = load4(ln_b + (lane + 32 * i) * 4);
          v[i].x += bb.x; v[i].y += bb.y; v[i].z += bb.z; v[i].w += bb.w;
        }
      }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int k = (lane + 32 * i) * 4;
    if (k < kc) *reinterpret_cast<uint2*>(xr + k) = make_uint2(m2_pack(v[i].x, v[i].y), m2_pack(v[i].z, v[i].w));
  }
}

// ---- banned next tokens of every sequence (transformers NoRepeatNGramLogitsProcessor), into shared memory ----
__device__ void m2_banned(const M2Args& a, M2Smem& S, int cur_len, int tid) {
  if (tid < M2_B) S.nbanned[tid] = 0;
  __syncthreads();
  for (int g = 0; g < a.n_ngrams; ++g) {
    const int n = a.ngrams[g];
    if (n <= 0 || cur_len + 1 < n) continue;
    const int tail = cur_len + 1 - n;
    const int span = cur_len - n + 1;                 // candidate start positions 0 .. cur_len - n
    for (int w = tid; w < span * a.B; w += M2_THREADS) {
      const int b = w / span, i = w - b * span;
      const int64_t* idr = a.ids + (int64_t)b * a.ids_ld;
      bool same = true;
      for (int j = 0; j < n - 1; ++j) same = same && (__ldcg(idr + i + j) == __ldcg(idr + tail + j));
      if (same) {
        const int slot = atomicAdd(&S.nbanned[b], 1);
        if (slot < M2_MAX_BANNED) S.banned[b][slot] = (int)__ldcg(idr + i + n - 1);
        else atomicExch(a.error_flag, 5);
      }
    }
  }
  __syncthreads();
}

// ---- one linear stage ----
__device__ void m2_linear(const M2Args& a, int op, M2Smem& S, M2Pref& pf, int pos, int keybuf, int keyout, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const int64_t* d = a.lin + (size_t)op * M2_LIN_FIELDS;
  const M2Op o = m2_op(d);
  const float* bias = reinterpret_cast<const float*>(d[1]);
  float* out = reinterpret_cast<float*>(d[5]);
  const float* residual = reinterpret_cast<const float*>(d[6]);
  const int act = (int)d[9], mode = (int)d[10], in_mode = (int)d[13];
  const int64_t ldo = d[15];
  const bool argmax = ((int)d[17] & 1) != 0 && a.top_k == 1;
  const bool atomic_out = o.nchunks > 1;
  const int G = gridDim.x;
  const bool tr = a.trace != nullptr && blockIdx.x == 0 && tid == 0;
  int u = m2_first_unit(op);
  if (u < o.total && !(pf.op == op && pf.unit == u)) m2_issue(o, u, pf.w, warp, lane);   // not prefetched: fetch now
  pf.op = -1;
  const int cur_len = pos + 1;
  if (argmax) m2_banned(a, S, cur_len, tid);
  // running arg-max of this thread's (row, batch) position over the CTA's units
  float best_v = -INFINITY;
  int best_n = 0x7fffffff;
  int staged_chunk = -1, ucount = 0;
  uint4 xf[M2_WB];
  const int g = lane >> 2, qd = lane & 3;
  const int er = tid >> 3, eb = tid & 7;       // epilogue position of threads 0..127: row er, batch eb
  for (; u < o.total; u += G) {
    uint4 w[M2_WB][2];
#pragma unroll
    for (int i = 0; i < M2_WB; ++i) { w[i][0] = pf.w[i][0]; w[i][1] = pf.w[i][1]; }
    if (u + G < o.total) m2_issue(o, u + G, pf.w, warp, lane);
    const int chunk = u / o.rows16, rg = u - chunk * o.rows16;
    const int n0 = rg * M2_ROWS;
    const int nblk = min(M2_CHUNK_BLKS, o.K / M2_BLK - chunk * M2_CHUNK_BLKS);
    // epilogue operands do not depend on the MMAs: fetch them now
    float e_bias = 0.f, e_res = 0.f;
    const bool e_on = tid < M2_ROWS * M2_B && n0 + er < o.N && eb < a.B;
    if (e_on) {
      if (bias != nullptr && (!atomic_out || chunk == 0)) e_bias = bias[n0 + er];
      if (!atomic_out && mode == 0 && residual != nullptr && in_mode == 0) e_res = __ldcg(residual + (int64_t)eb * ldo + n0 + er);
    }
    if (chunk != staged_chunk) {
      if (staged_chunk >= 0) __syncthreads();        // everybody finished reading the previous chunk's fragments
      m2_stage_x(a, d, S, chunk, pos, keybuf, warp, lane);
      __syncthreads();
      staged_chunk = chunk;
#pragma unroll
      for (int i = 0; i < M2_WB; ++i) {
        const int blk = warp + M2_WARPS * i;
        xf[i] = blk < nblk ? *reinterpret_cast<const uint4*>(S.xs + g * M2_XPITCH + blk * M2_BLK + qd * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
      if (tr && ucount == 0) a.trace[1] = clock64();
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < M2_WB; ++i) {
      m2_mma(acc, w[i][0].x, w[i][1].x, w[i][0].y, w[i][1].y, xf[i].x, xf[i].y);
      m2_mma(acc, w[i][0].z, w[i][1].z, w[i][0].w, w[i][1].w, xf[i].z, xf[i].w);
    }
    // acc: D[g][2qd], D[g][2qd+1], D[g+8][2qd], D[g+8][2qd+1]  ->  flat (row * 8 + batch)
    float* red = S.red[ucount & 1][warp];
    ++ucount;
    *reinterpret_cast<float2*>(red + lane * 2) = make_float2(acc[0], acc[1]);
    *reinterpret_cast<float2*>(red + 64 + lane * 2) = make_float2(acc[2], acc[3]);
    __syncthreads();
    if (e_on) {
      const float* rp = &S.red[(ucount - 1) & 1][0][tid];
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < M2_WARPS; ++i) v += rp[i * M2_ROWS * M2_B];
      const int n = n0 + er;
      v += e_bias;
      v = apply_act(v, act);
      if (atomic_out) {
        atomicAdd(out + (int64_t)eb * ldo + n, v);
      } else if (mode == 0) {
        if (residual != nullptr && in_mode == 0) v += e_res;
        if (argmax) {
          bool ban = false;
          const int nb = min(S.nbanned[eb], M2_MAX_BANNED);
          for (int i = 0; i < nb; ++i) ban = ban || (S.banned[eb][i] == n);
          if (!ban && (v > best_v || best_n == 0x7fffffff)) { best_v = v; best_n = n; }
        } else {
          out[(int64_t)eb * ldo + n] = v;
        }
      } else {   // packed q | k | v: q to the scratch, k / v appended to the cache at `pos`
        const int seg = n / a.C, nl = n - seg * a.C;
        if (seg == 0) {
          out[(int64_t)eb * ldo + nl] = v;
        } else {
          __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(seg == 1 ? d[11] : d[12]);
          base[(int64_t)eb * d[16] + (int64_t)pos * a.C + nl] = __float2bfloat16_rn(v);
        }
      }
    }
  }
  if (argmax) {
    // CTA-level arg-max per sequence: rows of a warp (lane bits 3,4), then the 4 epilogue warps, then one atomicMax
    unsigned long long key = 0ull;
    if (tid < M2_ROWS * M2_B && best_n != 0x7fffffff)
      key = ((unsigned long long)float_key(best_v) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)best_n);
    const unsigned long long k1 = __shfl_xor_sync(0xffffffffu, key, 8);
    key = k1 > key ? k1 : key;
    const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, 16);
    key = k2 > key ? k2 : key;
    if (warp < 4 && lane < M2_B) S.best[warp][lane] = key;
    __syncthreads();
    if (tid < a.B) {
      unsigned long long k = S.best[0][tid];
#pragma unroll
      for (int i = 1; i < 4; ++i) k = S.best[i][tid] > k ? S.best[i][tid] : k;
      if (k != 0ull) atomicMax(a.keys + keyout * M2_B + tid, k);
    }
  }
}

// ---- single-query attention for one (batch, head): one key per thread ----
template <int HS>
__device__ void m2_attention(const M2Args& a, int ai, M2Smem& S, int pos, int tid) {
  constexpr int NV = HS / 8;                 // 16-byte vectors per row
  const int lane = tid & 31, warp = tid >> 5;
  const int64_t* d = a.att + (size_t)ai * 8;
  const __nv_bfloat16* kc = reinterpret_cast<const __nv_bfloat16*>(d[0]);
  const __nv_bfloat16* vc = reinterpret_cast<const __nv_bfloat16*>(d[1]);
  const int64_t bs = d[2], rs = d[3];
  const int len = d[4] == 0 ? pos + 1 : (int)d[5];
  const float scale = 1.0f / sqrtf((float)HS);
  for (int unit = blockIdx.x; unit < a.B * a.H; unit += gridDim.x) {
    const int b = unit / a.H, h = unit - b * a.H;
    const bool on = tid < len;
    uint4 kr[NV], vr[NV];
    if (on) {
      const __nv_bfloat16* kp = kc + b * bs + (int64_t)tid * rs + h * HS;
      const __nv_bfloat16* vp = vc + b * bs + (int64_t)tid * rs + h * HS;
#pragma unroll
      for (int i = 0; i < NV; ++i) kr[i] = __ldcg(reinterpret_cast<const uint4*>(kp) + i);
#pragma unroll
      for (int i = 0; i < NV; ++i) vr[i] = __ldcg(reinterpret_cast<const uint4*>(vp) + i);
    }
    if (tid < HS) {
      float v = __ldcg(a.q + (int64_t)b * a.C + h * HS + tid);
      S.att_q[tid] = __bfloat162float(__float2bfloat16_rn(v)) * scale;      // autocast: SDPA sees a bf16 query
    }
    __syncthreads();
    float s = -INFINITY;
    if (on) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 q0 = *reinterpret_cast<const float4*>(&S.att_q[i * 8]);
        const float4 q1 = *reinterpret_cast<const float4*>(&S.att_q[i * 8 + 4]);
        acc = fmaf(__uint_as_float(kr[i].x << 16), q0.x, acc); acc = fmaf(__uint_as_float(kr[i].x & 0xffff0000u), q0.y, acc);
        acc = fmaf(__uint_as_float(kr[i].y << 16), q0.z, acc); acc = fmaf(__uint_as_float(kr[i].y & 0xffff0000u), q0.w, acc);
        acc = fmaf(__uint_as_float(kr[i].z << 16), q1.x, acc); acc = fmaf(__uint_as_float(kr[i].z & 0xffff0000u), q1.y, acc);
        acc = fmaf(__uint_as_float(kr[i].w << 16), q1.z, acc); acc = fmaf(__uint_as_float(kr[i].w & 0xffff0000u), q1.w, acc);
      }
      s = acc;
    }
    // softmax over the CTA
    float mx = warp_max(s);
    if (lane == 0) S.att_red[warp] = mx;
    __syncthreads();
    mx = S.att_red[0];
#pragma unroll
    for (int i = 1; i < M2_WARPS; ++i) mx = fmaxf(mx, S.att_red[i]);
    const float p = on ? expf(s - mx) : 0.f;
    float sum = warp_sum(p);
    __syncthreads();
    if (lane == 0) S.att_red[warp] = sum;
    // o[e] = sum_j p_j v_j[e]: transposing shuffle tree over the warp's 32 keys, 32 dims at a time (register budget);
    // after the 5 levels lane l holds dim `base` of the half, base = the lane bits weighted 16,8,4,2,1
    {
      int base = 0;
#pragma unroll
      for (int off = 16, span = 16; off >= 1; off >>= 1, span >>= 1)
        if (lane & off) base += span;
#pragma unroll
      for (int hf = 0; hf < HS / 32; ++hf) {
        float o[32];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 vv = vr[hf * 4 + i];
          const uint32_t r[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[i * 8 + 2 * j] = on ? p * __uint_as_float(r[j] << 16) : 0.f;
            o[i * 8 + 2 * j + 1] = on ? p * __uint_as_float(r[j] & 0xffff0000u) : 0.f;
          }
        }
#pragma unroll
        for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
          const int half = n >> 1;
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < half; ++i) {
            const float send = upper ? o[i] : o[i + half];
            const float keep = upper ? o[i + half] : o[i];
            o[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        S.att_o[warp][hf * 32 + base] = o[0];
      }
    }
    __syncthreads();
    if (tid < HS) {
      float tot = 0.f, ov = 0.f;
#pragma unroll
      for (int i = 0; i < M2_WARPS; ++i) { tot += S.att_red[i]; ov += S.att_o[i][tid]; }
      a.y[(int64_t)b * a.C + h * HS + tid] = tot > 0.f ? ov / tot : 0.f;
    }
    __syncthreads();
  }
}

template <int HS>
__global__ void __launch_bounds__(M2_THREADS, 2) decode_mega2_kernel(M2Args a) {
  __shared__ M2Smem S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pos0 = *a.pos;
  unsigned int epoch = 0;
  M2Pref pf;
  pf.op = -1;
  pf.unit = -1;
  const int n_steps = a.n_prefill + a.n_sample;
  for (int step = 0; step < n_steps; ++step) {
    const bool sampling = step >= a.n_prefill;
    const int32_t* sched = sampling ? a.sched_sample : a.sched_prefill;
    const int n_sched = sampling ? a.n_sched_sample : a.n_sched_prefill;
    const int pos = pos0 + step;
    const int sstep = step - a.n_prefill;                       // index among the sampled steps
    const bool greedy = a.top_k == 1;
    // token source of this step's embedding: the previous sampled step's arg-max key (greedy), else the ids buffer
    const int keybuf = (greedy && sampling && sstep > 0) ? (sstep - 1) % 3 : -1;
    const int keyout = sstep >= 0 ? sstep % 3 : 0;
    const bool trace_step = a.trace != nullptr && step == n_steps - 1 && sampling;
    if (greedy && sampling && blockIdx.x == 0 && tid < a.B) {
      // publish the previous pick for the host and the n-gram ban; recycle the key slot two steps ahead
      if (sstep > 0) a.ids[(int64_t)tid * a.ids_ld + pos] = (int64_t)m2_token(a, tid, pos, keybuf);
      a.keys[((sstep + 1) % 3) * M2_B + tid] = 0ull;
    }
    for (int s = 0; s < n_sched; ++s) {
      const int kind = sched[s * 4], idx = sched[s * 4 + 1];
      const bool tr = trace_step && blockIdx.x == 0 && tid == 0;
      M2Args at = a;
      if (tr) { at.trace = a.trace + s * 4; at.trace[0] = clock64(); } else at.trace = nullptr;
      bool need_sync = true;
      if (kind == 0) {
        m2_linear(at, idx, S, pf, pos, keybuf, keyout, tid);
      } else if (kind == 1) {
        m2_attention<HS>(at, idx, S, pos, tid);
      } else if (kind == 2) {
        if (!greedy) {
          // general sampler on B CTAs, vocabulary row processed in place (global / L2)
          if ((int)blockIdx.x < a.B) {
            const int b = blockIdx.x;
            float* row = a.logits + (int64_t)b * a.V;
            SampleScratch& samp = *reinterpret_cast<SampleScratch*>(&S);
            static_assert(sizeof(SampleScratch) <= sizeof(M2Smem), "sampler scratch must fit the stage buffers");
            const int choice = sample_row_smem(row, samp, row, a.V, a.ids + (int64_t)b * a.ids_ld, pos + 1, a.temperature,
                                               a.top_k, a.ngrams, a.n_ngrams, *a.seed_ptr, b, nullptr, tid, M2_THREADS);
            if (tid == 0) a.ids[(int64_t)b * a.ids_ld + pos + 1] = (int64_t)choice;
          }
        } else {
          need_sync = false;                  // the arg-max keys were completed by the LM-head stage's barrier
        }
      } else {
        need_sync = false;                    // ADVANCE: the position is a kernel-local counter here
      }
      if (tr) at.trace[2] = clock64();
      if (need_sync) {
        // weights of the next linear stage do not depend on anything: put them in flight before waiting
        if (pf.op < 0) {
          int s2 = s + 1, step2 = step;
          const int32_t* sc2 = sched;
          int n2 = n_sched;
          for (int guard = 0; guard < 8; ++guard) {
            if (s2 >= n2) {
              ++step2;
              if (step2 >= n_steps) break;
              const bool samp2 = step2 >= a.n_prefill;
              sc2 = samp2 ? a.sched_sample : a.sched_prefill;
              n2 = samp2 ? a.n_sched_sample : a.n_sched_prefill;
              s2 = 0;
            }
            if (sc2[s2 * 4] == 0) {
              const int op2 = sc2[s2 * 4 + 1];
              const M2Op o2 = m2_op(a.lin + (size_t)op2 * M2_LIN_FIELDS);
              const int u2 = m2_first_unit(op2);
              if (u2 < o2.total) {
                m2_issue(o2, u2, pf.w, warp, lane);
                pf.op = op2;
                pf.unit = u2;
              }
              break;
            }
            ++s2;
          }
        }
        m2_grid_sync(a.bar, epoch, a.error_flag, tid);
      }
      if (tr) at.trace[3] = clock64();
    }
  }
  // final bookkeeping: last pick -> ids, position counter
  const int pos_end = pos0 + n_steps;
  if (blockIdx.x == 0) {
    if (a.top_k == 1 && a.n_sample > 0 && tid < a.B)
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
                                const int32_t* sched_prefill, int64_t n_sched_prefill, int64_t n_prefill, int64_t n_sample,
                                int64_t B, int64_t C, int64_t H, int64_t V, int64_t n_prompt, int64_t* ids, int64_t ids_ld,
                                int32_t* pos, float* q, float* y, float* logits, uint32_t* bar, int32_t* error_flag,
                                uint64_t* keys, float temperature, int64_t top_k, const int32_t* ngrams, int64_t n_ngrams,
                                const uint64_t* seed_ptr, int32_t* ticket, int64_t max_k, int64_t max_len, int64_t* trace,
                                void* stream) {
  I2T_REQUIRE(lin && att && sched_sample && sched_prefill && ids && pos && q && y && logits && bar && error_flag && keys &&
                  seed_ptr && ticket, "decode_mega2: null pointer");
  I2T_REQUIRE(B > 0 && B <= M2_B, "decode_mega2: batch %lld outside 1..8", (long long)B);
  I2T_REQUIRE(H > 0 && C % H == 0 && (C / H == 64 || C / H == 32), "decode_mega2: head_dim must be 32 or 64");
  I2T_REQUIRE(C % M2_BLK == 0 && max_k % M2_BLK == 0, "decode_mega2: widths must be multiples of %d", M2_BLK);
  I2T_REQUIRE(C <= M2_CHUNK_BLKS * M2_BLK, "decode_mega2: n_embd=%lld above %d (LayerNorm rows are staged as one chunk)",
              (long long)C, M2_CHUNK_BLKS * M2_BLK);
  I2T_REQUIRE(max_len <= M2_MAX_KEYS, "decode_mega2: %lld cached positions exceed the one-key-per-thread limit %d",
              (long long)max_len, M2_MAX_KEYS);
  I2T_REQUIRE(temperature > 0.f && n_prefill >= 0 && n_sample >= 0 && n_prefill + n_sample > 0, "decode_mega2: bad step counts / temperature");
  cudaStream_t st = (cudaStream_t)stream;
  M2Args a;
  a.lin = lin; a.att = att; a.sched_sample = sched_sample; a.sched_prefill = sched_prefill;
  a.n_sched_sample = (int)n_sched_sample; a.n_sched_prefill = (int)n_sched_prefill;
  a.n_prefill = (int)n_prefill; a.n_sample = (int)n_sample;
  a.B = (int)B; a.C = (int)C; a.H = (int)H; a.V = (int)V; a.n_prompt = (int)n_prompt;
  a.ids = ids; a.ids_ld = ids_ld; a.pos = pos; a.q = q; a.y = y; a.logits = logits; a.bar = bar; a.error_flag = error_flag;
  a.keys = reinterpret_cast<unsigned long long*>(keys);
  a.temperature = temperature; a.top_k = (int)(top_k > 0 ? top_k : 0);
  a.ngrams = ngrams; a.n_ngrams = (int)n_ngrams; a.seed_ptr = seed_ptr; a.ticket = ticket;
  a.trace = reinterpret_cast<long long*>(trace);
  const void* kern = (C / H == 64) ? (const void*)decode_mega2_kernel<64> : (const void*)decode_mega2_kernel<32>;
  int per_sm = 0;
  I2T_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, M2_THREADS, 0));
  I2T_REQUIRE(per_sm >= 1, "decode_mega2: kernel does not fit on an SM");
  const int grid = num_sms() * (per_sm >= 2 ? 2 : 1);
  I2T_CUDA(cudaMemsetAsync(bar, 0, sizeof(uint32_t), st));
  I2T_CUDA(cudaMemsetAsync(keys, 0, sizeof(uint64_t) * 3 * M2_B, st));
  void* params[] = {&a};
  I2T_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(M2_THREADS), params, 0, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return I2T_OK;
}
