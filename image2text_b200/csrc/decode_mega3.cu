// Decode megakernel v3 (bf16 weights): the whole generate() loop as ONE cooperative launch WITHOUT grid barriers.
//
// Why v3.  v2 (decode_mega2.cu) ran a decode step as 81 dependent stages separated by grid barriers: 12 % of the HBM
// roofline, 42 % of the warp samples parked at the barrier (profiles/r01_ncu_mega2_stall_mix.txt), and the weight stream
// only ever one stage ahead.  v3 removes both limits:
//   * DATAFLOW instead of barriers.  Every activation that crosses CTAs lives in an exchange buffer that is pre-filled
//     with a POISON pattern (0xFFFFFFFF per fp32 word / 0xFFFF per bf16: NaNs no kernel arithmetic produces).  A producer
//     just stores its result; a consumer polls the words it needs until none is poison.  Data and "ready flag" are the
//     same word, so the release fence + atomic + acquire poll + reload of a barrier collapses into one L2 round trip, and a
//     CTA only waits for the stages it takes part in.  Each buffer has 3 generations (step % 3): the thread that writes an
//     element for step t also re-poisons the same element of generation t+1 (last read during step t-2).
//   * WEIGHTS NEVER WAIT FOR ACTIVATIONS.  The bf16 weights are re-packed once per weight version into one contiguous
//     stream PER CTA, in exactly the order (stage, tile, k-chunk) and the per-thread mma fragment layout the CTA consumes
//     them.  A dedicated producer warp walks that stream with cp.async.bulk (one 8-24 KB copy per chunk, L2 evict-first)
//     into a ring of 24 KB slots gated by full / empty mbarriers -- several stages ahead of the consumers, so HBM stays
//     busy while the dependency chain waits on L2 latency.
//   * the per-stage arithmetic is v2's: the batch (<= 8 sequences) is the N of mma.sync.m16n8k16, a tile = 16 weight rows,
//     8 consumer warps split K and reduce through shared memory in a fixed order (deterministic), LayerNorm on registers,
//     greedy n-gram ban + arg-max fused into the LM-head epilogue (per-CTA keys, no atomics).
// Replaces, for KV-cached decode, reference models/vision_encoder_decoder.py:144-180 (generate loop),
// models/decoder.py:214-256 and models/layers.py:447-486,565-614 (one-token forward).
//
// Tables (int64): lin[op][24] =
//   0 unused | 1 bias f32* | 2 ln gamma | 3 ln beta | 4 input (generation 0; wte when in_mode 1) | 5 output (generation 0)
//   6 residual (generation 0) or 0 | 7 N | 8 K | 9 activation | 10 mode (1 = packed q|k|v: q -> output, k / v -> cache row
//   `pos`) | 11 k cache | 12 v cache | 13 in_mode (1 = token embedding wte[tok] + wpe[n_prompt + pos]) | 14 wpe | 15 output
//   row pitch | 16 cache batch stride | 17 flags (1 LM head, 2 input is bf16, 4 output is bf16, 8 publish the embedding)
//   18 rot (tile u belongs to CTA (u + rot) % grid) | 19 where the embedding is published (generation 0) | 20 input row
//   pitch in elements (0 = K: lets a K-split op read a column range of a wider buffer) | 21 group size Gs (0 = grid): only
//   the first Gs CTAs of the rotation take part, tile u belongs to CTA ((u % Gs) + rot) % grid -- fewer CTAs fan the
//   activations out of L2 when an op has few tiles per CTA anyway | 22 1 = tcgen05 stage (weights packed as UMMA atoms) | 23 folded residual: 1 + combine-table row
// cmb[c][8] (combine: out = bias + residual + sum of n partial buffers, the second half of a K-split projection) =
//   0 n partials | 1 first partial f32 (generation 0) | 2 bytes between partials | 3 bias f32* | 4 residual (generation 0)
//   5 output (generation 0) | 6 N (row length) | 7 rot
// att[a][12] = 0 K | 1 V | 2 batch stride | 3 row stride | 4 len mode (0: pos + 1 keys, the last one appended this step;
//   1: constant) | 5 constant length | 6 q f32 (generation 0) | 7 y bf16 out (generation 0) | 8 rot | 9-11 unused
// sched[s][4] = {kind (0 linear, 1 attention, 2 sample, 3 combine), index, 0, 0}; LM-head ops and the sample entry are skipped in
// prefill steps.
#include "common.cuh"
#include "sampler.cuh"
#include "tc_common.cuh"

namespace i2t {

constexpr int M3_CWARPS = 8;                  // consumer warps
constexpr int M3_CTHREADS = M3_CWARPS * 32;
constexpr int M3_THREADS = M3_CTHREADS + 64;  // + the weight-stream producer warp + the tcgen05 issuer warp
constexpr int M3_TBUFS = 2;                   // TMEM tile buffers (tcgen05 path): 4 partial accumulators x 16 columns each
constexpr int M3_B = 8;                       // batch rows = MMA N
constexpr int M3_ROWS = 16;                   // weight rows per tile = MMA M
constexpr int M3_BLK = 32;                    // k elements per block (one 16-byte vector per thread and row)
constexpr int M3_KC = 768;                    // k elements per chunk
constexpr int M3_NB = M3_KC / M3_BLK / M3_CWARPS;            // blocks per warp and chunk (3)
constexpr int M3_BLOCK_BYTES = 2 * M3_CTHREADS * 16;         // one block index of every warp, both row halves: 8 KB
constexpr int M3_SLOT_BYTES = M3_NB * M3_BLOCK_BYTES;        // 24 KB
constexpr int M3_MAX_SLOTS = 8;
constexpr int M3_LIN_FIELDS = 24;
constexpr int M3_ATT_FIELDS = 12;
constexpr int M3_CMB_FIELDS = 8;
constexpr int M3_MAX_LM_ROWS = 1024;          // vocabulary rows one CTA can own in the LM head (64 tiles)
constexpr int M3_MAX_KEYS = M3_CTHREADS;      // attention: one key per thread
constexpr int M3_GENS = 3;
constexpr int M3_RED_T = 160;                 // floats per partial tile: [8 batch][20] (16 rows + 4 pad: conflict-free STS)
constexpr int M3_NVX = 6;                     // 16-byte vectors per lane and staging pass
constexpr int M3_TRACE = 8;                   // stamps per stage

struct M3Args {
  const int64_t* lin;
  const int64_t* att;
  const int64_t* cmb;
  const int32_t* sched;
  int n_sched, n_ops, n_att, n_cmb;
  int n_prefill, n_sample;
  int B, C, H, V, n_prompt, max_k, max_len, nslots;
  int64_t* ids;
  int64_t ids_ld;
  int32_t* pos;
  float* logits;
  int64_t ldl;
  unsigned int* bar;
  int32_t* error_flag;
  unsigned long long* ctakeys;   // [3][grid][8] per-CTA arg-max keys, 0 = not written yet
  const uint8_t* wpack;          // packed weight streams
  const int64_t* cta_base;       // [grid] byte offset of every CTA's stream
  int64_t gen_stride;            // bytes between the generations of an exchange buffer
  float temperature;
  int top_k;
  const int32_t* ngrams;
  int n_ngrams;
  const uint64_t* seed_ptr;
  int sleep_ns;                  // back-off between unsuccessful polls (0 = spin)
  int tc;                        // 1: linear stages on tcgen05 (weights packed in the 128B-swizzled K-major UMMA layout)
  long long* trace;              // optional [n_sched][4] clock64 stamps of CTA `trace_cta` for the LAST sampled step
  int trace_cta;
};

struct M3AttScratch {
  uint4 kfresh[8];                                 // the K row appended this step (polled by warp 7, read by thread `pos`)
  float att_q[64];
  float att_p[M3_MAX_KEYS];
  float att_red[M3_CWARPS];
  float att_o[M3_CWARPS][64];
};
struct __align__(16) M3Fixed {
  union {                                          // a CTA runs one stage at a time:
    float red[M3_CWARPS][2 * M3_RED_T];            //   linear: per-warp partial tiles of a pair (10 KB);
    M3AttScratch att;                              //   attention: query, probabilities, partial outputs
  };
  uint64_t full[M3_MAX_SLOTS], empty[M3_MAX_SLOTS];
  uint64_t tfull[M3_TBUFS], tempty[M3_TBUFS];    // tcgen05 path: accumulator buffer complete / drained
  uint64_t xsempty;                              // tcgen05 path: every MMA of the stage has read the staged activations
  uint32_t tmem_base, pad_;
  uint32_t banmask[M3_MAX_LM_ROWS / 4];          // no-repeat-n-gram ban: one byte per LM-head row this CTA owns, bit b = sequence b
  int hist[M3_B][M3_MAX_KEYS + 8];               // token history of every sequence (n-gram ban)
  int tok[M3_B];                                  // the tokens this step embeds
  unsigned long long best[M3_CWARPS][2];
  uint4 fold[M3_CWARPS * 8][3];                   // folded combine: residual / partial rows of an epilogue thread, fetched by cp.async
};

// Dynamic shared memory: [ring | M3Fixed | work | lin | att | cmb | sched | (sampler scratch, sampling mode only)].
// `work` is used by one stage at a time: a linear stage keeps the staged activation rows (bf16, 8 x xpitch) and its LayerNorm
// gamma / beta there, an attention stage the V rows of its (batch, head).
extern __shared__ __align__(128) uint8_t m3_smem[];
__host__ __device__ inline size_t m3_align16(size_t x) { return (x + 15) & ~(size_t)15; }
struct M3Sm {
  int xpitch;                         // bf16 elements between the staged activation rows
  uint32_t fixed_off, xs_off, ln_off, lin_off, att_off, cmb_off, sched_off, samp_off, total;
};
__host__ __device__ inline M3Sm m3_layout(int nslots, int max_k, int max_len, int hs, int n_ops, int n_att, int n_cmb, int n_sched,
                                          bool sampling) {
  M3Sm S;
  S.xpitch = max_k + 32;              // bytes = 2 * max_k + 64 = 64 (mod 128): conflict-free LDS.128 of the B fragments
  uint32_t off = (uint32_t)nslots * 24576u;
  S.fixed_off = off; off += (uint32_t)m3_align16(sizeof(M3Fixed));
  off = (off + 1023u) & ~1023u;      // the tcgen05 path reads the activations as 128B-swizzled atoms (1024-byte aligned)
  S.xs_off = off;
  const uint32_t xs_plain = (uint32_t)m3_align16((size_t)8 * S.xpitch * 2), xs_tc = (uint32_t)(max_k / 64 + 1) * 1024u;
  const uint32_t xs = xs_plain > xs_tc ? xs_plain : xs_tc;
  S.ln_off = off + xs;
  const uint32_t lin_work = xs + 2u * 768u * 4u, att_work = (uint32_t)m3_align16((size_t)max_len * hs * 2);
  off += lin_work > att_work ? lin_work : att_work;
  S.lin_off = off; off += (uint32_t)m3_align16((size_t)n_ops * 24 * 8);
  S.att_off = off; off += (uint32_t)m3_align16((size_t)n_att * 12 * 8);
  S.cmb_off = off; off += (uint32_t)m3_align16((size_t)n_cmb * 8 * 8);
  S.sched_off = off; off += (uint32_t)m3_align16((size_t)n_sched * 16);
  S.samp_off = off;
  if (sampling) off += (uint32_t)m3_align16(sizeof(SampleScratch));
  S.total = off;
  return S;
}
__device__ __forceinline__ uint4* m3_ring(int slot) { return reinterpret_cast<uint4*>(m3_smem + (size_t)slot * M3_SLOT_BYTES); }
__device__ __forceinline__ M3Fixed* m3_f(const M3Sm& S) { return reinterpret_cast<M3Fixed*>(m3_smem + S.fixed_off); }
__device__ __forceinline__ __nv_bfloat16* m3_xs(const M3Sm& S) { return reinterpret_cast<__nv_bfloat16*>(m3_smem + S.xs_off); }
// where the 8 (4) consecutive k of batch row `row` starting at k live: plain rows (mma.sync path) or the K-major,
// 128B-swizzled atoms the tcgen05 path reads as its B operand ([k / 64][row][16-byte chunk ((k % 64) / 8) ^ row])
__device__ __forceinline__ __nv_bfloat16* m3_xaddr(const M3Sm& S, bool tc, int row, int k) {
  __nv_bfloat16* xs = m3_xs(S);
  if (!tc) return xs + row * S.xpitch + k;
  return xs + (k >> 6) * 512 + row * 64 + ((((k >> 3) & 7) ^ row) << 3) + (k & 7);
}
__device__ __forceinline__ float* m3_ln(const M3Sm& S) { return reinterpret_cast<float*>(m3_smem + S.ln_off); }   // gamma[768] | beta[768]
__device__ __forceinline__ const int64_t* m3_lin(const M3Sm& S, int op) {
  return reinterpret_cast<const int64_t*>(m3_smem + S.lin_off) + (size_t)op * M3_LIN_FIELDS;
}
__device__ __forceinline__ const int64_t* m3_att(const M3Sm& S, int ai) {
  return reinterpret_cast<const int64_t*>(m3_smem + S.att_off) + (size_t)ai * M3_ATT_FIELDS;
}
__device__ __forceinline__ const int64_t* m3_cmb(const M3Sm& S, int ci) {
  return reinterpret_cast<const int64_t*>(m3_smem + S.cmb_off) + (size_t)ci * M3_CMB_FIELDS;
}
__device__ __forceinline__ const int32_t* m3_sched(const M3Sm& S) { return reinterpret_cast<const int32_t*>(m3_smem + S.sched_off); }

// ---- small PTX helpers ----
__device__ __forceinline__ void m3_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void m3_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void m3_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void m3_cp_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// barrier over the 8 consumer warps (the producer warp never joins)
__device__ __forceinline__ void m3_csync() { asm volatile("bar.sync 1, %0;" ::"n"(M3_CTHREADS) : "memory"); }
__device__ __forceinline__ void m3_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool m3_mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// polling loads: always served by L2 (the point of coherence), never hoisted or merged by the compiler
__device__ __forceinline__ uint4 m3_ld16(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t m3_ld4(const void* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long m3_ld8(const void* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// stores another CTA polls for: issued as strong gpu-scope stores (never parked in a write-combining buffer)
__device__ __forceinline__ void m3_st16(void* p, uint4 v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void m3_st8(void* p, uint2 v) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void m3_st8(void* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void m3_st4(void* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void m3_st2(void* p, unsigned short v) {
  asm volatile("st.relaxed.gpu.global.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory");
}
__device__ __forceinline__ uint4 m3_f4_bits(float4 v) {
  return make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
}
__device__ __forceinline__ bool m3_ok32(uint4 v) {   // no fp32 poison word
  return v.x != 0xFFFFFFFFu && v.y != 0xFFFFFFFFu && v.z != 0xFFFFFFFFu && v.w != 0xFFFFFFFFu;
}
__device__ __forceinline__ bool m3_ok16(uint4 v) {   // no bf16 poison half-word
  return (__vcmpeq2(v.x, 0xFFFFFFFFu) | __vcmpeq2(v.y, 0xFFFFFFFFu) | __vcmpeq2(v.z, 0xFFFFFFFFu) | __vcmpeq2(v.w, 0xFFFFFFFFu)) == 0u;
}
// A wait that can never hang the GPU: after 2^22 unsuccessful polls (~1 s) the error flag is raised; once it is up every
// wait of every CTA gives up at its first unsuccessful poll and the kernel drains (results are then garbage, the host raises).
__device__ __forceinline__ bool m3_giveup(uint32_t& spins, int32_t* error_flag) {
  ++spins;
  if (spins == 1u || (spins & 255u) == 0u) {
    if (m3_ld4(error_flag) != 0u) return true;
    if (spins > (1u << 22)) {
      atomicExch(error_flag, 2);
      return true;
    }
  }
  return false;
}
__device__ __forceinline__ void m3_mbar_wait(uint64_t* bar, uint32_t parity, int32_t* error_flag) {
  uint32_t spins = 0;
  while (!m3_mbar_try(bar, parity)) {
    if (++spins > (1u << 24)) {
      atomicExch(error_flag, 3);
      break;
    }
  }
}
template <typename T>
__device__ __forceinline__ T* m3_gen(int64_t base, int64_t gen_stride, int gen) {
  return reinterpret_cast<T*>(base + (int64_t)gen * gen_stride);
}

// grid-wide barrier (sampling mode only: all logits must exist before the sampler CTAs read them)
__device__ __forceinline__ void m3_grid_sync(unsigned int* bar, unsigned int& epoch, int32_t* error_flag, int tid) {
  m3_csync();
  if (tid == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
    unsigned int seen = 0;
    uint32_t spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if ((int)(seen - epoch) >= 0) break;
    } while (!m3_giveup(spins, error_flag));
  }
  m3_csync();
}

// ---- work decomposition of a linear op ----
__device__ __forceinline__ int m3_nkc(int K) { return (K + M3_KC - 1) / M3_KC; }
__device__ __forceinline__ int m3_chunk_nb(int K, int kc) {          // blocks per warp of chunk kc (1..3)
  const int span = min(M3_KC, K - kc * M3_KC);
  return ((span + M3_BLK - 1) / M3_BLK + M3_CWARPS - 1) / M3_CWARPS;
}
__device__ __forceinline__ int m3_kpad(int K) {                       // staged row length: every chunk padded to 256 k
  const int last = m3_nkc(K) - 1;
  return last * M3_KC + m3_chunk_nb(K, last) * M3_BLK * M3_CWARPS;
}
__device__ __forceinline__ int m3_first_unit(int rot) {          // (blockIdx - rot) mod grid, rot in [0, grid): no division
  const int G = (int)gridDim.x;
  const int t = (int)blockIdx.x + G - rot;
  return t >= G ? t - G : t;
}

struct M3Ring {
  int slot;
  uint32_t phase;
  __device__ __forceinline__ void advance(int nslots) {
    if (++slot == nslots) { slot = 0; phase ^= 1u; }
  }
};

// ---- the weight-stream producer: one thread walks this CTA's packed stream in schedule order ----
__device__ __forceinline__ void m3_producer(const M3Args& a, const M3Sm& S) {
  M3Fixed* f = m3_f(S);
  const uint8_t* base = a.wpack + a.cta_base[blockIdx.x];
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  const int32_t* sched = m3_sched(S);
  M3Ring R{0, 0u};
  const int G = (int)gridDim.x;
  const int n_steps = a.n_prefill + a.n_sample;
  for (int step = 0; step < n_steps; ++step) {
    const bool sampling = step >= a.n_prefill;
    const uint8_t* src = base;
    for (int s = 0; s < a.n_sched; ++s) {
      if (sched[s * 4] != 0) continue;
      const int64_t* d = m3_lin(S, sched[s * 4 + 1]);
      if (((int)d[17] & 1) != 0 && !sampling) continue;
      const int N = (int)d[7], K = (int)d[8];
      const int total = (N + M3_ROWS - 1) / M3_ROWS;
      const int nkc = m3_nkc(K);
      const int Gs = d[21] != 0 ? (int)d[21] : G;
      const int pu0 = m3_first_unit((int)d[18]);
      for (int u = pu0 < Gs ? pu0 : total; u < total; u += Gs) {
        for (int kc = 0; kc < nkc; ++kc) {
          const uint32_t bytes = (uint32_t)m3_chunk_nb(K, kc) * M3_BLOCK_BYTES;
          m3_mbar_wait(&f->empty[R.slot], R.phase ^ 1u, a.error_flag);
          mbar_expect_tx(&f->full[R.slot], bytes);
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                  smem_u32(m3_ring(R.slot))),
              "l"(src), "r"(bytes), "r"(smem_u32(&f->full[R.slot])), "l"(policy)
              : "memory");
          src += bytes;
          R.advance(a.nslots);
        }
      }
    }
  }
}

__device__ __forceinline__ uint32_t m3_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- stage the activations of a linear stage: warp b owns batch row b ----
// fp32 input (the residual stream, K <= 768): optimistic full read, then every lane polls ITS first missing vector (512 B
// per warp and round instead of the whole row), then the row is re-read; LayerNorm on registers.  bf16 input (attention
// output, MLP hidden): passes of 6 vectors per lane, copied as they are.
__device__ __forceinline__ void m3_stage_x(const M3Args& a, const int64_t* d, const M3Sm& S, int gen, int pos, int u0, int warp,
                                           int lane, long long* trace, const bool tc = false) {
  const int K = (int)d[8];
  const int flags = (int)d[17];
  const int in_mode = (int)d[13];
  const int kpad = m3_kpad(K);
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  if (warp >= a.B) {                         // unused batch rows: zeros
    for (int k = lane * 8; k < kpad; k += 256) *reinterpret_cast<uint4*>(m3_xaddr(S, tc, warp, k)) = zero4;
    if ((flags & 2) == 0 && d[2] != 0) {     // (the LayerNorm path has one CTA barrier: keep the count equal)
      m3_cp_wait0();
      m3_csync();
    }
    return;
  }
  for (int k = K + lane * 8; k < kpad; k += 256) *reinterpret_cast<uint4*>(m3_xaddr(S, tc, warp, k)) = zero4;     // (K % 8 == 0)
  if ((flags & 2) != 0) {                    // bf16 exchange buffer, no LayerNorm
    const uint8_t* src = m3_gen<const uint8_t>(d[4], a.gen_stride, gen) + (int64_t)warp * (d[20] != 0 ? d[20] : (int64_t)K) * 2;
    const int nvec = K / 8;
    for (int v0 = 0; v0 < nvec; v0 += 32 * M3_NVX) {
      uint4 v[M3_NVX];
      uint32_t spins = 0;
      if (v0 + lane < nvec) {
        while (!m3_ok16(m3_ld16(src + (size_t)(v0 + lane) * 16))) {
          if (m3_giveup(spins, a.error_flag)) break;
          if (a.sleep_ns > 0) __nanosleep(a.sleep_ns);
        }
      }
      while (true) {
        int bad = -1;
#pragma unroll
        for (int i = M3_NVX - 1; i >= 0; --i) {
          const int vi = min(v0 + lane + 32 * i, nvec - 1);      // (clamped: every element is defined, the array stays in registers)
          v[i] = m3_ld16(src + (size_t)vi * 16);
          if (!m3_ok16(v[i])) bad = vi;
        }
        if (__all_sync(0xffffffffu, bad < 0)) break;
        bool quit = false;
        while (bad >= 0 && !m3_ok16(m3_ld16(src + (size_t)bad * 16))) {
          if (m3_giveup(spins, a.error_flag)) { quit = true; break; }
          if (a.sleep_ns > 0) __nanosleep(a.sleep_ns);
        }
        if (__any_sync(0xffffffffu, quit)) break;
      }
#pragma unroll
      for (int i = 0; i < M3_NVX; ++i) {
        const int vi = v0 + lane + 32 * i;
        if (vi < nvec) *reinterpret_cast<uint4*>(m3_xaddr(S, tc, warp, vi * 8)) = v[i];
      }
    }
    return;
  }
  // fp32 row (K <= 768)
  const float* ln_g = reinterpret_cast<const float*>(d[2]);
  const float* ln_b = reinterpret_cast<const float*>(d[3]);
  float4 v[M3_NVX];
  if (in_mode == 1) {                        // x = wte[tok] + wpe[n_prompt + pos]
    const float* wte = reinterpret_cast<const float*>(d[4]) + (int64_t)m3_f(S)->tok[warp] * K;
    const float* wpe = reinterpret_cast<const float*>(d[14]) + (int64_t)(a.n_prompt + pos) * K;
#pragma unroll
    for (int i = 0; i < M3_NVX; ++i) {
      const int k = min((lane + 32 * i) * 4, K - 4);
      const float4 te = __ldcg(reinterpret_cast<const float4*>(wte + k));
      const float4 pe = __ldcg(reinterpret_cast<const float4*>(wpe + k));
      v[i] = make_float4(te.x + pe.x, te.y + pe.y, te.z + pe.z, te.w + pe.w);
    }
    if ((flags & 8) != 0 && u0 == 0) {       // the CTA that owns tile 0 publishes the embedding as the residual stream
      float* xo = m3_gen<float>(d[19], a.gen_stride, gen) + (int64_t)warp * K;
      float* xp = m3_gen<float>(d[19], a.gen_stride, (gen + 1) % M3_GENS) + (int64_t)warp * K;
#pragma unroll
      for (int i = 0; i < M3_NVX; ++i) {
        const int k = (lane + 32 * i) * 4;
        if (k < K) {
          m3_st16(xp + k, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
          m3_st16(xo + k, m3_f4_bits(v[i]));
        }
      }
    }
  } else {
    const float* src = m3_gen<const float>(d[4], a.gen_stride, gen) + (int64_t)warp * K;
    uint32_t spins = 0;
    // consumers are usually early: poll ONE vector per lane (512 B per warp and round) before reading the whole row
    if (lane * 4 < K) {
      while (!m3_ok32(m3_ld16(src + lane * 4))) {
        if (m3_giveup(spins, a.error_flag)) break;
        if (a.sleep_ns > 0) __nanosleep(a.sleep_ns);
      }
    }
    while (true) {
      int bad = -1;
#pragma unroll
      for (int i = M3_NVX - 1; i >= 0; --i) {
        const int k = min((lane + 32 * i) * 4, K - 4);
        const uint4 r = m3_ld16(src + k);
        v[i] = make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
        if (!m3_ok32(r)) bad = k;
      }
      if (__all_sync(0xffffffffu, bad < 0)) break;
      bool quit = false;
      while (bad >= 0 && !m3_ok32(m3_ld16(src + bad))) {
        if (m3_giveup(spins, a.error_flag)) { quit = true; break; }
        if (a.sleep_ns > 0) __nanosleep(a.sleep_ns);
      }
      if (__any_sync(0xffffffffu, quit)) break;
    }
    if (trace != nullptr) { trace[6] = clock64(); trace[7] = (long long)spins; }
  }
  if (ln_g != nullptr) {
    // LayerNorm parameters: requested with cp.async by the whole CTA before the polling started (m3_linear)
    m3_cp_wait0();
    m3_csync();
    const float has_b = ln_b != nullptr ? 1.f : 0.f;
    // one pass: sums of (x - s) and (x - s)^2 with s = the row's first element (no cancellation for rows with a large mean)
    const float sh = __shfl_sync(0xffffffffu, v[0].x, 0);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < M3_NVX; ++i)
      if ((lane + 32 * i) * 4 < K) {
        const float c0 = v[i].x - sh, c1 = v[i].y - sh, c2 = v[i].z - sh, c3 = v[i].w - sh;
        s1 += (c0 + c1) + (c2 + c3);
        s2 += (c0 * c0 + c1 * c1) + (c2 * c2 + c3 * c3);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 / (float)K;
    const float mu = sh + m1;
    const float rs = 1.0f / sqrtf(fmaxf(s2 / (float)K - m1 * m1, 0.f) + 1e-5f);
#pragma unroll
    for (int i = 0; i < M3_NVX; ++i)
      if ((lane + 32 * i) * 4 < K) {
        const float4 gm = *reinterpret_cast<const float4*>(m3_ln(S) + (lane + 32 * i) * 4);
        const float4 bt = *reinterpret_cast<const float4*>(m3_ln(S) + M3_KC + (lane + 32 * i) * 4);
        v[i].x = fmaf((v[i].x - mu) * rs, gm.x, has_b * bt.x); v[i].y = fmaf((v[i].y - mu) * rs, gm.y, has_b * bt.y);
        v[i].z = fmaf((v[i].z - mu) * rs, gm.z, has_b * bt.z); v[i].w = fmaf((v[i].w - mu) * rs, gm.w, has_b * bt.w);
      }
  }
#pragma unroll
  for (int i = 0; i < M3_NVX; ++i) {
    const int k = (lane + 32 * i) * 4;
    if (k < K) *reinterpret_cast<uint2*>(m3_xaddr(S, tc, warp, k)) = make_uint2(m3_pack(v[i].x, v[i].y), m3_pack(v[i].z, v[i].w));
  }
}

// ---- banned next tokens of every sequence (transformers NoRepeatNGramLogitsProcessor) from the in-kernel history ----
// The result is a bitmap over the vocabulary rows THIS CTA owns in the LM head (tiles u0, u0 + G, ...: local row =
// 16 * (u - u0) / G + token % 16), so the epilogue tests 4 rows with one shared-memory word instead of scanning a list.
__device__ __noinline__ void m3_banned(M3Fixed* f, int B, const int32_t* ngrams, int n_ngrams, int cur_len, int tid, int u0, int G) {
  for (int i = tid; i < M3_MAX_LM_ROWS / 4; i += M3_CTHREADS) f->banmask[i] = 0u;
  m3_csync();
  for (int g = 0; g < n_ngrams; ++g) {
    const int n = ngrams[g];
    if (n <= 0 || cur_len + 1 < n) continue;
    const int tail = cur_len + 1 - n;
    const int span = cur_len - n + 1;                 // candidate start positions 0 .. cur_len - n
    for (int w = tid; w < span * B; w += M3_CTHREADS) {
      const int b = w / span, i = w - b * span;
      const int* idr = f->hist[b];
      bool same = true;
      for (int j = 0; j < n - 1; ++j) same = same && (idr[i + j] == idr[tail + j]);
      if (same) {
        const int tok = idr[i + n - 1];
        const int du = (tok >> 4) - u0;
        if (du >= 0 && du % G == 0) {
          const int r = (du / G) * M3_ROWS + (tok & 15);
          atomicOr(&f->banmask[r >> 2], (1u << b) << (8 * (r & 3)));
        }
      }
    }
  }
  m3_csync();
}

// bf16 path: tanh.approx (|err| ~ 5e-4 relative) is far below the bf16 rounding of the value it feeds
__device__ __forceinline__ float m3_act(float x, int act) {
  if (act == I2T_ACT_GELU_TANH) {
    float t;
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    return 0.5f * x * (1.0f + t);
  }
  if (act == I2T_ACT_GELU_ERF) return gelu_erf_f(x);
  return x;
}

// ---- one linear stage ----
__device__ __forceinline__ void m3_linear(const M3Args& a, int op, const M3Sm& S, M3Ring& R, int gen, int pos, int keyslot,
                                          int tid, long long* trace, uint32_t xcount = 0) {
  const int lane = tid & 31, warp = tid >> 5;
  M3Fixed* f = m3_f(S);
  const int64_t* d = m3_lin(S, op);
  const int N = (int)d[7], K = (int)d[8];
  const int flags = (int)d[17];
  const int total = (N + M3_ROWS - 1) / M3_ROWS;
  const int nkc = m3_nkc(K);
  const int G = (int)gridDim.x;
  const int Gs = d[21] != 0 ? (int)d[21] : G;                   // CTAs taking part; a CTA's tiles are Gs apart
  const int uf = m3_first_unit((int)d[18]);
  const int u0 = uf < Gs ? uf : total;
  const bool lm_head = (flags & 1) != 0;
  const bool argmax = lm_head && a.top_k == 1;
  const int gen1 = (gen + 1) % M3_GENS;
  float best_v = -INFINITY;
  int best_n = 0x7fffffff;
  if (u0 < total) {
    if (argmax) m3_banned(f, a.B, a.ngrams, a.n_ngrams, pos + 1, tid, u0, Gs);
    if (d[2] != 0) {                           // LayerNorm gamma / beta -> shared memory, in flight while the inputs are polled
      const float* ln_g = reinterpret_cast<const float*>(d[2]);
      const float* ln_b = reinterpret_cast<const float*>(d[3]);
      for (int k = tid * 4; k < K; k += M3_CTHREADS * 4) {
        m3_cp_async16(m3_ln(S) + k, ln_g + k);
        m3_cp_async16(m3_ln(S) + M3_KC + k, (ln_b != nullptr ? ln_b : ln_g) + k);
      }
      m3_cp_commit();
    }
    if (a.tc) m3_mbar_wait(&f->xsempty, (xcount & 1u) ^ 1u, a.error_flag);   // a tcgen05 stage may still be reading the rows
    m3_stage_x(a, d, S, gen, pos, u0, warp, lane, trace);
    m3_csync();                                // xs visible
    if (trace != nullptr) trace[1] = clock64();
    const float* bias = reinterpret_cast<const float*>(d[1]);
    const int act = (int)d[9], mode = (int)d[10];
    const int64_t ldo = d[15], cache_bs = d[16];
    const bool out_bf16 = (flags & 4) != 0;
    const int g = lane >> 2, qd = lane & 3;
    const __nv_bfloat16* xp = m3_xs(S) + g * S.xpitch + qd * 8;
    // epilogue role: lanes 0..7 of every warp; warps 0..3 -> first tile of a pair, 4..7 -> second; a thread owns 4 rows of
    // one batch row (16-byte loads of the partial tiles, 16-byte stores of the result)
    const int et = warp >> 2, rg = lane & 3, eb = (lane >> 2) + 2 * (warp & 3);
    const bool pairs = nkc == 1;
    const int ustep = pairs ? 2 * Gs : Gs;
    int pidx = 0;
    long long wait_cycles = 0;
#pragma unroll 1
    for (int u = u0; u < total; u += ustep, ++pidx) {
      const bool two = pairs && u + Gs < total;
      const int n0 = (u + et * Gs) * M3_ROWS, nq = n0 + 4 * rg;
      const bool e_on = lane < 8 && (et == 0 || two) && nq < N && eb < a.B;
      float4 e_bias = make_float4(0.f, 0.f, 0.f, 0.f);
      uint4 e_res = make_uint4(0u, 0u, 0u, 0u);
      const float* resp = nullptr;
      bool folded = false;                             // folded combine: the residual is residual' + bias' + partial row(s) of
      if (e_on) {                                      // another op, fetched into f->fold by cp.async (no registers held meanwhile)
        // epilogue operands do not depend on the MMAs: request them now
        if (bias != nullptr) e_bias = __ldg(reinterpret_cast<const float4*>(bias + nq));
        if (mode == 0 && d[23] != 0) {
          const int64_t* c = m3_cmb(S, (int)d[23] - 1);
          folded = true;
          resp = m3_gen<const float>(c[4], a.gen_stride, gen) + (int64_t)eb * ldo + nq;
          const float* pap = m3_gen<const float>(c[1], a.gen_stride, gen) + (int64_t)eb * ldo + nq;
          uint4* slot = f->fold[warp * 8 + lane];
          m3_cp_async16(slot, resp);
          m3_cp_async16(slot + 1, pap);
          if (c[0] > 1) m3_cp_async16(slot + 2, reinterpret_cast<const uint8_t*>(pap) + c[2]);
          m3_cp_commit();
          const float4 b2 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(c[3]) + nq));
          e_bias.x += b2.x; e_bias.y += b2.y; e_bias.z += b2.z; e_bias.w += b2.w;     // (no activation on a residual op)
        } else if (mode == 0 && d[6] != 0) {
          resp = m3_gen<const float>(d[6], a.gen_stride, gen) + (int64_t)eb * ldo + nq;
          e_res = m3_ld16(resp);
        }
      }
      int rs0 = 0, rs1 = 0;
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      // (three independent accumulator chains per tile were measured: no gain in the MMA phase, 12 bytes of spills, +20 us / step)
#pragma unroll 1
      for (int kc = 0; kc < nkc; ++kc) {
        const int nbc = m3_chunk_nb(K, kc);
        const int s0 = R.slot;
        const long long tw0 = trace != nullptr ? clock64() : 0;
        m3_mbar_wait(&f->full[s0], R.phase, a.error_flag);
        R.advance(a.nslots);
        const int s1 = R.slot;
        if (two) {
          m3_mbar_wait(&f->full[s1], R.phase, a.error_flag);
          R.advance(a.nslots);
        }
        if (trace != nullptr) {
          const long long tw1 = clock64();
          if (pidx == 0 && kc == 0) trace[3] = tw1;
          else wait_cycles += tw1 - tw0;                // time parked on the weight ring after the first tile (LM head)
        }
        const uint4* w0p = m3_ring(s0) + tid;
        const uint4* w1p = m3_ring(s1) + tid;
        const __nv_bfloat16* xk = xp + kc * M3_KC + warp * M3_BLK;
#pragma unroll
        for (int i = 0; i < M3_NB; ++i) {
          if (i < nbc) {
            const uint4 xf = *reinterpret_cast<const uint4*>(xk + i * M3_CWARPS * M3_BLK);
            const uint4 w0 = w0p[(2 * i) * M3_CTHREADS], w1 = w0p[(2 * i + 1) * M3_CTHREADS];
            m3_mma(acc[0], w0.x, w1.x, w0.y, w1.y, xf.x, xf.y);
            m3_mma(acc[0], w0.z, w1.z, w0.w, w1.w, xf.z, xf.w);
            if (two) {
              const uint4 v0 = w1p[(2 * i) * M3_CTHREADS], v1 = w1p[(2 * i + 1) * M3_CTHREADS];
              m3_mma(acc[1], v0.x, v1.x, v0.y, v1.y, xf.x, xf.y);
              m3_mma(acc[1], v0.z, v1.z, v0.w, v1.w, xf.z, xf.w);
            }
          }
        }
        // hand the slot(s) back to the producer once all 8 warps are done with them: warps 0..3 arrive (the barrier counts 4)
        if (nkc > 1) m3_csync();                       // (single-chunk tiles: the barrier in front of the epilogue does it)
        if (nkc > 1 && lane == 0 && warp < 4) {
          m3_mbar_arrive(&f->empty[s0]);
          if (two) m3_mbar_arrive(&f->empty[s1]);
        }
        rs0 = s0; rs1 = s1;
      }
      if (trace != nullptr && pidx == 0) trace[4] = clock64() + (long long)(acc[0][0] == 123.f);
      // acc: D[g][2qd], D[g][2qd+1], D[g+8][2qd], D[g+8][2qd+1]  ->  red[tile][batch][row]
      if (pidx > 0) m3_csync();                          // the previous pair's epilogue is done with the partial tiles
      float* red = &f->red[warp][0];
      red[(2 * qd) * 20 + g] = acc[0][0];
      red[(2 * qd + 1) * 20 + g] = acc[0][1];
      red[(2 * qd) * 20 + g + 8] = acc[0][2];
      red[(2 * qd + 1) * 20 + g + 8] = acc[0][3];
      if (two) {
        red[M3_RED_T + (2 * qd) * 20 + g] = acc[1][0];
        red[M3_RED_T + (2 * qd + 1) * 20 + g] = acc[1][1];
        red[M3_RED_T + (2 * qd) * 20 + g + 8] = acc[1][2];
        red[M3_RED_T + (2 * qd + 1) * 20 + g + 8] = acc[1][3];
      }
      m3_csync();
      if (nkc == 1 && lane == 0 && warp < 4) {         // every warp has read the slot(s): 4 arrivals free them
        m3_mbar_arrive(&f->empty[rs0]);
        if (two) m3_mbar_arrive(&f->empty[rs1]);
      }
      if (trace != nullptr && pidx == 0) trace[5] = clock64();
      if (e_on) {
        const float* rp = &f->red[0][et * M3_RED_T + eb * 20 + 4 * rg];
        float4 v = *reinterpret_cast<const float4*>(rp);
#pragma unroll
        for (int i = 1; i < M3_CWARPS; ++i) {
          const float4 p = *reinterpret_cast<const float4*>(rp + i * 2 * M3_RED_T);
          v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
        v.x = m3_act(v.x + e_bias.x, act); v.y = m3_act(v.y + e_bias.y, act);
        v.z = m3_act(v.z + e_bias.z, act); v.w = m3_act(v.w + e_bias.w, act);
        if (resp != nullptr) {
          uint32_t spins = 0;
          if (folded) {
            m3_cp_wait0();
            e_res = f->fold[warp * 8 + lane][0];
          }
          while (!m3_ok32(e_res)) {
            if (m3_giveup(spins, a.error_flag)) break;
            e_res = m3_ld16(resp);
          }
          v.x += __uint_as_float(e_res.x); v.y += __uint_as_float(e_res.y);
          v.z += __uint_as_float(e_res.z); v.w += __uint_as_float(e_res.w);
          if (folded) {
            const int64_t* c = m3_cmb(S, (int)d[23] - 1);
            const float* pap = m3_gen<const float>(c[1], a.gen_stride, gen) + (int64_t)eb * ldo + nq;
            const int np = (int)c[0];                  // 1 or 2 partial rows
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (j >= np) break;
              const float* pp = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(pap) + j * c[2]);
              uint4 e = f->fold[warp * 8 + lane][1 + j];
              while (!m3_ok32(e)) {
                if (m3_giveup(spins, a.error_flag)) break;
                e = m3_ld16(pp);
              }
              v.x += __uint_as_float(e.x); v.y += __uint_as_float(e.y); v.z += __uint_as_float(e.z); v.w += __uint_as_float(e.w);
            }
          }
        }
        if (lm_head) {
          const float vv[4] = {v.x, v.y, v.z, v.w};
          if (argmax) {
            const uint32_t bw = f->banmask[(2 * pidx + et) * 4 + rg] >> eb;      // byte j = row nq + j, bit 0 = this sequence
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = nq + j;
              const bool ban = n >= N || ((bw >> (8 * j)) & 1u) != 0u;
              if (!ban && (vv[j] > best_v || best_n == 0x7fffffff)) { best_v = vv[j]; best_n = n; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (nq + j < N) a.logits[(int64_t)eb * a.ldl + nq + j] = vv[j];
          }
        } else if (mode == 1 && n0 >= a.C) {   // k / v rows of the packed q|k|v output: appended at `pos` (poisoned by the host)
          const int seg = n0 >= 2 * a.C ? 2 : 1, nl = nq - seg * a.C;
          __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(seg == 1 ? d[11] : d[12]);
          m3_st8(base + (int64_t)eb * cache_bs + (int64_t)pos * a.C + nl, make_uint2(m3_pack(v.x, v.y), m3_pack(v.z, v.w)));
        } else if (out_bf16) {
          const int64_t off = ((int64_t)eb * ldo + nq) * 2;
          m3_st8(m3_gen<uint8_t>(d[5], a.gen_stride, gen1) + off, make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu));
          m3_st8(m3_gen<uint8_t>(d[5], a.gen_stride, gen) + off, make_uint2(m3_pack(v.x, v.y), m3_pack(v.z, v.w)));
        } else {
          const int64_t off = ((int64_t)eb * ldo + nq) * 4;
          m3_st16(m3_gen<uint8_t>(d[5], a.gen_stride, gen1) + off, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
          m3_st16(m3_gen<uint8_t>(d[5], a.gen_stride, gen) + off, m3_f4_bits(v));
        }
      }
    }
    if (trace != nullptr && lm_head) trace[7] = wait_cycles;
  }
  if (argmax) {
    // CTA-level arg-max per sequence: the 4 row groups of a warp, then the two tiles (warps w, w + 4), then this CTA's key
    unsigned long long key = 0ull;
    if (best_n != 0x7fffffff)
      key = ((unsigned long long)float_key(best_v) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)best_n);
    const unsigned long long k1 = __shfl_xor_sync(0xffffffffu, key, 1);
    key = k1 > key ? k1 : key;
    const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, 2);
    key = k2 > key ? k2 : key;
    if (lane == 0 || lane == 4) f->best[warp][lane >> 2] = key;
    m3_csync();
    if (tid < M3_B) {
      const unsigned long long ka = f->best[tid >> 1][tid & 1], kb = f->best[(tid >> 1) + 4][tid & 1];
      unsigned long long k = ka > kb ? ka : kb;
      if (k == 0ull) k = 1ull;                          // "nothing from this CTA": still not the poison value
      const int64_t o = (int64_t)blockIdx.x * M3_B + tid;
      m3_st8(a.ctakeys + (int64_t)((keyslot + 1) % M3_GENS) * G * M3_B + o, 0ull);
      m3_st8(a.ctakeys + (int64_t)keyslot * G * M3_B + o, k);
    }
  }
}


// =====================================================================================================================
// tcgen05 path of a linear stage.  The weights of a tile (16 rows) sit in a ring slot as K-major, 128B-swizzled atoms
// ([k / 64][row][128 B]); the staged activations likewise ([k / 64][batch row][128 B]).  A single tcgen05.mma M128 N16 K16
// takes ~4-8 tensor-core cycles but tens of issue cycles, so FOUR consumer warps (1, 2, 3, 5) each issue the MMAs of every
// fourth 64-wide K block into their own TMEM accumulator (4 partials x 16 columns per tile, 2 tiles in flight); rows
// 16..127 of the A operand and columns 8..15 of the B operand read whatever follows in shared memory -- their results land
// in TMEM lanes / columns nobody reads.  tcgen05.commit hands the ring slot back to the weight producer and wakes the
// epilogue warps 0 and 4 (the warps that may read TMEM lanes 0..31): lane r owns row r of the tile, warp 0 the sequences
// 0..3, warp 4 the sequences 4..7; they add the four partials in a fixed order.  No partial tiles through shared memory.
// =====================================================================================================================
struct M3Tc {
  uint32_t tcount;      // tiles this CTA has pushed through TMEM so far (buffer = tcount % 2)
  uint32_t xcount;      // tcgen05 linear stages this CTA took part in so far (parity of xsempty)
};
constexpr int M3_TC_ISSUERS = 4;
constexpr uint32_t M3_TC_COLS = 128;          // 2 tile buffers x 4 partial accumulators x 16 columns

__device__ __forceinline__ void m3_tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ bool m3_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0u;
}

__device__ __forceinline__ void m3_linear_tc(const M3Args& a, int op, const M3Sm& S, M3Ring& R, M3Tc& T, int gen, int pos, int keyslot,
                                             int tid, long long* trace) {
  const int lane = tid & 31, warp = tid >> 5;
  M3Fixed* f = m3_f(S);
  const int64_t* d = m3_lin(S, op);
  const int N = (int)d[7], K = (int)d[8];
  const int flags = (int)d[17];
  const int total = (N + M3_ROWS - 1) / M3_ROWS;
  const int nkc = m3_nkc(K);
  const int G = (int)gridDim.x;
  const int Gs = d[21] != 0 ? (int)d[21] : G;
  const int uf = m3_first_unit((int)d[18]);
  const int u0 = uf < Gs ? uf : total;
  const int gen1 = (gen + 1) % M3_GENS;
  if (u0 >= total) return;
  if (d[2] != 0) {                           // LayerNorm gamma / beta -> shared memory, in flight while the inputs are polled
    const float* ln_g = reinterpret_cast<const float*>(d[2]);
    const float* ln_b = reinterpret_cast<const float*>(d[3]);
    for (int k = tid * 4; k < K; k += M3_CTHREADS * 4) {
      m3_cp_async16(m3_ln(S) + k, ln_g + k);
      m3_cp_async16(m3_ln(S) + M3_KC + k, (ln_b != nullptr ? ln_b : ln_g) + k);
    }
    m3_cp_commit();
  }
  m3_mbar_wait(&f->xsempty, (T.xcount & 1u) ^ 1u, a.error_flag);   // the MMAs of the previous stage are done with the rows
  m3_stage_x(a, d, S, gen, pos, u0, warp, lane, trace, true);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
  m3_csync();
  ++T.xcount;
  if (trace != nullptr) trace[1] = clock64();
  // roles: issuer q (warps 1, 2, 3, 5) takes the K blocks q, q + 4, ...; warps 0 / 4 the epilogue of sequences 0..3 / 4..7
  const int q = warp == 5 ? 3 : warp - 1;
  const bool issuer = warp == 1 || warp == 2 || warp == 3 || warp == 5;
  const bool epi = (warp & 3) == 0;
  const uint32_t tmem_base = f->tmem_base;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);   // f32 += bf16 x bf16, N 16, M 128
  const uint32_t xs_addr = smem_u32(m3_xs(S));
  const float* bias = reinterpret_cast<const float*>(d[1]);
  const int act = (int)d[9], mode = (int)d[10];
  const int64_t ldo = d[15], cache_bs = d[16];
  const bool out_bf16 = (flags & 4) != 0;
  const bool has_res = mode == 0 && d[6] != 0;
  const int b0 = (warp >> 2) * 4;                                  // first sequence of this epilogue warp
  int ti = 0;
#pragma unroll 1
  for (int u = u0; u < total; u += Gs, ++ti) {
    const uint32_t tc = T.tcount + (uint32_t)ti, buf = tc & 1u;
    if (issuer) {
      m3_mbar_wait(&f->tempty[buf], ((tc >> 1) & 1u) ^ 1u, a.error_flag);       // both epilogue warps have drained this buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int n = u * M3_ROWS + lane;
    const bool on = epi && lane < M3_ROWS && n < N;
    float e_bias = 0.f;
    uint32_t e_res[4] = {0u, 0u, 0u, 0u};
    const float* resp = nullptr;
    if (on) {                                        // epilogue operands do not depend on the MMAs: request them now
      if (bias != nullptr) e_bias = __ldg(bias + n);
      if (has_res) {
        resp = m3_gen<const float>(d[6], a.gen_stride, gen) + (int64_t)b0 * ldo + n;
#pragma unroll
        for (int j = 0; j < 4; ++j) e_res[j] = b0 + j < a.B ? m3_ld4(resp + (int64_t)j * ldo) : 0u;
      }
    }
    uint32_t accf = 0u;
#pragma unroll 1
    for (int kc = 0; kc < nkc; ++kc) {
      const int npq = m3_chunk_nb(K, kc);                          // 64-wide K blocks per issuer in this chunk (1..3)
      const int slot = R.slot;
      if (issuer) {
        m3_mbar_wait(&f->full[slot], R.phase, a.error_flag);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (m3_elect_one()) {
          const uint32_t tmem_d = tmem_base + buf * 64u + (uint32_t)q * 16u;
          const uint64_t adesc = umma_desc_sw128(smem_u32(m3_ring(slot))) + (uint64_t)(q * (2048 >> 4));
          const uint64_t bdesc = umma_desc_sw128(xs_addr + (uint32_t)kc * (M3_KC / 64) * 1024u) + (uint64_t)(q * (1024 >> 4));
#pragma unroll
          for (int i = 0; i < M3_NB; ++i) {
            if (i < npq) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {                         // 16 values of K = 32 bytes inside the 128-byte swizzle span
                umma_bf16(tmem_d, adesc + (uint64_t)(i * 4 * (2048 >> 4) + 2 * j), bdesc + (uint64_t)(i * 4 * (1024 >> 4) + 2 * j),
                          idesc, (accf | (uint32_t)(i | j)) != 0u ? 1u : 0u);
              }
            }
          }
          umma_commit(&f->empty[slot]);                             // one of 4 arrivals: the slot goes back to the weight producer
        }
        __syncwarp();
        accf = 1u;
      }
      R.advance(a.nslots);
    }
    if (issuer) {
      if (m3_elect_one()) umma_commit(&f->tfull[buf]);              // one of 4 arrivals: the partial accumulators are complete
      __syncwarp();
    }
    if (epi) {
      m3_mbar_wait(&f->tfull[buf], (tc >> 1) & 1u, a.error_flag);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r0[4], r1[4], r2[4], r3[4];
      const uint32_t taddr = tmem_base + buf * 64u + (uint32_t)b0;   // lane r of the warp = TMEM lane r = row r of the tile
      m3_tmem_ld4(taddr, r0);
      m3_tmem_ld4(taddr + 16u, r1);
      m3_tmem_ld4(taddr + 32u, r2);
      m3_tmem_ld4(taddr + 48u, r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) m3_mbar_arrive(&f->tempty[buf]);
      if (trace != nullptr && ti == 0) trace[5] = clock64();
      if (on) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[j] = m3_act(((__uint_as_float(r0[j]) + __uint_as_float(r1[j])) + (__uint_as_float(r2[j]) + __uint_as_float(r3[j]))) + e_bias, act);
        if (has_res) {
          uint32_t spins = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (b0 + j < a.B) {
              while (e_res[j] == 0xFFFFFFFFu) {
                if (m3_giveup(spins, a.error_flag)) break;
                e_res[j] = m3_ld4(resp + (int64_t)j * ldo);
              }
              v[j] += __uint_as_float(e_res[j]);
            }
        }
        if (mode == 1 && n >= a.C) {           // k / v rows of the packed q|k|v output: appended at `pos` (poisoned by the host)
          const int seg = n >= 2 * a.C ? 2 : 1, nl = n - seg * a.C;
          __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(seg == 1 ? d[11] : d[12]);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (b0 + j < a.B)
              m3_st2(base + (int64_t)(b0 + j) * cache_bs + (int64_t)pos * a.C + nl, __bfloat16_as_ushort(__float2bfloat16_rn(v[j])));
        } else if (out_bf16) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (b0 + j < a.B) {
              const int64_t off = ((int64_t)(b0 + j) * ldo + n) * 2;
              m3_st2(m3_gen<uint8_t>(d[5], a.gen_stride, gen1) + off, (unsigned short)0xFFFFu);
              m3_st2(m3_gen<uint8_t>(d[5], a.gen_stride, gen) + off, __bfloat16_as_ushort(__float2bfloat16_rn(v[j])));
            }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (b0 + j < a.B) {
              const int64_t off = ((int64_t)(b0 + j) * ldo + n) * 4;
              m3_st4(m3_gen<uint8_t>(d[5], a.gen_stride, gen1) + off, 0xFFFFFFFFu);
              m3_st4(m3_gen<uint8_t>(d[5], a.gen_stride, gen) + off, __float_as_uint(v[j]));
            }
        }
      }
    }
  }
  if (issuer) {
    if (m3_elect_one()) umma_commit(&f->xsempty);                   // one of 4 arrivals: every MMA of the stage has read the rows
    __syncwarp();
  }
  T.tcount += (uint32_t)ti;
}

// ---- single-query attention for one (batch, head) per CTA: one key per thread for q.k (K row in registers),
//      V rows staged in shared memory (aliasing the idle activation rows) for P.V ----
template <int HS>
__device__ __forceinline__ void m3_attention(const M3Args& a, int ai, const M3Sm& S, int gen, int pos, int tid, uint32_t xcount = 0) {
  constexpr int NV = HS / 8, DPL = HS / 32;          // 16-byte vectors per row; dims per lane in P.V
  const int lane = tid & 31, warp = tid >> 5;
  M3Fixed* f = m3_f(S);
  const int64_t* d = m3_att(S, ai);
  const int unit = m3_first_unit((int)d[8]);
  if (unit >= a.B * a.H) return;
  const bool self = d[4] == 0;
  const int len = self ? pos + 1 : (int)d[5];
  const float scale = 1.0f / sqrtf((float)HS);
  const int b = unit / a.H, h = unit - b * a.H;
  const bool on = tid < len;
  const bool fresh = self && tid == pos;             // the row the QKV stage of THIS step appends
  if (a.tc) m3_mbar_wait(&f->xsempty, (xcount & 1u) ^ 1u, a.error_flag);   // the V rows alias the rows the tensor core may still read
  uint4 kr[NV];
  uint4* vs = reinterpret_cast<uint4*>(m3_xs(S));
  const __nv_bfloat16* kp = reinterpret_cast<const __nv_bfloat16*>(d[0]) + b * d[2] + (int64_t)tid * d[3] + h * HS;
  const __nv_bfloat16* vp = reinterpret_cast<const __nv_bfloat16*>(d[1]) + b * d[2] + (int64_t)tid * d[3] + h * HS;
  if (on && !fresh) {                                // rows of earlier steps are final: request them before polling
#pragma unroll
    for (int i = 0; i < NV; ++i) kr[i] = __ldcg(reinterpret_cast<const uint4*>(kp) + i);
#pragma unroll
    for (int i = 0; i < NV; ++i) m3_cp_async16(vs + (size_t)tid * NV + i, reinterpret_cast<const uint4*>(vp) + i);
  }
  m3_cp_commit();
  m3_csync();                                // the scratch below aliases the partial tiles a linear stage may still be reading
  uint32_t spins = 0;
  if (tid < HS) {
    const float* qp = m3_gen<const float>(d[6], a.gen_stride, gen) + (int64_t)b * a.C + h * HS + tid;
    uint32_t r = m3_ld4(qp);
    while (r == 0xFFFFFFFFu) {
      if (m3_giveup(spins, a.error_flag)) break;
      r = m3_ld4(qp);
    }
    f->att.att_q[tid] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r))) * scale;      // autocast: SDPA sees a bf16 query
  }
  if (self && warp == M3_CWARPS - 1 && lane < 2 * NV) {
    // the K / V row the QKV stage of THIS step appends: 2 * NV vectors polled in parallel by the last warp
    const bool isv = lane >= NV;
    const int i = isv ? lane - NV : lane;
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(isv ? d[1] : d[0]) + b * d[2] +
                                                      (int64_t)pos * d[3] + h * HS) + i;
    uint4 vv = m3_ld16(src);
    while (!m3_ok16(vv)) {
      if (m3_giveup(spins, a.error_flag)) break;
      vv = m3_ld16(src);
    }
    if (isv) vs[(size_t)pos * NV + i] = vv;
    else f->att.kfresh[i] = vv;
  }
  m3_cp_wait0();
  m3_csync();                                // q and every K / V row visible
  if (fresh) {
#pragma unroll
    for (int i = 0; i < NV; ++i) kr[i] = f->att.kfresh[i];
  }
  float s = -INFINITY;
  if (on) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 q0 = *reinterpret_cast<const float4*>(&f->att.att_q[i * 8]);
      const float4 q1 = *reinterpret_cast<const float4*>(&f->att.att_q[i * 8 + 4]);
      acc = fmaf(__uint_as_float(kr[i].x << 16), q0.x, acc); acc = fmaf(__uint_as_float(kr[i].x & 0xffff0000u), q0.y, acc);
      acc = fmaf(__uint_as_float(kr[i].y << 16), q0.z, acc); acc = fmaf(__uint_as_float(kr[i].y & 0xffff0000u), q0.w, acc);
      acc = fmaf(__uint_as_float(kr[i].z << 16), q1.x, acc); acc = fmaf(__uint_as_float(kr[i].z & 0xffff0000u), q1.y, acc);
      acc = fmaf(__uint_as_float(kr[i].w << 16), q1.z, acc); acc = fmaf(__uint_as_float(kr[i].w & 0xffff0000u), q1.w, acc);
    }
    s = acc;
  }
  float mx = warp_max(s);
  if (lane == 0) f->att.att_red[warp] = mx;
  m3_csync();
  mx = f->att.att_red[0];
#pragma unroll
  for (int i = 1; i < M3_CWARPS; ++i) mx = fmaxf(mx, f->att.att_red[i]);
  f->att.att_p[tid] = on ? expf(s - mx) : 0.f;
  m3_csync();
  // P.V: warp w takes keys w, w+8, ...; lane l owns dims [l * DPL, (l + 1) * DPL); the softmax sum rides along
  const __nv_bfloat16* vsh = reinterpret_cast<const __nv_bfloat16*>(vs);
  float o0 = 0.f, o1 = 0.f, psum = 0.f;
  for (int j = warp; j < len; j += M3_CWARPS) {
    const float p = f->att.att_p[j];
    psum += p;
    if (DPL == 2) {
      const uint32_t r = *reinterpret_cast<const uint32_t*>(vsh + (size_t)j * HS + lane * 2);
      o0 = fmaf(p, __uint_as_float(r << 16), o0);
      o1 = fmaf(p, __uint_as_float(r & 0xffff0000u), o1);
    } else {
      o0 = fmaf(p, __bfloat162float(vsh[(size_t)j * HS + lane]), o0);
    }
  }
  if (DPL == 2) *reinterpret_cast<float2*>(&f->att.att_o[warp][lane * 2]) = make_float2(o0, o1);
  else f->att.att_o[warp][lane] = o0;
  if (lane == 0) f->att.att_red[warp] = psum;
  m3_csync();
  if (tid < HS) {
    float tot = 0.f, ov = 0.f;
#pragma unroll
    for (int i = 0; i < M3_CWARPS; ++i) { tot += f->att.att_red[i]; ov += f->att.att_o[i][tid]; }
    const int64_t off = (int64_t)b * a.C + h * HS + tid;
    m3_st2(m3_gen<__nv_bfloat16>(d[7], a.gen_stride, (gen + 1) % M3_GENS) + off, (unsigned short)0xFFFFu);
    m3_st2(m3_gen<__nv_bfloat16>(d[7], a.gen_stride, gen) + off, __bfloat16_as_ushort(__float2bfloat16_rn(tot > 0.f ? ov / tot : 0.f)));
  }
}

// ---- combine: the second half of a K-split projection.  out[b][n] = bias[n] + residual[b][n] + sum_j partial_j[b][n];
//      a unit = 256 float4 outputs (one per thread); every operand is polled (5 loads in flight), fixed summation order ----
__device__ __forceinline__ void m3_combine(const M3Args& a, int ci, const M3Sm& S, int gen, int tid) {
  const int64_t* d = m3_cmb(S, ci);
  const int N = (int)d[6];
  const int nvec_row = N / 4, total = (M3_B * nvec_row + M3_CTHREADS - 1) / M3_CTHREADS;
  const int G = (int)gridDim.x;
  const int np = (int)d[0];
  for (int u = m3_first_unit((int)d[7]); u < total; u += G) {
    const int vi = u * M3_CTHREADS + tid;
    int b = 0, rem = vi;
#pragma unroll
    for (int j = 0; j < M3_B; ++j)
      if (rem >= nvec_row) { rem -= nvec_row; ++b; }
    const int n = rem * 4;
    if (rem >= nvec_row) continue;
    if (b >= a.B) continue;
    const int64_t off = ((int64_t)b * N + n) * 4;
    const uint8_t* p0 = m3_gen<const uint8_t>(d[1], a.gen_stride, gen) + off;
    const uint8_t* rp = m3_gen<const uint8_t>(d[4], a.gen_stride, gen) + off;
    float4 acc = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d[3]) + n));
    uint4 r = m3_ld16(rp);
    uint4 pv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < np) pv[j] = m3_ld16(p0 + (int64_t)j * d[2]);
    uint32_t spins = 0;
    while (!m3_ok32(r)) {
      if (m3_giveup(spins, a.error_flag)) break;
      r = m3_ld16(rp);
    }
    acc.x += __uint_as_float(r.x); acc.y += __uint_as_float(r.y); acc.z += __uint_as_float(r.z); acc.w += __uint_as_float(r.w);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < np) {
        while (!m3_ok32(pv[j])) {
          if (m3_giveup(spins, a.error_flag)) break;
          pv[j] = m3_ld16(p0 + (int64_t)j * d[2]);
        }
        acc.x += __uint_as_float(pv[j].x); acc.y += __uint_as_float(pv[j].y);
        acc.z += __uint_as_float(pv[j].z); acc.w += __uint_as_float(pv[j].w);
      }
    m3_st16(m3_gen<uint8_t>(d[5], a.gen_stride, (gen + 1) % M3_GENS) + off, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
    m3_st16(m3_gen<uint8_t>(d[5], a.gen_stride, gen) + off, m3_f4_bits(acc));
  }
}

// general sampler on B CTAs, vocabulary row processed in place (global / L2).  Not inlined: its register needs (double
// precision prefix sums) and code size must not shape the hot loop.
__device__ __noinline__ void m3_sample_stage(SampleScratch* scratch, float* logits, int64_t ldl, int V, int B, int64_t* ids, int64_t ids_ld,
                                             float temperature, int top_k, const int32_t* ngrams, int n_ngrams,
                                             const uint64_t* seed_ptr, int pos, int tid) {
  if ((int)blockIdx.x >= B) return;
  const int b = blockIdx.x;
  float* row = logits + (int64_t)b * ldl;
  SampleScratch& samp = *scratch;
  const int choice = sample_row_smem(row, samp, row, V, ids + (int64_t)b * ids_ld, pos + 1, temperature, top_k, ngrams,
                                     n_ngrams, *seed_ptr, b, nullptr, tid, M3_CTHREADS);
  if (tid == 0) {
    __threadfence();
    *reinterpret_cast<volatile int64_t*>(ids + (int64_t)b * ids_ld + pos + 1) = (int64_t)choice;
  }
}

// ---- the tokens this step embeds (warp b = sequence b) ----
__device__ __forceinline__ void m3_tokens(const M3Args& a, const M3Sm& S, int pos, bool from_keys, int keyslot, bool poll_ids,
                                          int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  M3Fixed* f = m3_f(S);
  const int G = (int)gridDim.x;
  if (warp < a.B) {
    int tok;
    uint32_t spins = 0;
    if (from_keys) {                 // greedy: max over the per-CTA keys of the previous step's LM head
      unsigned long long best = 0ull;
      const unsigned long long* kp = a.ctakeys + (int64_t)keyslot * G * M3_B + warp;
      unsigned long long kk[8];                       // grid <= 256 CTAs: up to 8 keys per lane, all requested before any is checked
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        kk[j] = c < G ? m3_ld8(kp + (int64_t)c * M3_B) : 1ull;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        while (kk[j] == 0ull) {
          if (m3_giveup(spins, a.error_flag)) break;
          if (a.sleep_ns > 0) __nanosleep(a.sleep_ns);
          kk[j] = m3_ld8(kp + (int64_t)c * M3_B);
        }
        best = kk[j] > best ? kk[j] : best;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long k = __shfl_xor_sync(0xffffffffu, best, o);
        best = k > best ? k : best;
      }
      tok = (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull));
      if (best <= 1ull) tok = 0;     // every token banned: cannot happen with a finite ban list; keep the index in range
    } else {
      const int64_t* ip = a.ids + (int64_t)warp * a.ids_ld + pos;
      long long t = (long long)m3_ld8(ip);
      while (poll_ids && t < 0) {
        if (m3_giveup(spins, a.error_flag)) break;
        if (a.sleep_ns > 0) __nanosleep(a.sleep_ns);
        t = (long long)m3_ld8(ip);
      }
      tok = (int)(t < 0 ? 0 : t);
    }
    if (lane == 0) {
      f->tok[warp] = tok;
      f->hist[warp][pos] = tok;
      if (from_keys && blockIdx.x == 0) a.ids[(int64_t)warp * a.ids_ld + pos] = (int64_t)tok;   // publish the previous pick
    }
  }
  m3_csync();
}

template <int HS, bool TC>
__global__ void __launch_bounds__(M3_THREADS, 1) decode_mega3_kernel(M3Args a_in) {
  // TC = false: no tcgen05 code in the kernel at all (the default; the experimental tcgen05 instantiation costs the hot loop
  // registers and instruction-cache footprint: 281 -> 292 us per step when both paths shared one kernel)
  M3Args a = a_in;
  a.tc = TC ? 1 : 0;
  const int tid = threadIdx.x;
  // ---- lay out dynamic shared memory, copy the tables, arm the ring ----
  const M3Sm S = m3_layout(a.nslots, a.max_k, a.max_len, HS, a.n_ops, a.n_att, a.n_cmb, a.n_sched, a.top_k != 1);
  {
    int64_t* lin = reinterpret_cast<int64_t*>(m3_smem + S.lin_off);
    int64_t* att = reinterpret_cast<int64_t*>(m3_smem + S.att_off);
    int32_t* sc = reinterpret_cast<int32_t*>(m3_smem + S.sched_off);
    int64_t* cmb = reinterpret_cast<int64_t*>(m3_smem + S.cmb_off);
    for (int i = tid; i < a.n_cmb * M3_CMB_FIELDS; i += M3_THREADS) cmb[i] = a.cmb[i];
    for (int i = tid; i < a.n_ops * M3_LIN_FIELDS; i += M3_THREADS) lin[i] = a.lin[i];
    for (int i = tid; i < a.n_att * M3_ATT_FIELDS; i += M3_THREADS) att[i] = a.att[i];
    for (int i = tid; i < a.n_sched * 4; i += M3_THREADS) sc[i] = a.sched[i];
  }
  M3Fixed* f = m3_f(S);
  const int pos0 = *a.pos;
  if (tid == 0) {
    for (int i = 0; i < a.nslots; ++i) {
      mbar_init(&f->full[i], 1);
      mbar_init(&f->empty[i], 4);            // 4 arrivals free a slot: the 4 issuing warps' tcgen05.commit, or warps 0..3 (mma.sync path)
    }
    for (int i = 0; i < M3_TBUFS; ++i) {
      mbar_init(&f->tfull[i], M3_TC_ISSUERS);
      mbar_init(&f->tempty[i], 2);
    }
    mbar_init(&f->xsempty, M3_TC_ISSUERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (TC && (tid >> 5) == M3_CWARPS + 1) {   // warp 9 owns the tensor memory: 4 accumulator buffers of 16 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&f->tmem_base)), "r"(M3_TC_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  // token history before the first position this launch processes
  for (int w = tid; w < pos0 * a.B; w += M3_THREADS) {
    const int b = w / pos0, i = w - b * pos0;
    f->hist[b][i] = (int)a.ids[(int64_t)b * a.ids_ld + i];
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid >= M3_CTHREADS + 32) {             // warp 9: one thread issues every tcgen05.mma; the warp frees the tensor memory at the end
    if (!TC) return;
    asm volatile("bar.sync 3, %0;" ::"n"(M3_CTHREADS + 32) : "memory");      // every epilogue has read its accumulators
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(f->tmem_base), "r"(M3_TC_COLS) : "memory");
    return;
  }
  if (tid >= M3_CTHREADS) {                  // warp 8: one thread streams the weights, then the warp retires
    if (tid == M3_CTHREADS) m3_producer(a, S);
    return;
  }
  M3Tc T{0u, 0u};
  const int32_t* sched = m3_sched(S);
  M3Ring R{0, 0u};
  unsigned int epoch = 0;
  const int n_steps = a.n_prefill + a.n_sample;
  const bool greedy = a.top_k == 1;
#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
    const bool sampling = step >= a.n_prefill;
    const int pos = pos0 + step;
    const int sstep = step - a.n_prefill;                       // index among the sampled steps
    const int gen = step % M3_GENS;
    const bool trace_step = a.trace != nullptr && step == n_steps - 1 && (int)blockIdx.x == a.trace_cta && tid == 0;
    // token source: the previous sampled step's arg-max keys (greedy) or the ids row the sampler CTAs publish, else the prompt
    if (trace_step) a.trace[(a.n_sched - 1) * M3_TRACE + 3] = clock64();       // (the sample entry's slot: step begin / tokens known)
    m3_tokens(a, S, pos, greedy && sstep > 0, (sstep + M3_GENS - 1) % M3_GENS, !greedy && sstep > 0, tid);
    if (trace_step) a.trace[(a.n_sched - 1) * M3_TRACE + 4] = clock64();
#pragma unroll 1
    for (int s = 0; s < a.n_sched; ++s) {
      const int kind = sched[s * 4], idx = sched[s * 4 + 1];
      long long* trace = trace_step ? a.trace + s * M3_TRACE : nullptr;
      if (trace_step) trace[0] = clock64();
      if (kind == 0) {
        if (!sampling && ((int)m3_lin(S, idx)[17] & 1) != 0) continue;
        if (TC && m3_lin(S, idx)[22] != 0) m3_linear_tc(a, idx, S, R, T, gen, pos, sstep >= 0 ? sstep % M3_GENS : 0, tid, trace);
        else m3_linear(a, idx, S, R, gen, pos, sstep >= 0 ? sstep % M3_GENS : 0, tid, trace, T.xcount);
      } else if (kind == 1) {
        m3_attention<HS>(a, idx, S, gen, pos, tid, T.xcount);
      } else if (kind == 3) {
        m3_combine(a, idx, S, gen, tid);
      } else if (kind == 2 && sampling && !greedy) {
        m3_grid_sync(a.bar, epoch, a.error_flag, tid);          // every logit of this step is in L2
        m3_sample_stage(reinterpret_cast<SampleScratch*>(m3_smem + S.samp_off), a.logits, a.ldl, a.V, a.B, a.ids, a.ids_ld, a.temperature, a.top_k, a.ngrams, a.n_ngrams, a.seed_ptr,
                        pos, tid);
      }
      if (trace_step) trace[2] = clock64();
    }
  }
  // final bookkeeping by CTA 0: the last pick -> ids, the position counter
  const int pos_end = pos0 + n_steps;
  if (blockIdx.x == 0) {
    if (a.n_sample > 0) m3_tokens(a, S, pos_end, greedy, (a.n_sample - 1) % M3_GENS, !greedy, tid);
    if (tid == 0) *a.pos = pos_end;
  }
  if (TC) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 3, %0;" ::"n"(M3_CTHREADS + 32) : "memory");
  }
}

// ---- weight re-pack: one op's [N][K] bf16 matrix into the per-CTA streams (fragment order, zero padded) ----
// grid.x = tiles of 16 rows; a tile's chunks are contiguous at tile_off[tile]; within a chunk the 16-byte vector
// (2 * i + half) * 256 + t holds row tile * 16 + g + 8 * half, k = chunk * 768 + (warp + 8 * i) * 32 + qd * 8 .. + 8
// (t = warp * 32 + lane, g = lane / 4, qd = lane % 4) -- exactly what consumer thread t feeds to its two MMAs.
__global__ void __launch_bounds__(M3_CTHREADS) decode_mega3_pack_kernel(const __nv_bfloat16* __restrict__ W, int N, int K, int64_t ldw,
                                                                        uint8_t* __restrict__ dst, const int64_t* __restrict__ tile_off,
                                                                        int tc_layout) {
  const int tile = blockIdx.x, t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5, g = lane >> 2, qd = lane & 3;
  uint4* out = reinterpret_cast<uint4*>(dst + tile_off[tile]);
  const int nkc = (K + M3_KC - 1) / M3_KC;
  if (tc_layout) {
    // tcgen05 path: a chunk = [k / 64][row 0..15][128 bytes], the 16-byte chunk c of a row stored at c ^ (row % 8)
    // (the canonical K-major SWIZZLE_128B atom the UMMA shared-memory descriptor of tc_common.cuh describes)
    for (int kc = 0; kc < nkc; ++kc) {
      const int span = min(M3_KC, K - kc * M3_KC);
      const int nbc = ((span + M3_BLK - 1) / M3_BLK + M3_CWARPS - 1) / M3_CWARPS;
      const int nvec = nbc * 4 * 128;                               // 16-byte vectors of the chunk
      for (int v = t; v < nvec; v += M3_CTHREADS) {
        const int kb = v >> 7, row = (v >> 3) & 15, c = v & 7;
        const int grow = tile * M3_ROWS + row, k = kc * M3_KC + kb * 64 + c * 8;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (grow < N && k < K) val = *reinterpret_cast<const uint4*>(W + (size_t)grow * ldw + k);
        out[kb * 128 + row * 8 + (c ^ (row & 7))] = val;
      }
      out += nvec;
    }
    return;
  }
  for (int kc = 0; kc < nkc; ++kc) {
    const int span = min(M3_KC, K - kc * M3_KC);
    const int nbc = ((span + M3_BLK - 1) / M3_BLK + M3_CWARPS - 1) / M3_CWARPS;
    for (int i = 0; i < nbc; ++i) {
      const int k = kc * M3_KC + (warp + M3_CWARPS * i) * M3_BLK + qd * 8;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int row = tile * M3_ROWS + g + 8 * half;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row < N && k < K) v = *reinterpret_cast<const uint4*>(W + (size_t)row * ldw + k);
        out[(2 * i + half) * M3_CTHREADS + t] = v;
      }
    }
    out += 2 * nbc * M3_CTHREADS;
  }
}

}  // namespace i2t

using namespace i2t;

// ---- everything a launch needs poisoned / zeroed, in ONE launch (six torch fills cost ~60 us of host time per generate) ----
__global__ void __launch_bounds__(256) decode_mega3_prepare_kernel(uint4* exch, int64_t exch_vecs, uint8_t* kcache, uint8_t* vcache,
                                                                   int64_t n_rows, int64_t row_pitch, int64_t row_off, int64_t fill_vecs,
                                                                   int64_t* ids, int64_t ids_ld, int B, int P, int ids_cols,
                                                                   unsigned long long* ctakeys, int n_keys, int32_t* err) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const uint4 ff = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  for (int64_t i = tid; i < exch_vecs; i += nth) exch[i] = ff;
  for (int64_t i = tid; i < n_rows * fill_vecs; i += nth) {          // cache rows [pos0, pos0 + steps) of every (layer, sequence)
    const int64_t r = i / fill_vecs, c = i - r * fill_vecs;
    reinterpret_cast<uint4*>(kcache + r * row_pitch + row_off)[c] = ff;
    reinterpret_cast<uint4*>(vcache + r * row_pitch + row_off)[c] = ff;
  }
  for (int64_t i = tid; i < (int64_t)B * (ids_cols - P); i += nth) {
    const int64_t b = i / (ids_cols - P), c = i - b * (ids_cols - P);
    ids[b * ids_ld + P + c] = -1;
  }
  for (int64_t i = tid; i < n_keys; i += nth) ctakeys[i] = 0ull;
  if (tid == 0) *err = 0;
}

static std::atomic<int> g_m3_sleep_ns{0};
extern "C" void i2t_set_decode_poll_sleep(int ns) { g_m3_sleep_ns.store(ns < 0 ? 0 : ns); }

extern "C" int i2t_decode_mega3_max_keys(void) { return M3_MAX_KEYS; }

// bytes one tile (16 weight rows) of a K-wide op occupies in a packed stream
extern "C" int64_t i2t_decode_mega3_tile_bytes(int64_t K) {
  int64_t blocks = 0;
  for (int64_t k0 = 0; k0 < K; k0 += M3_KC) {
    const int64_t span = K - k0 < M3_KC ? K - k0 : M3_KC;
    blocks += ((span + M3_BLK - 1) / M3_BLK + M3_CWARPS - 1) / M3_CWARPS;
  }
  return blocks * M3_BLOCK_BYTES;
}

// number of CTAs the decode kernel runs (= SMs): the host lays the weight streams out for exactly this grid
extern "C" int i2t_decode_mega3_grid(void) { return num_sms(); }

extern "C" int i2t_decode_mega3_pack(const void* W, int64_t N, int64_t K, int64_t ldw, void* dst, const int64_t* tile_off,
                                     int64_t tc_layout, void* stream) {
  I2T_REQUIRE(W && dst && tile_off, "decode_mega3_pack: null pointer");
  I2T_REQUIRE(N > 0 && K > 0 && K % 8 == 0 && ldw >= K && ldw % 8 == 0, "decode_mega3_pack: K and the row pitch must be positive multiples of 8");
  I2T_REQUIRE(aligned16(W) && aligned16(dst), "decode_mega3_pack: pointers must be 16-byte aligned");
  const int tiles = (int)((N + M3_ROWS - 1) / M3_ROWS);
  decode_mega3_pack_kernel<<<tiles, M3_CTHREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(W), (int)N, (int)K, ldw,
                                                                            reinterpret_cast<uint8_t*>(dst), tile_off, (int)tc_layout);
  I2T_LAUNCHED();
  return I2T_OK;
}

// Poisons what i2t_decode_mega3 expects poisoned: exch (exch_bytes of 0xFF), rows [pos0, pos0 + steps) of the two caches
// (n_rows = layers x sequences rows of row_pitch bytes, a position = pos_bytes bytes), ids[:, P:] = -1, ctakeys = 0, *err = 0.
extern "C" int i2t_decode_mega3_prepare(void* exch, int64_t exch_bytes, void* kcache, void* vcache, int64_t n_rows, int64_t row_pitch,
                                        int64_t pos_bytes, int64_t pos0, int64_t steps, int64_t* ids, int64_t ids_ld, int64_t B,
                                        int64_t P, int64_t ids_cols, uint64_t* ctakeys, int64_t n_keys, int32_t* err, void* stream) {
  I2T_REQUIRE(exch && kcache && vcache && ids && ctakeys && err, "decode_mega3_prepare: null pointer");
  I2T_REQUIRE(exch_bytes % 16 == 0 && pos_bytes % 16 == 0 && row_pitch % 16 == 0 && aligned16(exch) && aligned16(kcache) && aligned16(vcache),
              "decode_mega3_prepare: buffers must be 16-byte granular");
  I2T_REQUIRE(P >= 0 && P <= ids_cols && steps >= 0 && (pos0 + steps) * pos_bytes <= row_pitch, "decode_mega3_prepare: bad ranges");
  decode_mega3_prepare_kernel<<<num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<uint4*>(exch), exch_bytes / 16, reinterpret_cast<uint8_t*>(kcache), reinterpret_cast<uint8_t*>(vcache), n_rows,
      row_pitch, pos0 * pos_bytes, steps * pos_bytes / 16, ids, ids_ld, (int)B, (int)P, (int)ids_cols,
      reinterpret_cast<unsigned long long*>(ctakeys), (int)n_keys, err);
  I2T_LAUNCHED();
  return I2T_OK;
}

// Runs n_prefill prompt steps (no LM head) followed by n_sample sampled steps, starting at the device-side position *pos.
// The caller poisons the exchange buffers (0xFF bytes), the cache rows [pos, pos + steps) (0xFF bytes) and ids beyond the
// prompt (-1), zeroes ctakeys, and packs the weights (i2t_decode_mega3_pack) with the tile -> CTA map `rot` of the tables.
extern "C" int i2t_decode_mega3(const int64_t* lin, const int64_t* att, const int64_t* cmb, const int32_t* sched, int64_t n_sched,
                                int64_t n_ops, int64_t n_att, int64_t n_cmb, int64_t n_prefill, int64_t n_sample, int64_t B, int64_t C, int64_t H, int64_t V,
                                int64_t n_prompt, int64_t* ids, int64_t ids_ld, int32_t* pos, float* logits, int64_t ldl,
                                uint32_t* bar, int32_t* error_flag, uint64_t* ctakeys, const void* wpack, const int64_t* cta_base,
                                int64_t gen_stride, float temperature, int64_t top_k, const int32_t* ngrams, int64_t n_ngrams,
                                const uint64_t* seed_ptr, int64_t max_k, int64_t max_len, int64_t* trace, int64_t trace_cta,
                                int64_t tc, void* stream) {
  I2T_REQUIRE(lin && att && cmb && sched && ids && pos && logits && bar && error_flag && ctakeys && wpack && cta_base && seed_ptr,
              "decode_mega3: null pointer");
  I2T_REQUIRE(B > 0 && B <= M3_B, "decode_mega3: batch %lld outside 1..8", (long long)B);
  I2T_REQUIRE(H > 0 && C % H == 0 && (C / H == 64 || C / H == 32), "decode_mega3: head_dim must be 32 or 64");
  I2T_REQUIRE(C % 64 == 0 && C <= M3_KC, "decode_mega3: n_embd=%lld must be a multiple of 64, at most %d", (long long)C, M3_KC);
  I2T_REQUIRE(max_k % 256 == 0 && max_k >= C, "decode_mega3: max_k must be the padded widest input (multiple of 256)");
  I2T_REQUIRE(max_len <= M3_MAX_KEYS, "decode_mega3: %lld cached positions exceed the one-key-per-thread limit %d",
              (long long)max_len, M3_MAX_KEYS);
  I2T_REQUIRE(temperature > 0.f && n_prefill >= 0 && n_sample >= 0 && n_prefill + n_sample > 0, "decode_mega3: bad step counts / temperature");
  I2T_REQUIRE(gen_stride % 16 == 0, "decode_mega3: generation stride must be a multiple of 16 bytes");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = num_sms();
  I2T_REQUIRE(((V + M3_ROWS - 1) / M3_ROWS + grid - 1) / grid * M3_ROWS <= M3_MAX_LM_ROWS,
              "decode_mega3: vocabulary %lld gives a CTA more than %d LM-head rows", (long long)V, M3_MAX_LM_ROWS);
  I2T_REQUIRE(B * H <= grid && grid <= 256, "decode_mega3: %lld (batch, head) pairs exceed the %d CTAs (or more than 256 SMs)",
              (long long)(B * H), grid);
  const bool sampling = top_k != 1;
  int nslots = 0;
  for (int n = M3_MAX_SLOTS; n >= 3; --n)
    if (m3_layout(n, (int)max_k, (int)max_len, (int)(C / H), (int)n_ops, (int)n_att, (int)n_cmb, (int)n_sched, sampling).total <= 227 * 1024) {
      nslots = n;
      break;
    }
  I2T_REQUIRE(nslots >= 3, "decode_mega3: tables + activations leave no room for a 3-slot weight ring");
  const size_t smem = m3_layout(nslots, (int)max_k, (int)max_len, (int)(C / H), (int)n_ops, (int)n_att, (int)n_cmb, (int)n_sched, sampling).total;
  M3Args a;
  a.lin = lin; a.att = att; a.cmb = cmb; a.sched = reinterpret_cast<const int32_t*>(sched);
  a.n_sched = (int)n_sched; a.n_ops = (int)n_ops; a.n_att = (int)n_att; a.n_cmb = (int)n_cmb;
  a.n_prefill = (int)n_prefill; a.n_sample = (int)n_sample;
  a.B = (int)B; a.C = (int)C; a.H = (int)H; a.V = (int)V; a.n_prompt = (int)n_prompt; a.max_k = (int)max_k; a.max_len = (int)max_len; a.nslots = nslots;
  a.ids = ids; a.ids_ld = ids_ld; a.pos = pos; a.logits = logits; a.ldl = ldl; a.bar = bar; a.error_flag = error_flag;
  a.ctakeys = reinterpret_cast<unsigned long long*>(ctakeys);
  a.wpack = reinterpret_cast<const uint8_t*>(wpack); a.cta_base = cta_base; a.gen_stride = gen_stride;
  a.temperature = temperature; a.top_k = (int)(top_k > 0 ? top_k : 0);
  a.ngrams = ngrams; a.n_ngrams = (int)n_ngrams; a.seed_ptr = seed_ptr;
  a.trace = reinterpret_cast<long long*>(trace); a.trace_cta = (int)trace_cta;
  a.sleep_ns = g_m3_sleep_ns.load();
  a.tc = tc != 0 ? 1 : 0;
  const void* kern = tc != 0 ? ((C / H == 64) ? (const void*)decode_mega3_kernel<64, true> : (const void*)decode_mega3_kernel<32, true>)
                             : ((C / H == 64) ? (const void*)decode_mega3_kernel<64, false> : (const void*)decode_mega3_kernel<32, false>);
  I2T_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  I2T_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, M3_THREADS, smem));
  I2T_REQUIRE(per_sm >= 1, "decode_mega3: kernel does not fit on an SM");
  I2T_CUDA(cudaMemsetAsync(bar, 0, sizeof(uint32_t), st));
  void* params[] = {&a};
  I2T_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(M3_THREADS), params, smem, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return I2T_OK;
}
