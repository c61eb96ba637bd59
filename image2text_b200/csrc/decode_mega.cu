// Decode megakernel: ONE cooperative launch executes a whole KV-cached decode step (all layers, LM head, sampler).
//
// Why: a decode step for 8 sequences is ~80 strictly dependent stages that each touch a few MB of weights.  As separate
// kernels every stage pays launch + drain + cold-start latency (7-13 us measured) while the bytes need 0.4-1.2 us.
// Here the 148 CTAs stay resident, stages are separated by a ~1 us grid barrier, and -- the point of the design --
// weight traffic is decoupled from the stage structure: a dedicated PRODUCER warp per CTA walks the CTA's static list
// of weight tiles for the WHOLE step and streams them HBM -> shared memory with bulk async copies (cp.async.bulk, the
// 1-D TMA path) through a 4..8 slot ring, gated only by ring space (mbarrier full/empty pairs), never by the grid
// barriers.  While the compute warps wait at a barrier or run attention, the next stages' weights are already landing.
//
// Work split: a linear stage with N outputs is cut into N/8 units of 8 weight rows; unit u of op `o` belongs to CTA
// (u + 53*o) mod G.  A unit's K dimension is consumed in tiles of <= 768 elements (one ring slot: 8 rows x 768).
// The 8 compute warps split K across all 256 threads (x chunk in registers for the 8 batch rows, weights from the
// slot), reduce 8x8 partial sums with a transposing shuffle tree, and 64 threads run the epilogue (bias, GELU,
// residual, KV-cache append).  Attention stages: one (batch, head) pair per CTA, keys split over the 8 warps.
// The sampler runs on B CTAs with the vocabulary row staged in the (then idle) ring + activation shared memory.
//
// Tables (device int64 / int32 arrays built by the host, plain numbers -- no structs cross the C ABI):
//   lin[op][20]  : 0 W, 1 bias, 2 ln_g, 3 ln_b, 4 in, 5 out, 6 residual (in_mode 1: buffer that receives the embedding),
//                  7 N, 8 K, 9 act, 10 mode(0 plain, 1 qkv split + KV append), 11 kcache, 12 vcache,
//                  13 in_mode(0 buffer, 1 token embedding: in = wte), 14 wpe, 15 ldo, 16 cache batch stride, 17-19 unused
//   att[a][8]    : k, v, batch_stride, row_stride, len_mode(0: pos+1, 1: const), len_const, 0, 0
//   sched[s][4]  : kind (0 LIN op | 1 ATTN a | 2 SAMPLE | 3 ADVANCE), index, 0, 0
#include "common.cuh"
#include "sampler.cuh"

namespace i2t {

constexpr int MK_CW = 8;                         // compute warps
constexpr int MK_CT = MK_CW * 32;                // compute threads
constexpr int MK_THREADS = MK_CT + 32;           // + producer warp
constexpr int MK_R = 8, MK_B = 8;                // weight rows per unit, batch rows
constexpr int MK_KT = 768;                       // K tile (elements) per ring slot
constexpr int MK_MAXK = 3072;                    // activation staging capacity (elements per batch row)
constexpr int MK_RING_BYTES = 4 * MK_R * MK_KT * 4;   // 96 KB: 4 fp32 slots / 8 bf16 slots
constexpr int MK_MAX_SLOTS = 8;
constexpr int MK_LIN_FIELDS = 20;
constexpr int MK_SMEM = MK_RING_BYTES + MK_B * MK_MAXK * 4 + 2 * MK_CW * 64 * 4 + 3 * 1024 * 4;   // ring + xs + 2 partial buffers + (gamma|beta|wpe)

struct MkArgs {
  const int64_t* lin;
  const int64_t* att;
  const int32_t* sched;
  int n_sched, n_ops;
  int B, C, H, hs, V, n_prompt, w_bf16;
  const int64_t* ids_c;
  int64_t* ids;
  int64_t ids_ld;
  int32_t* pos;
  float* q;        // (B, C) query scratch
  float* y;        // (B, C) attention output scratch
  float* logits;   // (B, V)
  unsigned int* bar;      // grid barrier counter (zeroed by the host before every launch)
  int32_t* error_flag;
  float temperature;
  int top_k;
  const int32_t* ngrams;
  int n_ngrams;
  const uint64_t* seed_ptr;
  int32_t* ticket;
  int samp_in_smem;
  int trace_row;
  long long* trace;   // optional [n_sched][4] clock64 stamps of CTA 0 (begin, staged, computed, synced)
};

__device__ __forceinline__ uint32_t mk_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mk_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mk_smem(bar)), "r"(count));
}
__device__ __forceinline__ void mk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mk_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mk_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mk_smem(bar)) : "memory");
}
__device__ __forceinline__ bool mk_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = mk_smem(bar);
  uint32_t done = 0;
  for (uint32_t spins = 0; spins < (1u << 26); ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;   // never hang the GPU on a protocol bug: the caller raises the error flag
}
__device__ __forceinline__ void mk_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mk_smem(dst)),
               "l"(src), "r"(bytes), "r"(mk_smem(bar))
               : "memory");
}
__device__ __forceinline__ void mk_csync() { asm volatile("bar.sync 1, %0;" ::"n"(MK_CT) : "memory"); }   // compute warps only

// grid-wide barrier over the compute warps of all CTAs (monotonic counter, zeroed by the host per launch)
__device__ __forceinline__ void mk_grid_sync(unsigned int* bar, unsigned int& epoch, int32_t* error_flag, int tid) {
  mk_csync();
  if (tid == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int seen = 0;
    uint32_t spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if (++spins > (1u << 27)) {
        atomicExch(error_flag, 2);
        break;
      }
    } while (seen < epoch);
    __threadfence();
  }
  mk_csync();
}

__device__ __forceinline__ void mk_unpack(const float4& raw, float (&o)[4]) {
  o[0] = raw.x; o[1] = raw.y; o[2] = raw.z; o[3] = raw.w;
}
__device__ __forceinline__ void mk_unpack(const uint2& raw, float (&o)[4]) {
  o[0] = __uint_as_float(raw.x << 16); o[1] = __uint_as_float(raw.x & 0xffff0000u);
  o[2] = __uint_as_float(raw.y << 16); o[3] = __uint_as_float(raw.y & 0xffff0000u);
}

struct MkRing {
  uint32_t slot, phase;       // running slot index and its parity
  int nslots;
};
__device__ __forceinline__ void mk_ring_advance(MkRing& r) {
  if (++r.slot == (uint32_t)r.nslots) { r.slot = 0; r.phase ^= 1u; }
}

// ---- producer warp: stream every weight tile this CTA will consume, in consumption order ----
template <typename TW>
__device__ void mk_producer(const MkArgs& a, uint8_t* ring, uint64_t* full, uint64_t* empty, int lane) {
  MkRing r{0u, 0u, (int)(MK_RING_BYTES / (MK_R * MK_KT * sizeof(TW)))};
  const int G = gridDim.x;
  const size_t slot_bytes = (size_t)MK_R * MK_KT * sizeof(TW);
  for (int s = 0; s < a.n_sched; ++s) {
    if (a.sched[s * 4] != 0) continue;
    const int op = a.sched[s * 4 + 1];
    const int64_t* d = a.lin + (size_t)op * MK_LIN_FIELDS;
    const TW* W = reinterpret_cast<const TW*>(d[0]);
    const int N = (int)d[7], K = (int)d[8];
    const int nunits = (N + MK_R - 1) / MK_R;
    const int ntiles = (K + MK_KT - 1) / MK_KT;
    const int first = (int)(((int64_t)blockIdx.x - (int64_t)op * 53 % G + G) % G);
    for (int u = first; u < nunits; u += G) {
      const int n0 = u * MK_R;
      const int rows = min(MK_R, N - n0);
      for (int kt = 0; kt < ntiles; ++kt) {
        const int k0 = kt * MK_KT;
        const int kc = min(MK_KT, K - k0);
        if (!mk_mbar_wait(&empty[r.slot], r.phase ^ 1u)) { atomicExch(a.error_flag, 3); return; }
        uint8_t* dst = ring + (size_t)r.slot * slot_bytes;
        const uint32_t row_bytes = (uint32_t)kc * (uint32_t)sizeof(TW);
        if (lane == 0) mk_mbar_expect_tx(&full[r.slot], row_bytes * (uint32_t)rows);
        __syncwarp();
        if (ntiles == 1) {
          // rows are contiguous in memory: one copy for the whole unit
          if (lane == 0) mk_bulk_g2s(dst, W + (size_t)n0 * K, row_bytes * (uint32_t)rows, &full[r.slot]);
        } else if (lane < rows) {
          mk_bulk_g2s(dst + (size_t)lane * row_bytes, W + (size_t)(n0 + lane) * K + k0, row_bytes, &full[r.slot]);
        }
        mk_ring_advance(r);
      }
    }
  }
}

__device__ __forceinline__ void mk_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(mk_smem(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void mk_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ---- stage the activations of a linear stage into shared memory (LayerNorm prologue / token embedding) ----
// Every global read of the stage (8 activation rows, gamma, beta, position embedding) is issued as a 16-byte cp.async
// (LDGSTS, L2 -> shared memory, no registers) before anything waits, so all L2 latencies overlap; the LayerNorm
// statistics then run on shared memory.  `aux` = 3 * 1024 floats of scratch (gamma | beta | wpe row).
template <typename TW>
__device__ void mk_stage_x(const MkArgs& a, const int64_t* d, float* xs, float* aux, int pos, int w, int lane) {
  const int K = (int)d[8];
  const float* ln_g = reinterpret_cast<const float*>(d[2]);
  const float* ln_b = reinterpret_cast<const float*>(d[3]);
  const float* in = reinterpret_cast<const float*>(d[4]);
  const int in_mode = (int)d[13];
  const int tid = w * 32 + lane;
  float* xr = xs + (size_t)w * K;
  float* sg = aux;
  float* sb = aux + 1024;
  float* spe = aux + 2048;
  if (ln_g != nullptr)
    for (int k = tid * 4; k < K; k += MK_CT * 4) mk_cp_async16(sg + k, ln_g + k);
  if (ln_b != nullptr)
    for (int k = tid * 4; k < K; k += MK_CT * 4) mk_cp_async16(sb + k, ln_b + k);
  if (w < a.B) {
    const float* src = in + (int64_t)w * K;
    if (in_mode == 1) {   // x = wte[ids[b, pos]] + wpe[n_prompt + pos]
      const int64_t tok = __ldcg(a.ids_c + (int64_t)w * a.ids_ld + pos);
      src = in + tok * K;
      const float* wpe = reinterpret_cast<const float*>(d[14]) + (int64_t)(a.n_prompt + pos) * K;
      if (w == 0)
        for (int k = lane * 4; k < K; k += 128) mk_cp_async16(spe + k, wpe + k);
    }
    for (int k = lane * 4; k < K; k += 128) mk_cp_async16(xr + k, src + k);
  }
  mk_cp_async_wait_all();
  mk_csync();                      // gamma / beta / wpe were fetched cooperatively
  if (w >= a.B) {
    for (int k = lane * 4; k < K; k += 128) *reinterpret_cast<float4*>(xr + k) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  if (in_mode == 1) {              // CTA 0 also publishes the embedding as the residual stream
    float* xout = reinterpret_cast<float*>(d[6]);
    for (int k = lane * 4; k < K; k += 128) {
      float4 v = *reinterpret_cast<const float4*>(xr + k);
      const float4 pe = *reinterpret_cast<const float4*>(spe + k);
      v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
      *reinterpret_cast<float4*>(xr + k) = v;
      if (blockIdx.x == 0) store4(xout + (int64_t)w * K + k, v);
    }
  }
  if (ln_g != nullptr) {
    float s = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + k);
      s += (v.x + v.y) + (v.z + v.w);
    }
    const float mu = warp_sum(s) / (float)K;
    float q = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + k);
      const float c0 = v.x - mu, c1 = v.y - mu, c2 = v.z - mu, c3 = v.w - mu;
      q += (c0 * c0 + c1 * c1) + (c2 * c2 + c3 * c3);
    }
    const float rs = 1.0f / sqrtf(warp_sum(q) / (float)K + 1e-5f);
    for (int k = lane * 4; k < K; k += 128) {
      float4 v = *reinterpret_cast<const float4*>(xr + k);
      const float4 g = *reinterpret_cast<const float4*>(sg + k);
      v.x = (v.x - mu) * rs * g.x; v.y = (v.y - mu) * rs * g.y;
      v.z = (v.z - mu) * rs * g.z; v.w = (v.w - mu) * rs * g.w;
      if (ln_b != nullptr) {
        const float4 be = *reinterpret_cast<const float4*>(sb + k);
        v.x += be.x; v.y += be.y; v.z += be.z; v.w += be.w;
      }
      if (sizeof(TW) == 2) {       // autocast semantics: the Linear sees bf16 activations
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
}

template <typename TW>
__device__ void mk_stage_x_registers_unused(const MkArgs& a, const int64_t* d, float* xs, int pos, int w, int lane) {
  const int K = (int)d[8];
  const float* ln_g = reinterpret_cast<const float*>(d[2]);
  const float* ln_b = reinterpret_cast<const float*>(d[3]);
  const float* in = reinterpret_cast<const float*>(d[4]);
  const int in_mode = (int)d[13];
  float* xr = xs + (size_t)w * K;
  // All global loads of a batch row are issued back to back into registers (8 x 16 B per lane per trip) so their L2
  // latencies overlap; a loop of load -> store-to-smem -> load serialises ~0.7 us round trips.
  constexpr int NV = 8;                        // float4 per lane per trip: 1024 elements per warp trip
  auto round_bf16 = [](float4 v) {
    v.x = __bfloat162float(__float2bfloat16_rn(v.x)); v.y = __bfloat162float(__float2bfloat16_rn(v.y));
    v.z = __bfloat162float(__float2bfloat16_rn(v.z)); v.w = __bfloat162float(__float2bfloat16_rn(v.w));
    return v;
  };
  if (w >= a.B) {
    for (int k = lane * 4; k < K; k += 128) *reinterpret_cast<float4*>(xr + k) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const bool single_trip = K <= NV * 128;      // LayerNorm / embedding inputs have K = C <= 1024
  if (ln_g != nullptr || in_mode == 1) {
    // (the host guarantees C <= 1024 for these inputs)
    float4 v[NV], g[NV], be[NV];
    const float* src = in + (int64_t)w * K;
    const float* wpe = nullptr;
    if (in_mode == 1) {   // x = wte[ids[b, pos]] + wpe[n_prompt + pos]; CTA 0 also publishes it as the residual stream
      const int64_t tok = __ldcg(a.ids_c + (int64_t)w * a.ids_ld + pos);
      src = in + tok * K;
      wpe = reinterpret_cast<const float*>(d[14]) + (int64_t)(a.n_prompt + pos) * K;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int k = (lane + 32 * i) * 4;
      if (k < K) {
        v[i] = in_mode == 1 ? load4(src + k) : __ldcg(reinterpret_cast<const float4*>(src + k));
        if (ln_g != nullptr) g[i] = load4(ln_g + k);
        if (ln_b != nullptr) be[i] = load4(ln_b + k);
      }
    }
    if (in_mode == 1) {
      float* xout = reinterpret_cast<float*>(d[6]);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int k = (lane + 32 * i) * 4;
        if (k < K) {
          const float4 pe = load4(wpe + k);
          v[i].x += pe.x; v[i].y += pe.y; v[i].z += pe.z; v[i].w += pe.w;
          if (blockIdx.x == 0) store4(xout + (int64_t)w * K + k, v[i]);
        }
      }
    }
    if (ln_g != nullptr) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if ((lane + 32 * i) * 4 < K) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      const float mu = warp_sum(s) / (float)K;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if ((lane + 32 * i) * 4 < K) {
          const float c0 = v[i].x - mu, c1 = v[i].y - mu, c2 = v[i].z - mu, c3 = v[i].w - mu;
          q += (c0 * c0 + c1 * c1) + (c2 * c2 + c3 * c3);
        }
      const float rs = 1.0f / sqrtf(warp_sum(q) / (float)K + 1e-5f);
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if ((lane + 32 * i) * 4 < K) {
          v[i].x = (v[i].x - mu) * rs * g[i].x; v[i].y = (v[i].y - mu) * rs * g[i].y;
          v[i].z = (v[i].z - mu) * rs * g[i].z; v[i].w = (v[i].w - mu) * rs * g[i].w;
          if (ln_b != nullptr) {
            v[i].x += be[i].x; v[i].y += be[i].y; v[i].z += be[i].z; v[i].w += be[i].w;
          }
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int k = (lane + 32 * i) * 4;
      if (k < K) *reinterpret_cast<float4*>(xr + k) = sizeof(TW) == 2 ? round_bf16(v[i]) : v[i];   // autocast: bf16 inputs
    }
    (void)single_trip;
  } else {
    const float* src = in + (int64_t)w * K;
    for (int k0 = 0; k0 < K; k0 += NV * 128) {
      float4 v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int k = k0 + (lane + 32 * i) * 4;
        if (k < K) v[i] = __ldcg(reinterpret_cast<const float4*>(src + k));
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int k = k0 + (lane + 32 * i) * 4;
        if (k < K) *reinterpret_cast<float4*>(xr + k) = sizeof(TW) == 2 ? round_bf16(v[i]) : v[i];
      }
    }
  }
}

// ---- one linear stage on the compute warps ----
template <typename TW>
__device__ bool mk_linear(const MkArgs& a, int op, uint8_t* ring, uint64_t* full, uint64_t* empty, MkRing& r, float* xs,
                          float* part2, int pos, int tid) {
  using WVec = typename std::conditional<sizeof(TW) == 4, float4, uint2>::type;   // 4 weights per load
  int ucount = 0;
  const int lane = tid & 31, w = tid >> 5;
  const int64_t* d = a.lin + (size_t)op * MK_LIN_FIELDS;
  const float* bias = reinterpret_cast<const float*>(d[1]);
  float* out = reinterpret_cast<float*>(d[5]);
  const float* residual = reinterpret_cast<const float*>(d[6]);
  const int N = (int)d[7], K = (int)d[8], act = (int)d[9], mode = (int)d[10];
  const int64_t ldo = d[15];
  const int G = gridDim.x;
  const size_t slot_bytes = (size_t)MK_R * MK_KT * sizeof(TW);
  mk_stage_x<TW>(a, d, xs, part2 + 2 * MK_CW * 64, pos, w, lane);
  mk_csync();
  if (a.trace != nullptr && blockIdx.x == 0 && tid == 0) a.trace[a.trace_row * 4 + 1] = clock64();
  const int nunits = (N + MK_R - 1) / MK_R;
  const int ntiles = (K + MK_KT - 1) / MK_KT;
  const int first = (int)(((int64_t)blockIdx.x - (int64_t)op * 53 % G + G) % G);
  for (int u = first; u < nunits; u += G) {
    const int n0 = u * MK_R;
    const int rows = min(MK_R, N - n0);
    // epilogue operands do not depend on this stage: fetch them now, use them after the reduction
    float e_bias = 0.f, e_res = 0.f;
    const int er = tid / MK_B, eb = tid % MK_B;
    const bool e_on = tid < MK_R * MK_B && er < rows && eb < a.B;
    if (e_on) {
      if (bias != nullptr) e_bias = bias[n0 + er];
      if (mode == 0 && residual != nullptr && (int)d[13] == 0) e_res = __ldcg(residual + (int64_t)eb * ldo + n0 + er);
    }
    float acc[MK_R * MK_B];
#pragma unroll
    for (int i = 0; i < MK_R * MK_B; ++i) acc[i] = 0.f;
    for (int kt = 0; kt < ntiles; ++kt) {
      const int k0 = kt * MK_KT;
      const int kc = min(MK_KT, K - k0);
      if (!mk_mbar_wait(&full[r.slot], r.phase)) { atomicExch(a.error_flag, 4); return false; }
      const TW* wp = reinterpret_cast<const TW*>(ring + (size_t)r.slot * slot_bytes);
      const int row_pitch = ntiles == 1 ? K : kc;      // elements between consecutive rows inside the slot
      for (int c = tid; c < kc / 4; c += MK_CT) {
        float xv[MK_B][4];
#pragma unroll
        for (int b = 0; b < MK_B; ++b) {
          const float4 t4 = *reinterpret_cast<const float4*>(xs + (size_t)b * K + k0 + c * 4);
          xv[b][0] = t4.x; xv[b][1] = t4.y; xv[b][2] = t4.z; xv[b][3] = t4.w;
        }
#pragma unroll
        for (int rr = 0; rr < MK_R; ++rr) {
          if (rr < rows) {
            float wv[4];
            mk_unpack(*reinterpret_cast<const WVec*>(wp + (size_t)rr * row_pitch + c * 4), wv);
#pragma unroll
            for (int b = 0; b < MK_B; ++b)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[rr * MK_B + b] = fmaf(wv[j], xv[b][j], acc[rr * MK_B + b]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mk_mbar_arrive(&empty[r.slot]);     // 8 arrivals (one per compute warp) free the slot
      mk_ring_advance(r);
    }
    // transposing shuffle reduction: afterwards lane l holds the warp totals of flat indices 2l and 2l+1
#pragma unroll
    for (int off = 16, n = MK_R * MK_B; off >= 1; off >>= 1, n >>= 1) {
      const int half = n >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float send = upper ? acc[i] : acc[i + half];
        const float keep = upper ? acc[i + half] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    // two partial buffers used alternately: one CTA barrier per unit is enough (the buffer written by unit i is next
    // written by unit i+2, after every thread has passed unit i+1's barrier, hence after unit i's readers finished)
    float* part = part2 + (ucount & 1) * (MK_CW * 64);
    ++ucount;
    part[w * 64 + 2 * lane] = acc[0];
    part[w * 64 + 2 * lane + 1] = acc[1];
    mk_csync();
    if (e_on) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < MK_CW; ++i) v += part[i * 64 + tid];
      const int n = n0 + er;
      v += e_bias;
      v = apply_act(v, act);
      if (mode == 0) {
        if (residual != nullptr) {
          if ((int)d[13] == 1) e_res = 0.f;
          v += e_res;
        }
        out[(int64_t)eb * ldo + n] = v;
      } else {
        const int seg = n / a.C, nl = n % a.C;
        if (seg == 0) {
          out[(int64_t)eb * ldo + nl] = v;
        } else {
          void* base = reinterpret_cast<void*>(seg == 1 ? d[11] : d[12]);
          const int64_t cache_bs = d[16];
          const int64_t off = (int64_t)eb * cache_bs + (int64_t)pos * a.C + nl;
          if (a.w_bf16) reinterpret_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(v);
          else reinterpret_cast<float*>(base)[off] = v;
        }
      }
    }
  }
  return true;
}

// ---- single-query attention for one (batch, head) per CTA; keys split over the 8 compute warps ----
template <typename TC, int HS>
__device__ void mk_attention(const MkArgs& a, int ai, float* scratch, int pos, int tid) {
  constexpr int EPL = HS / 32, G4 = 4;
  const int lane = tid & 31, w = tid >> 5;
  const int64_t* d = a.att + (size_t)ai * 8;
  const TC* kc = reinterpret_cast<const TC*>(d[0]);
  const TC* vc = reinterpret_cast<const TC*>(d[1]);
  const int64_t bs = d[2], rs = d[3];
  const int len = d[4] == 0 ? pos + 1 : (int)d[5];
  float* s_m = scratch;                 // [8]
  float* s_l = scratch + 8;             // [8]
  float* s_acc = scratch + 16;          // [8][HS]
  const float scale = 1.0f / sqrtf((float)HS);
  for (int unit = blockIdx.x; unit < a.B * a.H; unit += gridDim.x) {
    const int b = unit / a.H, h = unit % a.H;
    float qv[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      float v = __ldcg(a.q + (int64_t)b * a.C + h * HS + lane + 32 * e);
      if (sizeof(TC) == 2) v = __bfloat162float(__float2bfloat16_rn(v));
      qv[e] = v * scale;
    }
    const TC* kb = kc + b * bs + (int64_t)h * HS;
    const TC* vb = vc + b * bs + (int64_t)h * HS;
    float m = -INFINITY, l = 0.f, acc[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
    for (int j0 = w * G4; j0 < len; j0 += MK_CW * G4) {
      float kk[G4][EPL], vv[G4][EPL], dd[G4];
#pragma unroll
      for (int g = 0; g < G4; ++g) {
        const int j = min(j0 + g, len - 1);
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          kk[g][e] = to_f32(__ldcg(kb + (int64_t)j * rs + lane + 32 * e));
          vv[g][e] = to_f32(__ldcg(vb + (int64_t)j * rs + lane + 32 * e));
        }
      }
#pragma unroll
      for (int g = 0; g < G4; ++g) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) s = fmaf(qv[e], kk[g][e], s);
        dd[g] = s;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int g = 0; g < G4; ++g) dd[g] += __shfl_xor_sync(0xffffffffu, dd[g], o);
      float m_new = m;
#pragma unroll
      for (int g = 0; g < G4; ++g)
        if (j0 + g < len) m_new = fmaxf(m_new, dd[g]);
      const float corr = expf(m - m_new);
      l *= corr;
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[e] *= corr;
#pragma unroll
      for (int g = 0; g < G4; ++g) {
        if (j0 + g < len) {
          const float p = expf(dd[g] - m_new);
          l += p;
#pragma unroll
          for (int e = 0; e < EPL; ++e) acc[e] = fmaf(p, vv[g][e], acc[e]);
        }
      }
      m = m_new;
    }
    if (lane == 0) { s_m[w] = m; s_l[w] = l; }
#pragma unroll
    for (int e = 0; e < EPL; ++e) s_acc[w * HS + lane + 32 * e] = acc[e];
    mk_csync();
    if (w == 0) {
      float M = -INFINITY;
#pragma unroll
      for (int i = 0; i < MK_CW; ++i) M = fmaxf(M, s_m[i]);
      float L = 0.f, o[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) o[e] = 0.f;
#pragma unroll
      for (int i = 0; i < MK_CW; ++i) {
        const float c = (s_m[i] == -INFINITY) ? 0.f : expf(s_m[i] - M);
        L += s_l[i] * c;
#pragma unroll
        for (int e = 0; e < EPL; ++e) o[e] = fmaf(s_acc[i * HS + lane + 32 * e], c, o[e]);
      }
      const float inv = L > 0.f ? 1.0f / L : 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) a.y[(int64_t)b * a.C + h * HS + lane + 32 * e] = o[e] * inv;
    }
    mk_csync();
  }
}

template <typename TW, int HS>
__global__ void __launch_bounds__(MK_THREADS, 1) decode_mega_kernel(MkArgs a_in) {
  MkArgs a = a_in;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[MK_MAX_SLOTS];
  __shared__ __align__(8) uint64_t empty[MK_MAX_SLOTS];
  uint8_t* ring = smem;
  float* xs = reinterpret_cast<float*>(smem + MK_RING_BYTES);
  float* part = xs + MK_B * MK_MAXK;
  const int tid = threadIdx.x;
  const int nslots = (int)(MK_RING_BYTES / (MK_R * MK_KT * sizeof(TW)));
  if (tid == 0) {
    for (int i = 0; i < nslots; ++i) {
      mk_mbar_init(&full[i], 1);
      mk_mbar_init(&empty[i], MK_CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int pos = *a.pos;     // every CTA reads it before the final grid barrier; it changes only after that barrier
  if (tid >= MK_CT) {
    mk_producer<TW>(a, ring, full, empty, tid - MK_CT);
    return;
  }
  MkRing r{0u, 0u, nslots};
  unsigned int epoch = 0;
  for (int s = 0; s < a.n_sched; ++s) {
    const int kind = a.sched[s * 4], idx = a.sched[s * 4 + 1];
    const bool tr = a.trace != nullptr && blockIdx.x == 0 && tid == 0;
    a.trace_row = s;
    if (tr) a.trace[s * 4 + 0] = clock64();
    if (kind == 0) {
      if (!mk_linear<TW>(a, idx, ring, full, empty, r, xs, part, pos, tid)) return;
    } else if (kind == 1) {
      mk_attention<TW, HS>(a, idx, xs, pos, tid);
    } else if (kind == 2) {
      // sampler: B CTAs, vocabulary row staged in the now idle ring + activation shared memory
      if ((int)blockIdx.x < a.B) {
        const int b = blockIdx.x;
        const int cur_len = pos + 1;
        float* row = a.logits + (int64_t)b * a.V;
        // dynamic smem was sized by the host as max(stage buffers, V floats + scratch) when that fits (a.samp_in_smem)
        const size_t row_bytes = ((size_t)a.V * 4 + 15) & ~(size_t)15;
        float* sv = a.samp_in_smem ? reinterpret_cast<float*>(smem) : row;
        SampleScratch& samp = *reinterpret_cast<SampleScratch*>(smem + (a.samp_in_smem ? row_bytes : 0));
        const int choice = sample_row_smem(sv, samp, row, a.V, a.ids_c + (int64_t)b * a.ids_ld, cur_len, a.temperature,
                                           a.top_k, a.ngrams, a.n_ngrams, *a.seed_ptr, b, nullptr, tid, MK_CT);
        if (tid == 0) {
          a.ids[(int64_t)b * a.ids_ld + cur_len] = (int64_t)choice;
          __threadfence();
          const int fin = atomicAdd(a.ticket, 1);
          if (fin == a.B - 1) {
            *a.ticket = 0;
            *a.pos = pos + 1;
          }
        }
      }
      continue;   // nothing follows the sampler
    } else {
      if (blockIdx.x == 0 && tid == 0) *a.pos = pos + 1;
      continue;
    }
    if (tr) a.trace[s * 4 + 2] = clock64();
    mk_grid_sync(a.bar, epoch, a.error_flag, tid);
    if (tr) a.trace[s * 4 + 3] = clock64();
  }
}

}  // namespace i2t

using namespace i2t;

// lin / att / sched are DEVICE tables (see the file header).  bar: device uint32 (zeroed here, per launch);
// error_flag: device int32, set non-zero if a wait timed out (the step's results are then invalid).
extern "C" int i2t_decode_mega(const int64_t* lin, const int64_t* att, const int32_t* sched, int64_t n_sched, int64_t n_ops,
                               int64_t B, int64_t C, int64_t H, int64_t V, int64_t n_prompt, int w_dtype,
                               int64_t* ids, int64_t ids_ld, int32_t* pos, float* q, float* y, float* logits,
                               uint32_t* bar, int32_t* error_flag, float temperature, int64_t top_k,
                               const int32_t* ngrams, int64_t n_ngrams, const uint64_t* seed_ptr, int32_t* ticket,
                               int64_t max_k, int64_t* trace, void* stream) {
  I2T_REQUIRE(lin && att && sched && ids && pos && q && y && logits && bar && error_flag && seed_ptr && ticket,
              "decode_mega: null pointer");
  I2T_REQUIRE(B > 0 && B <= MK_B, "decode_mega: batch %lld outside 1..8", (long long)B);
  I2T_REQUIRE(H > 0 && C % H == 0 && (C / H == 64 || C / H == 32), "decode_mega: head_dim must be 32 or 64");
  I2T_REQUIRE(C <= 1024, "decode_mega: n_embd=%lld above 1024 (LayerNorm staging keeps the row in registers)", (long long)C);
  I2T_REQUIRE(max_k <= MK_MAXK && C % 8 == 0, "decode_mega: K=%lld exceeds the staging capacity %d", (long long)max_k, MK_MAXK);
  I2T_REQUIRE(valid_dtype(w_dtype) && temperature > 0.f, "decode_mega: bad dtype / temperature");
  cudaStream_t st = (cudaStream_t)stream;
  MkArgs a;
  a.lin = lin; a.att = att; a.sched = sched;
  a.n_sched = (int)n_sched; a.n_ops = (int)n_ops;
  a.B = (int)B; a.C = (int)C; a.H = (int)H; a.hs = (int)(C / H); a.V = (int)V; a.n_prompt = (int)n_prompt;
  a.w_bf16 = w_dtype == I2T_BF16;
  a.ids_c = ids; a.ids = ids; a.ids_ld = ids_ld; a.pos = pos;
  a.q = q; a.y = y; a.logits = logits; a.bar = bar; a.error_flag = error_flag;
  a.temperature = temperature; a.top_k = (int)(top_k > 0 ? top_k : 0);
  a.ngrams = ngrams; a.n_ngrams = (int)n_ngrams; a.seed_ptr = seed_ptr; a.ticket = ticket;
  const void* kern = nullptr;
  const int hs = (int)(C / H);
  if (w_dtype == I2T_F32) kern = hs == 64 ? (const void*)decode_mega_kernel<float, 64> : (const void*)decode_mega_kernel<float, 32>;
  else kern = hs == 64 ? (const void*)decode_mega_kernel<__nv_bfloat16, 64> : (const void*)decode_mega_kernel<__nv_bfloat16, 32>;
  // the sampler stage re-uses the (then idle) dynamic shared memory for the vocabulary row + its scratch
  size_t smem = MK_SMEM;
  const size_t samp_need = (((size_t)V * 4 + 15) & ~(size_t)15) + sizeof(SampleScratch);
  a.trace = reinterpret_cast<long long*>(trace);
  a.trace_row = 0;
  a.samp_in_smem = samp_need <= 226 * 1024 ? 1 : 0;
  if (a.samp_in_smem && samp_need > smem) smem = samp_need;
  I2T_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  I2T_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MK_THREADS, smem));
  I2T_REQUIRE(per_sm >= 1, "decode_mega: kernel does not fit on an SM");
  const int grid = num_sms();      // one CTA per SM: all co-resident (cooperative launch checks it)
  I2T_CUDA(cudaMemsetAsync(bar, 0, sizeof(uint32_t), st));
  void* params[] = {&a};
  I2T_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(MK_THREADS), params, smem, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return I2T_OK;
}
