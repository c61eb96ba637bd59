// On-device next-token sampler: temperature -> no-repeat-n-gram ban -> top-k threshold (ties kept) -> softmax ->
// multinomial draw -> append.  Replaces reference models/vision_encoder_decoder.py:152-180 and the host-side
// transformers NoRepeatNGramLogitsProcessor (generation/logits_process.py:1012-1135: python dict loops plus a
// .tolist() device sync per step).  One CTA per batch row; the vocabulary row (V fp32) is re-read from L2.
//   * top-k threshold: exact k-th largest value by 4-pass radix select on order-preserving uint keys, then
//     everything < threshold is dropped -- identical to `logits[logits < v[..., [-1]]] = -inf` (ties survive).
//   * top_k = 1 therefore is greedy decoding (the reference has no other greedy switch).
//   * draw: u ~ Philox4x32-10(seed, row, position); smallest index whose inclusive prefix mass >= u * total.
#include "common.cuh"

namespace i2t {

constexpr int SAMP_THREADS = 1024;
constexpr int SAMP_MAX_BANNED = 1024;

__device__ __forceinline__ uint32_t float_key(float f) {  // monotone: a < b  <=>  key(a) < key(b)
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t row, uint32_t pos) {
  uint32_t c[4] = {pos, row, 0x243F6A88u, 0x85A308D3u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
}

template <typename T>
__device__ __forceinline__ T block_reduce(T v, T* scratch, bool is_max) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? (other > v ? other : v) : v + other;
  }
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  T r = scratch[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? (scratch[i] > r ? scratch[i] : r) : r + scratch[i];
  return r;
}

// logits: (B, ldl) fp32, scaled / banned in place.  ids: (B, ids_ld) int64, tokens [0, cur_len) are valid;
// the sampled token is written to ids[b, cur_len].  cur_len = *pos_ptr + 1 when pos_ptr is given (decode loop:
// pos is the index of the token just processed) else the constant cur_len_const.
__global__ void __launch_bounds__(SAMP_THREADS)
sample_kernel(float* __restrict__ logits, int64_t ldl, int V, int64_t* __restrict__ ids, int64_t ids_ld,
              const int32_t* pos_ptr_c, int32_t* pos_ptr_adv, int cur_len_const,
              float temperature, int top_k, const int32_t* __restrict__ ngrams, int n_ngrams, uint64_t seed_arg,
              const uint64_t* __restrict__ seed_ptr, float* __restrict__ probs_out, int32_t* __restrict__ ticket, int write_token) {
  __shared__ int s_banned[SAMP_MAX_BANNED];
  __shared__ int s_nbanned;
  __shared__ uint32_t s_hist[256];
  __shared__ float s_redf[32];
  __shared__ double s_redd[32];
  __shared__ uint32_t s_prefix, s_kleft;
  __shared__ double s_scan[SAMP_THREADS / 32];
  __shared__ int s_choice, s_fallback;
  const int t = threadIdx.x;
  const int b = blockIdx.x;
  const uint64_t seed = seed_ptr != nullptr ? *seed_ptr : seed_arg;
  const int cur_len = pos_ptr_c != nullptr ? (*pos_ptr_c + 1) : cur_len_const;
  float* row = logits + (int64_t)b * ldl;
  const int64_t* idr = ids + (int64_t)b * ids_ld;

  // ---- 1. banned tokens (generation/logits_process.py:1012-1076) ----
  if (t == 0) { s_nbanned = 0; s_choice = 0x7fffffff; s_fallback = 0x7fffffff; }
  __syncthreads();
  for (int g = 0; g < n_ngrams; ++g) {
    const int n = ngrams[g];
    if (n <= 0 || cur_len + 1 < n) continue;
    const int tail = cur_len + 1 - n;  // start of the (n-1)-token suffix
    for (int i = t; i <= cur_len - n; i += blockDim.x) {
      bool same = true;
      for (int j = 0; j < n - 1; ++j) same = same && (idr[i + j] == idr[tail + j]);
      if (same) {
        const int slot = atomicAdd(&s_nbanned, 1);
        if (slot < SAMP_MAX_BANNED) s_banned[slot] = (int)idr[i + n - 1];
      }
    }
  }
  __syncthreads();
  const int nb = min(s_nbanned, SAMP_MAX_BANNED);

  // ---- 2. temperature (a true division like the reference: logits / temperature) ----
  for (int i = t; i < V; i += blockDim.x) row[i] = row[i] / temperature;
  __syncthreads();
  for (int i = t; i < nb; i += blockDim.x) {
    const int tok = s_banned[i];
    if (tok >= 0 && tok < V) row[tok] = -INFINITY;
  }
  __syncthreads();

  // ---- 3. top-k threshold by radix select on the keys (k-th largest) ----
  float thr = -INFINITY;
  if (top_k > 0 && top_k < V) {
    if (t == 0) { s_prefix = 0u; s_kleft = (uint32_t)top_k; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = t; i < 256; i += blockDim.x) s_hist[i] = 0u;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (int i = t; i < V; i += blockDim.x) {
        const uint32_t key = float_key(row[i]);
        if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (t == 0) {
        uint32_t left = s_kleft;
        int bin = 255;
        for (; bin > 0; --bin) {
          if (s_hist[bin] >= left) break;
          left -= s_hist[bin];
        }
        s_kleft = left;
        s_prefix = prefix | ((uint32_t)bin << shift);
      }
      __syncthreads();
    }
    const uint32_t kk = s_prefix;
    thr = __uint_as_float((kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk);
  }

  // ---- 4. softmax over the survivors ----
  float mx = -INFINITY;
  for (int i = t; i < V; i += blockDim.x) {
    const float v = row[i];
    if (v >= thr) mx = fmaxf(mx, v);
  }
  mx = block_reduce<float>(mx, s_redf, true);
  double part = 0.0;
  const int chunk = (V + blockDim.x - 1) / blockDim.x;   // contiguous chunk per thread (index order for the scan)
  const int beg = t * chunk, end = min(V, beg + chunk);
  for (int i = beg; i < end; ++i) {
    const float v = row[i];
    if (v >= thr) part += (double)expf(v - mx);
  }
  const double total = block_reduce<double>(part, s_redd, false);
  if (probs_out != nullptr) {
    for (int i = t; i < V; i += blockDim.x) {
      const float v = row[i];
      probs_out[(int64_t)b * V + i] = v >= thr ? (float)((double)expf(v - mx) / total) : 0.f;
    }
  }

  // ---- 5. multinomial draw: inverse CDF in index order ----
  const double target = (double)philox_uniform(seed, (uint32_t)b, (uint32_t)cur_len) * total;
  // block exclusive scan of `part`
  const int lane = t & 31, w = t >> 5;
  double incl = part;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double nbr = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += nbr;
  }
  if (lane == 31) s_scan[w] = incl;
  __syncthreads();
  double woff = 0.0;
  for (int i = 0; i < w; ++i) woff += s_scan[i];
  const double excl = woff + incl - part;
  if (part > 0.0 && target > excl && target <= excl + part) {
    double run = excl;
    int pick = -1;
    for (int i = beg; i < end; ++i) {
      const float v = row[i];
      if (v >= thr) {
        run += (double)expf(v - mx);
        pick = i;
        if (run >= target) break;
      }
    }
    if (pick >= 0) atomicMin(&s_choice, pick);
  }
  __syncthreads();
  if (s_choice == 0x7fffffff) {
    // rounding corner (target landed past the last survivor's prefix): take the first arg-max survivor
    int best = 0x7fffffff;
    for (int i = t; i < V; i += blockDim.x)
      if (row[i] == mx) best = min(best, i);
    if (best != 0x7fffffff) atomicMin(&s_fallback, best);
    __syncthreads();
    if (t == 0) s_choice = s_fallback;
    __syncthreads();
  }
  if (t == 0) {
    if (write_token) ids[(int64_t)b * ids_ld + cur_len] = (int64_t)s_choice;
    if (pos_ptr_adv != nullptr) {
      __threadfence();
      const int done = atomicAdd(ticket, 1);
      if (done == (int)gridDim.x - 1) {  // last row: every CTA has read *pos_ptr already
        *ticket = 0;
        *pos_ptr_adv += 1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Shared-memory-resident variant (V floats fit in the CTA's shared memory, e.g. GPT-2's 50257): the vocabulary row is
// read from memory ONCE; every later pass (radix select, max, sum, draw) runs on shared memory.  The top-k radix
// select uses 8 group histograms so the few hot exponent bins contend less.  top_k == 1 (greedy) skips the radix
// select: the threshold is the row maximum.  The draw walks the elements in (thread, stride) order -- any fixed order
// yields the same categorical distribution.
// ---------------------------------------------------------------------------------------------------------
constexpr int SAMP_HGROUPS = 8;

__global__ void __launch_bounds__(SAMP_THREADS)
sample_smem_kernel(float* __restrict__ logits, int64_t ldl, int V, int64_t* __restrict__ ids, int64_t ids_ld,
                   const int32_t* pos_ptr_c, int32_t* pos_ptr_adv, int cur_len_const, float temperature, int top_k,
                   const int32_t* __restrict__ ngrams, int n_ngrams, uint64_t seed_arg,
                   const uint64_t* __restrict__ seed_ptr, float* __restrict__ probs_out, int32_t* __restrict__ ticket,
                   int write_token) {
  extern __shared__ __align__(16) float sv[];   // [V]
  __shared__ int s_banned[SAMP_MAX_BANNED];
  __shared__ int s_nbanned;
  __shared__ uint32_t s_ghist[SAMP_HGROUPS][256];
  __shared__ float s_redf[32];
  __shared__ double s_redd[32];
  __shared__ uint32_t s_prefix, s_kleft;
  __shared__ double s_scan[SAMP_THREADS / 32];
  __shared__ int s_choice, s_fallback;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int b = blockIdx.x;
  const uint64_t seed = seed_ptr != nullptr ? *seed_ptr : seed_arg;
  const int cur_len = pos_ptr_c != nullptr ? (*pos_ptr_c + 1) : cur_len_const;
  float* row = logits + (int64_t)b * ldl;
  const int64_t* idr = ids + (int64_t)b * ids_ld;

  if (t == 0) { s_nbanned = 0; s_choice = 0x7fffffff; s_fallback = 0x7fffffff; }
  __syncthreads();
  for (int g = 0; g < n_ngrams; ++g) {
    const int n = ngrams[g];
    if (n <= 0 || cur_len + 1 < n) continue;
    const int tail = cur_len + 1 - n;
    for (int i = t; i <= cur_len - n; i += SAMP_THREADS) {
      bool same = true;
      for (int j = 0; j < n - 1; ++j) same = same && (idr[i + j] == idr[tail + j]);
      if (same) {
        const int slot = atomicAdd(&s_nbanned, 1);
        if (slot < SAMP_MAX_BANNED) s_banned[slot] = (int)idr[i + n - 1];
      }
    }
  }
  // one pass over global memory: scale, keep in shared memory, track the maximum
  float mx = -INFINITY;
  for (int i = t; i < V; i += SAMP_THREADS) sv[i] = row[i] / temperature;
  __syncthreads();
  const int nb = min(s_nbanned, SAMP_MAX_BANNED);
  for (int i = t; i < nb; i += SAMP_THREADS) {
    const int tok = s_banned[i];
    if (tok >= 0 && tok < V) sv[tok] = -INFINITY;
  }
  __syncthreads();
  if (temperature != 1.0f || nb > 0) {      // keep the documented in-place contract (scaled / banned logits)
    for (int i = t; i < V; i += SAMP_THREADS) row[i] = sv[i];
  }
  for (int i = t; i < V; i += SAMP_THREADS) mx = fmaxf(mx, sv[i]);
  mx = block_reduce<float>(mx, s_redf, true);

  float thr = -INFINITY;
  if (top_k == 1) {
    thr = mx;
  } else if (top_k > 1 && top_k < V) {
    if (t == 0) { s_prefix = 0u; s_kleft = (uint32_t)top_k; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = t; i < SAMP_HGROUPS * 256; i += SAMP_THREADS) (&s_ghist[0][0])[i] = 0u;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      uint32_t* hist = s_ghist[w & (SAMP_HGROUPS - 1)];
      for (int i = t; i < V; i += SAMP_THREADS) {
        const uint32_t key = float_key(sv[i]);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (t < 256) {
        uint32_t tot = 0;
#pragma unroll
        for (int i = 0; i < SAMP_HGROUPS; ++i) tot += s_ghist[i][t];
        s_ghist[0][t] = tot;
      }
      __syncthreads();
      if (t == 0) {
        uint32_t left = s_kleft;
        int bin = 255;
        for (; bin > 0; --bin) {
          if (s_ghist[0][bin] >= left) break;
          left -= s_ghist[0][bin];
        }
        s_kleft = left;
        s_prefix = prefix | ((uint32_t)bin << shift);
      }
      __syncthreads();
    }
    const uint32_t kk = s_prefix;
    thr = __uint_as_float((kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk);
  }

  double part = 0.0;
  for (int i = t; i < V; i += SAMP_THREADS) {
    const float x = sv[i];
    if (x >= thr) part += (double)expf(x - mx);
  }
  const double total = block_reduce<double>(part, s_redd, false);
  if (probs_out != nullptr) {
    for (int i = t; i < V; i += SAMP_THREADS) {
      const float x = sv[i];
      probs_out[(int64_t)b * V + i] = x >= thr ? (float)((double)expf(x - mx) / total) : 0.f;
    }
  }
  const double target = (double)philox_uniform(seed, (uint32_t)b, (uint32_t)cur_len) * total;
  double incl = part;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double nbr = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += nbr;
  }
  if (lane == 31) s_scan[w] = incl;
  __syncthreads();
  double woff = 0.0;
  for (int i = 0; i < w; ++i) woff += s_scan[i];
  const double excl = woff + incl - part;
  if (part > 0.0 && target > excl && target <= excl + part) {
    double run = excl;
    int pick = -1;
    for (int i = t; i < V; i += SAMP_THREADS) {
      const float x = sv[i];
      if (x >= thr) {
        run += (double)expf(x - mx);
        pick = i;
        if (run >= target) break;
      }
    }
    if (pick >= 0) atomicMin(&s_choice, pick);
  }
  __syncthreads();
  if (s_choice == 0x7fffffff) {
    int best = 0x7fffffff;
    for (int i = t; i < V; i += SAMP_THREADS)
      if (sv[i] == mx) best = min(best, i);
    if (best != 0x7fffffff) atomicMin(&s_fallback, best);
    __syncthreads();
    if (t == 0) s_choice = s_fallback;
    __syncthreads();
  }
  if (t == 0) {
    if (write_token) ids[(int64_t)b * ids_ld + cur_len] = (int64_t)s_choice;
    if (pos_ptr_adv != nullptr) {
      __threadfence();
      const int fin = atomicAdd(ticket, 1);
      if (fin == (int)gridDim.x - 1) {
        *ticket = 0;
        *pos_ptr_adv += 1;
      }
    }
  }
}

}  // namespace i2t

using namespace i2t;

extern "C" int i2t_sample(float* logits, int64_t ldl, int64_t B, int64_t V, int64_t* ids, int64_t ids_ld,
                          int32_t* pos_ptr, int advance_pos, int64_t cur_len, float temperature, int64_t top_k,
                          const int32_t* ngrams, int64_t n_ngrams, uint64_t seed, const uint64_t* seed_ptr, float* probs_out,
                          int32_t* ticket, int write_token, void* stream) {
  I2T_REQUIRE(logits && ids && B > 0 && V > 1, "sample: bad arguments");
  I2T_REQUIRE(temperature > 0.f, "sample: temperature must be positive");
  I2T_REQUIRE(n_ngrams == 0 || ngrams, "sample: n-gram list missing");
  I2T_REQUIRE(!advance_pos || (pos_ptr && ticket), "sample: advancing the position needs pos_ptr and a ticket counter");
  I2T_REQUIRE(pos_ptr || cur_len > 0, "sample: need a position");
  const size_t row_bytes = (size_t)V * sizeof(float);
  if (row_bytes <= 208 * 1024) {   // the row fits in shared memory next to ~17 KB of static scratch
    static std::atomic<size_t> attr{0};
    if (row_bytes > attr.load()) {
      I2T_CUDA(cudaFuncSetAttribute(sample_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_bytes));
      attr.store(row_bytes);
    }
    sample_smem_kernel<<<(unsigned)B, SAMP_THREADS, row_bytes, (cudaStream_t)stream>>>(
        logits, ldl, (int)V, ids, ids_ld, pos_ptr, advance_pos ? pos_ptr : nullptr, (int)cur_len, temperature,
        (int)(top_k > 0 ? top_k : 0), ngrams, (int)n_ngrams, seed, seed_ptr, probs_out, ticket, write_token);
  } else {
    sample_kernel<<<(unsigned)B, SAMP_THREADS, 0, (cudaStream_t)stream>>>(
        logits, ldl, (int)V, ids, ids_ld, pos_ptr, advance_pos ? pos_ptr : nullptr, (int)cur_len, temperature,
        (int)(top_k > 0 ? top_k : 0), ngrams, (int)n_ngrams, seed, seed_ptr, probs_out, ticket, write_token);
  }
  I2T_LAUNCHED();
  return I2T_OK;
}
