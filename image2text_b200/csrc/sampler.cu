// Stand-alone on-device sampler kernel: one CTA (1024 threads) per batch row; see sampler.cuh for the algorithm.
//   * the vocabulary row is read from global memory ONCE into shared memory when it fits (V * 4 B <= 208 KB, e.g.
//     GPT-2's 50257); every later pass (radix select, max, softmax mass, draw) runs on shared memory;
//   * larger vocabularies run the same code with the global row itself as the working copy (re-read from L2);
//   * top_k = 1 is greedy decoding (the reference has no other greedy switch);
//   * the last CTA to finish advances the device-side position counter (one CUDA graph serves every step).
#include "sampler.cuh"

namespace i2t {

__global__ void __launch_bounds__(SAMP_THREADS)
sample_kernel(float* __restrict__ logits, int64_t ldl, int V, int64_t* __restrict__ ids, int64_t ids_ld,
              const int32_t* pos_ptr_c, int32_t* pos_ptr_adv, int cur_len_const, float temperature, int top_k, float nucleus_p,
              const int32_t* __restrict__ ngrams, int n_ngrams, uint64_t seed_arg, const uint64_t* __restrict__ seed_ptr,
              float* __restrict__ probs_out, int32_t* __restrict__ ticket, int write_token, int use_smem) {
  extern __shared__ __align__(16) float sv_smem[];
  __shared__ SampleScratch S;
  const int t = threadIdx.x;
  const int b = blockIdx.x;
  const uint64_t seed = seed_ptr != nullptr ? *seed_ptr : seed_arg;
  const int cur_len = pos_ptr_c != nullptr ? (*pos_ptr_c + 1) : cur_len_const;
  float* row = logits + (int64_t)b * ldl;
  float* sv = use_smem ? sv_smem : row;
  const int choice = sample_row_smem(sv, S, row, V, ids + (int64_t)b * ids_ld, cur_len, temperature, top_k, ngrams, n_ngrams,
                                     seed, b, probs_out ? probs_out + (int64_t)b * V : nullptr, t, SAMP_THREADS, nucleus_p);
  if (t == 0) {
    if (write_token) ids[(int64_t)b * ids_ld + cur_len] = (int64_t)choice;
    if (pos_ptr_adv != nullptr) {
      __threadfence();
      const int fin = atomicAdd(ticket, 1);
      if (fin == (int)gridDim.x - 1) {  // last row: every CTA has read *pos_ptr already
        *ticket = 0;
        *pos_ptr_adv += 1;
      }
    }
  }
}

// Greedy fast path (top_k = 1, no nucleus filter, no probability output): the pick is the arg-max of the un-banned logits,
// so the row never needs to be resident -- one vectorised pass over global memory per row, no shared-memory copy, several
// rows per SM.  (The general kernel spends ~30 us per row in its shared-memory passes: 150 us per step at 512 sequences.)
// Same contract otherwise: banned positions are written to the row as -inf; a positive temperature does not change the
// arg-max and is not applied to the stored logits here.  Ties: the lowest token id wins.
constexpr int GREEDY_THREADS = 512;
__global__ void __launch_bounds__(GREEDY_THREADS)
sample_greedy_kernel(float* __restrict__ logits, int64_t ldl, int V, int64_t* __restrict__ ids, int64_t ids_ld,
                     const int32_t* pos_ptr_c, int32_t* pos_ptr_adv, int cur_len_const, const int32_t* __restrict__ ngrams,
                     int n_ngrams, int32_t* __restrict__ ticket, int write_token) {
  __shared__ unsigned long long best[GREEDY_THREADS / 32];
  const int t = threadIdx.x, b = blockIdx.x, lane = t & 31, w = t >> 5;
  const int cur_len = pos_ptr_c != nullptr ? (*pos_ptr_c + 1) : cur_len_const;
  float* row = logits + (int64_t)b * ldl;
  const int64_t* idr = ids + (int64_t)b * ids_ld;
  for (int g = 0; g < n_ngrams; ++g) {               // generation/logits_process.py:1012-1076, as in sample_row_smem
    const int n = ngrams[g];
    if (n <= 0 || cur_len + 1 < n) continue;
    const int tail = cur_len + 1 - n;
    for (int i = t; i <= cur_len - n; i += GREEDY_THREADS) {
      bool same = true;
      for (int j = 0; j < n - 1; ++j) same = same && (__ldcg(idr + i + j) == __ldcg(idr + tail + j));
      if (same) {                                     // banned straight in the row: no list, so no cap on how many
        const int tok = (int)__ldcg(idr + i + n - 1);
        if (tok >= 0 && tok < V) row[tok] = -INFINITY;
      }
    }
  }
  __syncthreads();                                    // the -inf stores are visible to this CTA's loads below
  // packed key: order-preserving float bits in the high word, ~index in the low word -> max = largest logit, lowest index
  unsigned long long bk = 0ull;
  const bool vec = (reinterpret_cast<uintptr_t>(row) & 15u) == 0;
  const int nv = vec ? V / 4 : 0;
  for (int i = t; i < nv; i += GREEDY_THREADS) {
    const float4 x = __ldcg(reinterpret_cast<const float4*>(row) + i);
    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const unsigned long long k = ((unsigned long long)float_key(xs[e]) << 32) | (uint32_t)(~(uint32_t)(4 * i + e));
      bk = k > bk ? k : bk;
    }
  }
  for (int i = nv * 4 + t; i < V; i += GREEDY_THREADS) {
    const unsigned long long k = ((unsigned long long)float_key(__ldcg(row + i)) << 32) | (uint32_t)(~(uint32_t)i);
    bk = k > bk ? k : bk;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, bk, o);
    bk = other > bk ? other : bk;
  }
  if (lane == 0) best[w] = bk;
  __syncthreads();
  if (t == 0) {
    for (int i = 1; i < GREEDY_THREADS / 32; ++i) bk = best[i] > bk ? best[i] : bk;
    const int choice = (int)(~(uint32_t)(bk & 0xffffffffull));
    if (write_token) ids[(int64_t)b * ids_ld + cur_len] = (int64_t)choice;
    if (pos_ptr_adv != nullptr) {
      __threadfence();
      const int fin = atomicAdd(ticket, 1);
      if (fin == (int)gridDim.x - 1) {  // last row: every CTA has read *pos_ptr already
        *ticket = 0;
        *pos_ptr_adv += 1;
      }
    }
  }
}

}  // namespace i2t

using namespace i2t;

static std::atomic<int> g_greedy_fast{1};
extern "C" void i2t_set_sampler_greedy_fast_path(int enabled) { g_greedy_fast.store(enabled ? 1 : 0); }

extern "C" int i2t_sample(float* logits, int64_t ldl, int64_t B, int64_t V, int64_t* ids, int64_t ids_ld,
                          int32_t* pos_ptr, int advance_pos, int64_t cur_len, float temperature, int64_t top_k, float nucleus_p,
                          const int32_t* ngrams, int64_t n_ngrams, uint64_t seed, const uint64_t* seed_ptr, float* probs_out,
                          int32_t* ticket, int write_token, void* stream) {
  I2T_REQUIRE(logits && ids && B > 0 && V > 1, "sample: bad arguments");
  I2T_REQUIRE(temperature > 0.f, "sample: temperature must be positive");
  I2T_REQUIRE(nucleus_p >= 0.f && nucleus_p <= 1.f, "sample: nucleus_p must be in [0, 1] (0 or 1 = off)");
  I2T_REQUIRE(n_ngrams == 0 || ngrams, "sample: n-gram list missing");
  I2T_REQUIRE(!advance_pos || (pos_ptr && ticket), "sample: advancing the position needs pos_ptr and a ticket counter");
  I2T_REQUIRE(pos_ptr || cur_len > 0, "sample: need a position");
  if (top_k == 1 && !(nucleus_p > 0.f && nucleus_p < 1.f) && probs_out == nullptr && g_greedy_fast.load() == 1) {
    sample_greedy_kernel<<<(unsigned)B, GREEDY_THREADS, 0, (cudaStream_t)stream>>>(
        logits, ldl, (int)V, ids, ids_ld, pos_ptr, advance_pos ? pos_ptr : nullptr, (int)cur_len, ngrams, (int)n_ngrams, ticket,
        write_token);
    I2T_LAUNCHED();
    return I2T_OK;
  }
  const size_t row_bytes = (size_t)V * sizeof(float);
  const int use_smem = row_bytes <= 208 * 1024 ? 1 : 0;
  const size_t smem = use_smem ? row_bytes : 0;
  static std::atomic<size_t> attr{0};
  if (smem > attr.load()) {
    I2T_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr.store(smem);
  }
  sample_kernel<<<(unsigned)B, SAMP_THREADS, smem, (cudaStream_t)stream>>>(
      logits, ldl, (int)V, ids, ids_ld, pos_ptr, advance_pos ? pos_ptr : nullptr, (int)cur_len, temperature,
      (int)(top_k > 0 ? top_k : 0), nucleus_p, ngrams, (int)n_ngrams, seed, seed_ptr, probs_out, ticket, write_token, use_smem);
  I2T_LAUNCHED();
  return I2T_OK;
}
