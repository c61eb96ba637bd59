// Device-side next-token sampler shared by the stand-alone kernel (sampler.cu) and the decode megakernel.
// temperature -> no-repeat-n-gram ban -> top-k threshold (ties kept) -> softmax -> multinomial draw -> append.
// Replaces reference models/vision_encoder_decoder.py:152-180 and transformers' NoRepeatNGramLogitsProcessor
// (generation/logits_process.py:1012-1135).
#pragma once
#include "common.cuh"

namespace i2t {

constexpr int SAMP_THREADS = 1024;     // stand-alone kernel block size
constexpr int SAMP_MAX_BANNED = 1024;
constexpr int SAMP_HGROUPS = 8;

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

// scratch every sampling CTA needs next to the vocabulary row
struct SampleScratch {
  int banned[SAMP_MAX_BANNED];
  uint32_t ghist[SAMP_HGROUPS][256];
  float mhist[256];            // nucleus: probability mass per radix bin
  double redd[32];
  double scan[32];
  float redf[32];
  int nbanned;
  uint32_t prefix, kleft;
  float nuc_above;             // nucleus: unnormalised mass of the keys above the boundary group
  int nuc_found, nuc_first;
  int choice, fallback;
};

// Barrier over the `nthreads` participating threads (named barrier 2, so a megakernel's other warps are not involved).
__device__ __forceinline__ void samp_sync(int nthreads) { asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); }

template <typename T>
__device__ __forceinline__ T samp_block_reduce(T v, T* scratch, bool is_max, int t, int nthreads) {
  const int lane = t & 31, w = t >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? (other > v ? other : v) : v + other;
  }
  samp_sync(nthreads);
  if (lane == 0) scratch[w] = v;
  samp_sync(nthreads);
  T r = scratch[0];
  for (int i = 1; i < (nthreads >> 5); ++i) r = is_max ? (scratch[i] > r ? scratch[i] : r) : r + scratch[i];
  return r;
}

// One vocabulary row, resident in shared memory `sv` (V floats).  Called by `nthreads` threads (multiple of 32, all
// warps complete), thread index t in [0, nthreads).  Returns the chosen token in every thread.
// `row` is the global logits row (read once with ld.global.cg; written back scaled/banned when modified).
__device__ __forceinline__ int sample_row_smem(float* __restrict__ sv, SampleScratch& S, float* __restrict__ row, int V,
                                               const int64_t* __restrict__ idr, int cur_len, float temperature,
                                               int top_k, const int32_t* __restrict__ ngrams, int n_ngrams,
                                               uint64_t seed, int row_index, float* __restrict__ probs_row, int t,
                                               int nthreads, float nucleus_p = 0.f) {
  const int lane = t & 31, w = t >> 5;
  if (t == 0) { S.nbanned = 0; S.choice = 0x7fffffff; S.fallback = 0x7fffffff; }
  samp_sync(nthreads);
  // ---- one pass over global memory: logits / temperature (a true division, like the reference) ----
  for (int i = t; i < V; i += nthreads) sv[i] = __ldcg(row + i) / temperature;
  samp_sync(nthreads);
  // ---- banned tokens (generation/logits_process.py:1012-1076): -inf straight into the resident row -- no list, no cap ----
  for (int g = 0; g < n_ngrams; ++g) {
    const int n = ngrams[g];
    if (n <= 0 || cur_len + 1 < n) continue;
    const int tail = cur_len + 1 - n;              // start of the (n-1)-token suffix
    for (int i = t; i <= cur_len - n; i += nthreads) {
      bool same = true;
      for (int j = 0; j < n - 1; ++j) same = same && (__ldcg(idr + i + j) == __ldcg(idr + tail + j));
      if (same) {
        const int tok = (int)__ldcg(idr + i + n - 1);
        if (tok >= 0 && tok < V) {
          sv[tok] = -INFINITY;
          S.nbanned = 1;                           // (benign race: every writer stores 1)
        }
      }
    }
  }
  samp_sync(nthreads);
  const int nb = S.nbanned;
  if (temperature != 1.0f || nb > 0) {             // documented in-place contract: scaled / banned logits
    for (int i = t; i < V; i += nthreads) row[i] = sv[i];
  }
  float mx = -INFINITY;
  for (int i = t; i < V; i += nthreads) mx = fmaxf(mx, sv[i]);
  mx = samp_block_reduce<float>(mx, S.redf, true, t, nthreads);

  // ---- top-k threshold: exact k-th largest by radix select on order-preserving keys; greedy = row maximum ----
  float thr = -INFINITY;
  if (top_k == 1) {
    thr = mx;
  } else if (top_k > 1 && top_k < V) {
    if (t == 0) { S.prefix = 0u; S.kleft = (uint32_t)top_k; }
    samp_sync(nthreads);
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = t; i < SAMP_HGROUPS * 256; i += nthreads) (&S.ghist[0][0])[i] = 0u;
      samp_sync(nthreads);
      const uint32_t prefix = S.prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      uint32_t* hist = S.ghist[w & (SAMP_HGROUPS - 1)];
      for (int i = t; i < V; i += nthreads) {
        const uint32_t key = float_key(sv[i]);
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
      samp_sync(nthreads);
      if (t < 256) {
        uint32_t tot = 0;
#pragma unroll
        for (int i = 0; i < SAMP_HGROUPS; ++i) tot += S.ghist[i][t];
        S.ghist[0][t] = tot;
      }
      samp_sync(nthreads);
      if (w == 0) {
        // which of the 256 bins holds the kleft-th largest key: warp 0 does the top-down scan in parallel (lane l owns bins
        // 8l .. 8l+7, a shuffle suffix-sum gives the count above them); a serial scan by one thread cost ~4 us per pass
        uint32_t hb[8], own = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          hb[i] = S.ghist[0][8 * lane + i];
          own += hb[i];
        }
        uint32_t suf = own;                                   // keys in this lane's bins and every higher bin
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t up = __shfl_down_sync(0xffffffffu, suf, o);
          if (lane + o < 32) suf += up;
        }
        const uint32_t above = suf - own, left = S.kleft;
        const bool mine = (above < left && left <= suf) || (lane == 0 && left > suf);      // second term: cannot happen (k <= count)
        __syncwarp();
        if (mine) {
          uint32_t l2 = left - above;
          int bsel = 7;
          for (; bsel > 0; --bsel) {
            if (hb[bsel] >= l2) break;
            l2 -= hb[bsel];
          }
          S.kleft = l2;
          S.prefix = prefix | ((uint32_t)(8 * lane + bsel) << shift);
        }
      }
      samp_sync(nthreads);
    }
    const uint32_t kk = S.prefix;
    thr = __uint_as_float((kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk);
  }

  // ---- softmax mass of the survivors and the draw: inverse CDF in INDEX order (thread t owns a contiguous chunk), so the
  //      chosen token does not depend on how many threads run the sampler ----
  const int chunk = (V + nthreads - 1) / nthreads;
  const int beg = min(V, t * chunk), end = min(V, beg + chunk);
  // nucleus filter state (models/vision_encoder_decoder.py:160-172): a survivor stays iff its logit is ABOVE xk, or it is
  // the designated first element of a tie group at the very top (the reference keeps exactly the first sorted entry then)
  float xk = -INFINITY;
  int tie_first = -1;
  auto keep = [&](int i, float x) { return x >= thr && (x > xk || i == tie_first); };
  double part = 0.0;
  for (int i = beg; i < end; ++i) {
    const float x = sv[i];
    if (keep(i, x)) part += (double)expf(x - mx);
  }
  double total = samp_block_reduce<double>(part, S.redd, false, t, nthreads);
  if (nucleus_p > 0.f && nucleus_p < 1.f) {
    // sorted descending, keep i iff cumsum_i <= max(p, p_max): find the boundary key by a radix descent over the
    // order-preserving keys with per-bin probability MASS (unnormalised: p_max = 1, budget = max(p * total, 1))
    const float budget = fmaxf(nucleus_p * (float)total, 1.0f);
    if (t == 0) { S.prefix = 0u; S.nuc_above = 0.f; S.nuc_found = 0; S.nuc_first = 0x7fffffff; }
    samp_sync(nthreads);
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = t; i < 256; i += nthreads) S.mhist[i] = 0.f;
      samp_sync(nthreads);
      const uint32_t prefix = S.prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      const bool live = pass == 0 || S.nuc_found;          // once every survivor fits there is nothing to refine
      if (live) {
        for (int i = t; i < V; i += nthreads) {
          const float x = sv[i];
          if (x >= thr) {
            const uint32_t key = float_key(x);
            if ((key & mask) == prefix) atomicAdd(&S.mhist[(key >> shift) & 255u], expf(x - mx));
          }
        }
      }
      samp_sync(nthreads);
      if (t == 0 && live) {
        float acc = S.nuc_above;
        int bin = 255, found = 0;
        for (; bin >= 0; --bin) {
          if (acc + S.mhist[bin] > budget) { found = 1; break; }
          acc += S.mhist[bin];
        }
        if (!found && pass > 0) {      // rounding: the sub-bins sum to less than their parent bin did -> lowest sub-bin is the boundary
          found = 1;
          bin = 0;
          acc -= S.mhist[0];
        }
        S.nuc_found = found;
        if (found) {
          S.nuc_above = acc;
          S.prefix = prefix | ((uint32_t)bin << shift);
        }
      }
      samp_sync(nthreads);
    }
    if (S.nuc_found) {
      const uint32_t kk = S.prefix;                         // key of the boundary group: it does not fit -> dropped
      xk = __uint_as_float((kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk);
      if (S.nuc_above == 0.f) {                             // the boundary group IS the top (ties at the maximum): keep its first
        int best = 0x7fffffff;
        for (int i = t; i < V; i += nthreads)
          if (sv[i] == xk) best = min(best, i);
        if (best != 0x7fffffff) atomicMin(&S.nuc_first, best);
        samp_sync(nthreads);
        tie_first = S.nuc_first;
      }
      part = 0.0;
      for (int i = beg; i < end; ++i) {
        const float x = sv[i];
        if (keep(i, x)) part += (double)expf(x - mx);
      }
      total = samp_block_reduce<double>(part, S.redd, false, t, nthreads);
    }
  }
  if (probs_row != nullptr) {
    for (int i = t; i < V; i += nthreads) {
      const float x = sv[i];
      probs_row[i] = keep(i, x) ? (float)((double)expf(x - mx) / total) : 0.f;
    }
  }
  const double target = (double)philox_uniform(seed, (uint32_t)row_index, (uint32_t)cur_len) * total;
  double incl = part;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double nbr = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += nbr;
  }
  if (lane == 31) S.scan[w] = incl;
  samp_sync(nthreads);
  double woff = 0.0;
  for (int i = 0; i < w; ++i) woff += S.scan[i];
  const double excl = woff + incl - part;
  if (part > 0.0 && target > excl && target <= excl + part) {
    double run = excl;
    int pick = -1;
    for (int i = beg; i < end; ++i) {
      const float x = sv[i];
      if (keep(i, x)) {
        run += (double)expf(x - mx);
        pick = i;
        if (run >= target) break;
      }
    }
    if (pick >= 0) atomicMin(&S.choice, pick);
  }
  samp_sync(nthreads);
  if (S.choice == 0x7fffffff) {    // rounding corner: the target landed past the last survivor's prefix
    int best = 0x7fffffff;
    for (int i = t; i < V; i += nthreads)
      if (sv[i] == mx) best = min(best, i);
    if (best != 0x7fffffff) atomicMin(&S.fallback, best);
    samp_sync(nthreads);
    if (t == 0) S.choice = S.fallback;
    samp_sync(nthreads);
  }
  return S.choice;
}

}  // namespace i2t
