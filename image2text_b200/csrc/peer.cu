// PEER tail core (product-key expert retrieval): reference models/layers.py:73-109 (PeerLookup.forward) after its four
// dense projections, i.e. everything that is not a GEMM:
//   top-k of the left / right query-unit scores (:31-34), the top-k of their k x k sums (:84-87), softmax (:88), the expert
//   ids `left * topk + right` (:90-96, restated literally), the gather of emb_in / emb_out rows (:98-99), the k dot products
//   with the key projection (:101), GELU(tanh) (:102) and the score-weighted sum of the output experts (:104-108).
// One CTA per (batch, slot) row; the heads run one after the other so that the 1 x out row is written once.
// fp32 throughout: the selection is discrete, a bf16 score could pick another expert (same reasoning as the LSH tail).
// Memory-bound and tiny (B * n_cls rows): latency matters, not throughput.
#include "common.cuh"

namespace i2t {

constexpr int PEER_THREADS = 256;
constexpr int PEER_MAX_K = 16;        // topk
constexpr int PEER_MAX_UNITS = 1024;  // sqrt(num_units): scores per query unit

__device__ __forceinline__ float peer_gelu(float x) { return gelu_tanh_f(x); }
__device__ __forceinline__ float peer_gelu_grad(float x) { return act_grad(x, I2T_ACT_GELU_TANH); }

// the k largest of vals[0..n) in descending order (ties: lower index first, like a stable sort); n <= PEER_MAX_UNITS.
// All PEER_THREADS threads call it; results in out_v / out_i (shared memory).  `scratch` holds n floats and is destroyed.
__device__ void peer_topk(float* scratch, int n, int k, float* out_v, int* out_i, float* red_v, int* red_i) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (int r = 0; r < k; ++r) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = t; i < n; i += PEER_THREADS) {
      const float v = scratch[i];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[w] = bv; red_i[w] = bi; }
    __syncthreads();
    if (t == 0) {
      for (int j = 1; j < PEER_THREADS / 32; ++j)
        if (red_v[j] > bv || (red_v[j] == bv && red_i[j] < bi)) { bv = red_v[j]; bi = red_i[j]; }
      out_v[r] = bv;
      out_i[r] = bi;
      if (bi != 0x7fffffff) scratch[bi] = -INFINITY;
    }
    __syncthreads();
  }
}

struct PeerShared {
  float scores[PEER_MAX_UNITS];
  float lv[PEER_MAX_K], rv[PEER_MAX_K], cv[PEER_MAX_K];
  int li[PEER_MAX_K], ri[PEER_MAX_K], ci[PEER_MAX_K];
  float cross[PEER_MAX_K * PEER_MAX_K];
  float red_v[PEER_THREADS / 32];
  int red_i[PEER_THREADS / 32];
  float w[PEER_MAX_K], score[PEER_MAX_K], dot[PEER_MAX_K];
  int idx[PEER_MAX_K];
};

// ql / qr: (M, H, U) scores of the left / right query units; key: (M, H, D) key projection; emb_in (E, D), emb_out (E, O);
// out (M, O) = sum over heads and experts (the caller adds the residual projection).  Saved for the backward pass (all
// (M, H, K)): expert id, softmax score, in_dot, and the positions of the chosen left / right scores inside ql / qr.
__global__ void __launch_bounds__(PEER_THREADS)
peer_lookup_fwd_kernel(const float* __restrict__ ql, const float* __restrict__ qr, const float* __restrict__ key,
                       const float* __restrict__ emb_in, const float* __restrict__ emb_out, float* __restrict__ out,
                       int32_t* __restrict__ s_idx, float* __restrict__ s_score, float* __restrict__ s_dot,
                       int32_t* __restrict__ s_lpos, int32_t* __restrict__ s_rpos, int H, int U, int K, int D, int O) {
  __shared__ PeerShared S;
  const int m = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  float acc[8];                                   // out columns t, t + 256, ... (O <= 2048)
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int h = 0; h < H; ++h) {
    const int64_t mh = (int64_t)m * H + h;
    for (int i = t; i < U; i += PEER_THREADS) S.scores[i] = ql[mh * U + i];
    __syncthreads();
    peer_topk(S.scores, U, K, S.lv, S.li, S.red_v, S.red_i);
    for (int i = t; i < U; i += PEER_THREADS) S.scores[i] = qr[mh * U + i];
    __syncthreads();
    peer_topk(S.scores, U, K, S.rv, S.ri, S.red_v, S.red_i);
    for (int i = t; i < K * K; i += PEER_THREADS) S.cross[i] = S.lv[i / K] + S.rv[i % K];
    __syncthreads();
    peer_topk(S.cross, K * K, K, S.cv, S.ci, S.red_v, S.red_i);
    if (t == 0) {                                 // softmax over the k chosen sums; expert ids
      float mx = S.cv[0], sum = 0.f;
      for (int k = 0; k < K; ++k) { S.score[k] = expf(S.cv[k] - mx); sum += S.score[k]; }
      for (int k = 0; k < K; ++k) {
        S.score[k] /= sum;
        const int lp = S.li[S.ci[k] / K], rp = S.ri[S.ci[k] % K];
        S.idx[k] = lp * K + rp;                   // models/layers.py:93-96 (not lp * U + rp)
        s_lpos[mh * K + k] = lp;
        s_rpos[mh * K + k] = rp;
        s_idx[mh * K + k] = S.idx[k];
        s_score[mh * K + k] = S.score[k];
      }
    }
    __syncthreads();
    // in_dot[k] = emb_in[idx_k] . key[m, h]: warp w takes experts w, w + 8, ...
    const float* kp = key + mh * D;
    for (int k = w; k < K; k += PEER_THREADS / 32) {
      const float* ep = emb_in + (int64_t)S.idx[k] * D;
      float d = 0.f;
      for (int i = lane; i < D; i += 32) d = fmaf(ep[i], kp[i], d);
      d = warp_sum(d);
      if (lane == 0) {
        S.dot[k] = d;
        S.w[k] = S.score[k] * peer_gelu(d);
        s_dot[mh * K + k] = d;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = t + j * PEER_THREADS;
      if (e < O) {
        float a = acc[j];
        for (int k = 0; k < K; ++k) a = fmaf(S.w[k], emb_out[(int64_t)S.idx[k] * O + e], a);
        acc[j] = a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = t + j * PEER_THREADS;
    if (e < O) out[(int64_t)m * O + e] = acc[j];
  }
}

// dout (M, O) -> dql, dqr (M, H, U; dense rows, zero outside the chosen positions), dkey (M, H, D), and ACCUMULATED (atomics)
// dense gradients of the two expert tables (the reference's nn.Embedding gradients are dense too).
__global__ void __launch_bounds__(PEER_THREADS)
peer_lookup_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ key, const float* __restrict__ emb_in,
                       const float* __restrict__ emb_out, const int32_t* __restrict__ s_idx, const float* __restrict__ s_score,
                       const float* __restrict__ s_dot, const int32_t* __restrict__ s_lpos, const int32_t* __restrict__ s_rpos,
                       float* __restrict__ dql, float* __restrict__ dqr, float* __restrict__ dkey, float* __restrict__ demb_in,
                       float* __restrict__ demb_out, int H, int U, int K, int D, int O) {
  __shared__ float dw[PEER_MAX_K], ddot_in[PEER_MAX_K], dsel[PEER_MAX_K], wgt[PEER_MAX_K];
  __shared__ int idx[PEER_MAX_K];
  const int m = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  const float* dop = dout + (int64_t)m * O;
  for (int h = 0; h < H; ++h) {
    const int64_t mh = (int64_t)m * H + h;
    if (t < K) idx[t] = s_idx[mh * K + t];
    for (int i = t; i < U; i += PEER_THREADS) { dql[mh * U + i] = 0.f; dqr[mh * U + i] = 0.f; }
    __syncthreads();
    // dw[k] = dout . emb_out[idx_k]
    for (int k = w; k < K; k += PEER_THREADS / 32) {
      const float* ep = emb_out + (int64_t)idx[k] * O;
      float d = 0.f;
      for (int i = lane; i < O; i += 32) d = fmaf(ep[i], dop[i], d);
      d = warp_sum(d);
      if (lane == 0) dw[k] = d;
    }
    __syncthreads();
    if (t == 0) {
      float dsc[PEER_MAX_K], inner = 0.f;
      for (int k = 0; k < K; ++k) {
        const float sc = s_score[mh * K + k], dt = s_dot[mh * K + k], a = peer_gelu(dt);
        wgt[k] = sc * a;
        dsc[k] = dw[k] * a;                            // d loss / d score_k
        ddot_in[k] = dw[k] * sc * peer_gelu_grad(dt);  // d loss / d in_dot_k
        inner += sc * dsc[k];
      }
      for (int k = 0; k < K; ++k) dsel[k] = s_score[mh * K + k] * (dsc[k] - inner);   // softmax backward -> the chosen sums
      for (int k = 0; k < K; ++k) {                    // the same left / right score may be part of several chosen sums
        dql[mh * U + s_lpos[mh * K + k]] += dsel[k];
        dqr[mh * U + s_rpos[mh * K + k]] += dsel[k];
      }
    }
    __syncthreads();
    const float* kp = key + mh * D;
    for (int i = t; i < D; i += PEER_THREADS) {        // dkey and the emb_in rows
      float a = 0.f;
      const float kv = kp[i];
      for (int k = 0; k < K; ++k) {
        a = fmaf(ddot_in[k], emb_in[(int64_t)idx[k] * D + i], a);
        atomicAdd(demb_in + (int64_t)idx[k] * D + i, ddot_in[k] * kv);
      }
      dkey[mh * D + i] = a;
    }
    for (int i = t; i < O; i += PEER_THREADS) {        // the emb_out rows
      const float g = dop[i];
      for (int k = 0; k < K; ++k) atomicAdd(demb_out + (int64_t)idx[k] * O + i, wgt[k] * g);
    }
    __syncthreads();
  }
}

}  // namespace i2t

using namespace i2t;

static int peer_check(int64_t M, int64_t H, int64_t U, int64_t K, int64_t D, int64_t O) {
  I2T_REQUIRE(M > 0 && H > 0 && D > 0 && O > 0, "peer_lookup: bad sizes");
  I2T_REQUIRE(K >= 1 && K <= PEER_MAX_K && K <= U && U <= PEER_MAX_UNITS, "peer_lookup: topk %lld / %lld scores per unit outside 1..%d / %d",
              (long long)K, (long long)U, PEER_MAX_K, PEER_MAX_UNITS);
  I2T_REQUIRE(O <= 8 * PEER_THREADS, "peer_lookup: out_features %lld above %d", (long long)O, 8 * PEER_THREADS);
  return I2T_OK;
}

extern "C" int i2t_peer_lookup_fwd(const float* ql, const float* qr, const float* key, const float* emb_in, const float* emb_out,
                                   float* out, int32_t* s_idx, float* s_score, float* s_dot, int32_t* s_lpos, int32_t* s_rpos,
                                   int64_t M, int64_t H, int64_t U, int64_t K, int64_t D, int64_t O, void* stream) {
  I2T_REQUIRE(ql && qr && key && emb_in && emb_out && out && s_idx && s_score && s_dot && s_lpos && s_rpos, "peer_lookup_fwd: null pointer");
  if (int rc = peer_check(M, H, U, K, D, O)) return rc;
  peer_lookup_fwd_kernel<<<(unsigned)M, PEER_THREADS, 0, (cudaStream_t)stream>>>(ql, qr, key, emb_in, emb_out, out, s_idx, s_score, s_dot,
                                                                                   s_lpos, s_rpos, (int)H, (int)U, (int)K, (int)D, (int)O);
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_peer_lookup_bwd(const float* dout, const float* key, const float* emb_in, const float* emb_out, const int32_t* s_idx,
                                   const float* s_score, const float* s_dot, const int32_t* s_lpos, const int32_t* s_rpos, float* dql,
                                   float* dqr, float* dkey, float* demb_in, float* demb_out, int64_t M, int64_t H, int64_t U,
                                   int64_t K, int64_t D, int64_t O, void* stream) {
  I2T_REQUIRE(dout && key && emb_in && emb_out && s_idx && s_score && s_dot && s_lpos && s_rpos && dql && dqr && dkey && demb_in && demb_out,
              "peer_lookup_bwd: null pointer");
  if (int rc = peer_check(M, H, U, K, D, O)) return rc;
  peer_lookup_bwd_kernel<<<(unsigned)M, PEER_THREADS, 0, (cudaStream_t)stream>>>(dout, key, emb_in, emb_out, s_idx, s_score, s_dot, s_lpos,
                                                                                   s_rpos, dql, dqr, dkey, demb_in, demb_out, (int)H, (int)U,
                                                                                   (int)K, (int)D, (int)O);
  I2T_LAUNCHED();
  return I2T_OK;
}
