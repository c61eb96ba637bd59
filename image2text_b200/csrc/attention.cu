// Fused (flash-style) attention forward/backward with fp32 arithmetic on the FMA pipe.
//   forward : O = softmax(Q K^T / sqrt(hs) + mask) V, online softmax over 64-key tiles, never materialises
//             the (T,T) score matrix nor the additive mask the reference rebuilds per layer
//             (reference models/layers.py:581-596 + F.scaled_dot_product_attention at :465; torchvision
//             nn.MultiheadAttention at vision_transformer.py:113).
//   masks   : closed forms only (I2T_MASK_NONE / CAUSAL / PROMPT) -- see DESIGN.md "mask algebra" (D9: the
//             reference's user mask is a no-op, so nothing else can occur).  A row with no visible key -> 0.
//   backward: recomputes P from (Q,K,lse); one CTA per (b,h,64-key tile) accumulates dK,dV in registers and
//             atomically adds dQ (fp32 buffer zeroed by the caller through the host wrapper).
// This file is the fp32 parity path and the bf16-IO fallback; tensor-core tiles live in attention_tc.cu.
#include "common.cuh"
#include "rng.cuh"

namespace i2t {

constexpr int ATT_BQ = 64, ATT_BK = 64, ATT_THREADS = 128, ATT_PSTRIDE = 68;

__device__ __forceinline__ bool key_visible(int mode, int n_prompt, int qi, int kj) {
  if (mode == I2T_MASK_NONE) return true;
  if (kj > qi) return false;
  if (mode == I2T_MASK_CAUSAL) return true;
  return qi < n_prompt ? true : kj >= n_prompt;  // I2T_MASK_PROMPT
}

template <typename TIN, typename TOUT, int HS>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const TIN* __restrict__ q, const TIN* __restrict__ k, const TIN* __restrict__ v, TOUT* __restrict__ out,
                float* __restrict__ lse, int H, int Tq, int Tk, int64_t q_bs, int64_t q_rs, int64_t kv_bs,
                int64_t kv_rs, int mode, int n_prompt, float scale, DropArgs drop) {
  extern __shared__ __align__(16) float smem[];
  float* Qt = smem;                       // [HS][64]  (transposed, pre-scaled)
  float* Kt = Qt + HS * ATT_BQ;           // [HS][64]
  float* Vs = Kt + HS * ATT_BK;           // [64][HS]
  float* Ps = Vs + ATT_BK * HS;           // [64][ATT_PSTRIDE]
  constexpr int EC = HS / 8;              // output columns per thread
  const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
  const int q0 = blockIdx.x * ATT_BQ, h = blockIdx.y, b = blockIdx.z;
  const TIN* qb = q + (int64_t)b * q_bs + (int64_t)h * HS;
  const TIN* kb = k + (int64_t)b * kv_bs + (int64_t)h * HS;
  const TIN* vb = v + (int64_t)b * kv_bs + (int64_t)h * HS;
  DropKey dkey{0u, 0u, 0u};
  if (drop.thr != 0u) dkey = drop_key(drop);

  // load Q tile transposed: thread -> (row = t % 64, 4-wide e chunk = t / 64 + 2*i)
  for (int c = t >> 6; c < HS / 4; c += ATT_THREADS / 64) {
    const int r = t & 63;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < Tq) x = load4(qb + (int64_t)(q0 + r) * q_rs + c * 4);
    Qt[(c * 4 + 0) * ATT_BQ + r] = x.x * scale;
    Qt[(c * 4 + 1) * ATT_BQ + r] = x.y * scale;
    Qt[(c * 4 + 2) * ATT_BQ + r] = x.z * scale;
    Qt[(c * 4 + 3) * ATT_BQ + r] = x.w * scale;
  }

  float m_i[4], l_i[4], o[4][EC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_i[i] = -INFINITY;
    l_i[i] = 0.f;
#pragma unroll
    for (int e = 0; e < EC; ++e) o[i][e] = 0.f;
  }

  int kend = Tk;
  if (mode != I2T_MASK_NONE) {
    const int last_q = min(q0 + ATT_BQ, Tq) - 1;
    kend = min(Tk, last_q + 1);
  }
  for (int k0 = 0; k0 < kend; k0 += ATT_BK) {
    __syncthreads();  // previous tile fully consumed (also orders the Q stores on the first trip)
    for (int c = t >> 6; c < HS / 4; c += ATT_THREADS / 64) {
      const int r = t & 63;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
      if (k0 + r < Tk) {
        x = load4(kb + (int64_t)(k0 + r) * kv_rs + c * 4);
        y = load4(vb + (int64_t)(k0 + r) * kv_rs + c * 4);
      }
      Kt[(c * 4 + 0) * ATT_BK + r] = x.x;
      Kt[(c * 4 + 1) * ATT_BK + r] = x.y;
      Kt[(c * 4 + 2) * ATT_BK + r] = x.z;
      Kt[(c * 4 + 3) * ATT_BK + r] = x.w;
      *reinterpret_cast<float4*>(&Vs[r * HS + c * 4]) = y;
    }
    __syncthreads();

    // S tile: rows ty*4..+3, cols tx*8..+7
    float s[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int e = 0; e < HS; ++e) {
      const float4 a = *reinterpret_cast<const float4*>(&Qt[e * ATT_BQ + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Kt[e * ATT_BK + tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Kt[e * ATT_BK + tx * 8 + 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
    }
    // mask + online softmax (row statistics shared by the 8 lanes that own a row group)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kj = k0 + tx * 8 + j;
        if (kj >= Tk || !key_visible(mode, n_prompt, qi, kj)) s[i][j] = -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      const float m_new = fmaxf(m_i[i], mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = expf(m_i[i] - m_use);   // m_i = -inf -> 0
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[i][j] = expf(s[i][j] - m_use);          // masked -> exp(-inf) = 0
        rs += s[i][j];
      }
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);
      rs += __shfl_xor_sync(0xffffffffu, rs, 4);
      l_i[i] = l_i[i] * corr + rs;
      m_i[i] = m_new;
      if (drop.thr != 0u) {   // dropout on the probabilities: the denominator keeps every key, dropped terms leave the sum
        const uint32_t row = (uint32_t)(((int64_t)b * H + h) * Tq + qi);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!drop_attn_keep(drop, dkey, row, k0 + tx * 8 + j)) s[i][j] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < EC; ++e) o[i][e] *= corr;
      float* pr = Ps + (ty * 4 + i) * ATT_PSTRIDE + tx * 8;
      *reinterpret_cast<float4*>(pr) = make_float4(s[i][0], s[i][1], s[i][2], s[i][3]);
      *reinterpret_cast<float4*>(pr + 4) = make_float4(s[i][4], s[i][5], s[i][6], s[i][7]);
    }
    __syncthreads();
    // O += P V : rows ty*4..+3, cols tx*EC..+EC-1
#pragma unroll 4
    for (int j = 0; j < ATT_BK; j += 4) {
      float p[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 pv = *reinterpret_cast<const float4*>(&Ps[(ty * 4 + i) * ATT_PSTRIDE + j]);
        p[i][0] = pv.x; p[i][1] = pv.y; p[i][2] = pv.z; p[i][3] = pv.w;
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float vv[EC];
#pragma unroll
        for (int e = 0; e < EC; e += 4) {
          const float4 x = *reinterpret_cast<const float4*>(&Vs[(j + jj) * HS + tx * EC + e]);
          vv[e] = x.x; vv[e + 1] = x.y; vv[e + 2] = x.z; vv[e + 3] = x.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < EC; ++e) o[i][e] = fmaf(p[i][jj], vv[e], o[i][e]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty * 4 + i;
    if (qi >= Tq) continue;
    const float inv = l_i[i] > 0.f ? drop.inv_keep / l_i[i] : 0.f;
    TOUT* op = out + ((int64_t)b * Tq + qi) * ((int64_t)H * HS) + (int64_t)h * HS + tx * EC;
#pragma unroll
    for (int e = 0; e < EC; e += 4)
      store4(op + e, make_float4(o[i][e] * inv, o[i][e + 1] * inv, o[i][e + 2] * inv, o[i][e + 3] * inv));
    if (lse != nullptr && tx == 0)
      lse[((int64_t)b * H + h) * Tq + qi] = l_i[i] > 0.f ? m_i[i] + logf(l_i[i]) : -INFINITY;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward.  delta[b,h,i] = sum_e dO[i,e] * O[i,e] is computed by a small pre-kernel.  Main kernel: CTA =
// (64-key tile, h, b); loops over query tiles; P = exp(S - lse); dV += P^T dO; dP = dO V^T;
// dS = P * (dP - delta) ; dK += dS^T Q * scale ; dQ += dS K * scale (atomicAdd, fp32).
// ---------------------------------------------------------------------------------------------------------
template <typename T, int HS>
__global__ void __launch_bounds__(128) attn_delta_kernel(const T* __restrict__ out, const T* __restrict__ dout,
                                                         float* __restrict__ delta, int H, int Tq, int64_t rows) {
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 4 + warp;  // r = (b*Tq + i)*H + h
  if (r >= rows) return;
  const int64_t h = r % H, bi = r / H;
  const T* o = out + bi * (int64_t)H * HS + h * HS;
  const T* d = dout + bi * (int64_t)H * HS + h * HS;
  float s = 0.f;
  for (int e = lane; e < HS; e += 32) s += to_f32(o[e]) * to_f32(d[e]);
  s = warp_sum(s);
  if (lane == 0) {
    const int64_t b = bi / Tq, i = bi % Tq;
    delta[(b * H + h) * Tq + i] = s;
  }
}

template <typename T, int HS>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ dout,
                const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dq,
                T* __restrict__ dk, T* __restrict__ dv, int H, int Tq, int Tk, int64_t q_bs, int64_t q_rs, int64_t kv_bs,
                int64_t kv_rs, int mode, int n_prompt, float scale, DropArgs drop) {
  extern __shared__ __align__(16) float smem[];
  float* Kt = smem;                        // [HS][64]   K^T (key tile, fixed)
  float* Vt = Kt + HS * ATT_BK;            // [HS][64]   V^T
  float* Qt = Vt + HS * ATT_BK;            // [HS][64]   Q^T (query tile)
  float* dOt = Qt + HS * ATT_BQ;           // [HS][64]   dO^T
  float* Ps = dOt + HS * ATT_BQ;           // [64 q][ATT_PSTRIDE]  P then dS (row = query, col = key)
  float* Qs = Ps + ATT_BQ * ATT_PSTRIDE;   // [64 q][HS] row-major Q  (for dK)
  float* dOs = Qs + ATT_BQ * HS;           // [64 q][HS] row-major dO (for dV)
  float* Ks = dOs + ATT_BQ * HS;           // [64 k][HS] row-major K  (for dQ)
  constexpr int EC = HS / 8;
  const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
  const int k0 = blockIdx.x * ATT_BK, h = blockIdx.y, b = blockIdx.z;
  const T* qb = q + (int64_t)b * q_bs + (int64_t)h * HS;
  const T* kb = k + (int64_t)b * kv_bs + (int64_t)h * HS;
  const T* vb = v + (int64_t)b * kv_bs + (int64_t)h * HS;
  const T* dob = dout + (int64_t)b * Tq * ((int64_t)H * HS) + (int64_t)h * HS;
  const int64_t do_rs = (int64_t)H * HS;
  DropKey dkey{0u, 0u, 0u};
  if (drop.thr != 0u) dkey = drop_key(drop);

  for (int c = t >> 6; c < HS / 4; c += ATT_THREADS / 64) {
    const int r = t & 63;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
    if (k0 + r < Tk) {
      x = load4(kb + (int64_t)(k0 + r) * kv_rs + c * 4);
      y = load4(vb + (int64_t)(k0 + r) * kv_rs + c * 4);
    }
    Kt[(c * 4 + 0) * ATT_BK + r] = x.x; Kt[(c * 4 + 1) * ATT_BK + r] = x.y;
    Kt[(c * 4 + 2) * ATT_BK + r] = x.z; Kt[(c * 4 + 3) * ATT_BK + r] = x.w;
    Vt[(c * 4 + 0) * ATT_BK + r] = y.x; Vt[(c * 4 + 1) * ATT_BK + r] = y.y;
    Vt[(c * 4 + 2) * ATT_BK + r] = y.z; Vt[(c * 4 + 3) * ATT_BK + r] = y.w;
    *reinterpret_cast<float4*>(&Ks[r * HS + c * 4]) = x;
  }
  // accumulators: thread owns key rows ty*4..+3, e columns tx*EC..
  float adk[4][EC], adv[4][EC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < EC; ++e) adk[i][e] = adv[i][e] = 0.f;

  int qstart = 0;
  if (mode != I2T_MASK_NONE) qstart = (k0 / ATT_BQ) * ATT_BQ;  // queries before the key tile never see it
  for (int q0 = qstart; q0 < Tq; q0 += ATT_BQ) {
    __syncthreads();
    for (int c = t >> 6; c < HS / 4; c += ATT_THREADS / 64) {
      const int r = t & 63;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
      if (q0 + r < Tq) {
        x = load4(qb + (int64_t)(q0 + r) * q_rs + c * 4);
        y = load4(dob + (int64_t)(q0 + r) * do_rs + c * 4);
      }
      Qt[(c * 4 + 0) * ATT_BQ + r] = x.x; Qt[(c * 4 + 1) * ATT_BQ + r] = x.y;
      Qt[(c * 4 + 2) * ATT_BQ + r] = x.z; Qt[(c * 4 + 3) * ATT_BQ + r] = x.w;
      dOt[(c * 4 + 0) * ATT_BQ + r] = y.x; dOt[(c * 4 + 1) * ATT_BQ + r] = y.y;
      dOt[(c * 4 + 2) * ATT_BQ + r] = y.z; dOt[(c * 4 + 3) * ATT_BQ + r] = y.w;
      *reinterpret_cast<float4*>(&Qs[r * HS + c * 4]) = x;
      *reinterpret_cast<float4*>(&dOs[r * HS + c * 4]) = y;
    }
    __syncthreads();
    // S and dP for (query rows ty*4..+3, key cols tx*8..+7)
    float s[4][8], dp[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[i][j] = dp[i][j] = 0.f;
#pragma unroll 4
    for (int e = 0; e < HS; ++e) {
      const float4 a = *reinterpret_cast<const float4*>(&Qt[e * ATT_BQ + ty * 4]);
      const float4 g = *reinterpret_cast<const float4*>(&dOt[e * ATT_BQ + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Kt[e * ATT_BK + tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Kt[e * ATT_BK + tx * 8 + 4]);
      const float4 c0 = *reinterpret_cast<const float4*>(&Vt[e * ATT_BK + tx * 8]);
      const float4 c1 = *reinterpret_cast<const float4*>(&Vt[e * ATT_BK + tx * 8 + 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[i][j] = fmaf(av[i], bv[j], s[i][j]);
          dp[i][j] = fmaf(gv[i], cv[j], dp[i][j]);
        }
    }
    // P -> Ps (for dV), then dS -> Ps (for dK, dQ)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      float L = 0.f, D = 0.f;
      const bool qok = qi < Tq;
      if (qok) {
        L = lse[((int64_t)b * H + h) * Tq + qi];
        D = delta[((int64_t)b * H + h) * Tq + qi];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kj = k0 + tx * 8 + j;
        float p = 0.f;
        if (qok && kj < Tk && key_visible(mode, n_prompt, qi, kj) && L != -INFINITY) p = expf(s[i][j] * scale - L);
        float mk = 1.f;       // dropout multiplier of P[q][key]: dV sees P*mk, dP = (dO V^T)*mk, delta already is sum P*dP
        if (drop.thr != 0u && p != 0.f)
          mk = drop_attn_keep(drop, dkey, (uint32_t)(((int64_t)b * H + h) * Tq + qi), kj) ? drop.inv_keep : 0.f;
        s[i][j] = p * mk;
        dp[i][j] = p * (dp[i][j] * mk - D) * scale;  // dS (already carries the 1/sqrt(hs) of S = scale * q.k)
      }
      float* pr = Ps + (ty * 4 + i) * ATT_PSTRIDE + tx * 8;
      *reinterpret_cast<float4*>(pr) = make_float4(s[i][0], s[i][1], s[i][2], s[i][3]);
      *reinterpret_cast<float4*>(pr + 4) = make_float4(s[i][4], s[i][5], s[i][6], s[i][7]);
    }
    __syncthreads();
    // dV[key ty*4+i][e] += sum_q P[q][key] * dO[q][e]
#pragma unroll 4
    for (int qq = 0; qq < ATT_BQ; ++qq) {
      const float4 pc4 = *reinterpret_cast<const float4*>(&Ps[qq * ATT_PSTRIDE + ty * 4]);
      const float pcol[4] = {pc4.x, pc4.y, pc4.z, pc4.w};
#pragma unroll
      for (int e = 0; e < EC; e += 4) {
        const float4 x = *reinterpret_cast<const float4*>(&dOs[qq * HS + tx * EC + e]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          adv[i][e] = fmaf(pcol[i], x.x, adv[i][e]);
          adv[i][e + 1] = fmaf(pcol[i], x.y, adv[i][e + 1]);
          adv[i][e + 2] = fmaf(pcol[i], x.z, adv[i][e + 2]);
          adv[i][e + 3] = fmaf(pcol[i], x.w, adv[i][e + 3]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* pr = Ps + (ty * 4 + i) * ATT_PSTRIDE + tx * 8;
      *reinterpret_cast<float4*>(pr) = make_float4(dp[i][0], dp[i][1], dp[i][2], dp[i][3]);
      *reinterpret_cast<float4*>(pr + 4) = make_float4(dp[i][4], dp[i][5], dp[i][6], dp[i][7]);
    }
    __syncthreads();
    // dK[key ty*4+i][e] += sum_q dS[q][key] * Q[q][e]
#pragma unroll 4
    for (int qq = 0; qq < ATT_BQ; ++qq) {
      const float4 pc4 = *reinterpret_cast<const float4*>(&Ps[qq * ATT_PSTRIDE + ty * 4]);
      const float pcol[4] = {pc4.x, pc4.y, pc4.z, pc4.w};
#pragma unroll
      for (int e = 0; e < EC; e += 4) {
        const float4 x = *reinterpret_cast<const float4*>(&Qs[qq * HS + tx * EC + e]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          adk[i][e] = fmaf(pcol[i], x.x, adk[i][e]);
          adk[i][e + 1] = fmaf(pcol[i], x.y, adk[i][e + 1]);
          adk[i][e + 2] = fmaf(pcol[i], x.z, adk[i][e + 2]);
          adk[i][e + 3] = fmaf(pcol[i], x.w, adk[i][e + 3]);
        }
      }
    }
    // dQ[q ty*4+i][e] += sum_key dS[q][key] * K[key][e]   (atomic: other key tiles add to the same rows)
    {
      float adq[4][EC];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < EC; ++e) adq[i][e] = 0.f;
#pragma unroll 4
      for (int j = 0; j < ATT_BK; ++j) {
        float dsr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) dsr[i] = Ps[(ty * 4 + i) * ATT_PSTRIDE + j];
#pragma unroll
        for (int e = 0; e < EC; e += 4) {
          const float4 x = *reinterpret_cast<const float4*>(&Ks[j * HS + tx * EC + e]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            adq[i][e] = fmaf(dsr[i], x.x, adq[i][e]);
            adq[i][e + 1] = fmaf(dsr[i], x.y, adq[i][e + 1]);
            adq[i][e + 2] = fmaf(dsr[i], x.z, adq[i][e + 2]);
            adq[i][e + 3] = fmaf(dsr[i], x.w, adq[i][e + 3]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qi = q0 + ty * 4 + i;
        if (qi >= Tq) continue;
        float* dst = dq + (((int64_t)b * H + h) * Tq + qi) * HS + tx * EC;
#pragma unroll
        for (int e = 0; e < EC; ++e) atomicAdd(dst + e, adq[i][e]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kj = k0 + ty * 4 + i;
    if (kj >= Tk) continue;
    T* dkp = dk + (int64_t)b * kv_bs + (int64_t)kj * kv_rs + (int64_t)h * HS + tx * EC;
    T* dvp = dv + (int64_t)b * kv_bs + (int64_t)kj * kv_rs + (int64_t)h * HS + tx * EC;
#pragma unroll
    for (int e = 0; e < EC; e += 4) {
      store4(dkp + e, make_float4(adk[i][e], adk[i][e + 1], adk[i][e + 2], adk[i][e + 3]));
      store4(dvp + e, make_float4(adv[i][e], adv[i][e + 1], adv[i][e + 2], adv[i][e + 3]));
    }
  }
}

// dq_acc (B,H,Tq,HS) fp32 -> dq in the strided q layout
template <typename T, int HS>
__global__ void attn_dq_scatter_kernel(const float* __restrict__ acc, T* __restrict__ dq, int H, int Tq, int64_t q_bs,
                                       int64_t q_rs, int64_t total) {
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 each
  if (i >= total) return;
  const int64_t e4 = i % (HS / 4), r = i / (HS / 4);
  const int64_t qi = r % Tq, bh = r / Tq, h = bh % H, b = bh / H;
  const float4 x = *reinterpret_cast<const float4*>(acc + r * HS + e4 * 4);
  store4(dq + b * q_bs + qi * q_rs + h * HS + e4 * 4, x);
}

template <typename TIN, typename TOUT, int HS>
static int launch_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H,
                           int64_t Tq, int64_t Tk, int64_t q_bs, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, int mode,
                           int64_t n_prompt, DropArgs drop, cudaStream_t st) {
  const size_t smem = (size_t)(HS * ATT_BQ + HS * ATT_BK + ATT_BK * HS + ATT_BQ * ATT_PSTRIDE) * sizeof(float);
  auto kern = attn_fwd_kernel<TIN, TOUT, HS>;
  I2T_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(Tq, ATT_BQ), (unsigned)H, (unsigned)B);
  kern<<<grid, ATT_THREADS, smem, st>>>((const TIN*)q, (const TIN*)k, (const TIN*)v, (TOUT*)out, lse, (int)H, (int)Tq,
                                        (int)Tk, q_bs, q_rs, kv_bs, kv_rs, mode, (int)n_prompt,
                                        1.0f / sqrtf((float)HS), drop);
  I2T_LAUNCHED();
  return I2T_OK;
}

// attention_tc.cu: bf16 tensor-core forward (returns 1 when it handled the call, 0 when the shape is not eligible)
int attn_fwd_tc(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H, int64_t Tq, int64_t Tk,
                int64_t head_dim, int64_t q_bs, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, int mode, int64_t n_prompt,
                DropArgs drop, cudaStream_t st);
int attn_bwd_tc(const void* q, const void* k, const void* v, const void* dout, const float* lse, const float* delta, float* dq_acc,
                void* dk, void* dv, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_bs, int64_t q_rs,
                int64_t kv_bs, int64_t kv_rs, int mode, int64_t n_prompt, DropArgs drop, cudaStream_t st);
int attn_fwd_tc5(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H, int64_t Tq, int64_t Tk,
                 int64_t head_dim, int64_t q_bs, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, int mode, int64_t n_prompt,
                 DropArgs drop, cudaStream_t st);
int attn_bwd_tc5(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, float* dq_acc,
                 void* dk, void* dv, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_bs, int64_t q_rs,
                 int64_t kv_bs, int64_t kv_rs, int mode, int64_t n_prompt, DropArgs drop, cudaStream_t st);
int attn_bwd_tc5r(const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse, void* dq, void* dk,
                  void* dv, int64_t B, int64_t H, int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_bs, int64_t q_rs, int64_t kv_bs,
                  int64_t kv_rs, int mode, int64_t n_prompt, DropArgs drop, DropArgs tok, cudaStream_t st);
// 0: fp32-math kernels; 1 (default): tensor cores -- tcgen05 forward where eligible, mma.sync otherwise; 2: mma.sync only
static std::atomic<int> g_attn_tc{1};

}  // namespace i2t

using namespace i2t;

extern "C" void i2t_set_tensor_core_attention(int mode) { g_attn_tc.store(mode < 0 ? 0 : (mode > 2 ? 1 : mode)); }

static int attn_fwd_impl(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H,
                         int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_batch_stride, int64_t q_row_stride,
                         int64_t kv_batch_stride, int64_t kv_row_stride, int mask_mode, int64_t n_prompt,
                         int in_dtype, int out_dtype, DropArgs drop, void* stream) {
  I2T_REQUIRE(q && k && v && out, "attn_fwd: null pointer");
  I2T_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0 && B <= 65535 && H <= 65535, "attn_fwd: bad sizes");
  I2T_REQUIRE(head_dim == 64 || head_dim == 32, "attn_fwd: head_dim %lld not built (32, 64)", (long long)head_dim);
  I2T_REQUIRE(mask_mode >= 0 && mask_mode <= 2, "attn_fwd: bad mask mode");
  I2T_REQUIRE(q_row_stride % 4 == 0 && kv_row_stride % 4 == 0 && q_batch_stride % 4 == 0 && kv_batch_stride % 4 == 0,
              "attn_fwd: strides must be multiples of 4 elements");
  I2T_REQUIRE(valid_dtype(in_dtype) && valid_dtype(out_dtype), "attn_fwd: bad dtype");
  I2T_REQUIRE(B * H * Tq < (int64_t)1 << 32, "attn_fwd: more than 2^32 query rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == I2T_BF16 && out_dtype == I2T_BF16 && g_attn_tc.load() == 1) {
    const int r5 = attn_fwd_tc5(q, k, v, out, lse, B, H, Tq, Tk, head_dim, q_batch_stride, q_row_stride, kv_batch_stride,
                                kv_row_stride, mask_mode, n_prompt, drop, st);
    if (r5 != 0) return r5 < 0 ? r5 : I2T_OK;
  }
  if (in_dtype == I2T_BF16 && out_dtype == I2T_BF16 && g_attn_tc.load() != 0) {
    const int r = attn_fwd_tc(q, k, v, out, lse, B, H, Tq, Tk, head_dim, q_batch_stride, q_row_stride, kv_batch_stride,
                              kv_row_stride, mask_mode, n_prompt, drop, st);
    if (r != 0) return r < 0 ? r : I2T_OK;
  }
#define I2T_ATT(TI, TO, HSV) \
  return launch_attn_fwd<TI, TO, HSV>(q, k, v, out, lse, B, H, Tq, Tk, q_batch_stride, q_row_stride, kv_batch_stride, kv_row_stride, mask_mode, n_prompt, drop, st)
  if (head_dim == 64) {
    if (in_dtype == I2T_F32 && out_dtype == I2T_F32) I2T_ATT(float, float, 64);
    if (in_dtype == I2T_BF16 && out_dtype == I2T_BF16) I2T_ATT(__nv_bfloat16, __nv_bfloat16, 64);
    if (in_dtype == I2T_F32 && out_dtype == I2T_BF16) I2T_ATT(float, __nv_bfloat16, 64);
  } else {
    if (in_dtype == I2T_F32 && out_dtype == I2T_F32) I2T_ATT(float, float, 32);
    if (in_dtype == I2T_BF16 && out_dtype == I2T_BF16) I2T_ATT(__nv_bfloat16, __nv_bfloat16, 32);
  }
#undef I2T_ATT
  return fail(I2T_ERR_INVALID, "attn_fwd: dtype combination (%d,%d) not built", in_dtype, out_dtype);
}

extern "C" int i2t_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H,
                            int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_batch_stride, int64_t q_row_stride,
                            int64_t kv_batch_stride, int64_t kv_row_stride, int mask_mode, int64_t n_prompt,
                            int in_dtype, int out_dtype, void* stream) {
  return attn_fwd_impl(q, k, v, out, lse, B, H, Tq, Tk, head_dim, q_batch_stride, q_row_stride, kv_batch_stride, kv_row_stride,
                       mask_mode, n_prompt, in_dtype, out_dtype, make_drop(0.f, nullptr, 0), stream);
}

extern "C" int i2t_attn_fwd_dropout(const void* q, const void* k, const void* v, void* out, float* lse, int64_t B, int64_t H,
                                    int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_batch_stride, int64_t q_row_stride,
                                    int64_t kv_batch_stride, int64_t kv_row_stride, int mask_mode, int64_t n_prompt,
                                    int in_dtype, int out_dtype, float p_drop, const void* rng_state, int64_t site,
                                    void* stream) {
  I2T_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attn_fwd_dropout: p must be in [0,1)");
  I2T_REQUIRE(p_drop == 0.f || rng_state != nullptr, "attn_fwd_dropout: rng_state is null");
  return attn_fwd_impl(q, k, v, out, lse, B, H, Tq, Tk, head_dim, q_batch_stride, q_row_stride, kv_batch_stride, kv_row_stride,
                       mask_mode, n_prompt, in_dtype, out_dtype, make_drop(p_drop, rng_state, site), stream);
}

namespace i2t {
template <typename T, int HS>
static int launch_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout,
                           const float* lse, void* dq, void* dk, void* dv, float* ws, int64_t B, int64_t H, int64_t Tq,
                           int64_t Tk, int64_t q_bs, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, int mode,
                           int64_t n_prompt, DropArgs drop, cudaStream_t st, DropArgs tok = make_drop(0.f, nullptr, 0),
                           bool* tok_done = nullptr) {
  // workspace: delta (B*H*Tq) then dq accumulator (B*H*Tq*HS), fp32
  float* delta = ws;
  float* dq_acc = ws + (B * H * Tq + 3) / 4 * 4;      // keep the accumulator 16-byte aligned (float4 reads in the scatter)
  if (sizeof(T) == 2 && g_attn_tc.load() == 1) {      // <= 256 rows: one tcgen05 kernel, no workspace
    const int r = attn_bwd_tc5r(q, k, v, out, dout, lse, dq, dk, dv, B, H, Tq, Tk, HS, q_bs, q_rs, kv_bs, kv_rs, mode, n_prompt, drop,
                                tok, st);
    if (r < 0) return r;
    if (r == 1) {
      if (tok_done != nullptr) *tok_done = true;       // the kernel scaled dq / dk / dv by the token masks on its way out
      return I2T_OK;
    }
  }
  I2T_CUDA(cudaMemsetAsync(dq_acc, 0, (size_t)(B * H * Tq * HS) * sizeof(float), st));
  if (sizeof(T) == 2 && g_attn_tc.load() == 1) {      // tcgen05 backward: computes delta itself
    const int r = attn_bwd_tc5(q, k, v, out, dout, lse, dq_acc, dk, dv, B, H, Tq, Tk, HS, q_bs, q_rs, kv_bs, kv_rs, mode, n_prompt,
                               drop, st);
    if (r < 0) return r;
    if (r == 1) {
      const int64_t total_tc = B * H * Tq * (HS / 4);
      I2T_CUDA(launch_pdl(attn_dq_scatter_kernel<T, HS>, dim3((unsigned)ceil_div(total_tc, 256)), dim3(256), 0, st, (const float*)dq_acc,
                          (T*)dq, (int)H, (int)Tq, q_bs, q_rs, total_tc));
      I2T_LAUNCHED();
      return I2T_OK;
    }
  }
  const int64_t rows = B * Tq * H;
  I2T_CUDA(launch_pdl(attn_delta_kernel<T, HS>, dim3((unsigned)ceil_div(rows, 4)), dim3(128), 0, st, (const T*)out, (const T*)dout, delta,
                      (int)H, (int)Tq, rows));
  I2T_LAUNCHED();
  if (sizeof(T) == 2 && g_attn_tc.load() != 0) {
    const int r = attn_bwd_tc(q, k, v, dout, lse, delta, dq_acc, dk, dv, B, H, Tq, Tk, HS, q_bs, q_rs, kv_bs, kv_rs, mode,
                              n_prompt, drop, st);
    if (r < 0) return r;
    if (r == 1) {
      const int64_t total_tc = B * H * Tq * (HS / 4);
      I2T_CUDA(launch_pdl(attn_dq_scatter_kernel<T, HS>, dim3((unsigned)ceil_div(total_tc, 256)), dim3(256), 0, st, (const float*)dq_acc,
                          (T*)dq, (int)H, (int)Tq, q_bs, q_rs, total_tc));
      I2T_LAUNCHED();
      return I2T_OK;
    }
  }
  const size_t smem = (size_t)(4 * HS * 64 + ATT_BQ * ATT_PSTRIDE + 3 * 64 * HS) * sizeof(float);
  auto kern = attn_bwd_kernel<T, HS>;
  I2T_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(Tk, ATT_BK), (unsigned)H, (unsigned)B);
  kern<<<grid, ATT_THREADS, smem, st>>>((const T*)q, (const T*)k, (const T*)v, (const T*)dout, lse, delta, dq_acc,
                                        (T*)dk, (T*)dv, (int)H, (int)Tq, (int)Tk, q_bs, q_rs, kv_bs, kv_rs, mode,
                                        (int)n_prompt, 1.0f / sqrtf((float)HS), drop);
  I2T_LAUNCHED();
  const int64_t total = B * H * Tq * (HS / 4);
  attn_dq_scatter_kernel<T, HS><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(dq_acc, (T*)dq, (int)H, (int)Tq, q_bs,
                                                                                q_rs, total);
  I2T_LAUNCHED();
  return I2T_OK;
}
}  // namespace i2t

extern "C" int64_t i2t_attn_bwd_workspace_bytes(int64_t B, int64_t H, int64_t Tq, int64_t head_dim) {
  return ((B * H * Tq + 3) / 4 * 4 + B * H * Tq * head_dim) * (int64_t)sizeof(float);
}

static int attn_bwd_impl(const void* q, const void* k, const void* v, const void* out, const void* dout,
                         const float* lse, void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t H,
                         int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_batch_stride, int64_t q_row_stride,
                         int64_t kv_batch_stride, int64_t kv_row_stride, int mask_mode, int64_t n_prompt, int dtype,
                         DropArgs drop, void* stream, DropArgs tok = make_drop(0.f, nullptr, 0), bool* tok_done = nullptr) {
  I2T_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv && workspace, "attn_bwd: null pointer");
  I2T_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0 && B <= 65535 && H <= 65535, "attn_bwd: bad sizes");
  I2T_REQUIRE(head_dim == 64 || head_dim == 32, "attn_bwd: head_dim %lld not built (32, 64)", (long long)head_dim);
  I2T_REQUIRE(mask_mode >= 0 && mask_mode <= 2, "attn_bwd: bad mask mode");
  I2T_REQUIRE(q_row_stride % 4 == 0 && kv_row_stride % 4 == 0 && q_batch_stride % 4 == 0 && kv_batch_stride % 4 == 0,
              "attn_bwd: strides must be multiples of 4 elements");
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
#define I2T_ATTB(T, HSV) \
  return launch_attn_bwd<T, HSV>(q, k, v, out, dout, lse, dq, dk, dv, ws, B, H, Tq, Tk, q_batch_stride, q_row_stride, kv_batch_stride, kv_row_stride, mask_mode, n_prompt, drop, st, tok, tok_done)
  if (dtype == I2T_F32) {
    if (head_dim == 64) I2T_ATTB(float, 64);
    I2T_ATTB(float, 32);
  } else if (dtype == I2T_BF16) {
    if (head_dim == 64) I2T_ATTB(__nv_bfloat16, 64);
    I2T_ATTB(__nv_bfloat16, 32);
  }
#undef I2T_ATTB
  return fail(I2T_ERR_INVALID, "attn_bwd: bad dtype %d", dtype);
}

extern "C" int i2t_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout,
                            const float* lse, void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t H,
                            int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_batch_stride, int64_t q_row_stride,
                            int64_t kv_batch_stride, int64_t kv_row_stride, int mask_mode, int64_t n_prompt, int dtype,
                            void* stream) {
  return attn_bwd_impl(q, k, v, out, dout, lse, dq, dk, dv, workspace, B, H, Tq, Tk, head_dim, q_batch_stride, q_row_stride,
                       kv_batch_stride, kv_row_stride, mask_mode, n_prompt, dtype, make_drop(0.f, nullptr, 0), stream);
}

extern "C" int i2t_token_dropout(void* x, int64_t rows, int64_t ld, int64_t seg, int64_t nseg, float p, const void* rng_state,
                                 int64_t site, int dtype, void* stream);

// i2t_attn_bwd_dropout for the packed self-attention buffer of reference models/layers.py:447-469, plus the BACKWARD of the token-level
// q / k / v dropout (:454-461) that was applied to that buffer in the forward: dq / dk / dv (the three C-wide segments of one
// (B*T, 3C) gradient buffer) come out multiplied by the (row, segment) masks of `tok_site`.  The tcgen05 kernel does it while it
// stores its accumulators; any other kernel is followed by the stand-alone token-dropout pass -- same result either way.
extern "C" int i2t_attn_bwd_dropout_tok(const void* q, const void* k, const void* v, const void* out, const void* dout,
                                        const float* lse, void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t H,
                                        int64_t T, int64_t head_dim, int64_t batch_stride, int64_t row_stride, int mask_mode,
                                        int64_t n_prompt, int dtype, float p_drop, const void* rng_state, int64_t site, float p_tok,
                                        int64_t tok_site, void* stream) {
  I2T_REQUIRE(p_drop >= 0.f && p_drop < 1.f && p_tok > 0.f && p_tok < 1.f, "attn_bwd_dropout_tok: p must be in [0,1) / (0,1)");
  I2T_REQUIRE(rng_state != nullptr, "attn_bwd_dropout_tok: rng_state is null");
  const int64_t C = H * head_dim;
  const size_t es = dtype == I2T_F32 ? 4 : 2;
  I2T_REQUIRE(dq && (const char*)dk == (const char*)dq + C * es && (const char*)dv == (const char*)dq + 2 * C * es && row_stride >= 3 * C,
              "attn_bwd_dropout_tok: dq / dk / dv must be the three segments of one packed gradient buffer");
  bool done = false;
  const int rc = attn_bwd_impl(q, k, v, out, dout, lse, dq, dk, dv, workspace, B, H, T, T, head_dim, batch_stride, row_stride,
                               batch_stride, row_stride, mask_mode, n_prompt, dtype, make_drop(p_drop, rng_state, site), stream,
                               make_drop(p_tok, rng_state, tok_site), &done);
  if (rc != I2T_OK || done) return rc;
  I2T_REQUIRE(batch_stride == T * row_stride, "attn_bwd_dropout_tok: batches must be row-contiguous");
  return i2t_token_dropout(dq, B * T, row_stride, C, 3, p_tok, rng_state, tok_site, dtype, stream);
}

extern "C" int i2t_attn_bwd_dropout(const void* q, const void* k, const void* v, const void* out, const void* dout,
                                    const float* lse, void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t H,
                                    int64_t Tq, int64_t Tk, int64_t head_dim, int64_t q_batch_stride, int64_t q_row_stride,
                                    int64_t kv_batch_stride, int64_t kv_row_stride, int mask_mode, int64_t n_prompt,
                                    int dtype, float p_drop, const void* rng_state, int64_t site, void* stream) {
  I2T_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "attn_bwd_dropout: p must be in [0,1)");
  I2T_REQUIRE(p_drop == 0.f || rng_state != nullptr, "attn_bwd_dropout: rng_state is null");
  return attn_bwd_impl(q, k, v, out, dout, lse, dq, dk, dv, workspace, B, H, Tq, Tk, head_dim, q_batch_stride, q_row_stride,
                       kv_batch_stride, kv_row_stride, mask_mode, n_prompt, dtype, make_drop(p_drop, rng_state, site), stream);
}
