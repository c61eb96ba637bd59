// Memory-bound pieces around the ViT trunk and the decoder input:
//   patch im2col + token assembly  (torchvision vision_transformer.py:268-296: conv_proj k=16,s=16 -> reshape ->
//                                   permute -> prepend class token -> + pos_embedding)
//   LSH tail                       (reference models/layers.py:139-144,211-219; models/encoder.py:116-117)
//   decoder input embedding        (reference models/vision_encoder_decoder.py:84-88; models/decoder.py:234-243)
// All are coalesced 128-bit copies / gathers; none has data reuse worth staging in shared memory except the LSH
// projection, whose 768-vector is kept in shared memory for the 96 dot products per (image, slot).
#include "common.cuh"

namespace i2t {

// One thread moves 4 horizontally adjacent pixels of one channel row of one patch.
template <typename TOUT>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ img, TOUT* __restrict__ out, int64_t B,
                                                     int H, int W, int p, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int p4 = p / 4;
  const int nw = W / p, nh = H / p;
  int64_t r = i;
  const int kx4 = (int)(r % p4); r /= p4;
  const int ky = (int)(r % p); r /= p;
  const int c = (int)(r % 3); r /= 3;
  const int pw = (int)(r % nw); r /= nw;
  const int ph = (int)(r % nh); r /= nh;
  const int64_t b = r;
  const float4 v = load4(img + ((b * 3 + c) * H + (int64_t)(ph * p + ky)) * W + pw * p + kx4 * 4);
  const int64_t row = (b * nh + ph) * nw + pw;
  const int64_t col = ((int64_t)c * p + ky) * p + kx4 * 4;
  store4(out + row * (3 * p * p) + col, v);
}

template <typename TP>
__global__ void __launch_bounds__(256) vit_assemble_kernel(const TP* __restrict__ patch_out, const float* __restrict__ cls,
                                                           const float* __restrict__ pos, float* __restrict__ x,
                                                           int64_t B, int np, int C, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = C / 4;
  const int cv = (int)(i % c4);
  const int64_t r = i / c4;
  const int tok = (int)(r % (np + 1));
  const int64_t b = r / (np + 1);
  float4 v = tok == 0 ? load4(cls + cv * 4) : load4(patch_out + (b * np + tok - 1) * C + cv * 4);
  const float4 pe = load4(pos + (int64_t)tok * C + cv * 4);
  v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
  store4(x + r * C + cv * 4, v);
}

// grid = (n_cls, B, ceil(E / 256)); block = 256.  Every (image, slot) hashes the same normalised feature against its own
// projections, bucketises, and averages the selected EmbeddingBag rows.
__global__ void __launch_bounds__(256)
lsh_tail_kernel(const float* __restrict__ feat, const float* const* __restrict__ proj, const float* const* __restrict__ grid_,
                const float* const* __restrict__ emb, const int32_t* __restrict__ num_bins, float* __restrict__ out,
                int32_t* __restrict__ bucket_out, int D, int n_cls, int n_res, int n_proj, int E) {
  extern __shared__ float sm[];
  float* f = sm;                   // [D] normalised feature
  float* part = f + D;             // [256] partial dot products ([256 / n_proj groups][n_proj])
  int* rowidx = (int*)(part + 256);  // [n_res][n_proj]
  __shared__ float s_norm;
  const int s = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  // F.normalize(x, p=2, dim=-1): x / max(||x||, 1e-12)
  float ss = 0.f;
  for (int k = t; k < D; k += 256) {
    const float v = feat[b * D + k];
    f[k] = v;
    ss += v * v;
  }
  ss = warp_sum(ss);
  if (lane == 0) part[w] = ss;
  __syncthreads();
  if (t == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += part[i];
    s_norm = fmaxf(sqrtf(tot), 1e-12f);
  }
  __syncthreads();
  const float inv = 1.0f / s_norm;
  for (int k = t; k < D; k += 256) f[k] = f[k] * inv;
  __syncthreads();
  for (int r = 0; r < n_res; ++r) {
    const float* P = proj[s * n_res + r];
    const float* G = grid_[s * n_res + r];
    const int nb = num_bins[r];
    // z[p] = sum_k f[k] P[k][p]; thread (g = t / n_proj, p = t % n_proj) strides k by 256 / n_proj
    const int groups = 256 / n_proj;
    const int g = t / n_proj, p = t % n_proj;
    float acc = 0.f;
    if (g < groups) {
#pragma unroll 8
      for (int k = g; k < D; k += groups) acc = fmaf(f[k], P[(int64_t)k * n_proj + p], acc);     // (loads in flight, FMAs in order)
    }
    if (g < groups) part[g * n_proj + p] = acc;
    __syncthreads();
    if (t < n_proj) {
      float z = 0.f;
      for (int i = 0; i < groups; ++i) z += part[i * n_proj + t];
      int bucket = 0;  // torch.bucketize(right=False): number of boundaries strictly below z
      for (int i = 0; i < nb; ++i) bucket += (G[i] < z) ? 1 : 0;
      const int idx = bucket + (nb + 1) * t;
      rowidx[r * n_proj + t] = idx;
      if (bucket_out && blockIdx.z == 0) bucket_out[((b * n_cls + s) * n_res + r) * n_proj + t] = idx;
    }
    __syncthreads();
  }
  // gather: grid.z slices of 256 output columns (one per thread), so the 64 (slot, image) pairs spread over 192 CTAs and a thread's
  // n_res * n_proj row reads are independent loads in flight (was: 3 columns per thread in one CTA per pair, 124 us of latency)
  const float invp = 1.0f / (float)n_proj;
  const int e = blockIdx.z * 256 + t;
  if (e < E) {
    float tot = 0.f;
    for (int r = 0; r < n_res; ++r) {
      const float* T = emb[s * n_res + r];
      const int* ri = rowidx + r * n_proj;
      float a = 0.f;                       // (one accumulator, rows in order: the summation order of the reference's EmbeddingBag)
#pragma unroll 8
      for (int p = 0; p < n_proj; ++p) a += T[(int64_t)ri[p] * E + e];
      tot += a * invp;
    }
    out[(b * n_cls + s) * E + e] = tot;
  }
}

__global__ void __launch_bounds__(256) embed_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ prompt,
                                                        const float* __restrict__ wte, const float* __restrict__ wpe,
                                                        float* __restrict__ x, int64_t B, int T, int n_prompt, int S,
                                                        int C, int64_t total4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = C / 4;
  const int cv = (int)(i % c4);
  const int64_t r = i / c4;
  const int tt = (int)(r % T);
  const int64_t b = r / T;
  float4 v;
  if (tt < n_prompt) v = load4(prompt + (b * n_prompt + tt) * C + cv * 4);
  else v = load4(wte + ids[b * S + (tt - n_prompt)] * C + cv * 4);
  const float4 pe = load4(wpe + (int64_t)tt * C + cv * 4);
  v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
  store4(x + r * C + cv * 4, v);
}

}  // namespace i2t

using namespace i2t;

extern "C" int i2t_patch_im2col(const float* images, void* patches, int64_t B, int64_t H, int64_t W, int64_t p,
                                int out_dtype, void* stream) {
  I2T_REQUIRE(images && patches && B > 0, "patch_im2col: null pointer / empty batch");
  I2T_REQUIRE(p > 0 && p % 4 == 0 && H % p == 0 && W % p == 0 && W % 4 == 0, "patch_im2col: patch %lld must divide %lldx%lld",
              (long long)p, (long long)H, (long long)W);
  I2T_REQUIRE(valid_dtype(out_dtype) && aligned16(images), "patch_im2col: dtype/alignment");
  const int64_t total4 = B * 3 * H * W / 4;
  const unsigned blocks = (unsigned)ceil_div(total4, 256);
  if (out_dtype == I2T_F32)
    im2col_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(images, (float*)patches, B, (int)H, (int)W, (int)p, total4);
  else
    im2col_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(images, (__nv_bfloat16*)patches, B, (int)H, (int)W, (int)p, total4);
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_vit_assemble(const void* patch_out, const float* cls, const float* pos, float* x, int64_t B,
                                int64_t np, int64_t C, int patch_dtype, void* stream) {
  I2T_REQUIRE(patch_out && cls && pos && x && B > 0 && np > 0 && C % 4 == 0, "vit_assemble: bad arguments");
  I2T_REQUIRE(valid_dtype(patch_dtype), "vit_assemble: bad dtype");
  const int64_t total4 = B * (np + 1) * C / 4;
  const unsigned blocks = (unsigned)ceil_div(total4, 256);
  if (patch_dtype == I2T_F32)
    vit_assemble_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)patch_out, cls, pos, x, B, (int)np, (int)C, total4);
  else
    vit_assemble_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)patch_out, cls, pos, x, B, (int)np, (int)C, total4);
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_lsh_tail(const float* feat, const void* const* proj, const void* const* grid, const void* const* emb,
                            const int32_t* num_bins, float* out, int32_t* bucket_out, int64_t B, int64_t D, int64_t n_cls,
                            int64_t n_res, int64_t n_proj, int64_t E, void* stream) {
  I2T_REQUIRE(feat && proj && grid && emb && num_bins && out, "lsh_tail: null pointer");
  I2T_REQUIRE(B > 0 && B <= 65535 && D > 0 && n_cls > 0 && n_res > 0 && E > 0, "lsh_tail: bad sizes");
  I2T_REQUIRE(n_proj > 0 && n_proj <= 256 && 256 % n_proj == 0, "lsh_tail: n_proj=%lld must divide 256", (long long)n_proj);
  const size_t smem = (size_t)(D + 256) * sizeof(float) + (size_t)(n_res * n_proj) * sizeof(int);
  I2T_REQUIRE(smem <= 48 * 1024, "lsh_tail: feature too wide for shared memory");
  dim3 g((unsigned)n_cls, (unsigned)B, (unsigned)ceil_div(E, 256));
  lsh_tail_kernel<<<g, 256, smem, (cudaStream_t)stream>>>(feat, (const float* const*)proj, (const float* const*)grid,
                                                          (const float* const*)emb, num_bins, out, bucket_out, (int)D,
                                                          (int)n_cls, (int)n_res, (int)n_proj, (int)E);
  I2T_LAUNCHED();
  return I2T_OK;
}

// EmbeddingBag(mean) backward of the LSH tail: d emb[i][row, :] += d out[b, s, :] / n_proj for the n_proj rows image b picked in
// table i = s * n_res + r.  grid = (n_cls * n_res, B).  The table pointers travel by value in the launch parameters (the gradient
// buffers belong to the caller and differ from call to call: no device-side pointer table to keep in step).
#define I2T_LSH_MAX_TABLES 64
struct LshGradTables { float* t[I2T_LSH_MAX_TABLES]; };

__global__ void __launch_bounds__(256)
lsh_tail_bwd_kernel(const float* __restrict__ dout, const int32_t* __restrict__ bucket, LshGradTables tabs, int n_cls, int n_res,
                    int n_proj, int E) {
  const int i = blockIdx.x, s = i / n_res;
  const int64_t b = blockIdx.y;
  float* T = tabs.t[i];
  const int32_t* rows = bucket + (b * n_cls * n_res + i) * n_proj;
  const float* g = dout + (b * n_cls + s) * E;
  const float invp = 1.0f / (float)n_proj;
  for (int e = threadIdx.x; e < E; e += 256) {
    const float v = g[e] * invp;
    for (int p = 0; p < n_proj; ++p) atomicAdd(T + (int64_t)rows[p] * E + e, v);
  }
}

extern "C" int i2t_lsh_tail_bwd(const float* dout, const int32_t* bucket, void* const* demb_host, int64_t B, int64_t n_cls,
                                int64_t n_res, int64_t n_proj, int64_t E, void* stream) {
  I2T_REQUIRE(dout && bucket && demb_host, "lsh_tail_bwd: null pointer");
  I2T_REQUIRE(B > 0 && B <= 65535 && n_cls > 0 && n_res > 0 && n_proj > 0 && E > 0, "lsh_tail_bwd: bad sizes");
  I2T_REQUIRE(n_cls * n_res <= I2T_LSH_MAX_TABLES, "lsh_tail_bwd: %lld tables (max %d)", (long long)(n_cls * n_res),
              I2T_LSH_MAX_TABLES);
  LshGradTables tabs;
  for (int64_t i = 0; i < n_cls * n_res; ++i) {
    I2T_REQUIRE(demb_host[i], "lsh_tail_bwd: null table gradient");
    tabs.t[i] = (float*)demb_host[i];
  }
  dim3 g((unsigned)(n_cls * n_res), (unsigned)B);
  lsh_tail_bwd_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(dout, bucket, tabs, (int)n_cls, (int)n_res, (int)n_proj, (int)E);
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_embed_fwd(const int64_t* ids, const float* prompt, const float* wte, const float* wpe, float* x,
                             int64_t B, int64_t T, int64_t n_prompt, int64_t S, int64_t C, void* stream) {
  I2T_REQUIRE(wte && wpe && x && B > 0 && T > 0 && C % 4 == 0, "embed_fwd: bad arguments");
  I2T_REQUIRE(n_prompt >= 0 && (n_prompt == 0 || prompt) && (T <= n_prompt || ids), "embed_fwd: missing prompt/ids");
  I2T_REQUIRE(T <= n_prompt + S, "embed_fwd: T=%lld exceeds n_prompt+S=%lld", (long long)T, (long long)(n_prompt + S));
  const int64_t total4 = B * T * C / 4;
  embed_fwd_kernel<<<(unsigned)ceil_div(total4, 256), 256, 0, (cudaStream_t)stream>>>(ids, prompt, wte, wpe, x, B, (int)T,
                                                                                   (int)n_prompt, (int)S, (int)C, total4);
  I2T_LAUNCHED();
  return I2T_OK;
}
