// Training-mode dropout for the memory-bound sites (HBM bound: one read + one write of the tensor):
//   * residual / embedding dropout  -- nn.Dropout at reference models/layers.py:469 (resid_dropout), :485 (_MLP.dropout),
//     models/decoder.py:236-243 (transformer.drop), HF GPT-2 resid_pdrop / embd_pdrop.  Forward fuses the residual add
//     (out = residual + keep * y / (1-p)); backward fuses the fp32 -> compute-dtype cast of the incoming gradient.
//   * token-level q/k/v dropout     -- reference models/layers.py:454-461: nn.Dropout on a (B,1,T,1) tensor of
//     ones, multiplied into q, k and v separately -> one Bernoulli per (row, {q,k,v}) scaling a C-wide segment of the packed
//     (B*T, 3C) buffer; the same kernel scales the gradient of that buffer in the backward.
// Masks come from rng.cuh (counter-based, regenerated in the backward, nothing stored).
#include "common.cuh"
#include "rng.cuh"

namespace i2t {

template <typename TY>
__global__ void __launch_bounds__(256) dropout_add_kernel(const TY* __restrict__ y, const float* __restrict__ res,
                                                          float* __restrict__ out, int64_t n4, DropArgs d) {
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  const DropKey key = drop_key(d);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const Philox4 r = drop_elem4(d, key, (uint64_t)i);
    float4 v = load4(y + i * 4);
    v.x = r.x >= d.thr ? v.x * d.inv_keep : 0.f;
    v.y = r.y >= d.thr ? v.y * d.inv_keep : 0.f;
    v.z = r.z >= d.thr ? v.z * d.inv_keep : 0.f;
    v.w = r.w >= d.thr ? v.w * d.inv_keep : 0.f;
    if (res != nullptr) {
      const float4 a = load4(res + i * 4);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    store4(out + i * 4, v);
  }
}

template <typename TIN, typename TG>
__global__ void __launch_bounds__(256) dropout_bwd_kernel(const TIN* __restrict__ dy, TG* __restrict__ g, int64_t n4, DropArgs d) {
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  const DropKey key = drop_key(d);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const Philox4 r = drop_elem4(d, key, (uint64_t)i);
    float4 v = load4(dy + i * 4);
    v.x = r.x >= d.thr ? v.x * d.inv_keep : 0.f;
    v.y = r.y >= d.thr ? v.y * d.inv_keep : 0.f;
    v.z = r.z >= d.thr ? v.z * d.inv_keep : 0.f;
    v.w = r.w >= d.thr ? v.w * d.inv_keep : 0.f;
    store4(g + i * 4, v);
  }
}

// x: (rows, nseg * seg) with row pitch ld; element (row, s*seg + c) *= keep(row, s) / (1-p).  One warp per row: ONE Philox call
// per row (the generator costs ~60 instructions; per 4 elements it was as expensive as the memory traffic), then 16-byte
// vectors -- a lane's vector never straddles a segment (seg is a multiple of the vector width).
template <typename T>
__global__ void __launch_bounds__(256) token_dropout_kernel(T* __restrict__ x, int64_t rows, int64_t ld, int seg, int nseg,
                                                            DropArgs d) {
  pdl_launch_dependents();     // programmatic dependent launch: this grid may have started before its predecessor finished
  pdl_wait();
  constexpr int VEC = 16 / (int)sizeof(T);
  const DropKey key = drop_key(d);
  const int lane = threadIdx.x & 31;
  const int nvec = seg * nseg / VEC;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
    const Philox4 r = drop_elem4(d, key, (uint64_t)row);
    float m[4];
    m[0] = r.x >= d.thr ? d.inv_keep : 0.f;
    m[1] = r.y >= d.thr ? d.inv_keep : 0.f;
    m[2] = r.z >= d.thr ? d.inv_keep : 0.f;
    m[3] = r.w >= d.thr ? d.inv_keep : 0.f;
    T* xr = x + row * ld;
    for (int v = lane; v < nvec; v += 32) {
      const int sgm = v * VEC / seg;
      const float mk = sgm == 0 ? m[0] : sgm == 1 ? m[1] : sgm == 2 ? m[2] : m[3];
      T* p = xr + (int64_t)v * VEC;
#pragma unroll
      for (int h = 0; h < VEC / 4; ++h) {
        float4 val = load4(p + 4 * h);
        val.x *= mk; val.y *= mk; val.z *= mk; val.w *= mk;
        store4(p + 4 * h, val);
      }
    }
  }
}

static unsigned grid_for(int64_t work) {
  const int64_t blocks = ceil_div(work, 256);
  const int64_t cap = (int64_t)num_sms() * 8;
  return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace i2t

using namespace i2t;

extern "C" int i2t_dropout_add_fwd(const void* y, const float* residual, float* out, int64_t n, float p, const void* rng_state,
                                   int64_t site, int y_dtype, void* stream) {
  I2T_REQUIRE(y && out && rng_state, "dropout_add_fwd: null pointer");
  I2T_REQUIRE(n > 0 && n % 4 == 0, "dropout_add_fwd: n must be a positive multiple of 4");
  I2T_REQUIRE(p > 0.f && p < 1.f, "dropout_add_fwd: p must be in (0,1)");
  I2T_REQUIRE(valid_dtype(y_dtype), "dropout_add_fwd: bad dtype");
  const DropArgs d = make_drop(p, rng_state, site);
  cudaStream_t st = (cudaStream_t)stream;
  if (y_dtype == I2T_F32)
    I2T_CUDA(launch_pdl(dropout_add_kernel<float>, dim3(grid_for(n / 4)), dim3(256), 0, st, (const float*)y, residual, out, n / 4, d));
  else
    I2T_CUDA(launch_pdl(dropout_add_kernel<__nv_bfloat16>, dim3(grid_for(n / 4)), dim3(256), 0, st, (const __nv_bfloat16*)y, residual, out,
                        n / 4, d));
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_dropout_bwd(const void* dy, void* g, int64_t n, float p, const void* rng_state, int64_t site, int dy_dtype,
                               int g_dtype, void* stream) {
  I2T_REQUIRE(dy && g && rng_state, "dropout_bwd: null pointer");
  I2T_REQUIRE(n > 0 && n % 4 == 0, "dropout_bwd: n must be a positive multiple of 4");
  I2T_REQUIRE(p > 0.f && p < 1.f, "dropout_bwd: p must be in (0,1)");
  I2T_REQUIRE(valid_dtype(dy_dtype) && valid_dtype(g_dtype), "dropout_bwd: bad dtype");
  const DropArgs d = make_drop(p, rng_state, site);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(n / 4);
  if (dy_dtype == I2T_F32 && g_dtype == I2T_F32)
    I2T_CUDA(launch_pdl(dropout_bwd_kernel<float, float>, dim3(grid), dim3(256), 0, st, (const float*)dy, (float*)g, n / 4, d));
  else if (dy_dtype == I2T_F32 && g_dtype == I2T_BF16)
    I2T_CUDA(launch_pdl(dropout_bwd_kernel<float, __nv_bfloat16>, dim3(grid), dim3(256), 0, st, (const float*)dy, (__nv_bfloat16*)g, n / 4, d));
  else if (dy_dtype == I2T_BF16 && g_dtype == I2T_BF16)
    I2T_CUDA(launch_pdl(dropout_bwd_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid), dim3(256), 0, st, (const __nv_bfloat16*)dy,
                        (__nv_bfloat16*)g, n / 4, d));
  else
    return fail(I2T_ERR_INVALID, "dropout_bwd: dtype combination (%d,%d) not built", dy_dtype, g_dtype);
  I2T_LAUNCHED();
  return I2T_OK;
}

extern "C" int i2t_token_dropout(void* x, int64_t rows, int64_t ld, int64_t seg, int64_t nseg, float p, const void* rng_state,
                                 int64_t site, int dtype, void* stream) {
  I2T_REQUIRE(x && rng_state, "token_dropout: null pointer");
  I2T_REQUIRE(rows > 0 && seg > 0 && seg % 8 == 0 && nseg >= 1 && nseg <= 4 && ld >= seg * nseg && ld % 8 == 0 &&
                  ((uintptr_t)x & 15u) == 0,
              "token_dropout: bad sizes (16-byte aligned rows, segments of a multiple of 8 elements, at most 4 per row)");
  I2T_REQUIRE(p > 0.f && p < 1.f, "token_dropout: p must be in (0,1)");
  const DropArgs d = make_drop(p, rng_state, site);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(rows * 32);        // one warp per row
  if (dtype == I2T_F32)
    I2T_CUDA(launch_pdl(token_dropout_kernel<float>, dim3(grid), dim3(256), 0, st, (float*)x, rows, ld, (int)seg, (int)nseg, d));
  else if (dtype == I2T_BF16)
    I2T_CUDA(launch_pdl(token_dropout_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, st, (__nv_bfloat16*)x, rows, ld, (int)seg, (int)nseg, d));
  else
    return fail(I2T_ERR_INVALID, "token_dropout: bad dtype %d", dtype);
  I2T_LAUNCHED();
  return I2T_OK;
}

// state[1] += 1 on the stream (so a captured training step draws fresh masks on every replay)
namespace i2t {
__global__ void rng_advance_kernel(unsigned long long* state) { state[1] += 1ull; }
}
extern "C" int i2t_rng_advance(void* rng_state, void* stream) {
  I2T_REQUIRE(rng_state, "rng_advance: null pointer");
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)rng_state);
  I2T_LAUNCHED();
  return I2T_OK;
}
