// Micro-benchmark: latency of grid-wide barrier variants on 148 co-resident CTAs (no skew: every CTA arrives at once).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o grid_barrier grid_barrier.cu && ./grid_barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int V>
__device__ __forceinline__ void barrier(unsigned int* bar, unsigned int* flags, unsigned int& epoch, int tid) {
  if (V == 0) {          // release-add + acquire-poll (what decode_mega2 uses)
    __syncthreads();
    if (tid == 0) {
      epoch += gridDim.x;
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
      unsigned int seen;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory"); } while ((int)(seen - epoch) < 0);
    }
    __syncthreads();
  } else if (V == 1) {   // classic: threadfence + atomicAdd + volatile poll + threadfence
    __syncthreads();
    if (tid == 0) {
      epoch += gridDim.x;
      __threadfence();
      atomicAdd(bar, 1u);
      while ((int)(*((volatile unsigned int*)bar) - epoch) < 0) {}
      __threadfence();
    }
    __syncthreads();
  } else if (V == 2) {   // relaxed poll + one acquire fence at the end
    __syncthreads();
    if (tid == 0) {
      epoch += gridDim.x;
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
      unsigned int seen;
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory"); } while ((int)(seen - epoch) < 0);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
  } else if (V == 3) {   // per-CTA flags (distinct words), one warp polls all of them
    __syncthreads();
    epoch += 1;
    if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
    if (tid < 32) {
      bool done;
      do {
        done = true;
        for (int i = tid; i < (int)gridDim.x; i += 32) {
          unsigned int seen;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + i) : "memory");
          done = done && (int)(seen - epoch) >= 0;
        }
        done = __all_sync(0xffffffffu, done);
      } while (!done);
    }
    __syncthreads();
  } else if (V == 4) {   // cooperative groups
    cg::this_grid().sync();
  } else if (V == 5) {   // two-level: 4 sub-counters on separate lines + root
    __syncthreads();
    if (tid == 0) {
      epoch += 1;
      const int grp = blockIdx.x & 3;
      const unsigned int per = (gridDim.x + 3 - grp) / 4;           // CTAs in this group
      unsigned int prev;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(prev) : "l"(bar + 32 * (grp + 1)), "r"(1u) : "memory");
      if (prev + 1 == per * epoch) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
      unsigned int seen;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory"); } while ((int)(seen - 4 * epoch) < 0);
    }
    __syncthreads();
  }
}

template <int V>
__global__ void __launch_bounds__(256, 1) k(unsigned int* bar, unsigned int* flags, float* data, int iters, long long* out) {
  unsigned int epoch = 0;
  const int tid = threadIdx.x;
  long long t0 = 0;
  for (int it = 0; it < iters + 10; ++it) {
    if (it == 10) t0 = clock64();
    data[blockIdx.x * 256 + tid] = (float)it;      // something for the release to publish
    barrier<V>(bar, flags, epoch, tid);
  }
  if (tid == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

int main() {
  unsigned int *bar, *flags;
  float* data;
  long long* out;
  cudaMalloc(&bar, 4096);
  cudaMalloc(&flags, 4096);
  cudaMalloc(&data, 148 * 256 * 4);
  cudaMalloc(&out, 8);
  const int iters = 2000;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const char* names[] = {"red.release + ld.acquire poll", "threadfence + atomicAdd + volatile poll + threadfence",
                         "red.release + ld.relaxed poll + fence", "per-CTA flags, warp polls all", "cooperative_groups grid.sync",
                         "two-level (4 sub-counters)"};
  for (int v = 0; v < 6; ++v) {
    cudaMemset(bar, 0, 4096);
    cudaMemset(flags, 0, 4096);
    void* args[] = {&bar, &flags, &data, (void*)&iters, &out};
    const void* f = v == 0 ? (const void*)k<0> : v == 1 ? (const void*)k<1> : v == 2 ? (const void*)k<2> : v == 3 ? (const void*)k<3>
                    : v == 4 ? (const void*)k<4> : (const void*)k<5>;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(bar, 0, 4096);
      cudaMemset(flags, 0, 4096);
      cudaEventRecord(e0);
      cudaError_t err = cudaLaunchCooperativeKernel(f, dim3(148), dim3(256), args, 0, 0);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("variant %d failed\n", v); break; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      long long cyc;
      cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
      if (rep == 1) printf("V%d %-55s %7.0f cycles/barrier  %6.3f us/barrier (events)\n", v, names[v], (double)cyc / iters, ms * 1e3 / (iters + 10));
    }
  }
  return 0;
}
