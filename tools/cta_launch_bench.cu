// Development aid: how fast does the block scheduler start CTAs?  Every CTA records %globaltimer at entry and exit; the kernel body
// spins for `work_ns`.  Prints, per configuration, when the n-th CTA started relative to the first.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cta_launch_bench tools/cta_launch_bench.cu && tools/cta_launch_bench
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

__global__ void probe(unsigned long long *entry, unsigned long long *exit_, unsigned *sm, long work_ns, int smem) {
  extern __shared__ unsigned char dyn[];
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  if (threadIdx.x == 0) {
    unsigned s;
    asm volatile("mov.u32 %0, %smid;" : "=r"(s));
    sm[blockIdx.x] = s;
    entry[blockIdx.x] = t0;
    if (smem > 0) dyn[0] = 1;
  }
  unsigned long long t = t0;
  while ((long)(t - t0) < work_ns) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  __syncthreads();
  if (threadIdx.x == 0) exit_[blockIdx.x] = t;
}

// the same probe with ~128 live registers per thread and (TM = 1) a 128-column tensor-memory allocation, as window_attention_tc_kernel has
template <int TM>
__global__ void __launch_bounds__(128, 4) probe_fat(unsigned long long *entry, unsigned long long *exit_, const float *src, float *dst, long work_ns) {
  extern __shared__ unsigned char dyn[];
  __shared__ unsigned slot;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  if (threadIdx.x == 0) entry[blockIdx.x] = t0;
  float v[96];
#pragma unroll
  for (int i = 0; i < 96; ++i) v[i] = src[threadIdx.x + 128 * i];
  if (TM) {
    if (threadIdx.x < 32) {
      unsigned a = (unsigned)__cvta_generic_to_shared(&slot);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(128) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
  }
  unsigned long long t = t0;
  while ((long)(t - t0) < work_ns) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 96; ++i) acc += v[i] * (float)(i + 1);
  dst[blockIdx.x * 128 + threadIdx.x] = acc + dyn[threadIdx.x];
  __syncthreads();
  if (TM && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(128) : "memory");
  if (threadIdx.x == 0) exit_[blockIdx.x] = t;
}

int main() {
  const int maxg = 1 << 16;
  unsigned long long *entry, *exit_;
  unsigned *sm;
  cudaMalloc(&entry, maxg * 8);
  cudaMalloc(&exit_, maxg * 8);
  cudaMalloc(&sm, maxg * 4);
  std::vector<unsigned long long> he(maxg), hx(maxg);
  struct Cfg { int grid, threads, smem; long work; };
  const Cfg cfgs[] = {{148, 128, 0, 2000},     {592, 128, 0, 2000},     {592, 128, 29000, 2000}, {592, 128, 29000, 6000}, {2368, 128, 0, 2000},
                      {2368, 256, 0, 500},     {4736, 256, 0, 500},     {4736, 256, 0, 2000},    {392, 256, 0, 1500},     {784, 256, 0, 1500},
                      {148, 512, 100000, 2000}, {296, 512, 100000, 2000}, {1184, 64, 0, 2000},    {4736, 32, 0, 1000}};
  for (const Cfg &c : cfgs) {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    for (int rep = 0; rep < 3; ++rep) {
      probe<<<c.grid, c.threads, c.smem>>>(entry, exit_, sm, c.work, c.smem);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    }
    cudaMemcpy(he.data(), entry, c.grid * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hx.data(), exit_, c.grid * 8, cudaMemcpyDeviceToHost);
    std::vector<unsigned long long> e(he.begin(), he.begin() + c.grid);
    std::sort(e.begin(), e.end());
    const unsigned long long t0 = e[0];
    unsigned long long last_exit = 0;
    for (int i = 0; i < c.grid; ++i) last_exit = std::max(last_exit, hx[i]);
    printf("grid %5d x %3d thr, smem %6d, work %4ld ns: entry of CTA #148 +%5llu ns, #296 +%5llu, #592 +%5llu, #1184 +%5llu, last +%5llu ns; kernel span %6llu ns\n", c.grid,
           c.threads, c.smem, c.work, c.grid > 148 ? e[148] - t0 : 0ull, c.grid > 296 ? e[296] - t0 : 0ull, c.grid > 592 ? e[592] - t0 : 0ull,
           c.grid > 1184 ? e[1184] - t0 : 0ull, e[c.grid - 1] - t0, last_exit - t0);
  }
  float *src, *dst;
  cudaMalloc(&src, 128 * 96 * 4);
  cudaMalloc(&dst, 4096 * 128 * 4);
  cudaMemset(src, 0, 128 * 96 * 4);
  for (int tm = 0; tm < 2; ++tm)
    for (int grid : {148, 384, 592, 1184}) {
      for (int rep = 0; rep < 3; ++rep) {
        if (tm) probe_fat<1><<<grid, 128, 29000>>>(entry, exit_, src, dst, 4000);
        else probe_fat<0><<<grid, 128, 29000>>>(entry, exit_, src, dst, 4000);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
      }
      cudaMemcpy(he.data(), entry, grid * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(hx.data(), exit_, grid * 8, cudaMemcpyDeviceToHost);
      std::vector<unsigned long long> e(he.begin(), he.begin() + grid);
      std::sort(e.begin(), e.end());
      unsigned long long last_exit = 0;
      for (int i = 0; i < grid; ++i) last_exit = std::max(last_exit, hx[i]);
      printf("fat probe (128 thr, ~128 regs, 29 KB smem, tmem %d) grid %4d: entry of CTA #147 +%5llu ns, #148 +%5llu, #296 +%5llu, #444 +%5llu, last +%5llu; span %6llu ns\n", tm, grid,
             e[std::min(147, grid - 1)] - e[0], grid > 148 ? e[148] - e[0] : 0ull, grid > 296 ? e[296] - e[0] : 0ull, grid > 444 ? e[444] - e[0] : 0ull, e[grid - 1] - e[0], last_exit - e[0]);
    }
  return 0;
}
