// Development microbenchmark: issue rate of tcgen05.mma (kind::f16, M=128, cta_group::1, both operands in shared memory)
// without any TMA / producer handshake: one thread per CTA issues `iters` k-blocks of 4 MMAs on fixed shared-memory tiles.
#include <cstdio>
#include <cstdlib>
#include "../multilateral-temporal-view-pyramid-transformer-for-video-inpainting-detection_b200/csrc/tc_common.cuh"
namespace mumpy {
void set_error(const char *, ...) {}
int launch_status(const char *) { return 0; }
bool pdl_enabled() { return false; }
void register_f16_flag_setter(F16FlagSetter) {}
}
using namespace mumpy;
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}

// mode 0: one commit at the very end; 1: commit after every k-block (no waits); 2: commit + wait after every k-block (serial);
// 3: ring of `stages` barriers, wait for the k-block issued `stages` earlier (what a pipelined main loop does)
__global__ void __launch_bounds__(128, 1) umma_bench(int N, int iters, int mode, int stages, int a_stride, int b_stride, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[16];
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw)[i] = 0x3c003c00u;   // 1.0 halves
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (mode >= 6) {
    // minimal issue loops: fixed descriptors, nothing but the MMAs.  6: one thread; 7: the whole warp runs the loop, one elected lane issues
    if (threadIdx.x < 32) {
      const uint32_t idesc = make_idesc_16_f32(128, N, true);
      const uint64_t adesc = make_kmajor_sw128_desc(base), bdesc = make_kmajor_sw128_desc(base + 128 * 1024);
      const bool leader = threadIdx.x == 0;
      const long long t0 = clock64();
      if (mode >= 9) {
        // whole warp converged, elect.sync around the MMAs; descriptors advance by adds; 9: ring with waits, 10: + a second barrier wait per k-block (the "full" side)
        const uint32_t bstep = (uint32_t)b_stride >> 4;
        uint32_t s = 0, ph = 0;
        uint64_t bd = bdesc;
        const int nst = stages;
        for (int it = 0; it < iters; ++it) {
          if (it >= nst) mbar_wait(smem_u32(&bars[s]), ph ^ 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bd + 2 * k, idesc, (it | k) ? 1u : 0u);
            umma_commit(smem_u32(&bars[s]));
          }
          __syncwarp();
          bd += bstep;
          if (++s == (uint32_t)nst) { s = 0; ph ^= 1; bd = bdesc; }
        }
        if (elect_one()) umma_commit(smem_u32(&bars[15]));
        __syncwarp();
      } else if (mode == 6) {
        if (leader) {
          for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
          }
          umma_commit(smem_u32(&bars[15]));
        }
      } else {
        uint32_t alo = (uint32_t)adesc, blo = (uint32_t)bdesc;
        const uint32_t ahi = (uint32_t)(adesc >> 32), bhi = (uint32_t)(bdesc >> 32);
        uint32_t s = 0, ph = 0;
        for (int it = 0; it < iters; ++it) {
          if (mode == 8 && it >= stages) mbar_wait(smem_u32(&bars[s]), ph ^ 1);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem, ((uint64_t)ahi << 32) | (alo + 2 * k), ((uint64_t)bhi << 32) | (blo + 2 * k + s * (b_stride >> 4)), idesc, (it | k) ? 1u : 0u);
            if (mode == 8) umma_commit(smem_u32(&bars[s]));
          }
          __syncwarp();
          if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
        }
        if (leader) umma_commit(smem_u32(&bars[15]));
      }
      mbar_wait(smem_u32(&bars[15]), 0);
      const long long t1 = clock64();
      if (leader) out[blockIdx.x] = t1 - t0;
    }
  } else if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_16_f32(128, N, true);
    const uint32_t a0 = base, b0 = base + 128 * 1024;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t s = it % stages;
      if (mode == 3 && it >= stages) mbar_wait(smem_u32(&bars[s]), ((it / stages) - 1) & 1);
      const uint64_t adesc = make_kmajor_sw128_desc(a0 + (it % 8) * a_stride);
      const uint64_t bdesc = make_kmajor_sw128_desc(b0 + (s % (73728 / b_stride)) * b_stride);
      const uint32_t d = tmem + ((it / 8) & 1) * 256;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(mode == 4 ? tmem + (k & 1) * 256 : (mode == 5 ? tmem + k * 128 : d), adesc + 2 * k, bdesc + 2 * k, idesc, (it % 8 || k) ? 1u : 0u);
      if (mode == 1 || mode == 3) umma_commit(smem_u32(&bars[s]));
      if (mode == 2) {
        umma_commit(smem_u32(&bars[0]));
        mbar_wait(smem_u32(&bars[0]), it & 1);
      }
    }
    if (mode != 2) {
      umma_commit(smem_u32(&bars[15]));
      mbar_wait(smem_u32(&bars[15]), 0);
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main(int argc, char **argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 148;
  long long *out;
  cudaMallocManaged(&out, 148 * sizeof(long long));
  const int smem = 201 * 1024;
  cudaFuncSetAttribute(umma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  const int Ns[] = {64, 128, 256};
  for (int mode : {6, 9})
    for (int N : Ns)
      for (int stages : {2, 4, 8}) {
        if (mode != 3 && mode < 8 && stages != 4) continue;
        if (mode >= 8 && stages * N * 128 > 73728) continue;
        if (mode >= 1 && mode <= 5 && mode != 3) continue;
        if (mode == 5 && N > 128) continue;
        const int b_stride = N * 128;
        for (int rep = 0; rep < 2; ++rep) {
          umma_bench<<<ctas, 128, smem>>>(N, iters, mode, stages, 16384, b_stride, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long mx = 0, mn = 1ll << 60;
        for (int i = 0; i < ctas; ++i) { mx = out[i] > mx ? out[i] : mx; mn = out[i] < mn ? out[i] : mn; }
        printf("mode %d N=%3d stages=%d: %7.1f clk per MMA (min CTA %7.1f), floor %d  [%d CTAs]\n", mode, N, stages, (double)mx / (iters * 4), (double)mn / (iters * 4), N / 2, ctas);
      }
  return 0;
}
