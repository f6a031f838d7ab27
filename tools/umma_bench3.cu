// Development microbenchmark 3: CTA pair (cta_group::2, M = 256 over two CTAs), A resident in each CTA's shared memory, each CTA
// streams HALF of every B tile through a TMA ring: time per 64-wide k-block vs ring depth.
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench3(const __grid_constant__ CUtensorMap tmB, int BN, int stages, int tiles, int nkb,
                                                                           int n_tiles_total, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * 8 + 2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_res = base;
  const uint32_t ring = base + 8 * 16384;
  const uint32_t b_bytes = (uint32_t)(BN / 2) * 128u;          // this CTA's half of a B tile
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), done = smem_u32(&bars[16]);
  for (int i = threadIdx.x; i < 220 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(full0 + 8 * s, 2); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const long long t0 = clock64();
  const uint32_t pair = blockIdx.x >> 1;
  if (warp == 0) {
    uint32_t s = 0, ph = 0;
    const uint32_t full_leader = mapa_shared(full0, 0);
    for (int t = 0; t < tiles; ++t) {
      const int n0 = ((t + pair) % n_tiles_total) * BN + rank * (BN / 2);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx_cluster(full_leader + 8 * s, b_bytes);
          tma_load_2d_pair(ring + s * b_bytes, &tmB, full_leader + 8 * s, kb * 64, n0);
        }
        __syncwarp();
        if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_16_f32(256, BN, true);
    uint32_t s = 0, ph = 0;
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tmem + (t & 1) * 256;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint64_t adesc = make_kmajor_sw128_desc(a_res + kb * 16384);
        const uint64_t bdesc = make_kmajor_sw128_desc(ring + s * b_bytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_pair(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit_pair(empty0 + 8 * s);
        }
        __syncwarp();
        if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
      }
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    if (threadIdx.x == 32) out[pair] = clock64() - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int ctas = 148, K = 512, NW = 2048;
  void *B;
  cudaMalloc(&B, (size_t)NW * K * 2);
  cudaMemset(B, 0, (size_t)NW * K * 2);
  long long *out;
  cudaMallocManaged(&out, ctas * sizeof(long long));
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int smem = 225 * 1024;
  cudaFuncSetAttribute(bench3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int BN : {256, 128})
    for (int stages = 1; stages <= 8; ++stages) {
      const int stage_bytes = BN / 2 * 128;
      if (stages * stage_bytes > 220 * 1024 - 8 * 16384) continue;
      CUtensorMap tmB;
      cuuint64_t gdB[2] = {(cuuint64_t)K, (cuuint64_t)NW}, gs[1] = {(cuuint64_t)K * 2};
      cuuint32_t boxB[2] = {64, (cuuint32_t)(BN / 2)}, es[2] = {1, 1};
      enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B, gdB, gs, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      const int n_tiles_total = NW / BN, tiles = 2 * n_tiles_total, nkb = K / 64;
      for (int rep = 0; rep < 2; ++rep) {
        bench3<<<ctas, 128, smem>>>(tmB, BN, stages, tiles, nkb, n_tiles_total, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long mx = 0;
      for (int i = 0; i < ctas / 2; ++i) mx = out[i] > mx ? out[i] : mx;
      const double per_kb = (double)mx / (tiles * nkb);
      printf("pair BN=%3d stages=%d: %6.0f clk per k-block (256 x BN x 64; tensor floor %d clk)\n", BN, stages, per_kb, BN * 2);
    }
  return 0;
}
