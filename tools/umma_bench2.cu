// Development microbenchmark 2: TMA producer warp + MMA issuer warp (no epilogue): time per 64-wide k-block of a 128 x BN tile
// as a function of ring depth, with or without streaming the A tile (resident A = the fused LayerNorm GEMM's situation).
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

__global__ void __launch_bounds__(448, 1) bench2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int BN, int stages,
                                                 int load_a, int tiles, int nkb, int mma_on, int n_tiles_total, int pollers, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * 8 + 2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_res = base;                               // resident A: 8 k-blocks x 16 KB (load_a == 0)
  const uint32_t ring = load_a ? base : base + 8 * 16384;
  const uint32_t a_bytes = load_a ? 16384u : 0u, b_bytes = (uint32_t)BN * 128u, stage_bytes = a_bytes + b_bytes;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), done = smem_u32(&bars[16]);
  for (int i = threadIdx.x; i < 220 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const long long t0 = clock64();
  if (warp == 0) {
    uint32_t s = 0, ph = 0;
    for (int t = 0; t < tiles; ++t) {
      const int n0 = ((t + blockIdx.x) % n_tiles_total) * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full0 + 8 * s, stage_bytes);
          const uint32_t sa = ring + s * stage_bytes;
          if (load_a) tma_load_2d(sa, &tmA, full0 + 8 * s, kb * 64, blockIdx.x * 128);
          tma_load_2d(sa + a_bytes, &tmB, full0 + 8 * s, kb * 64, n0);
        }
        __syncwarp();
        if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_16_f32(128, BN, true);
    uint32_t s = 0, ph = 0;
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tmem + (t & 1) * 256;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t sa = ring + s * stage_bytes;
        const uint64_t adesc = make_kmajor_sw128_desc(load_a ? sa : a_res + kb * 16384);
        const uint64_t bdesc = make_kmajor_sw128_desc(sa + a_bytes);
        if (elect_one()) {
          if (mma_on) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
            umma_commit(empty0 + 8 * s);
          } else {
            mbar_arrive(empty0 + 8 * s);
          }
        }
        __syncwarp();
        if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
      }
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    if (threadIdx.x == 32) out[blockIdx.x] = clock64() - t0;
  }
  if (warp >= 2 && warp < 2 + (pollers & 15)) {
    // what idle epilogue warps do: poll a barrier that completes when the main loop is over.  bit 4: back off with nanosleep
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(done), "r"(0u) : "memory");
      if (!ok && (pollers & 16)) __nanosleep(200);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int ctas = 148, K = 512, NW = 2048;
  void *A, *B;
  cudaMalloc(&A, (size_t)ctas * 128 * K * 2);
  cudaMalloc(&B, (size_t)NW * K * 2);
  cudaMemset(A, 0, (size_t)ctas * 128 * K * 2);
  cudaMemset(B, 0, (size_t)NW * K * 2);
  long long *out;
  cudaMallocManaged(&out, ctas * sizeof(long long));
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int smem = 225 * 1024;
  cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int pollers : {0, 12, 28})
  for (int load_a : {1, 0})
    for (int BN : {256})
      for (int mma_on : {1})
        for (int stages = 2; stages <= 4; ++stages) {
          const int stage_bytes = (load_a ? 16384 : 0) + BN * 128;
          const int avail = 220 * 1024 - (load_a ? 0 : 8 * 16384);
          if (stages * stage_bytes > avail) continue;
          CUtensorMap tmA, tmB;
          cuuint64_t gdA[2] = {(cuuint64_t)K, (cuuint64_t)ctas * 128}, gsA[1] = {(cuuint64_t)K * 2};
          cuuint32_t boxA[2] = {64, 128}, es[2] = {1, 1};
          enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A, gdA, gsA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          cuuint64_t gdB[2] = {(cuuint64_t)K, (cuuint64_t)NW};
          cuuint32_t boxB[2] = {64, (cuuint32_t)BN};
          enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, B, gdB, gsA, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          const int n_tiles_total = NW / BN, tiles = 2 * n_tiles_total, nkb = K / 64;
          for (int rep = 0; rep < 2; ++rep) {
            bench2<<<ctas, 448, smem>>>(tmA, tmB, BN, stages, load_a, tiles, nkb, mma_on, n_tiles_total, pollers, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          }
          long long mx = 0;
          for (int i = 0; i < ctas; ++i) mx = out[i] > mx ? out[i] : mx;
          const double per_kb = (double)mx / (tiles * nkb);
          printf("pollers=%2d load_a=%d BN=%3d mma=%d stages=%d: %6.0f clk per k-block (tensor floor %d), %5.1f B/clk/SM loaded\n", pollers, load_a, BN, mma_on, stages, per_kb, BN * 2,
                 stage_bytes / per_kb);
        }
  return 0;
}
