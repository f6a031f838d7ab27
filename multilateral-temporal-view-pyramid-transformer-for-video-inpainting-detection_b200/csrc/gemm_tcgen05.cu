// bf16 GEMM on the 5th-gen tensor cores: TMA (128B-swizzled tiles) -> smem ring -> tcgen05.mma (one elected
// thread, fp32 accumulators in TMEM) -> tcgen05.ld epilogue with fused bias / GELU / fp32 residual.
//   D[m,n] = act( sum_k A[m,k] * W[n,k] + bias[n] ) + residual[m,n]      A (M,K), W (N,K) bf16, K-major both.
// This is the workhorse of the "bf16 mode": qkv / proj / fc1 / fc2 / pre / proj_{q,k,v,out} / reduction /
// globalembedding / global blocks / rgb_decoder linears (94% of the forward's FLOPs, SURVEY finding 3).
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lane quadrant = warp id), warp 4 TMA producer,
// warp 5 TMEM allocator + MMA issuer.  One 128 x BN output tile per CTA; BN and the ring depth are chosen
// per problem so that two CTAs are co-resident per SM (one's epilogue overlaps the other's main loop).
#include <cuda.h>

#include "common.cuh"

namespace mumpy {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;

int resolve_driver_entry_points() {
  if (g_encode_tiled) return MUMPY_OK;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return MUMPY_ERR_CUDA;
  }
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  return MUMPY_OK;
}

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;       // 64 bf16 = 128 B = one swizzle span
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_THREADS = 192;

struct TcParams {
  const float *bias;
  const float *residual;
  void *out;
  long ldo;
  long M;
  int N, K;
  int BN;
  int stages;
  int act;
  int out_bf16;
  uint32_t tmem_cols;
  uint32_t idesc;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Parity wait with a watchdog: a pipeline bug becomes a trapped launch (reported through the C ABI) instead
// of a hung GPU.  The timer is only read every 4096 failed probes, so the fast path is untouched.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (++spins & 4095u) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();     // 4 s without progress
    }
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128B-swizzled operand tile (rows of 128 B, 8-row groups 1024 B apart): UMMA::SmemDescriptor with
// start>>4 [0,14), LBO [16,30) (unused for swizzled K-major, 1), SBO=1024>>4 [32,46), version=1 [46,48),
// layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(TC_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * TC_MAX_STAGES + 1];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;        // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t a_bytes = TC_BM * 128;
  const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t full0 = smem_u32(&bars[0]);
  const uint32_t empty0 = smem_u32(&bars[TC_MAX_STAGES]);
  const uint32_t accum_bar = smem_u32(&bars[2 * TC_MAX_STAGES]);
  const int nkb = (p.K + TC_BK - 1) / TC_BK;
  const int n0 = blockIdx.x * p.BN;
  const long m0 = static_cast<long>(blockIdx.y) * TC_BM;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_arrive_expect_tx(full0 + 8 * s, stage_bytes);
        const uint32_t sa = tiles + s * stage_bytes;
        tma_load_2d(sa, &tmA, full0 + 8 * s, kb * TC_BK, static_cast<int>(m0));
        tma_load_2d(sa + a_bytes, &tmB, full0 + 8 * s, kb * TC_BK, n0);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(full0 + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = tiles + s * stage_bytes;
        const uint64_t adesc = make_kmajor_sw128_desc(sa);
        const uint64_t bdesc = make_kmajor_sw128_desc(sa + a_bytes);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the swizzle span: +2 in the (addr>>4) field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * s);      // frees the smem slot once these MMAs have read it
      }
      umma_commit(accum_bar);             // accumulator complete
    }
  } else {
    // ---------------- epilogue: thread <-> accumulator row, 32 columns per tcgen05.ld ----------------
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long gm = m0 + warp * 32 + lane;
    const bool row_ok = gm < p.M;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(lane_addr + c0, v);
      if (!row_ok) continue;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int cl = c0 + g * 8;
        const int n = n0 + cl;
        if (cl >= p.BN || n >= p.N) break;           // N % 8 == 0 is required by the host wrapper
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[g * 8 + i]);
        if (p.bias) {
          const float4 b0 = __ldg(reinterpret_cast<const float4 *>(p.bias + n));
          const float4 b1 = __ldg(reinterpret_cast<const float4 *>(p.bias + n + 4));
          f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
          f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
        }
        if (p.act != MUMPY_ACT_NONE) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = apply_act(f[i], p.act);
        }
        if (p.residual) {
          const float4 r0 = *reinterpret_cast<const float4 *>(p.residual + gm * p.ldo + n);
          const float4 r1 = *reinterpret_cast<const float4 *>(p.residual + gm * p.ldo + n + 4);
          f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
          f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
        }
        if (p.out_bf16) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(f[2], f[3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(f[6], f[7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t *>(&h0);
          u.y = *reinterpret_cast<uint32_t *>(&h1);
          u.z = *reinterpret_cast<uint32_t *>(&h2);
          u.w = *reinterpret_cast<uint32_t *>(&h3);
          *reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.out) + gm * p.ldo + n) = u;
        } else {
          float *o = reinterpret_cast<float *>(p.out) + gm * p.ldo + n;
          *reinterpret_cast<float4 *>(o) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4 *>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

static int encode_2d_bf16(CUtensorMap *map, const void *ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                          uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu stride=%llu box=%ux%u", (int)r, ptr,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_elems, box_inner, box_outer);
    return MUMPY_ERR_CUDA;
  }
  return MUMPY_OK;
}

static int pick_bn(long M, int N) {
  static const int cands[] = {256, 192, 128, 96, 64, 48, 32, 16};
  const long mt = cdiv(M, TC_BM);
  int largest = 0, smallest64 = 0;
  for (int c : cands) {                            // descending
    if (N % c != 0) continue;
    if (!largest) largest = c;
    if (mt * (N / c) >= 2 * 148) return c;         // widest tile that still gives two CTAs per SM
    if (c >= 64) smallest64 = c;
  }
  if (smallest64) return smallest64;               // small problem: favour parallelism, keep N >= 64
  if (largest) return largest;
  for (int c : cands)
    if (c <= N) return c;                          // no divisor: the tail tile is masked
  return 16;
}

static bool g_attr_set = false;

int linear_bf16(const void *A, long lda, const void *W, const float *bias, const float *residual, void *out, long ldo,
                long M, int N, int K, int out_dtype, int act, cudaStream_t st) {
  if (!g_encode_tiled) {
    int rc = resolve_driver_entry_points();
    if (rc) return rc;
  }
  MUMPY_REQUIRE(N % 8 == 0 && K % 8 == 0 && lda % 8 == 0, "linear(bf16): N, K, lda must be multiples of 8 (N=%d K=%d lda=%ld)", N, K, lda);
  MUMPY_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "linear(bf16): A, W, out must be 16-byte aligned");
  MUMPY_REQUIRE(out_dtype == MUMPY_BF16 ? (ldo % 8 == 0) : (ldo % 4 == 0), "linear(bf16): ldo alignment");
  MUMPY_REQUIRE(M < (1l << 31), "linear(bf16): M too large");
  TcParams p;
  p.bias = bias;
  p.residual = residual;
  p.out = out;
  p.ldo = ldo;
  p.M = M;
  p.N = N;
  p.K = K;
  p.BN = pick_bn(M, N);
  p.act = act;
  p.out_bf16 = (out_dtype == MUMPY_BF16);
  uint32_t cols = 32;
  while (cols < (uint32_t)p.BN) cols <<= 1;
  p.tmem_cols = cols;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(p.BN >> 3) << 17) | (static_cast<uint32_t>(TC_BM >> 4) << 24);
  const int stage_bytes = TC_BM * 128 + p.BN * 128;
  const int nkb = (K + TC_BK - 1) / TC_BK;
  int stages = (100 * 1024) / stage_bytes;           // ~100 KB of tiles -> two CTAs per SM
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > nkb) stages = nkb;
  if (stages < 1) stages = 1;
  p.stages = stages;
  const int smem = stages * stage_bytes + 1024;
  if (!g_attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_tc_kernel): %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    g_attr_set = true;
  }
  CUtensorMap tmA, tmB;
  int rc = encode_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TC_BK, TC_BM);
  if (rc) return rc;
  rc = encode_2d_bf16(&tmB, W, (uint64_t)K, (uint64_t)N, (uint64_t)K, TC_BK, (uint32_t)p.BN);
  if (rc) return rc;
  dim3 grid((unsigned)cdiv(N, p.BN), (unsigned)cdiv(M, TC_BM));
  gemm_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tmA, tmB, p);
  return launch_status("gemm_tc_kernel");
}

}  // namespace mumpy
