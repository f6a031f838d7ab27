// Fused Swin MLP for the narrow (HBM-bound) stages, C <= 256 (swinTransformer.py:305 + 45-51):
//   out = x + fc2( GELU( fc1( LayerNorm(x) ) ) )            x, out (M, C) fp32 residual stream; W1 (4C, C), W2 (C, 4C) 16-bit
// One kernel per block instead of LayerNorm + fc1 + fc2: neither the normalised rows nor the 4C-wide hidden activations ever
// reach global memory (at stage 0 of view 3, batch 32, the hidden matrix alone is 308 MB written and 308 MB read back per block).
//
// Per 128-row tile (persistent CTAs, one per SM):
//   prologue    the 16 epilogue warps normalise the tile's rows into a resident 16-bit K-major SWIZZLE_128B A operand
//               (lg_normalise_rows' arithmetic = layernorm_vec_kernel's: bit-identical operand);
//   hidden chunk j (128 of the 4C columns):
//     MMA1      acc1[j&1] (TMEM, 128 columns) = A . W1[128 j : 128 j + 128, :]^T          K = C
//     epi1      + b1, GELU, pack to 16 bits, written by the row's own thread straight into the K-major SWIZZLE_128B operand
//               H[j&1] in shared memory (two k-blocks of 128 rows x 128 B) -- the hidden tile never leaves the SM
//     MMA2      acc2 (TMEM, C columns) += H[j&1] . W2[:, 128 j : 128 j + 128]^T              K = 128
//   epi2        acc2 + b2 + x (fp32 residual re-read, coalesced) -> out.
// The MMA warp issues MMA1(j+1) before MMA2(j), so the tensor pipe works on the next chunk while the epilogue warps run the
// GELU of this one; weight tiles of both matrices stream through one TMA ring in exactly that consumption order.
// Accumulation order equals the unfused kernels' (k-blocks in ascending order), so the result is bit-identical to
// mumpy_layernorm + mumpy_linear(GELU, 16-bit) + mumpy_linear(residual).
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace mumpy {

int resolve_driver_entry_points();
int tc_encode_2d_16(CUtensorMap *map, const void *ptr, bool f16, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                    uint32_t box_inner, uint32_t box_outer);
int tc_num_sms();

constexpr int ML_BM = 128;
constexpr int ML_MAX_STAGES = 6;
constexpr int ML_KB_BYTES = ML_BM * 128;           // one 64-column k-block of a 128-row operand
constexpr int ML_SMEM_TOTAL = 226 * 1024;
// Two shapes of the same kernel (template parameters HC = hidden columns per chunk, EW = epilogue warps, four per TMEM lane
// quadrant: the EW (quadrant, 32-column) units of a hidden chunk map one to one onto the warps):
//   <128, 16, 1>  one CTA per SM, 128-column chunks, 512 TMEM columns, up to 6 ring stages                       (C = 192, 256)
//   < 64,  8, 2>  TWO CTAs per SM, 64-column chunks, 256 TMEM columns (acc1 2 x 64 + acc2 <= 128), 97 KB of shared memory,
//                 both k-blocks of a W1 chunk in one ring stage                                                  (C = 96, 128)
// In the one-CTA shape the 16 warps run the LayerNorm prologue (load latency), the GELU passes (issue) and the output pass (HBM)
// back to back while the tensor pipe and the memory system idle in turn (28-30 k clk per tile, 15 % tensor-pipe activity); two
// resident CTAs interleave those phases on the SM's schedulers.
constexpr int ML_SMEM_HALF = 112 * 1024;           // per CTA when two share an SM (227 KB - 2 x 1 KB reserved - static, halved)

struct MlpParams {
  const float *x;
  const float *gamma;
  const float *beta;
  const float *b1;
  const float *b2;
  float *out;
  long M;
  long num_tiles;
  int C;              // width (K of fc1, N of fc2)
  int nkb;            // 64-column k-blocks of the A operand (C = 96: the second holds 32 columns)
  int n_chunks;       // 4C / 128
  int stages;
  int a_bufs, acc2_bufs;      // pipelined shape: A operand / fc2 accumulator buffers (2 when C <= 128)
  int f16;
  uint32_t stage_bytes;
  uint32_t idesc1, idesc2;
  float eps;
};

__device__ __forceinline__ void ml_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// LayerNorm of rows [m0, m0 + 128) into the resident A operand (same scheme as gemm_ln_tcgen05.cu / norm.cu)
template <typename OutT, int LPR, int NV, int RI>
__device__ __forceinline__ float ml_normalise_rows(const MlpParams &p, uint32_t a_base, long m0, int warp, int lane, int n_warps, const float *gamma,
                                                   const float *beta) {
  constexpr int G = 32 / LPR;
  constexpr int RPW = G * RI;
  const int sub = lane % LPR, grp = lane / LPR;
  const int C = p.C;
  float amax = 0.0f;
  for (int r0 = warp * RPW; r0 < ML_BM; r0 += n_warps * RPW) {
    float4 v[RI][NV];
    float s[RI], q[RI];
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      const long row = m0 + r0 + r * G + grp;
#pragma unroll
      for (int u = 0; u < NV; ++u)
        v[r][u] = row < p.M ? *(reinterpret_cast<const float4 *>(p.x + row * C) + sub + LPR * u) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      s[r] = 0.0f;
#pragma unroll
      for (int u = 0; u < NV; ++u) s[r] += (v[r][u].x + v[r][u].y) + (v[r][u].z + v[r][u].w);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < RI; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      s[r] = s[r] / C;
      q[r] = 0.0f;
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const float a = v[r][u].x - s[r], b = v[r][u].y - s[r], c = v[r][u].z - s[r], d = v[r][u].w - s[r];
        q[r] += (a * a + b * b) + (c * c + d * d);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < RI; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
#pragma unroll
    for (int r = 0; r < RI; ++r) q[r] = 1.0f / sqrtf(q[r] / C + p.eps);
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int i = sub + LPR * u;
      const float4 g4 = *(reinterpret_cast<const float4 *>(gamma) + i), b4 = *(reinterpret_cast<const float4 *>(beta) + i);
      const uint32_t col_off = static_cast<uint32_t>(i >> 4) * ML_KB_BYTES + static_cast<uint32_t>(i & 1) * 8u;
      const uint32_t chunk = static_cast<uint32_t>(i & 15) >> 1;
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        const uint32_t rt = static_cast<uint32_t>(r0 + r * G + grp);
        const float rstd = q[r];
        const float o0 = (v[r][u].x - s[r]) * rstd * g4.x + b4.x, o1 = (v[r][u].y - s[r]) * rstd * g4.y + b4.y;
        const float o2 = (v[r][u].z - s[r]) * rstd * g4.z + b4.z, o3 = (v[r][u].w - s[r]) * rstd * g4.w + b4.w;
        if constexpr (is_half_t<OutT>::value) amax = fmaxf(fmaxf(amax, fabsf(o0)), fmaxf(fmaxf(fabsf(o1), fabsf(o2)), fabsf(o3)));
        const uint32_t w0 = pack2<OutT>(o0, o1), w1 = pack2<OutT>(o2, o3);
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_base + col_off + rt * 128u + ((chunk ^ (rt & 7u)) << 4)), "r"(w0), "r"(w1)
                     : "memory");
      }
    }
  }
  return amax;
}

// (out of line, like the output epilogue below: both run once per tile, and inlined next to the hidden epilogue they make it spill)
template <typename OutT>
__device__ __noinline__ float ml_normalise(const MlpParams &p, uint32_t a_base, long m0, int warp, int lane, int n_warps, const float *gamma, const float *beta) {
  switch (p.C) {
    case 96: return ml_normalise_rows<OutT, 8, 3, 2>(p, a_base, m0, warp, lane, n_warps, gamma, beta);
    case 128: return ml_normalise_rows<OutT, 8, 4, 2>(p, a_base, m0, warp, lane, n_warps, gamma, beta);
    case 192: return ml_normalise_rows<OutT, 16, 3, 2>(p, a_base, m0, warp, lane, n_warps, gamma, beta);
    default: return ml_normalise_rows<OutT, 16, 4, 2>(p, a_base, m0, warp, lane, n_warps, gamma, beta);
  }
}

// epi1: one warp's unit of a 128 x 128 fc1 accumulator (lane quadrant warp&3, 32-column chunk warp>>2; 16 warps = 16 units): + b1,
// GELU, pack, written as the K-major SWIZZLE_128B A operand of fc2 (row = this thread's accumulator row: no transposition needed).
// `bias` = the unit's 32 bias values, loaded by the caller BEFORE it waits for the accumulator (ncu: the FADD2 consuming a bias
// loaded after the wait was the top non-barrier stall of the kernel).
template <typename OutT>
__device__ __forceinline__ float ml_hidden_epilogue(const float4 (&bias)[8], uint32_t h_base, uint32_t acc, int warp, int lane) {
  const int quad = warp & 3, c0 = (warp >> 2) * 32;
  const uint32_t lane_addr = acc + (static_cast<uint32_t>(quad * 32) << 16);
  const uint32_t row = static_cast<uint32_t>(quad * 32 + lane);
  float amax = 0.0f;
  uint32_t v[32];
  tmem_ld32(lane_addr + c0, v);
  const uint32_t kb_base = h_base + static_cast<uint32_t>(c0 >> 6) * ML_KB_BYTES + row * 128u;
  const uint32_t chunk0 = static_cast<uint32_t>(c0 & 63) >> 3;                 // first 16-byte chunk of these 32 columns in the 128-byte row
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float2 f[4];
    f[0] = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1])), make_float2(bias[2 * g].x, bias[2 * g].y));
    f[1] = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])), make_float2(bias[2 * g].z, bias[2 * g].w));
    f[2] = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])), make_float2(bias[2 * g + 1].x, bias[2 * g + 1].y));
    f[3] = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])), make_float2(bias[2 * g + 1].z, bias[2 * g + 1].w));
    gelu_fast2_x4(f);
    if constexpr (is_half_t<OutT>::value) {
#pragma unroll
      for (int h = 0; h < 4; ++h) amax = fmaxf(amax, fmaxf(fabsf(f[h].x), fabsf(f[h].y)));
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(kb_base + (((chunk0 + g) ^ (row & 7u)) << 4)), "r"(pack2<OutT>(f[0].x, f[0].y)),
                 "r"(pack2<OutT>(f[1].x, f[1].y)), "r"(pack2<OutT>(f[2].x, f[2].y)), "r"(pack2<OutT>(f[3].x, f[3].y))
                 : "memory");
  }
  return amax;
}

// epi2: one warp's share of the 128 x C fc2 accumulator: + b2 + residual -> fp32 out, through a 4 KB swizzled staging tile so
// that residual loads and output stores are full 128-byte row segments
__device__ __noinline__ void ml_output_epilogue(const MlpParams &p, uint32_t st_base, uint32_t acc, int warp, int lane, long m0, int n_groups) {
  const int quad = warp & 3, grp = warp >> 2;
  const uint32_t lane_addr = acc + (static_cast<uint32_t>(quad * 32) << 16);
  const long row0 = m0 + quad * 32;
  const int c4 = lane & 7;
  for (int c0 = grp * 32; c0 < p.C; c0 += 32 * n_groups) {
    const int col = c0 + c4 * 4;
    float4 res[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long gm = row0 + i * 4 + (lane >> 3);
      res[i] = gm < p.M ? *reinterpret_cast<const float4 *>(p.x + gm * p.C + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t v[32];
    tmem_ld32(lane_addr + c0, v);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 b = __ldg(reinterpret_cast<const float4 *>(p.b2 + c0 + 4 * g));
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_base + lane * 128 + ((g ^ (lane & 7)) << 4)), "f"(__uint_as_float(v[4 * g]) + b.x),
                   "f"(__uint_as_float(v[4 * g + 1]) + b.y), "f"(__uint_as_float(v[4 * g + 2]) + b.z), "f"(__uint_as_float(v[4 * g + 3]) + b.w)
                   : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      const long gm = row0 + r;
      float4 o;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(st_base + r * 128 + ((c4 ^ (r & 7)) << 4)));
      if (gm < p.M) {
        o.x += res[i].x; o.y += res[i].y; o.z += res[i].z; o.w += res[i].w;
        *reinterpret_cast<float4 *>(p.out + gm * p.C + col) = o;
      }
    }
    __syncwarp();
  }
}

// (Register files are per scheduler partition, 16 K registers each, and a CTA's warps are dealt round-robin: the 10 warps of the
// two-per-SM shape put 3 on a partition, so two CTAs need 6 x 32 x regs <= 16384, i.e. 80 registers -- the bound is declared as
// 384 threads x 2 to make ptxas target that, the kernel is launched with 320.)
template <int ML_HC, int ML_EPI_WARPS, int MIN_CTAS>
__global__ void __launch_bounds__(MIN_CTAS == 2 ? 384 : (ML_EPI_WARPS + 2) * 32, MIN_CTAS) mlp_fused_tc_kernel(const __grid_constant__ CUtensorMap tmW1,
                                                                                         const __grid_constant__ CUtensorMap tmW2, const MlpParams p) {
  constexpr int ML_EPI_GROUPS = ML_EPI_WARPS / 4;
  static_assert(ML_EPI_GROUPS * 32 == ML_HC, "one 32-column unit of a hidden chunk per epilogue warp");
  constexpr int ML_H_KB = ML_HC / 64;                       // k-blocks of one hidden chunk (the A operand of fc2)
  constexpr int ML_H_BYTES = ML_H_KB * ML_KB_BYTES;         // one hidden chunk: 128 rows x HC columns x 16 bit
  static_assert(2 * ML_H_BYTES >= ML_EPI_WARPS * 4096, "epi2 staging (4 KB per warp) lives in the two hidden-chunk buffers");
  constexpr bool kGroupW1 = ML_HC == 64;                    // all k-blocks of a W1 chunk share one ring stage (8 KB each)
  constexpr uint32_t ML_TMEM_COLS = ML_HC == 64 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * ML_MAX_STAGES + 11];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t a_base = (raw + 1023u) & ~1023u;
  const uint32_t h_base = a_base + static_cast<uint32_t>(p.nkb) * ML_KB_BYTES;          // two hidden-chunk operands (also the epi2 staging)
  const uint32_t ring = h_base + 2 * ML_H_BYTES;
  const uint32_t w_full0 = smem_u32(&bars[0]);
  const uint32_t w_empty0 = smem_u32(&bars[ML_MAX_STAGES]);
  const uint32_t a_full = smem_u32(&bars[2 * ML_MAX_STAGES]);
  const uint32_t acc1_full0 = smem_u32(&bars[2 * ML_MAX_STAGES + 1]);       // [2]
  const uint32_t acc1_empty0 = smem_u32(&bars[2 * ML_MAX_STAGES + 3]);      // [2]
  const uint32_t h_full0 = smem_u32(&bars[2 * ML_MAX_STAGES + 5]);          // [2]
  const uint32_t h_empty0 = smem_u32(&bars[2 * ML_MAX_STAGES + 7]);         // [2]
  const uint32_t acc2_full = smem_u32(&bars[2 * ML_MAX_STAGES + 9]);
  const uint32_t acc2_empty = smem_u32(&bars[2 * ML_MAX_STAGES + 10]);

  if (warp == ML_EPI_WARPS && lane == 0) {
    prefetch_tensormap(&tmW1);
    prefetch_tensormap(&tmW2);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(w_full0 + 8 * s, 1);
      mbar_init(w_empty0 + 8 * s, 1);
    }
    mbar_init(a_full, ML_EPI_WARPS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc1_full0 + 8 * s, 1);
      mbar_init(acc1_empty0 + 8 * s, ML_EPI_WARPS);
      mbar_init(h_full0 + 8 * s, ML_EPI_WARPS);
      mbar_init(h_empty0 + 8 * s, 1);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, ML_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == ML_EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ML_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t acc2 = tmem_base + 2 * ML_HC;                   // acc1[0] | acc1[1] | acc2 (C <= 256 columns)
  pdl_grid_sync();

  const int n = p.n_chunks;
  if (warp == ML_EPI_WARPS) {
    // ------------------------------------------------ TMA producer: weight tiles in the MMA warp's consumption order ----
    uint32_t s = 0, ph = 0;
    const uint32_t w1_bytes = ML_HC * 128u, w2_bytes = static_cast<uint32_t>(p.C) * 128u;
    auto load = [&](const CUtensorMap *map, uint32_t bytes, int kx, int ry) {
      mbar_wait(w_empty0 + 8 * s, ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(w_full0 + 8 * s, bytes);
        tma_load_2d(ring + s * p.stage_bytes, map, w_full0 + 8 * s, kx, ry);
      }
      __syncwarp();
      if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
    };
    auto load_w1 = [&](int chunk) {
      if constexpr (kGroupW1) {
        mbar_wait(w_empty0 + 8 * s, ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(w_full0 + 8 * s, static_cast<uint32_t>(p.nkb) * w1_bytes);
          for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(ring + s * p.stage_bytes + kb * w1_bytes, &tmW1, w_full0 + 8 * s, kb * 64, chunk * ML_HC);
        }
        __syncwarp();
        if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
      } else {
        for (int kb = 0; kb < p.nkb; ++kb) load(&tmW1, w1_bytes, kb * 64, chunk * ML_HC);
      }
    };
    for (long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      load_w1(0);                                                                                             // MMA1(0)
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) load_w1(j + 1);                                                                        // MMA1(j+1)
        for (int kb = 0; kb < ML_H_KB; ++kb) load(&tmW2, w2_bytes, j * ML_HC + kb * 64, 0);                   // MMA2(j)
      }
    }
  } else if (warp == ML_EPI_WARPS + 1) {
    // ------------------------------------------------ MMA issuer ---------------------------------------------------------
    uint32_t s = 0, ph = 0;
    uint32_t c1 = 0;          // fc1 chunks issued so far (acc1 slot = c1 & 1, phase (c1 >> 1) & 1)
    uint32_t c2 = 0;          // fc2 chunks issued so far (H slot = c2 & 1)
    uint32_t ti = 0;
    auto mma1 = [&]() {
      const uint32_t slot = c1 & 1, aph = (c1 >> 1) & 1;
      mbar_wait(acc1_empty0 + 8 * slot, aph ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + slot * ML_HC;
      if constexpr (kGroupW1) {
        mbar_wait(w_full0 + 8 * s, ph);
        tc_fence_after();
      }
      for (int kb = 0; kb < p.nkb; ++kb) {
        if constexpr (!kGroupW1) {
          mbar_wait(w_full0 + 8 * s, ph);
          tc_fence_after();
        }
        const uint64_t adesc = make_kmajor_sw128_desc(a_base + kb * ML_KB_BYTES);
        const uint64_t bdesc = make_kmajor_sw128_desc(ring + s * p.stage_bytes + (kGroupW1 ? kb * ML_HC * 128 : 0));
        const bool tail = (kb + 1) * 64 > p.C;
        const bool release = !kGroupW1 || kb + 1 == p.nkb;
        if (elect_one()) {
          umma_bf16(d, adesc, bdesc, p.idesc1, kb > 0 ? 1u : 0u);
          umma_bf16(d, adesc + 2, bdesc + 2, p.idesc1, 1u);
          if (!tail) {
            umma_bf16(d, adesc + 4, bdesc + 4, p.idesc1, 1u);
            umma_bf16(d, adesc + 6, bdesc + 6, p.idesc1, 1u);
          }
          if (release) umma_commit(w_empty0 + 8 * s);
        }
        __syncwarp();
        if (release && ++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(acc1_full0 + 8 * slot);
      __syncwarp();
      ++c1;
    };
    for (long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++ti) {
      mbar_wait(a_full, ti & 1);
      tc_fence_after();
      mma1();
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) mma1();
        const uint32_t hs = c2 & 1, hph = (c2 >> 1) & 1;
        mbar_wait(h_full0 + 8 * hs, hph);                         // GELU(fc1 chunk j) is in shared memory
        tc_fence_after();
        if (j == 0) {
          mbar_wait(acc2_empty, (ti & 1) ^ 1);                    // the previous tile's output has left the accumulator
          tc_fence_after();
        }
        for (int kb = 0; kb < ML_H_KB; ++kb) {
          mbar_wait(w_full0 + 8 * s, ph);
          tc_fence_after();
          const uint64_t adesc = make_kmajor_sw128_desc(h_base + hs * ML_H_BYTES + kb * ML_KB_BYTES);
          const uint64_t bdesc = make_kmajor_sw128_desc(ring + s * p.stage_bytes);
          if (elect_one()) {
            umma_bf16(acc2, adesc, bdesc, p.idesc2, (j > 0 || kb > 0) ? 1u : 0u);
            umma_bf16(acc2, adesc + 2, bdesc + 2, p.idesc2, 1u);
            umma_bf16(acc2, adesc + 4, bdesc + 4, p.idesc2, 1u);
            umma_bf16(acc2, adesc + 6, bdesc + 6, p.idesc2, 1u);
            umma_commit(w_empty0 + 8 * s);
          }
          __syncwarp();
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) {
          umma_commit(h_empty0 + 8 * hs);                         // H[hs] may be overwritten once these MMAs have read it
          if (j == n - 1) umma_commit(acc2_full);
        }
        __syncwarp();
        ++c2;
      }
    }
  } else {
    // ------------------------------------------------ LayerNorm prologue, hidden epilogue, output epilogue ---------------
    uint32_t c = 0, ti = 0;        // hidden chunks processed so far
    float amax = 0.0f;
    const uint32_t st_base = h_base + warp * 4096;             // epi2 staging lives in the hidden-chunk buffers (idle by then)
#ifdef ML_TIMING
    long long mt[16];
#endif
    for (long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++ti) {
      const long m0 = tile * ML_BM;
#ifdef ML_TIMING
      mt[0] = clock64();
#endif
      // (every fc1 MMA of the previous tile has completed: this warp waited for its last acc1_full)
      amax = fmaxf(amax, p.f16 ? ml_normalise<__half>(p, a_base, m0, warp, lane, ML_EPI_WARPS, p.gamma, p.beta) : ml_normalise<__nv_bfloat16>(p, a_base, m0, warp, lane, ML_EPI_WARPS, p.gamma, p.beta));
      ml_fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
#ifdef ML_TIMING
      mt[1] = clock64();
#endif
      for (int j = 0; j < n; ++j, ++c) {
        const uint32_t slot = c & 1, cph = (c >> 1) & 1;
        float4 bias[8];                                          // this warp's 32 fc1 bias values of chunk j: in flight during the wait
#pragma unroll
        for (int i = 0; i < 8; ++i) bias[i] = __ldg(reinterpret_cast<const float4 *>(p.b1 + j * ML_HC + (warp >> 2) * 32) + i);
        mbar_wait(acc1_full0 + 8 * slot, cph);
        mbar_wait(h_empty0 + 8 * slot, cph ^ 1);                 // fc2 of the chunk two back has read H[slot]
        tc_fence_after();
#ifdef ML_TIMING
        if (j < 4) mt[2 + 2 * j] = clock64();
#endif
        const uint32_t acc1 = tmem_base + slot * ML_HC;
        const float a = p.f16 ? ml_hidden_epilogue<__half>(bias, h_base + slot * ML_H_BYTES, acc1, warp, lane)
                              : ml_hidden_epilogue<__nv_bfloat16>(bias, h_base + slot * ML_H_BYTES, acc1, warp, lane);
        amax = fmaxf(amax, a);
        tc_fence_before();
        ml_fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(acc1_empty0 + 8 * slot);
          mbar_arrive(h_full0 + 8 * slot);
        }
#ifdef ML_TIMING
        if (j < 4) mt[3 + 2 * j] = clock64();
#endif
      }
      mbar_wait(acc2_full, ti & 1);                               // all fc2 MMAs done: H buffers are free for the staging tiles
      tc_fence_after();
#ifdef ML_TIMING
      mt[10] = clock64();
#endif
      ml_output_epilogue(p, st_base, acc2, warp, lane, m0, ML_EPI_GROUPS);
#ifdef ML_TIMING
      mt[11] = clock64();
      if (blockIdx.x == 0 && (warp == 0 || warp == ML_EPI_WARPS - 1) && lane == 0 && ti >= 1 && ti < 4)
        printf("tile %u warp %d: prologue %lld | c0 wait %lld work %lld | c1 wait %lld work %lld | c2 wait %lld work %lld | c3 wait %lld work %lld | wait acc2 %lld | epi2 %lld | total %lld\n",
               ti, warp, mt[1] - mt[0], mt[2] - mt[1], mt[3] - mt[2], mt[4] - mt[3], mt[5] - mt[4], mt[6] - mt[5], mt[7] - mt[6], mt[8] - mt[7], mt[9] - mt[8],
               mt[10] - mt[9], mt[11] - mt[10], mt[11] - mt[0]);
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
    }
    f16_guard(amax);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ML_EPI_WARPS + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ML_TMEM_COLS) : "memory");
}

// =====================================================================================================================
// Pipelined shape (default): the same arithmetic with the three kinds of epilogue work on SEPARATE warp groups and two tiles in
// flight.  In the kernel above one group of 16 warps runs, per 128-row tile, the LayerNorm prologue (global-load latency), the
// GELU passes (issue / MUFU) and the output pass (global-load latency + HBM) strictly one after the other: 28-31 k clk per tile
// of which the tensor pipe works 4 k (-DML_TIMING).  Here
//     warps  0- 7  GELU group    hidden chunk c: two (quadrant, 32-column) units per warp, biases from shared memory
//     warps  8-11  output group  tile i-1: acc2[(i-1)&1] + b2 + x -> out       (a lane quadrant each, every column unit)
//     warps 12-17  LayerNorm     tile i+1: rows -> A[(i+1)&1]  (rows prefetched into L2 a tile earlier by the producer warp)
//     warp  18     TMA producer, warp 19 MMA issuer (one continuous chunk stream: MMA1 of the next chunk -- also across the
//                  tile boundary -- is issued before MMA2 of this one)
// so each group's latency hides behind the other groups' work on neighbouring tiles.  20 warps = 640 threads keep 96 registers per
// thread (22 or more warps round up to 24 in the register file: 80 registers and spills in the GELU loop).  Measured at M = 301056,
// C = 128 (serial shape 255-260 us): LayerNorm 4 / GELU 8 warps 215 us (LayerNorm group saturated, 17.7 k clk per tile), 8 / 8
// 203 us, all 16 compute warps alternating GELU and the next tile's LayerNorm 240-252 us (the two are serial again).
// Double-buffered A operand and fc2
// accumulator where they fit (C <= 128: 2 x 32 KB, 2 x 128 TMEM columns next to the two fc1 accumulators); single buffers for
// C = 192 / 256 (the groups still overlap within and across the tile boundary, the LayerNorm of tile i+1 then starts when the
// last MMA1 of tile i has read A).  Same k-block order, same per-element arithmetic: bit-identical to the kernel above.
constexpr int MP_GELU_WARPS = 8, MP_OUT_WARPS = 4, MP_LN_WARPS = 6;
constexpr int MP_WARP_OUT = MP_GELU_WARPS, MP_WARP_LN = MP_WARP_OUT + MP_OUT_WARPS, MP_WARP_TMA = MP_WARP_LN + MP_LN_WARPS, MP_WARP_MMA = MP_WARP_TMA + 1;
constexpr int MP_THREADS = (MP_WARP_MMA + 1) * 32;
constexpr int MP_HC = 128;
constexpr int MP_H_BYTES = 2 * ML_KB_BYTES;
constexpr int MP_STAGING_BYTES = MP_OUT_WARPS * 4096;

// tcgen05.ld of 32 accumulator columns without the wait (the caller issues several, then one tcgen05.wait::ld)
__device__ __forceinline__ void mp_tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
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
}

// one (quadrant, 32-column) unit of a hidden chunk (accumulator columns already in v): ml_hidden_epilogue's arithmetic with the bias
// read from shared memory
template <typename OutT>
__device__ __forceinline__ float mp_hidden_unit(const uint32_t (&v)[32], const float *bias_s, uint32_t h_base, int quad, int c0, int lane) {
  const uint32_t row = static_cast<uint32_t>(quad * 32 + lane);
  float amax = 0.0f;
  const uint32_t kb_base = h_base + static_cast<uint32_t>(c0 >> 6) * ML_KB_BYTES + row * 128u;
  const uint32_t chunk0 = static_cast<uint32_t>(c0 & 63) >> 3;
  const float4 *b4 = reinterpret_cast<const float4 *>(bias_s);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 ba = b4[2 * g], bb = b4[2 * g + 1];
    float2 f[4];
    f[0] = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1])), make_float2(ba.x, ba.y));
    f[1] = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])), make_float2(ba.z, ba.w));
    f[2] = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])), make_float2(bb.x, bb.y));
    f[3] = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])), make_float2(bb.z, bb.w));
    gelu_fast2_x4(f);
    if constexpr (is_half_t<OutT>::value) {
#pragma unroll
      for (int h = 0; h < 4; ++h) amax = fmaxf(amax, fmaxf(fabsf(f[h].x), fabsf(f[h].y)));
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(kb_base + (((chunk0 + g) ^ (row & 7u)) << 4)), "r"(pack2<OutT>(f[0].x, f[0].y)),
                 "r"(pack2<OutT>(f[1].x, f[1].y)), "r"(pack2<OutT>(f[2].x, f[2].y)), "r"(pack2<OutT>(f[3].x, f[3].y))
                 : "memory");
  }
  return amax;
}

// output pass of one lane quadrant over every 32-column unit (ml_output_epilogue with b2 in shared memory): the residual rows of
// the NEXT unit are requested before this unit's accumulator is read, so their latency hides behind the staging round trip
__device__ __forceinline__ void mp_output_quadrant(const MlpParams &p, const float *b2_s, uint32_t st_base, uint32_t acc, int quad, int lane, long m0) {
  const uint32_t lane_addr = acc + (static_cast<uint32_t>(quad * 32) << 16);
  const long row0 = m0 + quad * 32;
  const int c4 = lane & 7;
  for (int c0 = 0; c0 < p.C; c0 += 32) {
    const int col = c0 + c4 * 4;
    float4 res[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long gm = row0 + i * 4 + (lane >> 3);
      res[i] = gm < p.M ? *reinterpret_cast<const float4 *>(p.x + gm * p.C + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t v[32];
    tmem_ld32(lane_addr + c0, v);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 b = *reinterpret_cast<const float4 *>(b2_s + c0 + 4 * g);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_base + lane * 128 + ((g ^ (lane & 7)) << 4)), "f"(__uint_as_float(v[4 * g]) + b.x),
                   "f"(__uint_as_float(v[4 * g + 1]) + b.y), "f"(__uint_as_float(v[4 * g + 2]) + b.z), "f"(__uint_as_float(v[4 * g + 3]) + b.w)
                   : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3);
      const long gm = row0 + r;
      float4 o;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(st_base + r * 128 + ((c4 ^ (r & 7)) << 4)));
      if (gm < p.M) {
        o.x += res[i].x; o.y += res[i].y; o.z += res[i].z; o.w += res[i].w;
        *reinterpret_cast<float4 *>(p.out + gm * p.C + col) = o;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(MP_THREADS, 1) mlp_pipe_tc_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                                                                   const MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * ML_MAX_STAGES + 16];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t a_base = (raw + 1023u) & ~1023u;
  const uint32_t a_bytes = static_cast<uint32_t>(p.nkb) * ML_KB_BYTES;
  const uint32_t h_base = a_base + static_cast<uint32_t>(p.a_bufs) * a_bytes;
  const uint32_t stage_base = h_base + 2 * MP_H_BYTES;                      // output staging, 4 KB per output warp
  const uint32_t bias_base = stage_base + MP_STAGING_BYTES;                 // b1 (4C floats) | b2 (C) | gamma (C) | beta (C)
  const uint32_t ring = (bias_base + static_cast<uint32_t>(7 * p.C) * 4u + 1023u) & ~1023u;
  float *b1_s = reinterpret_cast<float *>(smem_raw + (bias_base - raw));
  float *b2_s = b1_s + 4 * p.C;
  float *g_s = b2_s + p.C;
  const uint32_t w_full0 = smem_u32(&bars[0]);
  const uint32_t w_empty0 = smem_u32(&bars[ML_MAX_STAGES]);
  const uint32_t bb = smem_u32(&bars[2 * ML_MAX_STAGES]);
  const uint32_t a_full0 = bb, a_empty0 = bb + 16, acc1_full0 = bb + 32, acc1_empty0 = bb + 48, h_full0 = bb + 64, h_empty0 = bb + 80, acc2_full0 = bb + 96,
                 acc2_empty0 = bb + 112;                                   // two barriers each

  if (warp == MP_WARP_MMA) {          // tensor memory first (see attention_tc.cu: the SM holds back later CTAs until the permit is relinquished)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == MP_WARP_TMA && lane == 0) {
    prefetch_tensormap(&tmW1);
    prefetch_tensormap(&tmW2);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(w_full0 + 8 * s, 1);
      mbar_init(w_empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full0 + 8 * s, MP_LN_WARPS);
      mbar_init(a_empty0 + 8 * s, 1);
      mbar_init(acc1_full0 + 8 * s, 1);
      mbar_init(acc1_empty0 + 8 * s, MP_GELU_WARPS);
      mbar_init(h_full0 + 8 * s, MP_GELU_WARPS);
      mbar_init(h_empty0 + 8 * s, 1);
      mbar_init(acc2_full0 + 8 * s, 1);
      mbar_init(acc2_empty0 + 8 * s, MP_OUT_WARPS);
    }
    fence_barrier_init();
  }
  // the biases are parameters, not a predecessor's output: staged ahead of the dependency wait
  for (int i = threadIdx.x; i < 7 * p.C; i += MP_THREADS)
    b1_s[i] = i < 4 * p.C ? __ldg(p.b1 + i) : (i < 5 * p.C ? __ldg(p.b2 + (i - 4 * p.C)) : (i < 6 * p.C ? __ldg(p.gamma + (i - 5 * p.C)) : __ldg(p.beta + (i - 6 * p.C))));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t acc2_0 = tmem_base + 2 * MP_HC;                  // acc1[0] | acc1[1] | acc2[0] (| acc2[1] at +128 when C <= 128)
  pdl_grid_sync();

  const int n = p.n_chunks;
  const uint32_t A = static_cast<uint32_t>(p.a_bufs), Q = static_cast<uint32_t>(p.acc2_bufs);
  const uint32_t n_my = blockIdx.x < p.num_tiles ? static_cast<uint32_t>((p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
  const bool ahead = A == 2;          // MMA1 of the next tile's first chunk goes ahead of this tile's last MMA2 only with two A buffers

  if (warp == MP_WARP_TMA) {
    // ------------------------------------------------ TMA producer: weight tiles in the MMA warp's consumption order ----
    uint32_t s = 0, ph = 0;
    const uint32_t w1_bytes = MP_HC * 128u, w2_bytes = static_cast<uint32_t>(p.C) * 128u;
    auto load = [&](const CUtensorMap *map, uint32_t bytes, int kx, int ry) {
      mbar_wait(w_empty0 + 8 * s, ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(w_full0 + 8 * s, bytes);
        tma_load_2d(ring + s * p.stage_bytes, map, w_full0 + 8 * s, kx, ry);
      }
      __syncwarp();
      if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
    };
    auto load_w1 = [&](int chunk) {
      for (int kb = 0; kb < p.nkb; ++kb) load(&tmW1, w1_bytes, kb * 64, chunk * MP_HC);
    };
    // the rows of a tile are contiguous (128 x C floats): a bulk L2 prefetch a tile period ahead of the LayerNorm that reads them
    // turns its exposed global-load latency from HBM (2-4 k clk under load) into an L2 hit
    auto prefetch_rows = [&](uint32_t i) {
      if (i < n_my && elect_one()) {
        const long m0 = (static_cast<long>(blockIdx.x) + static_cast<long>(i) * gridDim.x) * ML_BM;
        const long rows = p.M - m0 < ML_BM ? p.M - m0 : ML_BM;
        const uint32_t bytes = static_cast<uint32_t>(rows * p.C * 4);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.x + m0 * p.C), "r"(bytes) : "memory");
      }
      __syncwarp();
    };
    prefetch_rows(1);
    for (uint32_t i = 0; i < n_my; ++i) {
      prefetch_rows(i + 2);
      if (i == 0 || !ahead) load_w1(0);
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) load_w1(j + 1);
        else if (ahead && i + 1 < n_my) load_w1(0);
        for (int kb = 0; kb < 2; ++kb) load(&tmW2, w2_bytes, j * MP_HC + kb * 64, 0);
      }
    }
  } else if (warp == MP_WARP_MMA) {
    // ------------------------------------------------ MMA issuer ---------------------------------------------------------
    uint32_t s = 0, ph = 0;
    uint32_t c1 = 0, c2 = 0;          // fc1 / fc2 chunks issued so far
    auto mma1 = [&](uint32_t i, int j) {          // fc1 chunk j of this CTA's i-th tile
      const uint32_t ab = i % A;
      if (j == 0) {
        mbar_wait(a_full0 + 8 * ab, (i / A) & 1);
        tc_fence_after();
      }
      const uint32_t slot = c1 & 1, aph = (c1 >> 1) & 1;
      mbar_wait(acc1_empty0 + 8 * slot, aph ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + slot * MP_HC;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(w_full0 + 8 * s, ph);
        tc_fence_after();
        const uint64_t adesc = make_kmajor_sw128_desc(a_base + ab * a_bytes + kb * ML_KB_BYTES);
        const uint64_t bdesc = make_kmajor_sw128_desc(ring + s * p.stage_bytes);
        const bool tail = (kb + 1) * 64 > p.C;
        if (elect_one()) {
          umma_bf16(d, adesc, bdesc, p.idesc1, kb > 0 ? 1u : 0u);
          umma_bf16(d, adesc + 2, bdesc + 2, p.idesc1, 1u);
          if (!tail) {
            umma_bf16(d, adesc + 4, bdesc + 4, p.idesc1, 1u);
            umma_bf16(d, adesc + 6, bdesc + 6, p.idesc1, 1u);
          }
          umma_commit(w_empty0 + 8 * s);
        }
        __syncwarp();
        if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) {
        umma_commit(acc1_full0 + 8 * slot);
        if (j == n - 1) umma_commit(a_empty0 + 8 * ab);          // every MMA that reads this tile's A operand has been issued
      }
      __syncwarp();
      ++c1;
    };
    for (uint32_t i = 0; i < n_my; ++i) {
      if (i == 0 || !ahead) mma1(i, 0);
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) mma1(i, j + 1);
        else if (ahead && i + 1 < n_my) mma1(i + 1, 0);
        const uint32_t hs = c2 & 1, hph = (c2 >> 1) & 1;
        mbar_wait(h_full0 + 8 * hs, hph);                         // GELU(fc1 chunk) is in shared memory
        tc_fence_after();
        const uint32_t qb = i % Q;
        if (j == 0) {
          mbar_wait(acc2_empty0 + 8 * qb, ((i / Q) & 1) ^ 1);     // the output group has drained this accumulator
          tc_fence_after();
        }
        const uint32_t acc2 = acc2_0 + qb * 128;
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(w_full0 + 8 * s, ph);
          tc_fence_after();
          const uint64_t adesc = make_kmajor_sw128_desc(h_base + hs * MP_H_BYTES + kb * ML_KB_BYTES);
          const uint64_t bdesc = make_kmajor_sw128_desc(ring + s * p.stage_bytes);
          if (elect_one()) {
            umma_bf16(acc2, adesc, bdesc, p.idesc2, (j > 0 || kb > 0) ? 1u : 0u);
            umma_bf16(acc2, adesc + 2, bdesc + 2, p.idesc2, 1u);
            umma_bf16(acc2, adesc + 4, bdesc + 4, p.idesc2, 1u);
            umma_bf16(acc2, adesc + 6, bdesc + 6, p.idesc2, 1u);
            umma_commit(w_empty0 + 8 * s);
          }
          __syncwarp();
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) {
          umma_commit(h_empty0 + 8 * hs);
          if (j == n - 1) umma_commit(acc2_full0 + 8 * qb);
        }
        __syncwarp();
        ++c2;
      }
    }
  } else if (warp >= MP_WARP_LN) {
    // ------------------------------------------------ LayerNorm group: one tile ahead --------------------------------------
    const int lw = warp - MP_WARP_LN;
    float amax = 0.0f;
#ifdef ML_TIMING
    long long tw_ = 0, tb_ = 0;
#endif
    for (uint32_t i = 0; i < n_my; ++i) {
      const long m0 = (static_cast<long>(blockIdx.x) + static_cast<long>(i) * gridDim.x) * ML_BM;
      const uint32_t ab = i % A;
#ifdef ML_TIMING
      const long long tq0 = clock64();
#endif
      mbar_wait(a_empty0 + 8 * ab, ((i / A) & 1) ^ 1);              // the MMAs of the tile that used this buffer have completed
#ifdef ML_TIMING
      const long long tq1 = clock64();
#endif
      const uint32_t dst = a_base + ab * a_bytes;
      amax = fmaxf(amax, p.f16 ? ml_normalise<__half>(p, dst, m0, lw, lane, MP_LN_WARPS, g_s, g_s + p.C) : ml_normalise<__nv_bfloat16>(p, dst, m0, lw, lane, MP_LN_WARPS, g_s, g_s + p.C));
      ml_fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full0 + 8 * ab);
#ifdef ML_TIMING
      if (i > 0) { tw_ += tq1 - tq0; tb_ += clock64() - tq1; }
#endif
    }
#ifdef ML_TIMING
    if (blockIdx.x == 3 && lane == 0 && n_my > 1) printf("LN warp %d: per tile wait %lld work %lld\n", lw, tw_ / (n_my - 1), tb_ / (n_my - 1));
#endif
    f16_guard(amax);
  } else if (warp >= MP_WARP_OUT) {
    // ------------------------------------------------ output group -----------------------------------------------------------
    const int quad = warp & 3;
    const uint32_t st_base = stage_base + static_cast<uint32_t>(warp - MP_WARP_OUT) * 4096u;
#ifdef ML_TIMING
    long long tw_ = 0, tb_ = 0;
#endif
    for (uint32_t i = 0; i < n_my; ++i) {
      const long m0 = (static_cast<long>(blockIdx.x) + static_cast<long>(i) * gridDim.x) * ML_BM;
      const uint32_t qb = i % Q;
#ifdef ML_TIMING
      const long long tq0 = clock64();
#endif
      mbar_wait(acc2_full0 + 8 * qb, (i / Q) & 1);
      tc_fence_after();
#ifdef ML_TIMING
      const long long tq1 = clock64();
#endif
      mp_output_quadrant(p, b2_s, st_base, acc2_0 + qb * 128, quad, lane, m0);
#ifdef ML_TIMING
      if (i > 0) { tw_ += tq1 - tq0; tb_ += clock64() - tq1; }
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty0 + 8 * qb);
    }
#ifdef ML_TIMING
    if (blockIdx.x == 3 && lane == 0 && n_my > 1) printf("OUT warp %d: per tile wait %lld work %lld\n", warp - MP_WARP_OUT, tw_ / (n_my - 1), tb_ / (n_my - 1));
#endif
  } else {
    // ------------------------------------------------ GELU group ---------------------------------------------------------------
    const int quad = warp & 3, u0 = warp >> 2;          // units u0 and u0 + 2 of the chunk's four 32-column units
    float amax = 0.0f;
#ifdef ML_TIMING
    long long tw_ = 0, tb_ = 0;
#endif
    const uint32_t total = n_my * static_cast<uint32_t>(n);
    int j = 0;
    for (uint32_t c = 0; c < total; ++c) {
      const uint32_t slot = c & 1, cph = (c >> 1) & 1;
#ifdef ML_TIMING
      const long long tq0 = clock64();
#endif
      mbar_wait(acc1_full0 + 8 * slot, cph);
      mbar_wait(h_empty0 + 8 * slot, cph ^ 1);                   // fc2 of the chunk two back has read H[slot]
      tc_fence_after();
#ifdef ML_TIMING
      const long long tq1 = clock64();
#endif
      // both units' accumulator columns are requested before either is used: one tensor-memory round trip per chunk, not two
      const uint32_t acc1 = tmem_base + slot * MP_HC + (static_cast<uint32_t>(quad * 32) << 16);
      uint32_t va[32], vb[32];
      mp_tmem_ld32_nowait(acc1 + u0 * 32, va);
      mp_tmem_ld32_nowait(acc1 + (u0 + 2) * 32, vb);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const float *bj = b1_s + j * MP_HC;
      const uint32_t hb = h_base + slot * MP_H_BYTES;
      const float a0 = p.f16 ? mp_hidden_unit<__half>(va, bj + u0 * 32, hb, quad, u0 * 32, lane) : mp_hidden_unit<__nv_bfloat16>(va, bj + u0 * 32, hb, quad, u0 * 32, lane);
      const float a1 = p.f16 ? mp_hidden_unit<__half>(vb, bj + (u0 + 2) * 32, hb, quad, (u0 + 2) * 32, lane)
                             : mp_hidden_unit<__nv_bfloat16>(vb, bj + (u0 + 2) * 32, hb, quad, (u0 + 2) * 32, lane);
      amax = fmaxf(amax, fmaxf(a0, a1));
      tc_fence_before();
      ml_fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc1_empty0 + 8 * slot);
        mbar_arrive(h_full0 + 8 * slot);
      }
      if (++j == n) j = 0;
#ifdef ML_TIMING
      if (c >= (uint32_t)n) { tw_ += tq1 - tq0; tb_ += clock64() - tq1; }
#endif
    }
#ifdef ML_TIMING
    if (blockIdx.x == 3 && lane == 0 && n_my > 1 && (warp & 3) == 0) printf("GELU warp %d: per tile wait %lld work %lld\n", warp, tw_ / (n_my - 1), tb_ / (n_my - 1));
#endif
    f16_guard(amax);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MP_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

bool mlp_fused_supported(int C) { return C == 96 || C == 128 || C == 192 || C == 256; }

// shape policy: 0 = the pipelined shape (default), 1 = the serial one-CTA shape, 2 = serial, two CTAs per SM where it fits (C <= 128).
// Environment MUMPY_MLP_SHAPE.
// Measured (B = 32 stage-0 shapes, ncu: 19.4 of the 20 theoretical warps resident): the two-per-SM shape takes the same time, 256-262
// vs 255-261 us at M = 301056, C = 128 -- every phase of a CTA simply takes twice as long (prologue 7 -> 13 k clk, output pass
// 7.5 -> 17 k clk, -DML_TIMING): the SM's 16 epilogue warps are the limit in either arrangement, not the phase order.
static int g_mlp_shape = -1;
void set_mlp_fused_shape(int shape) { g_mlp_shape = shape < 0 ? 0 : (shape > 2 ? 2 : shape); }

template <int HC, int EW, int MIN_CTAS>
static int launch_mlp_fused(MlpParams &p, const void *W1, const void *W2, int C, cudaStream_t st) {
  p.n_chunks = 4 * C / HC;
  p.idesc1 = make_idesc_16_f32(ML_BM, HC, p.f16 != 0);
  p.idesc2 = make_idesc_16_f32(ML_BM, C, p.f16 != 0);
  const int w2_bytes = C * 128;                                             // one 64-column k-block of W2: C rows
  const int w1_bytes = (HC == 64 ? p.nkb : 1) * HC * 128;                   // HC = 64: every k-block of the W1 chunk in one stage
  p.stage_bytes = (uint32_t)(w2_bytes > w1_bytes ? w2_bytes : w1_bytes);
  p.stage_bytes = (p.stage_bytes + 1023u) & ~1023u;
  const int fixed = 1024 + p.nkb * ML_KB_BYTES + 2 * (HC / 64) * ML_KB_BYTES;
  const int budget = MIN_CTAS == 2 ? ML_SMEM_HALF : ML_SMEM_TOTAL;
  int stages = (budget - fixed) / (int)p.stage_bytes;
  if (stages > ML_MAX_STAGES) stages = ML_MAX_STAGES;
  MUMPY_REQUIRE(stages >= 2, "mlp_fused: shared memory budget (C=%d)", C);
  p.stages = stages;
  CUtensorMap tmW1, tmW2;
  int rc = tc_encode_2d_16(&tmW1, W1, p.f16 != 0, (uint64_t)C, (uint64_t)4 * C, (uint64_t)C, 64, HC);
  if (rc) return rc;
  rc = tc_encode_2d_16(&tmW2, W2, p.f16 != 0, (uint64_t)4 * C, (uint64_t)C, (uint64_t)4 * C, 64, (uint32_t)C);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_tc_kernel<HC, EW, MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget);
    // (without the carve-out preference the driver sizes the shared-memory partition for ONE CTA of the two-per-SM shape)
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fused_tc_kernel<HC, EW, MIN_CTAS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(mlp_fused_tc_kernel): %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    attr_set = true;
  }
  const long slots = (long)tc_num_sms() * MIN_CTAS;
  const int smem = fixed + p.stages * (int)p.stage_bytes;
  const unsigned grid = (unsigned)(p.num_tiles < slots ? p.num_tiles : slots);
  launch_kernel(mlp_fused_tc_kernel<HC, EW, MIN_CTAS>, grid, (EW + 2) * 32, smem, st, tmW1, tmW2, p);
  return launch_status("mlp_fused_tc_kernel");
}

static int launch_mlp_pipe(MlpParams &p, const void *W1, const void *W2, int C, cudaStream_t st) {
  p.n_chunks = 4 * C / MP_HC;
  p.idesc1 = make_idesc_16_f32(ML_BM, MP_HC, p.f16 != 0);
  p.idesc2 = make_idesc_16_f32(ML_BM, C, p.f16 != 0);
  p.a_bufs = p.acc2_bufs = C <= 128 ? 2 : 1;
  const int w2_bytes = C * 128;
  p.stage_bytes = (uint32_t)(w2_bytes > MP_HC * 128 ? w2_bytes : MP_HC * 128);
  p.stage_bytes = (p.stage_bytes + 1023u) & ~1023u;
  const int fixed = 1024 + p.a_bufs * p.nkb * ML_KB_BYTES + 2 * MP_H_BYTES + MP_STAGING_BYTES + ((7 * C * 4 + 1023) & ~1023);
  int stages = (ML_SMEM_TOTAL - fixed) / (int)p.stage_bytes;
  if (stages > ML_MAX_STAGES) stages = ML_MAX_STAGES;
  MUMPY_REQUIRE(stages >= 2, "mlp_fused(pipelined): shared memory budget (C=%d)", C);
  p.stages = stages;
  CUtensorMap tmW1, tmW2;
  int rc = tc_encode_2d_16(&tmW1, W1, p.f16 != 0, (uint64_t)C, (uint64_t)4 * C, (uint64_t)C, 64, MP_HC);
  if (rc) return rc;
  rc = tc_encode_2d_16(&tmW2, W2, p.f16 != 0, (uint64_t)4 * C, (uint64_t)C, (uint64_t)4 * C, 64, (uint32_t)C);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_pipe_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_SMEM_TOTAL);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(mlp_pipe_tc_kernel): %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    attr_set = true;
  }
  const long sms = tc_num_sms();
  const int smem = fixed + p.stages * (int)p.stage_bytes;
  const unsigned grid = (unsigned)(p.num_tiles < sms ? p.num_tiles : sms);
  launch_kernel(mlp_pipe_tc_kernel, grid, MP_THREADS, smem, st, tmW1, tmW2, p);
  return launch_status("mlp_pipe_tc_kernel");
}

int mlp_fused_16(const float *x, const float *gamma, const float *beta, float eps, const void *W1, const float *b1, const void *W2, const float *b2,
                 float *out, long M, int C, int w_dtype, cudaStream_t st) {
  int rc = resolve_driver_entry_points();
  if (rc) return rc;
  MUMPY_REQUIRE(mlp_fused_supported(C), "mlp_fused: unsupported width C=%d (96, 128, 192, 256)", C);
  MUMPY_REQUIRE(M > 0 && M < (1l << 31), "mlp_fused: 0 < M < 2^31 required");
  MUMPY_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(W1) |
                  reinterpret_cast<uintptr_t>(b1) | reinterpret_cast<uintptr_t>(W2) | reinterpret_cast<uintptr_t>(b2) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "mlp_fused: all buffers must be 16-byte aligned");
  if (g_mlp_shape < 0) {
    const char *v = getenv("MUMPY_MLP_SHAPE");
    g_mlp_shape = v ? atoi(v) : 0;
  }
  MlpParams p = {};
  p.x = x;
  p.gamma = gamma;
  p.beta = beta;
  p.b1 = b1;
  p.b2 = b2;
  p.out = out;
  p.M = M;
  p.num_tiles = cdiv(M, ML_BM);
  p.C = C;
  p.nkb = (C + 63) / 64;
  p.f16 = w_dtype == MUMPY_F16;
  p.eps = eps;
  if (g_mlp_shape == 0) return launch_mlp_pipe(p, W1, W2, C, st);
  if (C <= 128 && g_mlp_shape == 2) return launch_mlp_fused<64, 8, 2>(p, W1, W2, C, st);
  return launch_mlp_fused<128, 16, 1>(p, W1, W2, C, st);
}

}  // namespace mumpy
