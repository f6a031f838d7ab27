// LayerNorm fused into the GEMM that consumes it (16-bit modes):
//   out[m, n] = act( LN(x[m, :]) . W[n, :] + bias[n] )          x (M,K) fp32 residual stream, W (N,K) 16-bit, out 16-bit
// i.e. norm1 -> qkv (swinTransformer.py:266 + :142) and norm2 -> fc1 + GELU (:305 + :47-48) of every Swin block whose
// width fits (K = C in {96, 128, 192, 256, 384, 512}).  The normalised operand never exists in global memory:
//
//   * a work item is (128-row tile, group of `ng` N-tiles).  The 16 epilogue warps first normalise the item's 128 rows
//     (exact two-pass fp32 statistics, the arithmetic of layernorm_vec_kernel: same lanes-per-row split, same summation
//     order, so the operand is bit-identical to the unfused path) and write them as 16-bit K-major SWIZZLE_128B tiles
//     straight into shared memory -- the whole 128 x K A operand stays resident (K = 512: 128 KB) for all N-tiles of the item;
//   * warp 16 streams only the weight tiles (BN x 64) through a TMA / mbarrier ring, warp 17 issues tcgen05.mma
//     (M = 128, N = BN, K = 16) into double-buffered TMEM accumulators, the epilogue warps drain them (bias, GELU, 16-bit
//     pack through a 2 KB per-warp swizzled staging tile, 64-byte coalesced row segments).
//
// Versus layernorm_vec_kernel + gemm_tc_kernel this removes one kernel, the 16-bit write + re-reads of the normalised
// matrix, and the A half of the main loop's L2 -> SM traffic (the GEMM main loop is bound by the ~42.5 B/clk/SM L2 path).
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace mumpy {

int resolve_driver_entry_points();
int tc_encode_2d_16(CUtensorMap *map, const void *ptr, bool f16, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                    uint32_t box_inner, uint32_t box_outer);
int tc_num_sms();

constexpr int LG_BM = 128;
constexpr int LG_BK = 64;
constexpr int LG_MAX_STAGES = 8;
constexpr int LG_EPI_WARPS = 16;                   // (12 until round 2: 16 normalise 32 rows per round -- 4 rounds instead of 6 -- and drain the GELU tiles faster: fc1 57.3 -> 56.4 us, qkv 44.5 -> 43.9 us at M = 18816, K = 512; 96 registers)
constexpr int LG_EPI_GROUPS = LG_EPI_WARPS / 4;
constexpr int LG_THREADS = (LG_EPI_WARPS + 2) * 32;
constexpr int LG_STAGING = EPI16_STAGING;           // per-warp staging tile: 32 rows x 32 columns of 16-bit outputs
constexpr int LG_BIAS = 128 * 4;                    // per-warp bias slice (up to 4 chunks of 32 columns)
constexpr int LG_A_KB_BYTES = LG_BM * 128;          // one 64-column k-block of the resident A operand
constexpr int LG_SMEM_TOTAL = 226 * 1024;       // dynamic part; the barriers / TMEM slot are static shared memory

struct LgParams {
  const float *x;
  const float *gamma;
  const float *beta;
  const float *bias;
  uint16_t *out;
  long M;
  long ldo;
  long num_items;
  int N, K;
  int BN;
  int stages;
  int act;
  int f16;
  int ng;             // N-tiles per work item
  int n_groups;       // work items per 128-row tile (ng * n_groups = N / BN)
  int nkb;            // 64-column k-blocks (the last one may be partial: K = 96)
  uint32_t acc_cols;
  uint32_t idesc;
  float eps;
  int debug;          // development: bit 0 skips the LayerNorm prologue, bit 1 the epilogue, bit 2 the MMAs, bit 3 the TMA loads (timing attribution only)
};

// N-tile visited at step `nt` of work item `item`: every 128-row tile starts its sweep over the group's N-tiles at a different
// offset.  Without it all ~148 CTAs request the SAME weight tile from L2 at the same time (one wave, identical schedules) and
// the few L2 slices holding it serialise them: measured 1 us per k-block instead of 0.2.
__device__ __forceinline__ int lg_ntile(const LgParams &p, long item, int nt) {
  const int rot = static_cast<int>((item / p.n_groups) % p.ng);
  const int g = static_cast<int>(item % p.n_groups);
  int t = nt + rot;
  if (t >= p.ng) t -= p.ng;
  return g * p.ng + t;
}

__device__ __forceinline__ void lg_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Normalises rows [m0, m0 + 128) into the resident A operand.  LPR lanes share a row and hold NV float4 each (K = 16 LPR NV),
// 32/LPR rows side by side in a warp, RI such groups in flight -- layernorm_vec_kernel's scheme (norm.cu).
template <typename OutT, int LPR, int NV, int RI>
__device__ __forceinline__ float lg_normalise_rows(const LgParams &p, uint32_t a_base, long m0, int warp, int lane) {
  constexpr int G = 32 / LPR;
  constexpr int RPW = G * RI;
  static_assert(LG_BM % RPW == 0, "row groups must tile the 128-row block");
  const int sub = lane % LPR, grp = lane / LPR;
  const int C = p.K;
  float amax = 0.0f;
  for (int r0 = warp * RPW; r0 < LG_BM; r0 += LG_EPI_WARPS * RPW) {
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
      const int i = sub + LPR * u;                  // float4 index in the row: columns 4i .. 4i+3
      const float4 g4 = __ldg(reinterpret_cast<const float4 *>(p.gamma) + i), b4 = __ldg(reinterpret_cast<const float4 *>(p.beta) + i);
      // k-block i/16, 16-byte chunk (i%16)/2 of the 128-byte row, low / high half of the chunk
      const uint32_t col_off = static_cast<uint32_t>(i >> 4) * LG_A_KB_BYTES + static_cast<uint32_t>(i & 1) * 8u;
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

template <typename OutT>
__device__ __forceinline__ float lg_normalise(const LgParams &p, uint32_t a_base, long m0, int warp, int lane) {
  switch (p.K) {
    case 96: return lg_normalise_rows<OutT, 8, 3, 2>(p, a_base, m0, warp, lane);
    case 128: return lg_normalise_rows<OutT, 8, 4, 2>(p, a_base, m0, warp, lane);
    case 192: return lg_normalise_rows<OutT, 16, 3, 2>(p, a_base, m0, warp, lane);
    case 256: return lg_normalise_rows<OutT, 16, 4, 2>(p, a_base, m0, warp, lane);
    case 384: return lg_normalise_rows<OutT, 32, 3, 2>(p, a_base, m0, warp, lane);
    default: return lg_normalise_rows<OutT, 32, 4, 2>(p, a_base, m0, warp, lane);
  }
}

// kPair: the two CTAs of a (2,1,1) cluster work on two adjacent 128-row tiles with tcgen05.mma.cta_group::2 (M = 256): each CTA
// normalises and keeps its own 128 rows, loads HALF of every weight tile (BN/2 rows) and the tensor cores of both SMs read both
// halves, so a k-block costs a CTA 16 KB of ring space and L2 traffic instead of 32 KB at BN = 256 -- a 4-stage ring next to the
// 128 KB operand at K = 512 (the 1-CTA kernel has room for two stages and is latency bound: 0.64 us per k-block against 0.27 us
// here, tools/umma_bench3.cu).  Protocol as in gemm_tc_kernel<kPair>: `full`, `acc_empty` and `a_full` live on the leader (rank 0)
// and collect arrivals from both CTAs, the leader issues the MMAs, its commits arrive on `empty` / `acc_full` in both CTAs.
template <bool kPair>
__global__ void __launch_bounds__(LG_THREADS, 1) ln_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmB, const LgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * LG_MAX_STAGES + 5];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t slot_id = kPair ? blockIdx.x >> 1 : blockIdx.x;          // scheduler slot (a CTA or a CTA pair)
  const uint32_t n_slots = kPair ? gridDim.x >> 1 : gridDim.x;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t a_base = (raw + 1023u) & ~1023u;                         // resident A: nkb k-blocks of 128 rows x 128 B
  const uint32_t b_base = a_base + static_cast<uint32_t>(p.nkb) * LG_A_KB_BYTES;
  const uint32_t b_rows = static_cast<uint32_t>(kPair ? p.BN / 2 : p.BN);  // weight rows this CTA loads per k-block
  const uint32_t b_bytes = b_rows * 128u;
  const uint32_t epi_base = b_base + static_cast<uint32_t>(p.stages) * b_bytes;
  const uint32_t full0 = smem_u32(&bars[0]);
  const uint32_t empty0 = smem_u32(&bars[LG_MAX_STAGES]);
  const uint32_t acc_full0 = smem_u32(&bars[2 * LG_MAX_STAGES]);
  const uint32_t acc_empty0 = smem_u32(&bars[2 * LG_MAX_STAGES + 2]);
  const uint32_t a_full = smem_u32(&bars[2 * LG_MAX_STAGES + 4]);

  if (warp == LG_EPI_WARPS && lane == 0) {
    prefetch_tensormap(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, kPair ? 2 : 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full0 + 8 * s, 1);
      mbar_init(acc_empty0 + 8 * s, kPair ? 2 * LG_EPI_WARPS : LG_EPI_WARPS);
    }
    mbar_init(a_full, kPair ? 2 * LG_EPI_WARPS : LG_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == LG_EPI_WARPS + 1) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(2 * p.acc_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(2 * p.acc_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_grid_sync();

  if (warp == LG_EPI_WARPS) {
    // ------------------------------------------------ TMA producer (weights only) ------------------------------------
    // (whole warp converged, one elected lane issues: elect_one(), tc_common.cuh)
    uint32_t s = 0, ph = 0;
    const uint32_t full_leader = kPair ? mapa_shared(full0, 0) : full0;
    for (long item = slot_id; item < p.num_items; item += n_slots) {
      for (int nt = 0; nt < p.ng; ++nt) {
        const int n0 = lg_ntile(p, item, nt) * p.BN + static_cast<int>(rank * b_rows);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          if (elect_one()) {
            if (kPair) {
              mbar_arrive_expect_tx_cluster(full_leader + 8 * s, b_bytes);
              tma_load_2d_pair(b_base + s * b_bytes, &tmB, full_leader + 8 * s, kb * LG_BK, n0);
            } else if (p.debug & 8) {
              mbar_arrive(full0 + 8 * s);
            } else {
              mbar_arrive_expect_tx(full0 + 8 * s, b_bytes);
              tma_load_2d(b_base + s * b_bytes, &tmB, full0 + 8 * s, kb * LG_BK, n0);
            }
          }
          __syncwarp();
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == LG_EPI_WARPS + 1) {
    // ------------------------------------------------ MMA issuer (the leader of a pair) ------------------------------
    if (rank == 0) {
      uint32_t s = 0, ph = 0, t = 0, wi = 0;
#ifdef LG_TIMING
      const long long lt0 = clock64();
      long long lt[8][3];
      long long lt_a = 0;
#endif
      for (long item = slot_id; item < p.num_items; item += n_slots, ++wi) {
        mbar_wait(a_full, wi & 1);                          // the item's normalised rows are in shared memory (of both CTAs)
        tc_fence_after();
#ifdef LG_TIMING
        if (wi == 0) lt_a = clock64();
#endif
        for (int nt = 0; nt < p.ng; ++nt, ++t) {
          const uint32_t slot = t & 1, aph = (t >> 1) & 1;
#ifdef LG_TIMING
          if (t < 8) lt[t][0] = clock64();
#endif
          mbar_wait(acc_empty0 + 8 * slot, aph ^ 1);
          tc_fence_after();
#ifdef LG_TIMING
          if (t < 8) lt[t][1] = clock64();
#endif
          const uint32_t d_tmem = tmem_base + slot * p.acc_cols;
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
            const uint64_t adesc = make_kmajor_sw128_desc(a_base + kb * LG_A_KB_BYTES);
            const uint64_t bdesc = make_kmajor_sw128_desc(b_base + s * b_bytes);
            const bool tail = (kb + 1) * LG_BK > p.K;           // K = 96: the second k-block holds 32 columns
            if (elect_one()) {
              if (kPair) {
                umma_bf16_pair(d_tmem, adesc, bdesc, p.idesc, kb > 0 ? 1u : 0u);
                umma_bf16_pair(d_tmem, adesc + 2, bdesc + 2, p.idesc, 1u);
                if (!tail) {
                  umma_bf16_pair(d_tmem, adesc + 4, bdesc + 4, p.idesc, 1u);
                  umma_bf16_pair(d_tmem, adesc + 6, bdesc + 6, p.idesc, 1u);
                }
                umma_commit_pair(empty0 + 8 * s);
              } else if (p.debug & 4) {
                mbar_arrive(empty0 + 8 * s);
              } else {
                umma_bf16(d_tmem, adesc, bdesc, p.idesc, kb > 0 ? 1u : 0u);
                umma_bf16(d_tmem, adesc + 2, bdesc + 2, p.idesc, 1u);
                if (!tail) {
                  umma_bf16(d_tmem, adesc + 4, bdesc + 4, p.idesc, 1u);
                  umma_bf16(d_tmem, adesc + 6, bdesc + 6, p.idesc, 1u);
                }
                umma_commit(empty0 + 8 * s);
              }
            }
            __syncwarp();
            if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
          }
          if (elect_one()) {
            if (kPair) umma_commit_pair(acc_full0 + 8 * slot); else umma_commit(acc_full0 + 8 * slot);
          }
          __syncwarp();
#ifdef LG_TIMING
          if (t < 8) lt[t][2] = clock64();
#endif
        }
      }
#ifdef LG_TIMING
      if ((slot_id == 0 || slot_id == 40) && lane == 0) {
        printf("slot %u N=%d K=%d BN=%d stages=%d pair=%d: a_full at %lld;", slot_id, p.N, p.K, p.BN, p.stages, (int)kPair, lt_a - lt0);
        for (uint32_t i = 0; i < t && i < 8; ++i) printf(" [wait_acc %lld loop %lld]", lt[i][1] - lt[i][0], lt[i][2] - lt[i][1]);
        printf(" end %lld\n", clock64() - lt0);
      }
#endif
    }
  } else {
    // ------------------------------------------------ LayerNorm prologue + epilogue ----------------------------------
    const uint32_t st_base = epi_base + warp * LG_STAGING;
    float *bias_s = reinterpret_cast<float *>(smem_raw + (epi_base - raw) + LG_EPI_WARPS * LG_STAGING) + warp * (LG_BIAS / 4);
    const uint32_t a_full_leader = kPair ? mapa_shared(a_full, 0) : a_full;
    const uint32_t acc_empty_leader = kPair ? mapa_shared(acc_empty0, 0) : acc_empty0;
    uint32_t t = 0;
    float amax = 0.0f;
    for (long item = slot_id; item < p.num_items; item += n_slots) {
      const long m0 = ((item / p.n_groups) * (kPair ? 2 : 1) + rank) * LG_BM;
      // every MMA that read the previous item's rows has completed: this warp waited for the accumulator of that item's last
      // N-tile, which tcgen05.commit publishes (in both CTAs of a pair) after all earlier MMAs
      if (!(p.debug & 1)) amax = fmaxf(amax, p.f16 ? lg_normalise<__half>(p, a_base, m0, warp, lane) : lg_normalise<__nv_bfloat16>(p, a_base, m0, warp, lane));
      lg_fence_proxy_async();                                // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(a_full_leader); else mbar_arrive(a_full);
      }
      for (int nt = 0; nt < p.ng; ++nt, ++t) {
        const uint32_t slot = t & 1, aph = (t >> 1) & 1;
        const int n0 = lg_ntile(p, item, nt) * p.BN;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = (warp >> 2) * 32 + 32 * LG_EPI_GROUPS * ci + lane;
          bias_s[ci * 32 + lane] = (p.bias && c < p.BN && n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.0f;
        }
        __syncwarp();
        mbar_wait(acc_full0 + 8 * slot, aph);
        tc_fence_after();
        const uint32_t acc = tmem_base + slot * p.acc_cols;
        float a = 0.0f;
        if (p.debug & 2) {
        } else if (p.f16) {
          a = p.act == MUMPY_ACT_GELU ? epilogue16_tile<__half, 1>(p.out, p.ldo, p.M, p.N, p.BN, p.act, LG_EPI_GROUPS, st_base, bias_s, acc, warp, lane, m0, n0)
                                      : epilogue16_tile<__half, 0>(p.out, p.ldo, p.M, p.N, p.BN, p.act, LG_EPI_GROUPS, st_base, bias_s, acc, warp, lane, m0, n0);
        } else {
          a = p.act == MUMPY_ACT_GELU ? epilogue16_tile<__nv_bfloat16, 1>(p.out, p.ldo, p.M, p.N, p.BN, p.act, LG_EPI_GROUPS, st_base, bias_s, acc, warp, lane, m0, n0)
                                      : epilogue16_tile<__nv_bfloat16, 0>(p.out, p.ldo, p.M, p.N, p.BN, p.act, LG_EPI_GROUPS, st_base, bias_s, acc, warp, lane, m0, n0);
        }
        amax = fmaxf(amax, a);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_cluster(acc_empty_leader + 8 * slot); else mbar_arrive(acc_empty0 + 8 * slot);
        }
      }
    }
    f16_guard(amax);
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();      // nobody leaves while the partner can still touch its barriers / tiles
  if (warp == LG_EPI_WARPS + 1) {
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * p.acc_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * p.acc_cols) : "memory");
  }
}

// CTA-pair policy of mumpy_ln_linear: 0 never, 1 the cost model decides (default), 2 always.  Environment MUMPY_LG_PAIR or
// mumpy_set_ln_linear_pair_mode().
static int g_lg_pair = -1;
void set_ln_linear_pair_mode(int mode) { g_lg_pair = mode < 0 ? 1 : (mode > 2 ? 2 : mode); }

static int lg_env(const char *name) {
  const char *v = getenv(name);
  return v ? atoi(v) : 0;
}

bool ln_linear_supported(int N, int K) {
  return (K == 96 || K == 128 || K == 192 || K == 256 || K == 384 || K == 512) && (N % 64 == 0 || N % 96 == 0);
}

// Tile width and N-grouping: minimise  waves x (prologue + ng x max(main loop, epilogue)) + one epilogue  over the divisors of N.
// Development overrides: MUMPY_LG_BN, MUMPY_LG_GROUPS, MUMPY_LG_STAGES, MUMPY_LG_PAIR.
int ln_linear_16(const float *x, const float *gamma, const float *beta, float eps, const void *W, const float *bias, void *out, long ldo,
                 long M, int N, int K, int w_dtype, int act, cudaStream_t st) {
  int rc = resolve_driver_entry_points();
  if (rc) return rc;
  MUMPY_REQUIRE(ln_linear_supported(N, K), "ln_linear: unsupported shape N=%d K=%d (K in {96,128,192,256,384,512}, N %% 64 == 0 or N %% 96 == 0)", N, K);
  MUMPY_REQUIRE(act == MUMPY_ACT_NONE || act == MUMPY_ACT_GELU, "ln_linear: activation must be none or GELU");
  MUMPY_REQUIRE(ldo % 8 == 0 && M > 0 && M < (1l << 31), "ln_linear: ldo %% 8 == 0 and 0 < M < 2^31 required");
  MUMPY_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15) == 0,
                "ln_linear: x, gamma, beta, W, out must be 16-byte aligned");
  static int dbg_bn = -1, dbg_groups = -1, dbg_stages = -1, dbg_mode = 0;
  if (dbg_bn < 0) {
    const char *pv = getenv("MUMPY_LG_PAIR");
    if (pv && g_lg_pair < 0) g_lg_pair = atoi(pv);
    dbg_mode = lg_env("MUMPY_LG_DEBUG");
    dbg_bn = lg_env("MUMPY_LG_BN");
    dbg_groups = lg_env("MUMPY_LG_GROUPS");
    dbg_stages = lg_env("MUMPY_LG_STAGES");
  }
  const int nkb = (K + LG_BK - 1) / LG_BK;
  const long m_tiles = cdiv(M, LG_BM);
  const int sms = tc_num_sms();
  const int fixed = 1024 + nkb * LG_A_KB_BYTES + LG_EPI_WARPS * (LG_STAGING + LG_BIAS);
  static const int cands[] = {256, 192, 128, 96, 64};
  int best_bn = 0, best_groups = 1, best_stages = 0, best_pair = 0;
  double best_cost = 1e30;
  const bool gelu = act == MUMPY_ACT_GELU;
  // Cost in clocks, from tools/umma_bench2/3.cu and tools/ln_gemm_bench.py: a k-block takes max(tensor / issue floor, ring latency /
  // stages) with the floor max(2 BN, 420) clk and ~1300 clk of latency per stage; an epilogue pass over 32 columns costs a warp
  // ~1500 clk (+700 with GELU); the LayerNorm prologue ~25 clk per column.  minimise  waves x (prologue + ng x max(tile, epilogue)).
  for (int pair = 0; pair < 2; ++pair) {
    const int pair_mode = g_lg_pair < 0 ? 1 : g_lg_pair;
    if (pair && (pair_mode == 0 || sms < 2)) continue;
    if (!pair && pair_mode == 2) continue;
    const long m_units = pair ? cdiv(m_tiles, 2) : m_tiles;
    const long slots = pair ? sms / 2 : sms;
    for (int bn : cands) {
      if (N % bn) continue;
      if (pair && bn % 32) continue;
      if (dbg_bn > 0 && N % dbg_bn == 0 && bn != dbg_bn) continue;
      const int stage_bytes = (pair ? bn / 2 : bn) * 128;
      int stages = (LG_SMEM_TOTAL - fixed) / stage_bytes;
      if (stages > LG_MAX_STAGES) stages = LG_MAX_STAGES;
      if (stages < 2) continue;
      const int tiles_n = N / bn;
      const double per_kb = fmax(fmax(bn * 2.0, 420.0), 1300.0 / stages + 90.0);
      const double tile = nkb * per_kb;
      const double epi = ((bn + 32 * LG_EPI_GROUPS - 1) / (32 * LG_EPI_GROUPS)) * (gelu ? 2200.0 : 1500.0);
      const double prologue = K * 25.0;
      for (int groups = 1; groups <= tiles_n; ++groups) {
        if (tiles_n % groups) continue;
        if (dbg_groups > 0 && tiles_n % dbg_groups == 0 && groups != dbg_groups) continue;
        const int ng = tiles_n / groups;
        const double waves = (double)cdiv(m_units * groups, slots);
        const double cost = waves * (prologue + ng * fmax(tile, epi)) + fmin(tile, epi);
        if (cost < best_cost * 0.99) {
          best_cost = cost;
          best_bn = bn;
          best_groups = groups;
          best_stages = stages;
          best_pair = pair;
        }
      }
    }
  }
  MUMPY_REQUIRE(best_bn > 0, "ln_linear: no tile fits (N=%d K=%d)", N, K);
  if (dbg_stages > 0 && dbg_stages < best_stages) best_stages = dbg_stages;
  LgParams p = {};
  p.x = x;
  p.gamma = gamma;
  p.beta = beta;
  p.bias = bias;
  p.out = static_cast<uint16_t *>(out);
  p.M = M;
  p.ldo = ldo;
  p.N = N;
  p.K = K;
  p.BN = best_bn;
  p.act = act;
  p.f16 = w_dtype == MUMPY_F16;
  p.n_groups = best_groups;
  p.ng = N / best_bn / best_groups;
  p.num_items = (best_pair ? cdiv(m_tiles, 2) : m_tiles) * best_groups;
  p.nkb = nkb;
  p.eps = eps;
  p.debug = dbg_mode;
  uint32_t cols = 32;
  while (cols < (uint32_t)p.BN) cols <<= 1;
  p.acc_cols = cols;
  p.idesc = make_idesc_16_f32(best_pair ? 2 * LG_BM : LG_BM, p.BN, p.f16 != 0);
  const long slots = best_pair ? sms / 2 : sms;
  const long kb_per_cta = (long)nkb * p.ng * cdiv(p.num_items, slots);
  if (best_stages > kb_per_cta) best_stages = (int)kb_per_cta;
  p.stages = best_stages;
  const int b_rows = best_pair ? p.BN / 2 : p.BN;
  CUtensorMap tmB;
  rc = tc_encode_2d_16(&tmB, W, p.f16 != 0, (uint64_t)K, (uint64_t)N, (uint64_t)K, LG_BK, (uint32_t)b_rows);
  if (rc) return rc;
  static bool attr_set[2] = {false, false};
  if (!attr_set[best_pair]) {
    cudaError_t e = best_pair ? cudaFuncSetAttribute(ln_gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LG_SMEM_TOTAL)
                              : cudaFuncSetAttribute(ln_gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LG_SMEM_TOTAL);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(ln_gemm_tc_kernel): %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    attr_set[best_pair] = true;
  }
  const int smem = fixed + p.stages * b_rows * 128;
  const unsigned units = (unsigned)(p.num_items < slots ? p.num_items : slots);
  if (best_pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * units);
    cfg.blockDim = dim3(LG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, ln_gemm_tc_kernel<true>, tmB, p);
  } else {
    launch_kernel(ln_gemm_tc_kernel<false>, units, LG_THREADS, smem, st, tmB, p);
  }
  return launch_status("ln_gemm_tc_kernel");
}

}  // namespace mumpy
