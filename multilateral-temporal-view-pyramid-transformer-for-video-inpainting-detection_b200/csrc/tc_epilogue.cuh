// Lean epilogue for 16-bit outputs (no residual) of the tcgen05 GEMMs: one warp's share of a 128 x BN fp32 accumulator tile.
//
// The ncu profile of the previous epilogue on an epilogue-bound shape (fc1 + GELU, K = 256: the MMA warp waits ~70 % of the
// time for an accumulator slot) showed it issue-bound: 17 warp instructions per output element, of which only ~10 are the math
// (bias, GELU, range guard, pack).  This version drops the rest: packed 16-bit staging (half the shared-memory instructions of
// an fp32 tile), one row pointer per lane advanced by adds (no 64-bit multiply / compare per store), and an unpredicated path
// for interior tiles.
#pragma once
#include "tc_common.cuh"

namespace mumpy {

constexpr int EPI16_STAGING = 32 * 64;      // per-warp staging tile: 32 rows x 32 columns of 16-bit outputs

// ACT: 0 none, 1 GELU (gelu_fast2), 2 anything else through apply_act(act_code).
// Lane quadrant warp&3 of the accumulator, every `groups`-th 32-column chunk starting at chunk warp>>2.
// Phase 1 (thread <-> accumulator row): tcgen05.ld, + bias (staged in shared memory by the caller, zero when absent),
// activation, pack, into the warp's staging tile (rows of 64 B, 16-byte chunks XOR-swizzled by (row>>1)&3: conflict-free both
// ways).  Phase 2: four lanes per row, every store instruction writes eight full 64-byte row segments.
// Returns the largest |value| converted to IEEE half (0 for bf16): the caller feeds it to f16_guard().
template <typename OutT, int ACT>
__device__ __forceinline__ float epilogue16_tile(uint16_t *__restrict__ out, long ldo, long M, int N, int BN, int act_code, int groups, uint32_t st_base,
                                                 const float *bias_s, uint32_t acc, int warp, int lane, long m0, int n0) {
  const int quad = warp & 3, grp = warp >> 2;
  const uint32_t lane_addr = acc + (static_cast<uint32_t>(quad * 32) << 16);
  const long row = m0 + quad * 32 + (lane >> 2);                 // this lane's first output row in phase 2 (then +8 per pass)
  const int c8 = lane & 3;
  uint16_t *optr = out + row * ldo + n0 + c8 * 8;
  const long rows_left = M - row;
  const bool interior = m0 + 128 <= M && n0 + BN <= N && (BN & 31) == 0;      // (BN = 16 / 48: the last chunk is partial)
  const uint32_t st_w = st_base + lane * 64, sw_w = (lane >> 1) & 3;
  const uint32_t st_r = st_base + (lane >> 2) * 64, sw_r = (lane >> 3) & 3;      // rows i*8 + (lane>>2): ((row>>1)&3) == (lane>>3)&3
  float amax = 0.0f;
  for (int c0 = grp * 32, ci = 0; c0 < BN; c0 += 32 * groups, ++ci) {
    uint32_t v[32];
    tmem_ld32(lane_addr + c0, v);
    const float4 *bias4 = reinterpret_cast<const float4 *>(bias_s + ci * 32);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float2 f[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 b = bias4[2 * g + h];
        f[2 * h] = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 4 * h]), __uint_as_float(v[8 * g + 4 * h + 1])), make_float2(b.x, b.y));
        f[2 * h + 1] = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 4 * h + 2]), __uint_as_float(v[8 * g + 4 * h + 3])), make_float2(b.z, b.w));
      }
      uint32_t w[4];
      if (ACT == 1) gelu_fast2_x4(f);          // four interleaved dependency chains (tc_common.cuh)
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        if (ACT == 2) {
          f[h].x = apply_act(f[h].x, act_code);
          f[h].y = apply_act(f[h].y, act_code);
        }
        if constexpr (is_half_t<OutT>::value) amax = fmaxf(amax, fmaxf(fabsf(f[h].x), fabsf(f[h].y)));
        w[h] = pack2<OutT>(f[h].x, f[h].y);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_w + ((g ^ sw_w) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    }
    __syncwarp();
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u[i].x), "=r"(u[i].y), "=r"(u[i].z), "=r"(u[i].w) : "r"(st_r + i * 512 + ((c8 ^ sw_r) << 4)));
    uint16_t *o = optr + c0;
    if (interior) {
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4 *>(o + i * 8 * ldo) = u[i];
    } else if (c0 + c8 * 8 < BN && n0 + c0 + c8 * 8 < N) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i * 8 < rows_left) *reinterpret_cast<uint4 *>(o + i * 8 * ldo) = u[i];
    }
    __syncwarp();
  }
  return amax;
}

}  // namespace mumpy
