// Window attention on the tensor pipe for the bf16 mode: one warp per (window, head), q/k/v staged in swizzled shared
// memory with 16-byte gathers straight from the canvas-ordered qkv matrix (window partition / cyclic shift folded into
// the row index), S = QK^T and O = PV as mma.sync m16n8k16 bf16 with fp32 accumulators, probabilities kept in
// registers (S accumulator fragments are re-used as the A fragments of PV), fp32 softmax with the relative-position
// bias and shift mask added in fp32.  49 (or 64) tokens x d=32 per head: ~1.7% of the forward's FLOPs, so the kernel is
// HBM/L2-bound (reads 3C, writes C bf16 per token) -- the tiles are far too small for a 128-row tcgen05 atom.
// The same core serves the deformable cross-view attention (q from the query view's fp32 canvas, k/v from the sampled
// windows, outputs summed over the temporal ratio).
#include <type_traits>

#include "common.cuh"

namespace mumpy {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// T = __nv_bfloat16 or __half: the 16-bit operand type of q/k/v (and of the probabilities fed to P.V)
template <typename T>
__device__ __forceinline__ void mma_16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}

// [64 tokens][32 dims] bf16 tile, 64 B rows, 16-byte chunks XOR-swizzled by (row>>1)&3 (conflict-free ldmatrix)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

constexpr int ATT_TILE_BYTES = 64 * 64;     // one 64x32 bf16 tile
constexpr int ATT_WARPS = 4;

// Additive score terms (relative-position bias + shift mask), returned pre-divided by the qk scale so that they can be
// used as the initial value of the score accumulators: scale * (q.k + init) = scale * q.k + bias + mask.
struct NoBias {
  __device__ __forceinline__ float operator()(int, int, int) const { return 0.0f; }
};
// gathered bias (nH,N,N) and optional mask tensor (nW,N,N): clamped, branch-free loads issued before the MMAs
template <int N>
struct LoadedBias {
  const float *bias_h, *mask_w;
  int i_lo, i_hi, t;
  float inv_scale;
  __device__ __forceinline__ float operator()(int hi, int nt, int e) const {
    const int j = min(nt * 8 + 2 * t + (e & 1), N - 1);
    const int idx = (hi ? i_hi : i_lo) * N + j;
    float v = __ldg(bias_h + idx);
    if (mask_w) v += __ldg(mask_w + idx);
    return v * inv_scale;
  }
};
// bias from the (2ws-1)^2-entry relative-position table of this head (shared memory, pre-divided by the scale):
// index a(i) - a(j) + OFF with a(p) = (p/ws)(2ws-1) + p%ws; standard Swin shift mask from region ids (no loads at all).
template <int WS>
struct TableBias {
  const float *tbl;          // smem, pre-divided by scale
  const int *aj;             // per-lane packed (a(j) | region(j) << 16) for its 16 key columns [nt*2 + e]
  int a_lo, a_hi, reg_lo, reg_hi;
  float mask_val;            // -100 / scale, or 0 when the block is not shifted
  __device__ __forceinline__ float operator()(int hi, int nt, int e) const {
    const int pj = aj[nt * 2 + (e & 1)];
    const int a_i = hi ? a_hi : a_lo, r_i = hi ? reg_hi : reg_lo;
    float v = tbl[a_i - (pj & 0xffff)];
    if ((pj >> 16) != r_i) v += mask_val;
    return v;
  }
};

// scores for 16 query rows (m-tile mt) against the key n-tiles, then softmax -> un-normalised probabilities in s (base-2
// exponentials of scale*log2e*(q.k + init) minus the row maximum); the row sums of rows g / g+8 come back in sum_lo / sum_hi.
template <typename T, int N, typename Init>
__device__ __forceinline__ void scores_softmax(uint32_t sQ, uint32_t sK, int mt, int lane, float scale, const Init &init, float (&s)[8][4],
                                               float &sum_lo, float &sum_hi) {
  constexpr int NT = (N + 7) / 8;
  const int t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[nt][e] = (nt < NT) ? init(e >> 1, nt, e) : 0.0f;
  uint32_t a[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    ldsm_x4(sQ + tile_off(row, ks * 2 + (lane >> 4)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    uint32_t b0, b1, b2, b3;
    ldsm_x4(sK + tile_off(nt * 8 + (lane & 7), lane >> 3), b0, b1, b2, b3);
    mma_16<T>(s[nt], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
    mma_16<T>(s[nt], a[1][0], a[1][1], a[1][2], a[1][3], b2, b3);
  }
  const float c = scale * 1.4426950408889634f;     // exp(x) = 2^(x log2 e)
  float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (nt * 8 + 7 >= N) {                        // only the last n-tile can hold padded key columns
        const int j = nt * 8 + 2 * t + (e & 1);
        if (j >= N) s[nt][e] = -INFINITY;
      }
      if (e < 2) m_lo = fmaxf(m_lo, s[nt][e]); else m_hi = fmaxf(m_hi, s[nt][e]);
    }
  }
  m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
  m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
  m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
  m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
  const float mc_lo = m_lo * c, mc_hi = m_hi * c;
  sum_lo = 0.0f;
  sum_hi = 0.0f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float p;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(s[nt][e], c, -((e < 2) ? mc_lo : mc_hi))));   // 2^-inf = 0 for padding
      s[nt][e] = p;
      if (e < 2) sum_lo += p; else sum_hi += p;
    }
  }
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
}

// o[dn][4] += P(16 x 64) . V(64 x 32) for one m-tile; un-normalised probabilities come straight from the score fragments
template <typename T>
__device__ __forceinline__ void pv_accumulate(uint32_t sV, int lane, const float (&s)[8][4], float (&o)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t a0 = pack2<T>(s[2 * kk][0], s[2 * kk][1]);
    const uint32_t a1 = pack2<T>(s[2 * kk][2], s[2 * kk][3]);
    const uint32_t a2 = pack2<T>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    const uint32_t a3 = pack2<T>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    const int tok = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int dp = 0; dp < 2; ++dp) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(sV + tile_off(tok, dp * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma_16<T>(o[dp * 2], a0, a1, a2, a3, b0, b1);
      mma_16<T>(o[dp * 2 + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

// canvas row of token p of window n for a compile-time window size (divisions by constants)
template <int WS>
__device__ __forceinline__ int token_row(int wr, int wc, int p, int TH, int W, int shift) {
  int r = wr * WS + p / WS + shift;
  int c = wc * WS + p % WS + shift;
  if (r >= TH) r -= TH;
  if (c >= W) c -= W;
  return r * W + c;
}

constexpr int ATT_TABLE_FLOATS = 232;       // (2*8-1)^2 = 225 entries for ws 8, 169 for ws 7
constexpr int ATT_WARP_SMEM = 3 * ATT_TILE_BYTES + 64 * sizeof(long) + ATT_TABLE_FLOATS * sizeof(float);   // cva kernel

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Persistent window attention: 8 warps per CTA, one CTA per SM; every warp walks its own strided list of (window, head)
// tasks with a two-stage cp.async ring, so the q/k/v gather of task i+1 (49 x 64 B row segments each, straight from
// the canvas-ordered qkv matrix) is in flight while task i runs on the tensor pipe.  K and V fragments are loaded
// once per task into registers and reused by the four 16-row query tiles.
constexpr int WATT_WARPS = 16;
constexpr int WATT_STAGES = 1;           // q/k/v tile sets per warp (2 = the warp prefetches its next task while computing)
constexpr bool WATT_KV_REGS = false;     // keep the K / V fragments of a task in registers across its four query tiles
constexpr int WATT_STAGE_BYTES = 3 * ATT_TILE_BYTES;
constexpr int WATT_WARP_BYTES = WATT_STAGES * WATT_STAGE_BYTES + WATT_STAGES * 64 * (int)sizeof(int);

// MODE 0: gathered bias (nH,N,N) + optional mask tensor (nW,N,N) through global loads.
// MODE 1: bias from the raw relative_position_bias_table (T,nH), staged for all heads in shared memory; the mask is the
//         standard Swin shift mask recomputed from region ids on the stacked canvas (swinTransformer.py:233-252) and is
//         only evaluated for the windows of the last window row / column (the only ones that hold masked pairs).
template <typename T, int WS, int MODE>
__global__ void __launch_bounds__(WATT_WARPS * 32, 1) window_attention_mma_kernel(const T *__restrict__ qkv, const float *__restrict__ bias,
                                                                                  const float *__restrict__ mask, const float *__restrict__ rel_table,
                                                                                  T *__restrict__ out, int TH, int W, int C, int heads,
                                                                                  int shift, int mshift, long n_tasks) {
  pdl_grid_sync();
  constexpr int N = WS * WS;
  constexpr int NT = (N + 7) / 8;
  constexpr int TBL = (2 * WS - 1) * (2 * WS - 1);
  constexpr float kScale = 0.17677669529663687f;    // 32^-0.5
  constexpr float kC = kScale * 1.4426950408889634f;
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *wbase = att_smem + warp * WATT_WARP_BYTES;
  int *rows_base = reinterpret_cast<int *>(wbase + WATT_STAGES * WATT_STAGE_BYTES);
  float *tbl_all = reinterpret_cast<float *>(att_smem + WATT_WARPS * WATT_WARP_BYTES);      // [heads][T], pre-divided by the scale
  if (MODE == 1) {
    for (int i = threadIdx.x; i < TBL * heads; i += blockDim.x) {
      const int hh = i / TBL, e = i - hh * TBL;
      tbl_all[i] = __ldg(rel_table + (long)e * heads + hh) * (1.0f / kScale);
    }
  }
  // rows >= N of every tile stay zero for the whole kernel (only rows < N are ever copied)
  for (int i = lane; i < WATT_STAGES * WATT_STAGE_BYTES / 16; i += 32) reinterpret_cast<uint4 *>(wbase)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  const int wpr = W / WS, wrows = TH / WS;
  const int nW = wrows * wpr;
  const long L = (long)TH * W;
  const int g = lane >> 2, t = lane & 3;
  // per-lane constants: key columns j = nt*8 + 2t + e  ->  a(j) | (j / WS) << 8 | (j % WS) << 12
  auto a_of = [](int p) { return (p / WS) * (2 * WS - 1) + p % WS; };
  int jpack[16];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = min(nt * 8 + 2 * t + e, N - 1);
      jpack[nt * 2 + e] = a_of(j) | ((j / WS) << 8) | ((j % WS) << 12);
    }
  const long total_warps = (long)gridDim.x * WATT_WARPS;
  const long first = (long)blockIdx.x * WATT_WARPS + warp;

  auto issue = [&](long task, int stage) {
    // 32-bit index arithmetic (the host checks n_tasks < 2^31): a 64-bit division by a run-time value is a ~100-instruction call
    const unsigned win = (unsigned)task / (unsigned)heads;
    const int h = (int)((unsigned)task - win * (unsigned)heads);
    const unsigned b = win / (unsigned)nW;
    const int n = (int)(win - b * (unsigned)nW);
    const int wr = n / wpr, wc = n - wr * wpr;
    int *rows = rows_base + stage * 64;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int p = lane + 32 * u;
      if (p < N) rows[p] = (int)(b * L) + token_row<WS>(wr, wc, p, TH, W, shift);
    }
    __syncwarp();
    const T *base = qkv + h * 32 + (lane & 3) * 8;
    const uint32_t sbase = smem_addr(wbase + stage * WATT_STAGE_BYTES);
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      const int p = pass * 8 + (lane >> 2);
      if (pass * 8 < N && p < N) {
        const T *src = base + (long)rows[p] * 3 * C;
        const uint32_t dst = sbase + tile_off(p, lane & 3);
        cp_async_16(dst, src);
        cp_async_16(dst + ATT_TILE_BYTES, src + C);
        cp_async_16(dst + 2 * ATT_TILE_BYTES, src + 2 * C);
      }
    }
  };

  if (WATT_STAGES == 2) {
    if (first < n_tasks) issue(first, 0);
    cp_async_commit();
  }
  int stage = 0;
  for (long task = first; task < n_tasks; task += total_warps, stage ^= (WATT_STAGES - 1)) {
    if (WATT_STAGES == 2) {
      const long next = task + total_warps;
      if (next < n_tasks) issue(next, stage ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {                        // single tile set: the other 15 warps of the SM cover this warp's load latency
      issue(task, 0);
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncwarp();

    const unsigned win = (unsigned)task / (unsigned)heads;
    const int h = (int)((unsigned)task - win * (unsigned)heads);
    const int n = (int)(win % (unsigned)nW);
    const int wr = n / wpr, wc = n - wr * wpr;
    const int *rows = rows_base + stage * 64;
    const uint32_t sQ = smem_addr(wbase + stage * WATT_STAGE_BYTES), sK = sQ + ATT_TILE_BYTES, sV = sK + ATT_TILE_BYTES;
    const bool last_r = wr == wrows - 1, last_c = wc == wpr - 1;
    const bool masked = MODE == 1 && mshift > 0 && (last_r || last_c);
    const float *tbl = tbl_all + h * TBL;

    uint32_t kf[WATT_KV_REGS ? NT : 1][4], vf[WATT_KV_REGS ? 4 : 1][2][4];
    if (WATT_KV_REGS) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) ldsm_x4(sK + tile_off(nt * 8 + (lane & 7), lane >> 3), kf[nt % (WATT_KV_REGS ? NT : 1)][0], kf[nt % (WATT_KV_REGS ? NT : 1)][1], kf[nt % (WATT_KV_REGS ? NT : 1)][2], kf[nt % (WATT_KV_REGS ? NT : 1)][3]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int tok = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) ldsm_x4_t(sV + tile_off(tok, dp * 2 + (lane >> 4)), vf[kk % (WATT_KV_REGS ? 4 : 1)][dp][0], vf[kk % (WATT_KV_REGS ? 4 : 1)][dp][1], vf[kk % (WATT_KV_REGS ? 4 : 1)][dp][2], vf[kk % (WATT_KV_REGS ? 4 : 1)][dp][3]);
      }
    }
    // region ids of this lane's key columns (standard shift mask), only for windows that hold masked pairs
    int regj[16];
    if (masked) {
#pragma unroll
      for (int s = 0; s < 16; ++s) {
        const int jr = (jpack[s] >> 8) & 15, jc = (jpack[s] >> 12) & 15;
        const int hr = last_r ? (jr < WS - mshift ? 1 : 2) : 0;
        const int wreg = last_c ? (jc < WS - mshift ? 1 : 2) : 0;
        regj[s] = hr * 3 + wreg;
      }
    }
    const float mask_val = -100.0f / kScale;

#pragma unroll 1
    for (int mt = 0; mt * 16 < N; ++mt) {
      float s[8][4];
      const int i_lo = mt * 16 + g, i_hi = i_lo + 8;
      const int ic_lo = min(i_lo, N - 1), ic_hi = min(i_hi, N - 1);
      if (MODE == 1) {
        constexpr int OFF = (WS - 1) * 2 * WS;
        const float *t_lo = tbl + a_of(ic_lo) + OFF, *t_hi = tbl + a_of(ic_hi) + OFF;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[nt][e] = ((e >> 1) ? t_hi : t_lo)[-(jpack[nt * 2 + (e & 1)] & 0xff)];
        if (masked) {
          const int r_lo = (last_r ? (ic_lo / WS < WS - mshift ? 1 : 2) : 0) * 3 + (last_c ? (ic_lo % WS < WS - mshift ? 1 : 2) : 0);
          const int r_hi = (last_r ? (ic_hi / WS < WS - mshift ? 1 : 2) : 0) * 3 + (last_c ? (ic_hi % WS < WS - mshift ? 1 : 2) : 0);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (regj[nt * 2 + (e & 1)] != ((e >> 1) ? r_hi : r_lo)) s[nt][e] += mask_val;
        }
      } else {
        LoadedBias<N> init{bias + (long)h * N * N, mask ? mask + (long)n * N * N : nullptr, ic_lo, ic_hi, t, 1.0f / kScale};
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[nt][e] = init(e >> 1, nt, e);
      }
      uint32_t a[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4(sQ + tile_off(row, ks * 2 + (lane >> 4)), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (!WATT_KV_REGS) ldsm_x4(sK + tile_off(nt * 8 + (lane & 7), lane >> 3), kf[0][0], kf[0][1], kf[0][2], kf[0][3]);
        const int ki = WATT_KV_REGS ? nt : 0;
        mma_16<T>(s[nt], a[0][0], a[0][1], a[0][2], a[0][3], kf[ki][0], kf[ki][1]);
        mma_16<T>(s[nt], a[1][0], a[1][1], a[1][2], a[1][3], kf[ki][2], kf[ki][3]);
      }
      float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (nt * 8 + 7 >= N) {                        // only the last n-tile can hold padded key columns
            const int j = nt * 8 + 2 * t + (e & 1);
            if (j >= N) s[nt][e] = -INFINITY;
          }
          if (e < 2) m_lo = fmaxf(m_lo, s[nt][e]); else m_hi = fmaxf(m_hi, s[nt][e]);
        }
      }
      m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
      m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
      m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
      m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
      const float mc_lo = m_lo * kC, mc_hi = m_hi * kC;
      float sum_lo = 0.0f, sum_hi = 0.0f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float p;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(s[nt][e], kC, -((e < 2) ? mc_lo : mc_hi))));   // 2^-inf = 0 for padding
          s[nt][e] = p;
          if (e < 2) sum_lo += p; else sum_hi += p;
        }
      }
      sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
      sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
      sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
      sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
      float o[4][4];
#pragma unroll
      for (int dn = 0; dn < 4; ++dn)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[dn][e] = 0.0f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kk * 16 < N) {
          const bool two = 2 * kk + 1 < NT;             // second key n-tile of this k-step exists
          const uint32_t a0 = pack2<T>(s[2 * kk][0], s[2 * kk][1]);
          const uint32_t a1 = pack2<T>(s[2 * kk][2], s[2 * kk][3]);
          const uint32_t a2 = two ? pack2<T>(s[(2 * kk + 1) & 7][0], s[(2 * kk + 1) & 7][1]) : 0u;
          const uint32_t a3 = two ? pack2<T>(s[(2 * kk + 1) & 7][2], s[(2 * kk + 1) & 7][3]) : 0u;
#pragma unroll
          for (int dp = 0; dp < 2; ++dp) {
            const int vi = WATT_KV_REGS ? kk : 0;
            if (!WATT_KV_REGS) {
              const int tok = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
              ldsm_x4_t(sV + tile_off(tok, dp * 2 + (lane >> 4)), vf[0][dp][0], vf[0][dp][1], vf[0][dp][2], vf[0][dp][3]);
            }
            mma_16<T>(o[dp * 2], a0, a1, a2, a3, vf[vi][dp][0], vf[vi][dp][1]);
            mma_16<T>(o[dp * 2 + 1], a0, a1, a2, a3, vf[vi][dp][2], vf[vi][dp][3]);
          }
        }
      }
      const float inv_lo = 1.0f / sum_lo, inv_hi = 1.0f / sum_hi;
      if (i_lo < N) {
        T *dst = out + (long)rows[i_lo] * C + h * 32 + 2 * t;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) *reinterpret_cast<uint32_t *>(dst + dn * 8) = pack2<T>(o[dn][0] * inv_lo, o[dn][1] * inv_lo);
      }
      if (i_hi < N) {
        T *dst = out + (long)rows[i_hi] * C + h * 32 + 2 * t;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) *reinterpret_cast<uint32_t *>(dst + dn * 8) = pack2<T>(o[dn][2] * inv_hi, o[dn][3] * inv_hi);
      }
    }
    __syncwarp();      // every lane is done with this stage's tiles and rows before the next iteration refills them
  }
  cp_async_wait<0>();
}

__device__ __forceinline__ int cva_query_window_m(int j, int r, int N1, int nW1, int per_clip) {
  if (!per_clip) return j % N1;
  const int i = j / r;
  const int clip = i / nW1;
  return clip * nW1 + ((i % nW1) * r + j % r) % nW1;
}

// deformable cross-view attention core: o[i] = sum_t softmax(q[qidx(r i + t)] k[r i + t]^T * d^-1/2) v[r i + t]
template <typename T, int WS>
__global__ void __launch_bounds__(ATT_WARPS * 32) cva_attention_mma_kernel(const float *__restrict__ q, const T *__restrict__ kv,
                                                                           T *__restrict__ o_out, int N1, int TH1, int W, int C,
                                                                           int heads, int r, int per_clip, long n_tasks) {
  pdl_grid_sync();
  constexpr int N = WS * WS;
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long task = (long)blockIdx.x * ATT_WARPS + warp;
  if (task >= n_tasks) return;
  const int wpr = W / WS;
  const int nW1 = (TH1 / WS) * wpr;
  const int i = (int)(task / heads);
  const int h = (int)(task - (long)i * heads);
  const long L1 = (long)TH1 * W;
  uint8_t *tq = att_smem + warp * ATT_WARP_SMEM, *tk = tq + ATT_TILE_BYTES, *tv = tk + ATT_TILE_BYTES;
  long *rows = reinterpret_cast<long *>(tv + ATT_TILE_BYTES);
  const uint32_t sQ = smem_addr(tq), sK = smem_addr(tk), sV = smem_addr(tv);
  const int g = lane >> 2, t4 = lane & 3;
  float o[4][4][4];                       // [m-tile][d n-tile][frag]
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int dn = 0; dn < 4; ++dn)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][dn][e] = 0.0f;
  for (int t = 0; t < r; ++t) {
    const int j = r * i + t;
    const int qw = cva_query_window_m(j, r, N1, nW1, per_clip);
    const long qb = qw / nW1;
    const int qn = qw - (int)qb * nW1;
    const int wr = qn / wpr, wc = qn - wr * wpr;
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int p = lane + 32 * u;
      if (p < N) rows[p] = qb * L1 + token_row<WS>(wr, wc, p, TH1, W, 0);
    }
    __syncwarp();
    // k / v: already in operand precision -- asynchronous 16-byte copies straight into the tiles (no registers), in flight while
    // the q rows below are loaded, converted and stored (ncu: the kernel waited on its global loads at 17 % occupancy)
    const T *kvb = kv + ((long)j * N) * 2 * C + h * 32;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t tile = which == 0 ? sK : sV;
#pragma unroll
      for (int pass = 0; pass < 8; ++pass) {
        const int p = pass * 8 + (lane >> 2);
        const int chunk = lane & 3;
        if (p < N) cp_async_16(tile + tile_off(p, chunk), kvb + (long)p * 2 * C + which * C + chunk * 8);
        else *reinterpret_cast<uint4 *>((which == 0 ? tk : tv) + tile_off(p, chunk)) = make_uint4(0, 0, 0, 0);
      }
    }
    cp_async_commit();
    // q: fp32 canvas rows -> bf16 tile
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      const int p = pass * 8 + (lane >> 2);
      const int chunk = lane & 3;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (p < N) {
        const float *src = q + rows[p] * C + h * 32 + chunk * 8;
        const float4 f0 = __ldg(reinterpret_cast<const float4 *>(src)), f1 = __ldg(reinterpret_cast<const float4 *>(src + 4));
        v.x = pack2<T>(f0.x, f0.y); v.y = pack2<T>(f0.z, f0.w); v.z = pack2<T>(f1.x, f1.y); v.w = pack2<T>(f1.z, f1.w);
      }
      *reinterpret_cast<uint4 *>(tq + tile_off(p, chunk)) = v;
    }
    cp_async_wait<0>();
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      if (mt * 16 < N) {
        float s[8][4], sum_lo, sum_hi;
        scores_softmax<T, N>(sQ, sK, mt, lane, 0.17677669529663687f, NoBias{}, s, sum_lo, sum_hi);
        float ot[4][4];
#pragma unroll
        for (int dn = 0; dn < 4; ++dn)
#pragma unroll
          for (int e = 0; e < 4; ++e) ot[dn][e] = 0.0f;
        pv_accumulate<T>(sV, lane, s, ot);
        const float inv_lo = 1.0f / sum_lo, inv_hi = 1.0f / sum_hi;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          o[mt][dn][0] = fmaf(ot[dn][0], inv_lo, o[mt][dn][0]);
          o[mt][dn][1] = fmaf(ot[dn][1], inv_lo, o[mt][dn][1]);
          o[mt][dn][2] = fmaf(ot[dn][2], inv_hi, o[mt][dn][2]);
          o[mt][dn][3] = fmaf(ot[dn][3], inv_hi, o[mt][dn][3]);
        }
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int i_lo = mt * 16 + g, i_hi = i_lo + 8;
    if (i_lo < N) {
      T *dst = o_out + ((long)i * N + i_lo) * C + h * 32 + 2 * t4;
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) *reinterpret_cast<uint32_t *>(dst + dn * 8) = pack2<T>(o[mt][dn][0], o[mt][dn][1]);
    }
    if (i_hi < N) {
      T *dst = o_out + ((long)i * N + i_hi) * C + h * 32 + 2 * t4;
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) *reinterpret_cast<uint32_t *>(dst + dn * 8) = pack2<T>(o[mt][dn][2], o[mt][dn][3]);
    }
  }
}

// opt the kernel into > 48 KB of dynamic shared memory, once per kernel (keyed by entry address: several template
// instantiations share one function-pointer type)
template <typename K>
static int att_smem_attr(K kernel) {
  static const void *seen[16];
  static int n_seen = 0;
  const void *key = reinterpret_cast<const void *>(kernel);
  for (int i = 0; i < n_seen; ++i)
    if (seen[i] == key) return MUMPY_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_WARPS * ATT_WARP_SMEM);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e));
    return MUMPY_ERR_CUDA;
  }
  if (n_seen < 16) seen[n_seen++] = key;
  return MUMPY_OK;
}

template <typename K>
static int watt_smem_attr(K kernel, size_t bytes) {
  static const void *seen[8];
  static size_t granted[8];
  static int n_seen = 0;
  const void *key = reinterpret_cast<const void *>(kernel);
  int slot = -1;
  for (int i = 0; i < n_seen; ++i)
    if (seen[i] == key) slot = i;
  if (slot >= 0 && bytes <= granted[slot]) return MUMPY_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(window attention, %zu B): %s", bytes, cudaGetErrorString(e));
    return MUMPY_ERR_CUDA;
  }
  if (slot < 0 && n_seen < 8) slot = n_seen++;
  if (slot >= 0) {
    seen[slot] = key;
    granted[slot] = bytes;
  }
  return MUMPY_OK;
}

int window_attention_tc(const void *qkv, const float *rel_table, int standard_mask, void *out, int dtype, int B, int TH, int W, int C, int heads,
                        int ws, int shift, cudaStream_t st);

static int g_att_tc = -1;      // -1: read MUMPY_ATT_TC on first use (default on)
void set_attention_tc(int enabled) { g_att_tc = enabled ? 1 : 0; }
static bool attention_tc_enabled() {
  if (g_att_tc < 0) {
    const char *e = getenv("MUMPY_ATT_TC");
    g_att_tc = (e && e[0] == '0') ? 0 : 1;
  }
  return g_att_tc == 1;
}

template <typename T>
static int window_attention_mma_t(const void *qkv, const float *bias, const float *mask, const float *rel_table, int standard_mask, void *out, int B,
                                  int TH, int W, int C, int heads, int ws, int shift, cudaStream_t st) {
  const long n_tasks = (long)B * (TH / ws) * (W / ws) * heads;
  MUMPY_REQUIRE((long)B * TH * W < (1l << 31) && n_tasks < (1l << 31), "window_attention(16-bit): too many tokens");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const long want = cdiv(n_tasks, WATT_WARPS);
  const unsigned grid = (unsigned)(want < num_sms ? want : num_sms);
  const bool table_mode = rel_table != nullptr && (mask == nullptr || standard_mask);
  // tcgen05 kernel unless the per-CTA copy of the bias table gets large (>= 32 heads: few windows, the staging dominates)
  if (table_mode && (ws == 7 || ws == 8) && heads * 32 == C && (2 * ws - 1) * (2 * ws - 1) * heads * 4 <= 16 * 1024 && attention_tc_enabled())
    return window_attention_tc(qkv, rel_table, mask != nullptr, out, std::is_same<T, __half>::value ? MUMPY_F16 : MUMPY_BF16, B, TH, W, C, heads, ws, shift, st);
  const int Tn = (2 * ws - 1) * (2 * ws - 1);
  const size_t smem = (size_t)WATT_WARPS * WATT_WARP_BYTES + (table_mode ? (size_t)Tn * heads * sizeof(float) : 0);
  MUMPY_REQUIRE(smem <= 227 * 1024, "window_attention(16-bit): %d heads need %zu B of shared memory", heads, smem);
  const T *q = static_cast<const T *>(qkv);
  T *o = static_cast<T *>(out);
  int rc;
  const int mshift = (mask != nullptr) ? shift : 0;      // region-id mask only when the caller passed the (standard) mask
#define WATT_LAUNCH(WS_, MODE_)                                                                                               \
  {                                                                                                                           \
    if ((rc = watt_smem_attr(window_attention_mma_kernel<T, WS_, MODE_>, smem))) return rc;                                   \
    launch_kernel(window_attention_mma_kernel<T, WS_, MODE_>, grid, WATT_WARPS * 32, smem, st, q, bias, mask, rel_table, o, TH, W, C, heads, shift, mshift, n_tasks); \
  }
  if (ws == 7) {
    if (table_mode) WATT_LAUNCH(7, 1) else WATT_LAUNCH(7, 0)
  } else if (ws == 8) {
    if (table_mode) WATT_LAUNCH(8, 1) else WATT_LAUNCH(8, 0)
  } else {
    set_error("window_attention(16-bit): window size %d unsupported (7 or 8)", ws);
    return MUMPY_ERR_UNSUPPORTED;
  }
#undef WATT_LAUNCH
  return launch_status("window_attention_mma");
}

int window_attention_mma(const void *qkv, const float *bias, const float *mask, const float *rel_table, int standard_mask, void *out, int dtype,
                         int B, int TH, int W, int C, int heads, int ws, int shift, cudaStream_t st) {
  if (dtype == MUMPY_F16) return window_attention_mma_t<__half>(qkv, bias, mask, rel_table, standard_mask, out, B, TH, W, C, heads, ws, shift, st);
  return window_attention_mma_t<__nv_bfloat16>(qkv, bias, mask, rel_table, standard_mask, out, B, TH, W, C, heads, ws, shift, st);
}

template <typename T>
static int cva_attention_mma_t(const float *q, const void *kv, void *o, int B, int TH1, int TH2, int W, int C, int heads, int ws, int per_clip,
                               cudaStream_t st) {
  const int N1 = B * (TH1 / ws) * (W / ws);
  const long n_tasks = (long)N1 * heads;
  const unsigned grid = (unsigned)cdiv(n_tasks, ATT_WARPS);
  const size_t smem = ATT_WARPS * ATT_WARP_SMEM;
  const int r = TH2 / TH1;
  int rc;
  if (ws == 7) {
    if ((rc = att_smem_attr(cva_attention_mma_kernel<T, 7>))) return rc;
    launch_kernel(cva_attention_mma_kernel<T, 7>, grid, ATT_WARPS * 32, smem, st, q, static_cast<const T *>(kv), static_cast<T *>(o), N1, TH1, W, C, heads, r,
                  per_clip, n_tasks);
  } else if (ws == 8) {
    if ((rc = att_smem_attr(cva_attention_mma_kernel<T, 8>))) return rc;
    launch_kernel(cva_attention_mma_kernel<T, 8>, grid, ATT_WARPS * 32, smem, st, q, static_cast<const T *>(kv), static_cast<T *>(o), N1, TH1, W, C, heads, r,
                  per_clip, n_tasks);
  } else {
    set_error("cva_attention(16-bit): window size %d unsupported (7 or 8)", ws);
    return MUMPY_ERR_UNSUPPORTED;
  }
  return launch_status("cva_attention_mma");
}

int cva_attention_mma(const float *q, const void *kv, void *o, int dtype, int B, int TH1, int TH2, int W, int C, int heads, int ws, int per_clip,
                      cudaStream_t st) {
  if (dtype == MUMPY_F16) return cva_attention_mma_t<__half>(q, kv, o, B, TH1, TH2, W, C, heads, ws, per_clip, st);
  return cva_attention_mma_t<__nv_bfloat16>(q, kv, o, B, TH1, TH2, W, C, heads, ws, per_clip, st);
}

}  // namespace mumpy
