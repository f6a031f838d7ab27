// Swin window attention on tcgen05 (16-bit modes, relative-position table + standard shift mask; swinTransformer.py:117-149,
// 233-252).  One CTA of 128 threads works on a PAIR of windows and one head at a time: the 2 x 64 (49 or 64 live) query rows
// fill the 128 lanes of a tensor-memory accumulator, thread r owns query row r from the softmax to the store.
//
//   gather     : q, k and v rows (64 B per token and head) come straight from the canvas-ordered qkv matrix with 16-byte
//                cp.async, four lanes per token; window partition and the cyclic shift are folded into the row index.  All
//                three tiles use a 64-byte row pitch with the 64-byte swizzle: q and k are K-major operands, v is read as an
//                MN-major B operand (keys x dims, no transpose anywhere).
//   S  = Q K^T : two 128x64x32 MMAs (keys of window A -> columns 0-63, keys of window B -> columns 64-127).  Row r only
//                reads the 64 columns of its own window; the other half is the price of using one 128-row atom.
//   softmax    : the thread reads its row with tcgen05.ld and walks the keys with compile-time indices, so the
//                relative-position bias is one shared-memory load at an immediate offset per key (no index arithmetic),
//                the shift mask one bit test; exp2 with log2e folded into the scale and the table.
//   O  = P V   : P (16-bit pairs) is written back into the thread's own lane of tensor memory (tcgen05.st) and the MMAs take
//                their A operand from there; two 128x32x64 MMAs (V of window A -> columns 0-31, of window B -> 32-63), row r
//                reads its window's half, scales by 1/sum and stores through a shared-memory transpose (64 contiguous bytes
//                per four lanes).
//
// Software pipeline per CTA: the next tile's q/k gather is issued when S completes, its v gather when O has been read.
// 24 KB of tiles + the per-head bias table per CTA, 128 TMEM columns, 128 registers: four CTAs per SM cover each other's
// MMA round trips and memory latencies.  (-DWTC_TIMING prints per-phase clock64() deltas of CTA 0.)
#include <cstdio>
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace mumpy {

#ifdef WTC_TIMING
#define WTC_T(i) ts_[i] = clock64()
#else
#define WTC_T(i)
#endif

constexpr int WTC_THREADS = 128;
constexpr int WTC_Q_BYTES = 128 * 64;         // query rows of both windows: 128 rows x 64 B (d = 32), 64-byte swizzle
constexpr int WTC_K_BYTES = 64 * 64;          // keys of one window: 64 rows x 64 B, 64-byte swizzle (K-major B operand)
constexpr int WTC_V_BYTES = 64 * 64;          // values of one window, same layout (read as an MN-major B operand)
constexpr int WTC_TILE_BYTES = WTC_Q_BYTES + 2 * WTC_K_BYTES + 2 * WTC_V_BYTES;
constexpr int WTC_P_COL = 64;                 // probabilities (16-bit pairs, 32 columns) inside the accumulator allocation
constexpr int WTC_TMEM_COLS = 128;

__device__ __forceinline__ void wtc_cp_async_16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void wtc_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void wtc_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// waits until at most the most recent commit group of this thread is still in flight
__device__ __forceinline__ void wtc_cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (row i in lane i, 16-bit elements packed two per column) comes from tensor memory
__device__ __forceinline__ void umma_bf16_tmem_a(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 64-byte-swizzled operand tile with a 64-byte row pitch (8-row groups 512 B apart).  As a K-major operand (q, k) the 32
// dims of a row are the K extent (two k-steps 32 B apart); read as an MN-major B operand (v) the rows are the K index (keys)
// and the 32 dims the N extent: canonical layouts ((8,m),(T,2)):((4T,SBO),(1,T)) and ((T,4,m),(8,k)):((1,T,LBO),(4T,SBO)),
// T = 8 elements, SBO = 512 B, layout type SWIZZLE_64B = 4.
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

template <int WS>
__device__ __forceinline__ int wtc_token_row(int wr, int wc, int p, int TH, int W, int shift) {
  int r = wr * WS + p / WS + shift;
  int c = wc * WS + p % WS + shift;
  if (r >= TH) r -= TH;
  if (c >= W) c -= W;
  return r * W + c;
}

template <typename T, int WS>
__global__ void __launch_bounds__(WTC_THREADS, 4) window_attention_tc_kernel(const T *__restrict__ qkv, const float *__restrict__ rel_table,
                                                                              T *__restrict__ out, int TH, int W, int C, int heads, int shift,
                                                                              int mshift, int n_win, int n_tiles, int per_cta) {
  constexpr int N = WS * WS;
  constexpr int TBL = (2 * WS - 1) * (2 * WS - 1);
  constexpr int OFF = (WS - 1) * 2 * WS;
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr float kC = 0.17677669529663687f * kLog2e;          // 32^-0.5 * log2(e)
  constexpr bool kF16 = std::is_same<T, __half>::value;
  extern __shared__ uint8_t wtc_raw[];
  const uint32_t base = (smem_u32(wtc_raw) + 1023u) & ~1023u;
  uint8_t *gen = wtc_raw + (base - smem_u32(wtc_raw));
  const uint32_t sQ = base, sK = sQ + WTC_Q_BYTES, sV = sK + 2 * WTC_K_BYTES;
  float *tbl_all = reinterpret_cast<float *>(gen + WTC_TILE_BYTES);              // [heads][TBL], times log2(e)
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int half = tid >> 6, p = tid & 63;
#ifdef WTC_TIMING
  const long long t_entry = clock64();
  unsigned long long g_entry;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_entry));
#endif

  // Tensor memory first: the SM does not start the next CTA of a kernel that allocates tensor memory until this one has given up
  // its allocation permit (tools/cta_launch_bench.cu: CTAs 2, 3, 4 of an SM enter 0.9 / 1.7 / 2.4 us after the first when the
  // allocation follows ~1 us of prologue, all within 64 ns of each other without tensor memory) -- with four CTAs per SM the
  // zero fill and table staging below used to delay the fourth CTA by three prologues.
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(WTC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < WTC_TILE_BYTES / 16; i += WTC_THREADS) reinterpret_cast<uint4 *>(gen)[i] = make_uint4(0, 0, 0, 0);
  // bias table (TBL, heads) -> shared [heads][TBL] x log2(e).  Coalesced 16-byte reads, four in flight per thread: the former
  // transposing loop issued one dependent 4-byte load per iteration (21 L2 round trips, ~8 us before a CTA's first tile)
  if ((heads & 3) == 0) {
    const int n4 = TBL * heads / 4;
    const float4 *t4 = reinterpret_cast<const float4 *>(rel_table);
    for (int base = tid; base < n4; base += 4 * WTC_THREADS) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * WTC_THREADS;
        v[u] = i < n4 ? __ldg(t4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * WTC_THREADS;
        if (i < n4) {
          const int e = (4 * i) / heads, hh = 4 * i - e * heads;          // four consecutive heads of table entry e
          float *d = tbl_all + hh * TBL + e;
          d[0] = v[u].x * kLog2e;
          d[TBL] = v[u].y * kLog2e;
          d[2 * TBL] = v[u].z * kLog2e;
          d[3 * TBL] = v[u].w * kLog2e;
        }
      }
    }
  } else {
    const int n = TBL * heads;
    for (int base = tid; base < n; base += 4 * WTC_THREADS) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * WTC_THREADS;
        v[u] = i < n ? __ldg(rel_table + i) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * WTC_THREADS;
        if (i < n) {
          const int e = i / heads, hh = i - e * heads;
          tbl_all[hh * TBL + e] = v[u] * kLog2e;
        }
      }
    }
  }
  const uint32_t bar_s = smem_u32(&bars[0]), bar_o = smem_u32(&bars[1]);
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();          // the zero fill is read by the MMAs
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);

  const int wpr = W / WS, wrows = TH / WS;
  const int nW = wrows * wpr;
  const int L = TH * W;
  const uint32_t idesc_s = make_idesc_16_f32(128, 64, kF16);
  const uint32_t idesc_o = make_idesc_16_f32(128, 32, kF16) | (1u << 16);       // B (= V) is MN-major: dims contiguous per key
  const uint64_t dQ = make_sw64_desc(sQ);
  auto a_of = [](int t) { return (t / WS) * (2 * WS - 1) + t % WS; };
  const int pc = min(p, N - 1);
  const int a_i = a_of(pc) + OFF;
  const uint32_t o_row = sV + tid * 64, o_sw = static_cast<uint32_t>((tid >> 1) & 3);      // output staging (see the epilogue)

  // per-tile state of this thread's row: canvas token, liveness, shift-mask bits
  struct RowState {
    int pair, h;
    long row;
    bool valid, masked;
    unsigned long long mbits;
  };
  // canvas rows of the four tokens this lane gathers for the tile located last (lane>>2 + 8 * pass of the warp's 32 rows; -1 = dead
  // row): refreshed by shuffle only when the window pair changes, i.e. once per `heads` tiles instead of 12 times per tile
  const int lane = tid & 31;
  int grow[4] = {-1, -1, -1, -1};
  auto locate = [&](int pair, int h, RowState &rs) {
    rs.h = h;
    if (pair == rs.pair) return;
    rs.pair = pair;
    const int win = 2 * pair + half;
    rs.valid = p < N && win < n_win;
    const int wclamped = min(win, n_win - 1);
    const int b = wclamped / nW, n = wclamped - b * nW;
    const int wr = n / wpr, wc = n - wr * wpr;
    rs.row = (long)b * L + wtc_token_row<WS>(wr, wc, pc, TH, W, shift);
    const bool last_r = wr == wrows - 1, last_c = wc == wpr - 1;
    rs.masked = mshift > 0 && (last_r || last_c);
    if (rs.masked) {
      // keys in the same shift region as this query (region = row part x column part of the window, swinTransformer.py:236-247)
      const unsigned long long r1 = (1ull << (WS * (WS - mshift))) - 1ull;             // keys with jr < WS - mshift
      unsigned long long c1 = 0;
      for (int r = 0; r < WS; ++r) c1 |= ((1ull << (WS - mshift)) - 1ull) << (r * WS);  // keys with jc < WS - mshift
      const bool i_r1 = pc / WS < WS - mshift, i_c1 = pc % WS < WS - mshift;
      const unsigned long long same_r = last_r ? (i_r1 ? r1 : ~r1) : ~0ull;
      const unsigned long long same_c = last_c ? (i_c1 ? c1 : ~c1) : ~0ull;
      rs.mbits = ~(same_r & same_c);
    }
    const int enc = rs.valid ? (int)rs.row : -1;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) grow[pass] = __shfl_sync(0xffffffffu, enc, pass * 8 + (lane >> 2));
  };
  // Gather: a warp fetches the q / k / v segments of its own 32 rows, four lanes per row (64 contiguous bytes, 8 rows per
  // instruction) -- the token row of the lane's gather row comes from its owner by shuffle.  Destination rows are 64 B with
  // 16-byte chunks XOR (row >> 1) & 3: the SWIZZLE_64B pattern for a 64-byte row pitch.
  const uint32_t g_chunk = static_cast<uint32_t>(lane & 3);
  const uint32_t g_dst = static_cast<uint32_t>((tid & 32) + (lane >> 2));          // pass 0 row inside the window's 64-row tile (+8 per pass)
  const long row_pitch = 3l * C;
  auto gather = [&](const RowState &rs, int which, uint32_t tile_base) {       // which: 0 q, 1 k, 2 v (element offset which * C)
    const T *src0 = qkv + which * C + rs.h * 32 + g_chunk * 8;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      if (grow[pass] >= 0) {
        const uint32_t r = g_dst + pass * 8;
        wtc_cp_async_16(tile_base + r * 64 + ((g_chunk ^ ((r >> 1) & 3)) << 4), src0 + grow[pass] * row_pitch);
      }
    }
  };
  auto issue_qk = [&](const RowState &rs) {
    gather(rs, 0, sQ + half * (64 * 64));
    gather(rs, 1, sK + half * WTC_K_BYTES);
  };
  auto issue_v = [&](const RowState &rs) { gather(rs, 2, sV + half * WTC_V_BYTES); };

  // Everything above touches only shared / tensor memory and the (constant) bias table: under programmatic dependent launch it
  // overlaps the tail of the kernel that produces qkv; the first read of qkv comes after this wait.
#ifdef WTC_TIMING
  const long long t_prologue = clock64();
#endif
  pdl_grid_sync();
#ifdef WTC_TIMING
  const long long t_released = clock64();
#endif
  const int t_begin = blockIdx.x * per_cta;
  const int t_end = min(n_tiles, t_begin + per_cta);
  RowState cur;
  cur.pair = -1;
  cur.valid = cur.masked = false;
  cur.mbits = 0;
  cur.row = 0;
  cur.h = 0;
  if (t_begin < t_end) {
    locate(t_begin / heads, t_begin % heads, cur);
    issue_qk(cur);
    wtc_cp_async_commit();
    issue_v(cur);
    wtc_cp_async_commit();
  }
  uint32_t phase = 0;
#ifdef WTC_TIMING
  long long ts_[11];
  long long acc_[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
  for (int tile = t_begin; tile < t_end; ++tile) {
    // Software pipeline: the next tile's q/k gather is issued as soon as S is complete (it lands during the softmax), its v
    // gather as soon as O is complete (it lands during the next tile's S and softmax).
    WTC_T(0);
    const bool has_next = tile + 1 < t_end;
    RowState nxt = cur;
    if (has_next) {
      const bool wrap = cur.h + 1 == heads;            // heads of one window pair are consecutive tiles
      locate(cur.pair + (wrap ? 1 : 0), wrap ? 0 : cur.h + 1, nxt);
    }
    // q / k of this tile (the older commit group) must have landed; its v (the newest group, issued at the very end of the previous
    // tile) is only needed by the PV MMAs and is awaited there -- it used to be the exposed part of this wait (~1000 clk per tile
    // on the stage-0 canvas)
    wtc_cp_async_wait_but_one();
    fence_proxy_async_smem();
    __syncthreads();
    WTC_T(1);
    // ---- S = Q K^T
    if (warp == 0) {              // converged warp, one elected lane issues (tc_common.cuh elect_one: descriptors stay in uniform registers)
      tc_fence_after();
      const uint64_t dK0 = make_sw64_desc(sK), dK1 = make_sw64_desc(sK + WTC_K_BYTES);
      if (elect_one()) {
        umma_bf16(tmem, dQ, dK0, idesc_s, 0u);
        umma_bf16(tmem, dQ + 2, dK0 + 2, idesc_s, 1u);
        umma_bf16(tmem + 64, dQ, dK1, idesc_s, 0u);
        umma_bf16(tmem + 64, dQ + 2, dK1 + 2, idesc_s, 1u);
        umma_commit(bar_s);
      }
      __syncwarp();
    }
    mbar_wait(bar_s, phase);
    tc_fence_after();
    WTC_T(2);
    if (has_next) {
      issue_qk(nxt);
      wtc_cp_async_commit();
    }
    WTC_T(3);
    // ---- softmax of this thread's row (exp2 domain)
    uint32_t s_lo[32], s_hi[32];
    tmem_ld32(t_row + half * 64, s_lo);
    tmem_ld32(t_row + half * 64 + 32, s_hi);
    float inv = 0.0f;
    uint32_t pk[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) pk[e] = 0u;
    if (cur.valid) {
      // Packed fp32x2 arithmetic (FFMA2 / FADD2) on key pairs (2e, 2e+1): the same roundings and the same four interleaved partial
      // sums as the scalar form, two thirds of its issue slots (ncu: the kernel is bound by its instruction stream at ~3.6 warps
      // per scheduler).  Key N (odd window sizes) is a pad: -1e30 leaves the maximum alone and its probability is exactly zero.
      const float *tb = tbl_all + cur.h * TBL + a_i;
      constexpr int NP = (N + 1) / 2;
      float2 s2[NP];
#pragma unroll
      for (int e = 0; e < NP; ++e) {
        const int j = 2 * e;
        const float r0 = __uint_as_float(j < 32 ? s_lo[j & 31] : s_hi[j & 31]);
        if (j + 1 < N) {
          const float r1 = __uint_as_float(j + 1 < 32 ? s_lo[(j + 1) & 31] : s_hi[(j + 1) & 31]);
          s2[e] = __ffma2_rn(make_float2(r0, r1), make_float2(kC, kC), make_float2(tb[-a_of(j)], tb[-a_of(j + 1)]));
        } else {
          s2[e] = make_float2(fmaf(r0, kC, tb[-a_of(j)]), -1e30f);
        }
      }
      if (cur.masked) {
#pragma unroll
        for (int j = 0; j < N; ++j)
          if ((cur.mbits >> j) & 1ull) {
            if (j & 1) s2[j >> 1].y += -100.0f * kLog2e;
            else s2[j >> 1].x += -100.0f * kLog2e;
          }
      }
      // four interleaved partial maxima / sums: the reductions are dependent chains, and a CTA has one warp per scheduler
      float m4[4] = {s2[0].x, s2[0].y, s2[1].x, s2[1].y};
#pragma unroll
      for (int e = 2; e < NP; ++e) {
        m4[(2 * e) & 3] = fmaxf(m4[(2 * e) & 3], s2[e].x);
        m4[(2 * e + 1) & 3] = fmaxf(m4[(2 * e + 1) & 3], s2[e].y);
      }
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const float2 nm = make_float2(-m, -m);
      float2 sum2[2] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};
#pragma unroll
      for (int e = 0; e < NP; ++e) {
        const float2 d = __fadd2_rn(s2[e], nm);
        const float2 pr = make_float2(ex2_approx(d.x), ex2_approx(d.y));
        sum2[e & 1] = __fadd2_rn(sum2[e & 1], pr);
        pk[e] = pack2<T>(pr.x, pr.y);
      }
      inv = 1.0f / ((sum2[0].x + sum2[0].y) + (sum2[1].x + sum2[1].y));
    }
    // P goes into this thread's own lane of tensor memory (columns 64-95, inside the S block the row has finished reading):
    // the PV MMAs take their A operand from there, so the probabilities never touch shared memory.
    tmem_st32(t_row + WTC_P_COL, pk);
    WTC_T(4);
    // v of this tile: every group but the next tile's q / k gather (the newest one, when there is a next tile)
    if (has_next) wtc_cp_async_wait_but_one(); else wtc_cp_async_wait_all();
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    WTC_T(5);
    // ---- O = P V
    if (warp == 0) {
      tc_fence_after();
      const uint64_t dV0 = make_sw64_desc(sV), dV1 = make_sw64_desc(sV + WTC_V_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)      // per k-step: 16 keys = 8 packed columns of P, 1024 B of V
          umma_bf16_tmem_a(tmem, tmem + WTC_P_COL + 8 * k, dV0 + (16 * 64 / 16) * k, idesc_o, k);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_tmem_a(tmem + 32, tmem + WTC_P_COL + 8 * k, dV1 + (16 * 64 / 16) * k, idesc_o, k);
        umma_commit(bar_o);
      }
      __syncwarp();
    }
    mbar_wait(bar_o, phase);
    tc_fence_after();
    WTC_T(6);
    phase ^= 1;
    uint32_t o[32];
    tmem_ld32(t_row + half * 32, o);
    WTC_T(8);
    // The row's 32 outputs go through this warp's rows of the V tile (PV is complete, the next v gather not yet issued) so that
    // the global stores use the gather's mapping: four lanes write the 64 contiguous bytes of one token.
    if (cur.valid) {
#pragma unroll
      for (uint32_t c = 0; c < 4; ++c)
        sts_v4(o_row + ((c ^ o_sw) << 4), pack2<T>(__uint_as_float(o[c * 8 + 0]) * inv, __uint_as_float(o[c * 8 + 1]) * inv),
               pack2<T>(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv),
               pack2<T>(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv),
               pack2<T>(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv));
    }
    __syncwarp();
    {
      const int enc = cur.valid ? (int)cur.row : -1;
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int r_in_warp = pass * 8 + (lane >> 2);
        const int row = __shfl_sync(0xffffffffu, enc, r_in_warp);
        if (row >= 0) {
          const uint32_t r = static_cast<uint32_t>((tid & ~31) + r_in_warp);
          uint4 w;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(sV + r * 64 + ((g_chunk ^ ((r >> 1) & 3)) << 4)));
          *reinterpret_cast<uint4 *>(out + (long)row * C + cur.h * 32 + g_chunk * 8) = w;
        }
      }
    }
    __syncwarp();
    WTC_T(9);
    if (has_next) {
      issue_v(nxt);                   // (same warp, same rows: ordered after the staging reads)
      wtc_cp_async_commit();
    }
    WTC_T(7);
#ifdef WTC_TIMING
    if (tile > t_begin) {
      acc_[0] += ts_[1] - ts_[0]; acc_[1] += ts_[2] - ts_[1]; acc_[2] += ts_[3] - ts_[2]; acc_[3] += ts_[4] - ts_[3]; acc_[4] += ts_[5] - ts_[4];
      acc_[5] += ts_[6] - ts_[5]; acc_[6] += ts_[8] - ts_[6]; acc_[7] += ts_[9] - ts_[8]; acc_[8] += ts_[7] - ts_[9]; acc_[9] += 1;
    }
#endif
    tc_fence_before();      // (the barrier at the top of the next iteration orders these accumulator reads before the next S)
    cur = nxt;
  }
#ifdef WTC_TIMING
  if ((blockIdx.x % 97 == 5) && (tid == 0 || tid == 127) && acc_[9] > 0) {
    const long long n_ = acc_[9];
    printf("cta %d tid %d: per tile (avg of %lld): gather wait+sync %lld | S mma %lld | issue qk %lld | softmax %lld | sync2 %lld | PV mma %lld | ld O %lld | stage+store %lld | issue v %lld\n", blockIdx.x,
           tid, n_, acc_[0] / n_, acc_[1] / n_, acc_[2] / n_, acc_[3] / n_, acc_[4] / n_, acc_[5] / n_, acc_[6] / n_, acc_[7] / n_, acc_[8] / n_);
  }
  if (false) {
    unsigned long long g_exit;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_exit));
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    printf("cta %d sm %u: entry %llu ns  exit %llu ns | prologue %lld  pdl wait %lld  tiles %lld clk (%d tiles)\n", blockIdx.x, smid, g_entry % 10000000ull, g_exit % 10000000ull,
           t_prologue - t_entry, t_released - t_prologue, clock64() - t_released, t_end - t_begin);
  }
#endif
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(WTC_TMEM_COLS) : "memory");
  }
}

template <typename T, int WS>
static int launch_wtc(const T *qkv, const float *rel_table, T *out, int B, int TH, int W, int C, int heads, int shift, int mshift, cudaStream_t st) {
  const int nW = (TH / WS) * (W / WS);
  const long n_win = (long)B * nW;
  const long n_tiles = ((n_win + 1) / 2) * heads;
  MUMPY_REQUIRE(n_tiles < (1l << 30) && (long)B * TH * W < (1l << 31), "window_attention(tcgen05): too many tokens");
  const size_t smem = 1024 + WTC_TILE_BYTES + (size_t)(2 * WS - 1) * (2 * WS - 1) * heads * sizeof(float);
  static size_t granted = 0;
  if (smem > granted) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_tc_kernel<T, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("window_attention(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    granted = smem;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  // resident CTAs per SM: 512 TMEM columns / 128, 64 K registers / (128 threads x 128), 227 KB of shared memory (+1 KB reserved per CTA)
  const long by_smem = (227 * 1024) / (long)(smem + 1024);
  const long per_sm = by_smem < 1 ? 1 : (by_smem > 512 / WTC_TMEM_COLS ? 512 / WTC_TMEM_COLS : by_smem);
  const long slots = (long)num_sms * per_sm;
  const long per_cta = cdiv(n_tiles, n_tiles < slots ? n_tiles : slots);
  const unsigned grid = (unsigned)cdiv(n_tiles, per_cta);
  launch_kernel(window_attention_tc_kernel<T, WS>, grid, WTC_THREADS, smem, st, qkv, rel_table, out, TH, W, C, heads, shift, mshift, (int)n_win,
                (int)n_tiles, (int)per_cta);
  return launch_status("window_attention_tc");
}

// table-mode window attention (relative_position_bias_table + optional standard shift mask) on tcgen05; ws 7 or 8
int window_attention_tc(const void *qkv, const float *rel_table, int standard_mask, void *out, int dtype, int B, int TH, int W, int C, int heads,
                        int ws, int shift, cudaStream_t st) {
  const int mshift = standard_mask ? shift : 0;
  if (dtype == MUMPY_F16) {
    if (ws == 7) return launch_wtc<__half, 7>(static_cast<const __half *>(qkv), rel_table, static_cast<__half *>(out), B, TH, W, C, heads, shift, mshift, st);
    return launch_wtc<__half, 8>(static_cast<const __half *>(qkv), rel_table, static_cast<__half *>(out), B, TH, W, C, heads, shift, mshift, st);
  }
  if (ws == 7) return launch_wtc<__nv_bfloat16, 7>(static_cast<const __nv_bfloat16 *>(qkv), rel_table, static_cast<__nv_bfloat16 *>(out), B, TH, W, C, heads, shift, mshift, st);
  return launch_wtc<__nv_bfloat16, 8>(static_cast<const __nv_bfloat16 *>(qkv), rel_table, static_cast<__nv_bfloat16 *>(out), B, TH, W, C, heads, shift, mshift, st);
}

}  // namespace mumpy
