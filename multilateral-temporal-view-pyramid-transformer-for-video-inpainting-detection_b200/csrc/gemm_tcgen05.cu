// bf16 GEMM / implicit-GEMM convolution on the 5th-gen tensor cores.
//   D[m,n] = act( sum_k A[m,k] * W[n,k] + bias[n] ) + residual[m,n]      A (M,K), W (N,K) bf16, K-major both.
// Persistent, warp-specialised kernel, one CTA per SM, static round-robin tile scheduler (N-tile fastest so CTAs
// running together share the A tile through L2):
//   warp 8      TMA producer: 128B-swizzled A/B tiles into a 4-8 stage mbarrier ring.  In conv mode the A tile
//               comes from an im2col-mode tensor map over the NHWC activation (128 consecutive output pixels x 64
//               channels of one filter tap per k-block, zero-filled padding) -- no im2col buffer exists.
//   warp 9      TMEM allocator + MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN<=256, K=16, fp32
//               accumulators in TMEM, double buffered (2 x BN columns) so the epilogue of tile i overlaps the main
//               loop of tile i+1; tcgen05.commit releases smem slots and publishes accumulators.
//   warps 0-7   epilogue: tcgen05.ld (32 lanes x 32 columns per load; warp w reads lane quadrant w%4, column half
//               w/4), fused bias / GELU / ReLU / fp32 residual, 16-byte stores, fp32 or bf16 output.
// Workhorse of the bf16 mode: qkv / proj / fc1 / fc2 / pre / proj_{q,k,v,out} / reduction / globalembedding /
// global blocks / rgb_decoder linears and every decoder convolution (94% + 11% of the forward's FLOPs).
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace mumpy {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t, const cuuint32_t *,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_num_sms = 0;
static int g_driver_version = 0;

static void *driver_symbol(const char *name) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
  return fn;
}

int resolve_driver_entry_points() {
  if (g_encode_tiled && g_encode_im2col && g_num_sms) return MUMPY_OK;
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(driver_symbol("cuTensorMapEncodeTiled"));
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(driver_symbol("cuTensorMapEncodeIm2col"));
  if (!g_encode_tiled || !g_encode_im2col) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncode{Tiled,Im2col}) failed");
    return MUMPY_ERR_CUDA;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDriverGetVersion(&g_driver_version);
  if (g_num_sms <= 0) g_num_sms = 148;
  return MUMPY_OK;
}

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;       // 64 bf16 = 128 B = one swizzle span
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_EPI_WARPS = 12;
constexpr int TC_EPI_GROUPS = TC_EPI_WARPS / 4;      // warps per TMEM lane quadrant: each takes every TC_EPI_GROUPS-th 32-column chunk
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;
constexpr int TC_EPI_WARPS_LEAN = 16;
constexpr int TC_THREADS_LEAN = (TC_EPI_WARPS_LEAN + 2) * 32;
constexpr int TC_SMEM_BUDGET = 164 * 1024;      // operand ring; + 8 x (4 KB epilogue staging + 512 B bias) + 1 KB alignment slack

struct TcParams {
  const float *bias;
  const float *residual;
  void *out;
  uint16_t *aux;          // optional second output (fp32-out mode only): 16-bit act(acc + bias) BEFORE the residual add, row stride ldo
  long ldo;
  long M;
  int N, K;
  int BN;
  int stages;
  int act;
  int out_bf16;           // output is 16-bit (bf16, or f16 when `f16` is set)
  int f16;                // operands (and 16-bit outputs / the side output) are IEEE half instead of bf16
  int tiles_n;
  long num_tiles;
  uint32_t acc_cols;      // TMEM columns per accumulator slot (power of two >= BN)
  uint32_t idesc;
  // conv mode (A through the im2col tensor map)
  // (the struct is kept at its size: ptxas places part of a larger parameter block on the stack and the epilogue spills)
  short debug;
  short k_splits;         // split-K (kSplit kernels): a tile index also selects one of k_splits ranges of k-blocks; the partial
                          // sums of range s go to out + s * M * N (fp32, ldo = N)
  int conv;
  int Wout, Hout, lower_w, lower_h, kw, cblocks;
};

constexpr int EPI_STAGE_BYTES = 32 * 128;      // per-warp staging tile: 32 rows x 32 fp32, 16-byte chunks XOR-swizzled by row&7

// Epilogue of one 128 x BN accumulator tile for one warp (lane quadrant warp&3, column half warp>>2).
// Phase 1 (thread <-> accumulator row): tcgen05.ld of 32 columns, + bias, activation, into the swizzled staging tile.
// Phase 2 (coalesced): the warp re-reads the tile row-wise so that every global access is a full 128 B (fp32) or 64 B
// (bf16) row segment: residual loads, fp32->bf16 packing and the output stores are all coalesced.
template <int ACT, bool OUT_BF16, bool HAS_RES>
__device__ __forceinline__ void epilogue_tile(const TcParams &p, uint8_t *stage, const float *bias_s, uint32_t acc, int warp, int lane, long m0,
                                              int n0, long out_shift) {
  const int quad = warp & 3, half = warp >> 2;
  const uint32_t lane_addr = acc + (static_cast<uint32_t>(quad * 32) << 16);
  const long row0 = m0 + quad * 32;
  const uint32_t st_base = smem_u32(stage);
  float amax = 0.0f;          // largest |value| converted to IEEE half by this thread (fp16 range guard, common.cuh)
  for (int c0 = half * 32, ci = 0; c0 < p.BN; c0 += 32 * TC_EPI_GROUPS, ++ci) {
    // residual tile of this chunk, fetched in the coalesced phase-2 layout before anything else so that the global-load
    // latency overlaps the TMEM load and the phase-1 math
    float4 res[8];
    if (HAS_RES) {
      if (OUT_BF16) {
        const int col = c0 + (lane & 3) * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long gm = row0 + i * 8 + (lane >> 2);
          const bool ok = gm < p.M && col < p.BN && n0 + col < p.N;
          res[2 * i] = ok ? *reinterpret_cast<const float4 *>(p.residual + gm * p.ldo + n0 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
          res[2 * i + 1] = ok ? *reinterpret_cast<const float4 *>(p.residual + gm * p.ldo + n0 + col + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        const int col = c0 + (lane & 7) * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long gm = row0 + i * 4 + (lane >> 3);
          const bool ok = gm < p.M && col < p.BN && n0 + col < p.N;
          res[i] = ok ? *reinterpret_cast<const float4 *>(p.residual + gm * p.ldo + n0 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    uint32_t v[32];
    tmem_ld32(lane_addr + c0, v);
    if (p.debug == 2) continue;
    // ---- phase 1 ----  (bias of this chunk: staged in shared memory before the accumulator wait, zero when absent)
    const float4 *bias4 = reinterpret_cast<const float4 *>(bias_s + ci * 32);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float2 f0 = make_float2(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]));
      float2 f1 = make_float2(__uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
      {
        const float4 b = bias4[g];
        f0 = __fadd2_rn(f0, make_float2(b.x, b.y));
        f1 = __fadd2_rn(f1, make_float2(b.z, b.w));
      }
      if (ACT == 1) {
        f0 = gelu_fast2(f0);
        f1 = gelu_fast2(f1);
      } else if (ACT == 2) {
        f0.x = apply_act(f0.x, p.act); f0.y = apply_act(f0.y, p.act); f1.x = apply_act(f1.x, p.act); f1.y = apply_act(f1.y, p.act);
      }
      const float4 f = make_float4(f0.x, f0.y, f1.x, f1.y);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_base + lane * 128 + ((g ^ (lane & 7)) << 4)), "f"(f.x), "f"(f.y),
                   "f"(f.z), "f"(f.w)
                   : "memory");
    }
    __syncwarp();
    // ---- phase 2 ----
    if (p.debug != 1) {
      if (OUT_BF16) {
        const int c8 = lane & 3;                 // 8 columns (16 B of bf16) per lane, 4 lanes per row, 8 rows per pass
        const int col = c0 + c8 * 8;
        float4 xs[4], ys[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {          // all shared-memory reads first: their latencies overlap instead of chaining
          const int r = i * 8 + (lane >> 2);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(xs[i].x), "=f"(xs[i].y), "=f"(xs[i].z), "=f"(xs[i].w) : "r"(st_base + r * 128 + (((2 * c8) ^ (r & 7)) << 4)));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ys[i].x), "=f"(ys[i].y), "=f"(ys[i].z), "=f"(ys[i].w) : "r"(st_base + r * 128 + (((2 * c8 + 1) ^ (r & 7)) << 4)));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2);
          const long gm = row0 + r;
          float4 x = xs[i], y = ys[i];
          if (gm < p.M && col < p.BN && n0 + col < p.N) {
            if (HAS_RES) {
              const float4 r0 = res[2 * i], r1 = res[2 * i + 1];
              x.x += r0.x; x.y += r0.y; x.z += r0.z; x.w += r0.w;
              y.x += r1.x; y.y += r1.y; y.z += r1.z; y.w += r1.w;
            }
            uint4 u;
            if (p.f16) {
              amax = fmaxf(fmaxf(fmaxf(amax, fabsf(x.x)), fmaxf(fabsf(x.y), fabsf(x.z))), fmaxf(fmaxf(fabsf(x.w), fabsf(y.x)), fmaxf(fabsf(y.y), fmaxf(fabsf(y.z), fabsf(y.w)))));
              u.x = pack2<__half>(x.x, x.y); u.y = pack2<__half>(x.z, x.w); u.z = pack2<__half>(y.x, y.y); u.w = pack2<__half>(y.z, y.w);
            } else {
              u.x = pack2<__nv_bfloat16>(x.x, x.y); u.y = pack2<__nv_bfloat16>(x.z, x.w);
              u.z = pack2<__nv_bfloat16>(y.x, y.y); u.w = pack2<__nv_bfloat16>(y.z, y.w);
            }
            *reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(p.out) + out_shift + gm * p.ldo + n0 + col) = u;
          }
        }
      } else {
        const int c4 = lane & 7;                 // 4 columns (16 B of fp32) per lane, 8 lanes per row, 4 rows per pass
        const int col = c0 + c4 * 4;
        float4 xs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(xs[i].x), "=f"(xs[i].y), "=f"(xs[i].z), "=f"(xs[i].w) : "r"(st_base + r * 128 + ((c4 ^ (r & 7)) << 4)));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3);
          const long gm = row0 + r;
          float4 x = xs[i];
          if (gm < p.M && col < p.BN && n0 + col < p.N) {
            if (p.aux) {
              uint2 pk;
              if (p.f16) {
                amax = fmaxf(fmaxf(amax, fabsf(x.x)), fmaxf(fmaxf(fabsf(x.y), fabsf(x.z)), fabsf(x.w)));
                pk.x = pack2<__half>(x.x, x.y); pk.y = pack2<__half>(x.z, x.w);
              } else {
                pk.x = pack2<__nv_bfloat16>(x.x, x.y); pk.y = pack2<__nv_bfloat16>(x.z, x.w);
              }
              *reinterpret_cast<uint2 *>(p.aux + gm * p.ldo + n0 + col) = pk;
            }
            if (HAS_RES) {
              const float4 r0 = res[i];
              x.x += r0.x; x.y += r0.y; x.z += r0.z; x.w += r0.w;
            }
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out) + out_shift + gm * p.ldo + n0 + col) = x;
          }
        }
      }
    }
    __syncwarp();
  }
  f16_guard(amax);
}

// The lean 16-bit epilogue as an out-of-line call (once per warp and tile): inlined next to the twelve fp32 / residual variants it
// pushed gemm_tc_kernel over its 128-register limit (spills in every instantiation).
__device__ __noinline__ void epilogue16_call(uint16_t *out, long ldo, long M, int N, int BN, int act, int f16, uint32_t st_base, const float *bias_s,
                                             uint32_t acc, int warp, int lane, long m0, int n0) {
  float amax;
  if (f16) {
    amax = act == MUMPY_ACT_GELU   ? epilogue16_tile<__half, 1>(out, ldo, M, N, BN, act, TC_EPI_GROUPS, st_base, bias_s, acc, warp, lane, m0, n0)
           : act == MUMPY_ACT_NONE ? epilogue16_tile<__half, 0>(out, ldo, M, N, BN, act, TC_EPI_GROUPS, st_base, bias_s, acc, warp, lane, m0, n0)
                                   : epilogue16_tile<__half, 2>(out, ldo, M, N, BN, act, TC_EPI_GROUPS, st_base, bias_s, acc, warp, lane, m0, n0);
  } else {
    amax = act == MUMPY_ACT_GELU   ? epilogue16_tile<__nv_bfloat16, 1>(out, ldo, M, N, BN, act, TC_EPI_GROUPS, st_base, bias_s, acc, warp, lane, m0, n0)
           : act == MUMPY_ACT_NONE ? epilogue16_tile<__nv_bfloat16, 0>(out, ldo, M, N, BN, act, TC_EPI_GROUPS, st_base, bias_s, acc, warp, lane, m0, n0)
                                   : epilogue16_tile<__nv_bfloat16, 2>(out, ldo, M, N, BN, act, TC_EPI_GROUPS, st_base, bias_s, acc, warp, lane, m0, n0);
  }
  f16_guard(amax);
}

// FAF pass epilogue (kTS kernels): the DCT branch chains four GEMMs whose outputs are consumed TRANSPOSED per S x S image, band
// masked and split into [hi | hi | lo] 16-bit operands (faf.cu).  Thread <-> accumulator row m = (image, r): for a fixed column n
// the 32 lanes of a warp hold 32 consecutive r, so writing element (r, n) at transposed position (n, r) is a coalesced 64-byte
// store straight from registers -- the five repack kernels between the GEMMs become three 2-byte stores per element here.
// Geometry travels in the convolution fields of TcParams (unused by a plain linear; the struct must not grow): Wout = S,
// Hout = images per band, kw = bands written (1 or 3), lower_w / lower_h / cblocks = lo | hi << 16 of the three bands,
// act = MUMPY_FAF_FINAL for the last pass (fp32 result straight into the (B, 9, S, S) output).
constexpr int MUMPY_FAF_FINAL = 100;
template <typename T16>
__device__ __forceinline__ void faf_ts_epilogue(const TcParams &p, uint32_t acc, int warp, int lane, long m0, int n0, int groups) {
  const int quad = warp & 3, grp = warp >> 2;
  const uint32_t lane_addr = acc + (static_cast<uint32_t>(quad * 32) << 16);
  const long m = m0 + quad * 32 + lane;
  const int S = p.Wout, n_img = p.Hout, bands = p.kw & 0xff;          // (bit 8 of kw: band-limited INPUT, see faf_kblock_mask)
  const bool valid = m < p.M;
  const int img = valid ? static_cast<int>(m / S) : 0;
  const int r = valid ? static_cast<int>(m - (long)img * S) : 0;
  const bool final_pass = p.act == MUMPY_FAF_FINAL;
  float amax = 0.0f;
  // final pass: image (band, b, c) -> channel band * 3 + c of clip b
  const int f_band = img / n_img, f_im = img - f_band * n_img;
  float *out32 = reinterpret_cast<float *>(p.out) + (((long)(f_im / 3) * 9 + f_band * 3 + f_im % 3) * S) * S + r;
  T16 *out16 = reinterpret_cast<T16 *>(p.out) + ((long)img * S) * 3 * S + r;
  const long band_stride = (long)n_img * S * 3 * S;
  const int lo[3] = {p.lower_w & 0xffff, p.lower_h & 0xffff, p.cblocks & 0xffff};
  const int hi[3] = {p.lower_w >> 16, p.lower_h >> 16, p.cblocks >> 16};
  for (int c0 = grp * 32; c0 < p.BN; c0 += 32 * groups) {
    uint32_t v[32];
    tmem_ld32(lane_addr + c0, v);
    if (!valid) continue;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0 + c0 + j;
      if (c0 + j >= p.BN || n >= p.N) break;
      const float val = __uint_as_float(v[j]);
      if (final_pass) {
        out32[(long)n * S] = val;
      } else {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (q >= bands) break;
          const float mv = (bands > 1 && (r + n < lo[q] || r + n > hi[q])) ? 0.0f : val;
          if (is_half_t<T16>::value) amax = fmaxf(amax, fabsf(mv));
          const T16 h = from_f32<T16>(mv);
          const T16 l = from_f32<T16>(mv - to_f32(h));
          T16 *row = out16 + q * band_stride + (long)n * 3 * S;
          row[0] = h;
          row[S] = h;
          row[2 * S] = l;
        }
      }
    }
  }
  if (is_half_t<T16>::value) f16_guard(amax);
}

// FAF passes 3 and 4 (kTS kernels, TcParams::kw bit 8 set): the A operand's rows are (band, image, r) -- three bands of M / 3 rows --
// and band b's rows are zero outside columns [0, hi_b] of each of the three K segments ([hi | hi | lo], S columns each): the masked
// DCT coefficients of dct.py:66-68 vanish for k + k' > hi_b.  Returns the 64-column k-blocks that can hold non-zeros for the rows
// [m0, m0 + 128) (bit kb set = load and multiply it); a tile that straddles two bands takes the wider one's.
__device__ __forceinline__ uint32_t faf_kblock_mask(const TcParams &p, long m0, int nkb) {
  if (!(p.kw & 0x100) || nkb > 32) return 0xffffffffu;
  const long rpb = p.M / 3;
  long band = (m0 + TC_BM - 1) / rpb;
  if (band > 2) band = 2;
  const int hi = band == 0 ? (p.lower_w >> 16) : (band == 1 ? (p.lower_h >> 16) : (p.cblocks >> 16));
  const int S = p.Wout;
  uint32_t mask = 0;
  for (int kb = 0; kb < nkb; ++kb) {
    const int c0 = kb * TC_BK, c1 = c0 + TC_BK - 1;
    bool need = false;
    for (int seg = 0; seg < 3; ++seg) need = need || (c1 >= seg * S && c0 <= seg * S + hi);
    if (need) mask |= 1u << kb;
  }
  return mask;
}

// kPair: the two CTAs of a (2,1,1) cluster (one TPC) work on one 256 x BN tile with tcgen05.mma.cta_group::2: CTA r loads
// A rows [128 r, 128 r + 128) and B rows [BN/2 r, BN/2 r + BN/2) of the tile, the leader (rank 0) issues the M=256 MMAs,
// each CTA's TMEM receives its own 128 accumulator rows.  Per CTA and k-block that is (128 + BN/2) x 128 B from L2
// instead of (128 + BN) x 128 B -- the main loop of the 1-CTA kernel is bound by exactly that path (42.5 B/clk/SM).
// kSplit: split-K instantiation (conv only): kept apart so that the plain kernels carry none of its index arithmetic (the
// epilogue warps are at the register limit: one more live 64-bit value spills)
// kLean: 16-bit output without residual (qkv, fc1 + GELU, pre, k/v).  These GEMMs are epilogue-bound (per 128 x 256 tile the GELU
// epilogue issues ~10 k warp instructions at an IPC of ~1.7: ~7 k clk against a 5 k clk main loop at K = 512, 2.4 k at K = 256), so
// the variant carries ONLY the lean epilogue (tc_epilogue.cuh), which fits 112 registers, and runs 16 epilogue warps instead of
// 12: four per TMEM lane quadrant, i.e. exactly two 32-column chunks per warp of a 256-wide tile instead of 3 / 3 / 2.
template <bool kConv, bool kPair, bool kSplit = false, bool kLean = false, bool kTS = false>
__global__ void __launch_bounds__(kLean ? TC_THREADS_LEAN : TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                                         const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  constexpr int EW = kLean ? TC_EPI_WARPS_LEAN : TC_EPI_WARPS;        // epilogue warps; warp EW = TMA producer, EW + 1 = MMA issuer
  constexpr int EG = EW / 4;
  constexpr int EST = kLean ? EPI16_STAGING : EPI_STAGE_BYTES;       // per-warp staging tile
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * TC_MAX_STAGES + 4];
  __shared__ uint32_t tmem_slot;

#ifdef GEMM_TIMING
  const long long gt_entry = clock64();
#endif
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t group = kPair ? blockIdx.x >> 1 : blockIdx.x;          // tile-scheduler slot
  const uint32_t n_groups = kPair ? gridDim.x >> 1 : gridDim.x;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;        // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t a_bytes = TC_BM * 128;
  const uint32_t b_rows = static_cast<uint32_t>(kPair ? p.BN / 2 : p.BN);
  const uint32_t b_bytes = b_rows * 128;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t full0 = smem_u32(&bars[0]);
  const uint32_t empty0 = smem_u32(&bars[TC_MAX_STAGES]);
  const uint32_t acc_full0 = smem_u32(&bars[2 * TC_MAX_STAGES]);
  const uint32_t acc_empty0 = smem_u32(&bars[2 * TC_MAX_STAGES + 2]);
  const int nkb = kConv ? p.K : (p.K + TC_BK - 1) / TC_BK;      // conv: p.K already counts k-blocks (taps x cblocks)

  if (warp == EW && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    // pair mode: `full` and `acc_empty` live on the leader and collect arrivals from both CTAs
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, kPair ? 2 : 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full0 + 8 * s, 1);
      mbar_init(acc_empty0 + 8 * s, kPair ? 2 * EW : EW);
    }
    fence_barrier_init();
  }
  if (warp == EW + 1) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(2 * p.acc_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(2 * p.acc_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the tail of the previous kernel in the
  // stream; from here on global memory written by it is read
  pdl_grid_sync();

  if (warp == EW) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    // (the whole warp runs the loop converged, one elected lane issues: see elect_one())
    {
      uint32_t s = 0, ph = 0;
      const uint32_t full_leader = kPair ? mapa_shared(full0, 0) : full0;      // shared::cluster address on rank 0
      for (long tile_s = group; tile_s < p.num_tiles; tile_s += n_groups) {
        const long tile = kSplit ? tile_s / p.k_splits : tile_s;
        const int ks = kSplit ? static_cast<int>(tile_s - tile * p.k_splits) : 0;
        const int kb0 = kSplit ? static_cast<int>((long)ks * nkb / p.k_splits) : 0;
        const int kb1 = kSplit ? static_cast<int>((long)(ks + 1) * nkb / p.k_splits) : nkb;
        const int n0 = static_cast<int>(tile % p.tiles_n) * p.BN + static_cast<int>(rank * b_rows);
        const long m0 = (tile / p.tiles_n) * (kPair ? 2 * TC_BM : TC_BM) + rank * TC_BM;
        int cw = 0, ch = 0, cn = 0;
        if (kConv) {
          cw = static_cast<int>(m0 % p.Wout) + p.lower_w;
          const long r = m0 / p.Wout;
          ch = static_cast<int>(r % p.Hout) + p.lower_h;
          cn = static_cast<int>(r / p.Hout);
        }
        int tap = kConv ? kb0 / p.cblocks : 0, cb = kConv ? kb0 - tap * p.cblocks : 0;
        const uint32_t kmask = kTS ? faf_kblock_mask(p, m0, nkb) : 0xffffffffu;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (kTS && !((kmask >> (kb & 31)) & 1u)) continue;          // an all-zero k-block of a band-limited FAF operand
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t sa = tiles + s * stage_bytes;
          if (elect_one()) {
            if (kPair) {
              mbar_arrive_expect_tx_cluster(full_leader + 8 * s, stage_bytes);
              if (kConv) tma_load_im2col_4d_pair(sa, &tmA, full_leader + 8 * s, cb * TC_BK, cw, ch, cn, static_cast<uint16_t>(tap % p.kw), static_cast<uint16_t>(tap / p.kw));
              else tma_load_2d_pair(sa, &tmA, full_leader + 8 * s, kb * TC_BK, static_cast<int>(m0));
              tma_load_2d_pair(sa + a_bytes, &tmB, full_leader + 8 * s, kb * TC_BK, n0);
            } else {
              mbar_arrive_expect_tx(full0 + 8 * s, stage_bytes);
              if (kConv) tma_load_im2col_4d(sa, &tmA, full0 + 8 * s, cb * TC_BK, cw, ch, cn, static_cast<uint16_t>(tap % p.kw), static_cast<uint16_t>(tap / p.kw));
              else tma_load_2d(sa, &tmA, full0 + 8 * s, kb * TC_BK, static_cast<int>(m0));
              tma_load_2d(sa + a_bytes, &tmB, full0 + 8 * s, kb * TC_BK, n0);
            }
          }
          __syncwarp();
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
          if (kConv && ++cb == p.cblocks) { cb = 0; ++tap; }
        }
      }
    }
  } else if (warp == EW + 1) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    if (rank == 0) {
      uint32_t s = 0, ph = 0, t = 0;
#ifdef GEMM_TIMING
      const long long gt_ready = clock64();
      long long gt_tile[8][3];
#endif
      for (long tile_s = group; tile_s < p.num_tiles; tile_s += n_groups, ++t) {
#ifdef GEMM_TIMING
        if (t < 8) gt_tile[t][0] = clock64();
#endif
        const int ks = kSplit ? static_cast<int>(tile_s % p.k_splits) : 0;
        const int kb0 = kSplit ? static_cast<int>((long)ks * nkb / p.k_splits) : 0;
        const int kb1 = kSplit ? static_cast<int>((long)(ks + 1) * nkb / p.k_splits) : nkb;
        const uint32_t slot = t & 1, aph = (t >> 1) & 1;
        mbar_wait(acc_empty0 + 8 * slot, aph ^ 1);          // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + slot * p.acc_cols;
#ifdef GEMM_TIMING
        if (t < 8) gt_tile[t][1] = clock64();
#endif
        const uint32_t kmask = kTS ? faf_kblock_mask(p, (tile_s / p.tiles_n) * TC_BM, nkb) : 0xffffffffu;
        bool first_kb = true;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (kTS && !((kmask >> (kb & 31)) & 1u)) continue;          // (the producer skips the same k-blocks)
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t sa = tiles + s * stage_bytes;
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = make_kmajor_sw128_desc(sa + a_bytes);
          const bool accumulate = kTS ? !first_kb : kb > kb0;
          first_kb = false;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              // advance 16 bf16 = 32 B along K inside the swizzle span: +2 in the (addr>>4) field
              if (kPair) umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, p.idesc, (accumulate || k > 0) ? 1u : 0u);
              else umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, p.idesc, (accumulate || k > 0) ? 1u : 0u);
            }
            if (kPair) umma_commit_pair(empty0 + 8 * s); else umma_commit(empty0 + 8 * s);      // frees the smem slot(s) once these MMAs have read them
          }
          __syncwarp();
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) {
          if (kPair) umma_commit_pair(acc_full0 + 8 * slot); else umma_commit(acc_full0 + 8 * slot);  // accumulator complete
        }
        __syncwarp();
#ifdef GEMM_TIMING
        if (t < 8) gt_tile[t][2] = clock64();
#endif
      }
#ifdef GEMM_TIMING
      if ((blockIdx.x == 0 || blockIdx.x == 77) && lane == 0) {
        printf("cta %d M=%ld N=%d K=%d BN=%d stages=%d: setup %lld clk, tiles:", blockIdx.x, p.M, p.N, p.K, p.BN, p.stages, gt_ready - gt_entry);
        for (uint32_t i = 0; i < t && i < 8; ++i) printf(" [wait_acc %lld mainloop %lld | start %lld]", gt_tile[i][1] - gt_tile[i][0], gt_tile[i][2] - gt_tile[i][1], gt_tile[i][0] - gt_entry);
        printf(" end %lld\n", clock64() - gt_entry);
      }
#endif
    }
  } else {
    // ---------------- epilogue ----------------
    uint8_t *stage = smem_raw + (tiles - raw) + p.stages * stage_bytes + warp * EST;
    float *bias_s = reinterpret_cast<float *>(smem_raw + (tiles - raw) + p.stages * stage_bytes + EW * EST) + warp * 128;
    const int mode = (p.act == MUMPY_ACT_GELU ? 1 : (p.act == MUMPY_ACT_NONE ? 0 : 2)) | (p.out_bf16 ? 4 : 0) | (p.residual ? 8 : 0);
    uint32_t t = 0;
#ifdef GEMM_TIMING
    long long ge_tile[8][3];
#endif
    const uint32_t acc_empty_leader = kPair ? mapa_shared(acc_empty0, 0) : acc_empty0;
    for (long tile_s = group; tile_s < p.num_tiles; tile_s += n_groups, ++t) {
      const long tile = kSplit ? tile_s / p.k_splits : tile_s;
      const long out_shift = kSplit ? (tile_s - tile * p.k_splits) * (p.M * p.N) : 0;
      const uint32_t slot = t & 1, aph = (t >> 1) & 1;
      const int n0 = static_cast<int>(tile % p.tiles_n) * p.BN;
      const long m0 = (tile / p.tiles_n) * (kPair ? 2 * TC_BM : TC_BM) + rank * TC_BM;
      // bias values of this warp's column chunks -> shared memory while the accumulator is still being produced
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c = (warp >> 2) * 32 + 32 * EG * ci + lane;
        bias_s[ci * 32 + lane] = (p.bias && c < p.BN && n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.0f;
      }
      __syncwarp();
      // pull this warp's part of the fp32 residual tile into L2 while the accumulator is still being produced: one bulk
      // prefetch per row (the epilogue's residual loads are otherwise DRAM-latency bound, 4 dependent chunks per tile)
      if (p.residual) {
        const long gm = m0 + (warp & 3) * 32 + lane;
        const int cols = min(p.BN, p.N - n0), off = (warp >> 2) * 32;     // each warp group pulls 32-column slices, strided like its chunks
        if (gm < p.M && off < cols && (warp >> 2) == 0) {
          const uint32_t bytes = static_cast<uint32_t>(cols) * 4u;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.residual + gm * p.ldo + n0), "r"(bytes) : "memory");
        }
      }
#ifdef GEMM_TIMING
      if (t < 8) ge_tile[t][0] = clock64();
#endif
      mbar_wait(acc_full0 + 8 * slot, aph);
      tc_fence_after();
#ifdef GEMM_TIMING
      if (t < 8) ge_tile[t][1] = clock64();
#endif
      const uint32_t acc = tmem_base + slot * p.acc_cols;
      if constexpr (kTS) {
        if (p.f16) faf_ts_epilogue<__half>(p, acc, warp, lane, m0, n0, EG);
        else faf_ts_epilogue<__nv_bfloat16>(p, acc, warp, lane, m0, n0, EG);
      } else if constexpr (kLean) {
        float amax;
        if (p.f16) {
          amax = p.act == MUMPY_ACT_GELU ? epilogue16_tile<__half, 1>(reinterpret_cast<uint16_t *>(p.out), p.ldo, p.M, p.N, p.BN, p.act, EG, smem_u32(stage), bias_s, acc, warp, lane, m0, n0)
                                         : epilogue16_tile<__half, 0>(reinterpret_cast<uint16_t *>(p.out), p.ldo, p.M, p.N, p.BN, p.act, EG, smem_u32(stage), bias_s, acc, warp, lane, m0, n0);
        } else {
          amax = p.act == MUMPY_ACT_GELU ? epilogue16_tile<__nv_bfloat16, 1>(reinterpret_cast<uint16_t *>(p.out), p.ldo, p.M, p.N, p.BN, p.act, EG, smem_u32(stage), bias_s, acc, warp, lane, m0, n0)
                                         : epilogue16_tile<__nv_bfloat16, 0>(reinterpret_cast<uint16_t *>(p.out), p.ldo, p.M, p.N, p.BN, p.act, EG, smem_u32(stage), bias_s, acc, warp, lane, m0, n0);
        }
        f16_guard(amax);
      } else
      switch (mode) {
        case 0: epilogue_tile<0, false, false>(p, stage, bias_s, acc, warp, lane, m0, n0, out_shift); break;
        case 1: epilogue_tile<1, false, false>(p, stage, bias_s, acc, warp, lane, m0, n0, out_shift); break;
        case 2: epilogue_tile<2, false, false>(p, stage, bias_s, acc, warp, lane, m0, n0, out_shift); break;
        // 16-bit output, no residual (qkv, fc1 + GELU, pre, k/v, operand-precision convolutions): the lean epilogue of tc_epilogue.cuh
        case 4: case 5: case 6:
          epilogue16_call(reinterpret_cast<uint16_t *>(p.out) + out_shift, p.ldo, p.M, p.N, p.BN, p.act, p.f16, smem_u32(stage), bias_s, acc, warp, lane, m0, n0);
          break;
        case 8: epilogue_tile<0, false, true>(p, stage, bias_s, acc, warp, lane, m0, n0, out_shift); break;
        case 9: epilogue_tile<1, false, true>(p, stage, bias_s, acc, warp, lane, m0, n0, out_shift); break;
        case 10: epilogue_tile<2, false, true>(p, stage, bias_s, acc, warp, lane, m0, n0, out_shift); break;
        case 12: epilogue_tile<0, true, true>(p, stage, bias_s, acc, warp, lane, m0, n0, out_shift); break;
        case 13: epilogue_tile<1, true, true>(p, stage, bias_s, acc, warp, lane, m0, n0, out_shift); break;
        default: epilogue_tile<2, true, true>(p, stage, bias_s, acc, warp, lane, m0, n0, out_shift); break;
      }
      // this warp is done reading the accumulator: hand the TMEM slot back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(acc_empty_leader + 8 * slot); else mbar_arrive(acc_empty0 + 8 * slot);
      }
#ifdef GEMM_TIMING
      if (t < 8) ge_tile[t][2] = clock64();
#endif
    }
#ifdef GEMM_TIMING
    if (blockIdx.x == 0 && (warp == 0 || warp == 11) && lane == 0) {
      printf("cta 0 epilogue warp %d:", warp);
      for (uint32_t i = 0; i < t && i < 8; ++i) printf(" [wait %lld epi %lld | at %lld]", ge_tile[i][1] - ge_tile[i][0], ge_tile[i][2] - ge_tile[i][1], ge_tile[i][1] - gt_entry);
      printf("\n");
    }
#endif
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();     // nobody leaves while the partner can still touch its barriers / tiles
  if (warp == EW + 1) {
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * p.acc_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * p.acc_cols) : "memory");
  }
}

static int encode_2d_bf16(CUtensorMap *map, const void *ptr, bool f16, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                          uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu stride=%llu box=%ux%u", (int)r, ptr,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_elems, box_inner, box_outer);
    return MUMPY_ERR_CUDA;
  }
  return MUMPY_OK;
}

// shared with gemm_ln_tcgen05.cu
int tc_encode_2d_16(CUtensorMap *map, const void *ptr, bool f16, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                    uint32_t box_inner, uint32_t box_outer) {
  return encode_2d_bf16(map, ptr, f16, inner, outer, row_stride_elems, box_inner, box_outer);
}
int tc_num_sms() { return g_num_sms > 0 ? g_num_sms : 148; }

// development knobs (environment, read once): MUMPY_TC_BN / MUMPY_TC_STAGES override the tile heuristics,
// MUMPY_TC_DEBUG=1 skips the epilogue's global stores, =2 also skips its math (timing attribution only).
static int env_int(const char *name) {
  const char *v = getenv(name);
  return v ? atoi(v) : 0;
}
static int g_dbg_bn = -1, g_dbg_stages = -1, g_dbg_mode = -1;

// Tile shape: minimise  waves * max(mainloop, epilogue) + min(mainloop, epilogue)  over the divisors of N and over the
// 1-CTA (128 x BN) / CTA-pair (256 x BN) variants (accumulators are double buffered, so the epilogue of one tile overlaps the
// main loop of the next).  Calibration (round-1 measurements, gpurun_out/gemm_knobs*.txt, prof_gemm_fc*_v5): the main loop
// is bound by the L2->SM path, ~42.5 B/clk per SM: a CTA pulls (128 + rows of B it loads) * 128 B per k-block, i.e.
// (128 + BN) * 1.6 ns alone or (128 + BN/2) * 1.6 ns in a pair, never faster than the tensor pipe (BN * 1.06 ns); an
// epilogue pass over 32 columns costs a warp ~0.55 us (+0.25 us with an fp32 residual to fetch, +0.4 us for GELU); a warp
// takes every third chunk.  Checked against a sweep of all legal widths on the 26 heaviest shapes of the step
// (tools/gemm_bn_sweep.py): the model's choices cost 1188.5 us in total vs 1188.4 us for the per-shape optimum.
struct TileChoice {
  int bn;
  bool pair;
};
// CTA-pair policy: 0 never, 1 the cost model decides, 2 whenever legal, 3 cost model for K >= 1024 only, 4 (default) cost model
// except that 16-bit outputs without residual keep the lean 16-warp 1-CTA kernel (the pair kernel has no lean form).  Same-box
// A/B runs of the whole batch-32 step (tools/ab_env.sh, profiles/r2_ab_runs.txt): 0: 2398, 1: 2418-2449, 2: 2430, 3: 2402,
// 4: 2436-2437 against 2423-2432 for 1 on that box -- although the pairs rarely win in isolation (tools/gemm_bn_sweep.py) and the
// serial sum of kernel times is higher with 1: each CTA of a pair loads half of every weight tile, which pays when three lanes
// share L2.  Environment MUMPY_TC_PAIR or mumpy_set_gemm_pair_mode().
static int g_dbg_pair = -1;

static TileChoice pick_tile(long M, int N, int nkb, bool has_res, bool gelu, bool lean_ok = false) {
  if (g_dbg_bn < 0) {
    g_dbg_bn = env_int("MUMPY_TC_BN");
  }
  if (g_dbg_pair < 0) {
    const char *v = getenv("MUMPY_TC_PAIR");
    g_dbg_pair = v ? atoi(v) : 4;
  }
  static const int cands[] = {256, 192, 128, 96, 64, 48, 32, 16};
  const double t_chunk = 550.0 + (has_res ? 250.0 : 0.0) + (gelu ? 400.0 : 0.0);
  TileChoice best = {0, false};
  double best_cost = 1e30;
  // N values whose only divisors among the candidates are narrow (e.g. 224 = 7 x 32, the DCT passes) may also take a wide
  // tile with a masked tail: TMA zero-fills the weight rows beyond N and the epilogue masks the columns
  int widest_div = 0;
  for (int c : cands)
    if (N % c == 0 && c > widest_div) widest_div = c;
  for (int pair = 0; pair < 2; ++pair) {
    if (pair && (g_dbg_pair == 0 || g_num_sms < 2)) continue;
    if (pair && g_dbg_pair == 3 && nkb < 16) continue;          // 3: cost model, long reductions (K >= 1024) only
    if (pair && g_dbg_pair == 4 && lean_ok) continue;           // 4: cost model, but the 16-bit outputs keep the lean 1-CTA kernel
    const long mt = cdiv(M, pair ? 2 * TC_BM : TC_BM);
    const long slots = pair ? g_num_sms / 2 : g_num_sms;
    for (int c : cands) {                            // descending: ties go to the wider tile
      const bool divides = N % c == 0;
      if (!divides && !(widest_div < 64 && c >= 64 && c < N + 64)) continue;
      if (g_dbg_bn > 0 && N % g_dbg_bn == 0 && c != g_dbg_bn) continue;
      if (pair && c % 16 != 0) continue;
      const long tiles = mt * cdiv(N, c);
      const double waves = (double)cdiv(tiles, slots);
      const double load = (128.0 + (pair ? c / 2 : c)) * 1.6, pipe = c * 1.06;
      const double mma = (double)nkb * (load > pipe ? load : pipe) + (pair ? 500.0 : 300.0);
      const double epi = (double)((c + 32 * TC_EPI_GROUPS - 1) / (32 * TC_EPI_GROUPS)) * t_chunk;      // chunks per epilogue warp
      double cost = waves * (mma > epi ? mma : epi) + (mma > epi ? epi : mma);
      if (pair && g_dbg_pair == 2) cost *= 1e-3;
      if (cost < best_cost * 0.98) {
        best_cost = cost;
        best.bn = c;
        best.pair = pair != 0;
      }
    }
  }
  if (best.bn) return best;
  for (int c : cands)
    if (c <= N) return TileChoice{c, false};         // no divisor: the tail tile is masked
  return TileChoice{16, false};
}

void set_gemm_tile_override(int bn) { g_dbg_bn = bn < 0 ? 0 : bn; }
void set_gemm_pair_mode(int mode) { g_dbg_pair = mode < 0 ? 0 : (mode > 4 ? 4 : mode); }

template <typename... KArgs, typename... Args>
static void launch_pair_kernel(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
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
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static bool g_attr_set[8] = {false, false, false, false, false, false, false, false};
static int g_dbg_lean = -1;      // MUMPY_TC_LEAN=0 keeps the 12-warp kernel for 16-bit outputs (A/B runs)

static int launch_tc(const CUtensorMap &tmA, const CUtensorMap &tmB, TcParams &p, bool pair, cudaStream_t st) {
  uint32_t cols = 32;
  while (cols < (uint32_t)p.BN) cols <<= 1;
  p.acc_cols = cols;
  const int bm = pair ? 2 * TC_BM : TC_BM;
  p.idesc = make_idesc_16_f32(bm, p.BN, p.f16 != 0);
  p.tiles_n = (int)cdiv(p.N, p.BN);
  if (p.k_splits < 1) p.k_splits = 1;
  p.num_tiles = cdiv(p.M, bm) * p.tiles_n * p.k_splits;
  if (g_dbg_stages < 0) {
    const char *lv = getenv("MUMPY_TC_LEAN");
    g_dbg_lean = lv ? atoi(lv) : 1;
    g_dbg_stages = env_int("MUMPY_TC_STAGES");
    g_dbg_mode = env_int("MUMPY_TC_DEBUG");
  }
  p.debug = (short)g_dbg_mode;
  const int stage_bytes = TC_BM * 128 + (pair ? p.BN / 2 : p.BN) * 128;
  const int nkb = p.conv ? p.K : (p.K + TC_BK - 1) / TC_BK;
  const long slots = pair ? g_num_sms / 2 : g_num_sms;
  int stages = TC_SMEM_BUDGET / stage_bytes;
  if (g_dbg_stages > 0 && g_dbg_stages < stages) stages = g_dbg_stages;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  const long kb_per_cta = (nkb / p.k_splits) * cdiv(p.num_tiles, slots);
  if (stages > kb_per_cta) stages = (int)kb_per_cta;
  if (stages < 1) stages = 1;
  p.stages = stages;
  // lean variant: plain linear, 16-bit output, no residual / side output, activation none or GELU
  const bool lean = !p.conv && p.Wout == 0 && p.k_splits == 1 && p.out_bf16 && !p.residual && !p.aux && (p.act == MUMPY_ACT_NONE || p.act == MUMPY_ACT_GELU) &&
                    g_dbg_lean != 0;
  const int epi_smem = lean ? TC_EPI_WARPS_LEAN * (EPI16_STAGING + 512) : TC_EPI_WARPS * (32 * 128 + 512);
  const int smem = stages * stage_bytes + 1024 + epi_smem;
  const bool ts = !p.conv && p.Wout > 0;      // FAF pass: transposed / split store (geometry in the convolution fields)
  const int which = ts ? 6 : lean ? (pair ? 7 : 5) : p.k_splits > 1 ? 4 : (p.conv ? 1 : 0) + (pair ? 2 : 0);      // split-K: conv, single-CTA tiles only
  if (!g_attr_set[which]) {
    const int max_smem = TC_SMEM_BUDGET + 1024 + TC_EPI_WARPS * (32 * 128 + 512);
    cudaError_t e;
    switch (which) {
      case 0: e = cudaFuncSetAttribute(gemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); break;
      case 1: e = cudaFuncSetAttribute(gemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); break;
      case 2: e = cudaFuncSetAttribute(gemm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); break;
      case 3: e = cudaFuncSetAttribute(gemm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); break;
      case 4: e = cudaFuncSetAttribute(gemm_tc_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); break;
      case 5: e = cudaFuncSetAttribute(gemm_tc_kernel<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); break;
      case 7: e = cudaFuncSetAttribute(gemm_tc_kernel<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); break;
      default: e = cudaFuncSetAttribute(gemm_tc_kernel<false, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); break;
    }
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_tc_kernel): %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    g_attr_set[which] = true;
  }
  const unsigned groups = (unsigned)(p.num_tiles < slots ? p.num_tiles : slots);
  switch (which) {
    case 0: launch_kernel(gemm_tc_kernel<false, false>, groups, TC_THREADS, smem, st, tmA, tmB, p); break;
    case 1: launch_kernel(gemm_tc_kernel<true, false>, groups, TC_THREADS, smem, st, tmA, tmB, p); break;
    case 2: launch_pair_kernel(gemm_tc_kernel<false, true>, 2 * groups, TC_THREADS, smem, st, tmA, tmB, p); break;
    case 3: launch_pair_kernel(gemm_tc_kernel<true, true>, 2 * groups, TC_THREADS, smem, st, tmA, tmB, p); break;
    case 4: launch_kernel(gemm_tc_kernel<true, false, true>, groups, TC_THREADS, smem, st, tmA, tmB, p); break;
    case 5: launch_kernel(gemm_tc_kernel<false, false, false, true>, groups, TC_THREADS_LEAN, smem, st, tmA, tmB, p); break;
    case 7: launch_pair_kernel(gemm_tc_kernel<false, true, false, true>, 2 * groups, TC_THREADS_LEAN, smem, st, tmA, tmB, p); break;
    default: launch_kernel(gemm_tc_kernel<false, false, false, false, true>, groups, TC_THREADS, smem, st, tmA, tmB, p); break;
  }
  return launch_status("gemm_tc_kernel");
}

int linear_bf16(const void *A, long lda, const void *W, const float *bias, const float *residual, void *out, void *aux, long ldo,
                long M, int N, int K, int ab_dtype, int out_dtype, int act, cudaStream_t st) {
  int rc = resolve_driver_entry_points();
  if (rc) return rc;
  MUMPY_REQUIRE(N % 8 == 0 && K % 8 == 0 && lda % 8 == 0, "linear(bf16): N, K, lda must be multiples of 8 (N=%d K=%d lda=%ld)", N, K, lda);
  MUMPY_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "linear(bf16): A, W, out must be 16-byte aligned");
  MUMPY_REQUIRE(out_dtype == MUMPY_F32 || out_dtype == ab_dtype, "linear(16-bit): the output is fp32 or the operand type");
  MUMPY_REQUIRE(out_dtype != MUMPY_F32 ? (ldo % 8 == 0) : (ldo % 4 == 0), "linear(bf16): ldo alignment");
  MUMPY_REQUIRE(M < (1l << 31), "linear(bf16): M too large");
  MUMPY_REQUIRE(!aux || (out_dtype == MUMPY_F32 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0 && ldo % 8 == 0),
                "linear(bf16): the bf16 side output needs an fp32 main output, 16-byte alignment and ldo %% 8 == 0");
  TcParams p = {};
  p.bias = bias;
  p.residual = residual;
  p.out = out;
  p.aux = static_cast<uint16_t *>(aux);
  p.f16 = ab_dtype == MUMPY_F16;
  p.ldo = ldo;
  p.M = M;
  p.N = N;
  p.K = K;
  const bool lean_ok = out_dtype != MUMPY_F32 && !residual && !aux && (act == MUMPY_ACT_NONE || act == MUMPY_ACT_GELU);
  const TileChoice tc = pick_tile(M, N, (K + TC_BK - 1) / TC_BK, residual != nullptr, act == MUMPY_ACT_GELU, lean_ok);
  p.BN = tc.bn;
  p.act = act;
  p.out_bf16 = (out_dtype != MUMPY_F32);
  p.conv = 0;
  CUtensorMap tmA, tmB;
  rc = encode_2d_bf16(&tmA, A, p.f16 != 0, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TC_BK, TC_BM);
  if (rc) return rc;
  rc = encode_2d_bf16(&tmB, W, p.f16 != 0, (uint64_t)K, (uint64_t)N, (uint64_t)K, TC_BK, (uint32_t)(tc.pair ? p.BN / 2 : p.BN));
  if (rc) return rc;
  return launch_tc(tmA, tmB, p, tc.pair, st);
}

// One pass of the tensor-core FAF (faf.cu): out(transposed per S x S image, band masked, split) = A . W^T.
// A (M = images * S, K) and W (S, K) 16-bit K-major.  final_pass: fp32 result into the (B, 9, S, S) output instead.
int linear_faf_pass(const void *A, const void *W, void *out, long M, int K, int S, int imgs_per_band, int bands, const int *lo_hi6, int final_pass,
                    int ab_dtype, cudaStream_t st, int in_sparse) {
  int rc = resolve_driver_entry_points();
  if (rc) return rc;
  MUMPY_REQUIRE(S % 8 == 0 && K % 8 == 0 && S < 65536 && M % S == 0, "faf pass: S, K multiples of 8 and M a multiple of S required");
  TcParams p = {};
  p.out = out;
  p.f16 = ab_dtype == MUMPY_F16;
  p.ldo = S;
  p.M = M;
  p.N = S;
  p.K = K;
  const TileChoice tc = pick_tile(M, S, (K + TC_BK - 1) / TC_BK, false, false);
  p.BN = tc.bn;
  p.act = final_pass ? MUMPY_FAF_FINAL : MUMPY_ACT_NONE;
  p.out_bf16 = final_pass ? 0 : 1;
  p.Wout = S;
  p.Hout = imgs_per_band;
  p.kw = bands | ((in_sparse && lo_hi6 && M % 3 == 0) ? 0x100 : 0);          // bit 8: A rows are (band, image, r), band-limited (faf_kblock_mask)
  p.lower_w = lo_hi6 ? (lo_hi6[0] | (lo_hi6[1] << 16)) : 0;
  p.lower_h = lo_hi6 ? (lo_hi6[2] | (lo_hi6[3] << 16)) : 0;
  p.cblocks = lo_hi6 ? (lo_hi6[4] | (lo_hi6[5] << 16)) : 0;
  CUtensorMap tmA, tmB;
  rc = encode_2d_bf16(&tmA, A, p.f16 != 0, (uint64_t)K, (uint64_t)M, (uint64_t)K, TC_BK, TC_BM);
  if (rc) return rc;
  rc = encode_2d_bf16(&tmB, W, p.f16 != 0, (uint64_t)K, (uint64_t)S, (uint64_t)K, TC_BK, (uint32_t)p.BN);
  if (rc) return rc;
  return launch_tc(tmA, tmB, p, false, st);
}

// Second half of a split-K GEMM: out = act(sum_s partial[s] + bias) (+ residual), four columns per thread.  The partial
// sums (k_splits x M x N fp32, written by gemm_tc_kernel's plain epilogue) are still in L2 when this runs.
template <typename OutT>
__global__ void splitk_reduce_kernel(const float *__restrict__ partial, int S, long stride, const float *__restrict__ bias,
                                     const float *__restrict__ residual, OutT *__restrict__ out, long ldo, long M, int N, int act) {
  pdl_grid_sync();
  const int n4 = N >> 2;
  const long total = M * n4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long m = i / n4;
    const int c = (int)(i - m * n4) << 2;
    float4 a = *reinterpret_cast<const float4 *>(partial + m * N + c);
    for (int s = 1; s < S; ++s) {
      const float4 b = *reinterpret_cast<const float4 *>(partial + s * stride + m * N + c);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (bias) {
      const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (act == MUMPY_ACT_GELU) {
      const float2 lo = gelu_fast2(make_float2(a.x, a.y)), hi = gelu_fast2(make_float2(a.z, a.w));
      a = make_float4(lo.x, lo.y, hi.x, hi.y);
    } else if (act != MUMPY_ACT_NONE) {
      a.x = apply_act(a.x, act); a.y = apply_act(a.y, act); a.z = apply_act(a.z, act); a.w = apply_act(a.w, act);
    }
    if (residual) {
      const float4 r = *reinterpret_cast<const float4 *>(residual + m * ldo + c);
      a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
    }
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + m * ldo + c) = a;
    } else {
      if (is_half_t<OutT>::value) f16_guard(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
      uint2 u;
      u.x = pack2<OutT>(a.x, a.y);
      u.y = pack2<OutT>(a.z, a.w);
      *reinterpret_cast<uint2 *>(out + m * ldo + c) = u;
    }
  }
}

// Implicit-GEMM convolution (stride 1): in (B,H,W,Cin) bf16 NHWC with pixel stride ld_in elements; wpk (Cout, taps*cblocks*64)
// bf16 with K order (ky,kx,c) and every tap's channels zero padded to a multiple of 64; out (B*H*W, Cout).
// splitk_ws / splitk_ws_bytes: optional caller-owned fp32 workspace; when the tile grid would leave most SMs idle (small
// maps with a long reduction, e.g. 32 x 7 x 7 pixels and K = 7 * 2560) the reduction is split across CTAs and reduced by a
// second kernel.
int conv_bf16(const void *in, long ld_in, const void *wpk, const float *bias, const float *residual, void *out, long ldo, int B,
              int H, int W, int Cin, int Cout, int kh, int kw, int ph, int pw, int in_dtype, int out_dtype, int act, float *splitk_ws,
              long splitk_ws_bytes, cudaStream_t st) {
  int rc = resolve_driver_entry_points();
  if (rc) return rc;
  MUMPY_REQUIRE(Cout % 8 == 0 && ld_in % 8 == 0, "conv(bf16): Cout and ld_in must be multiples of 8");
  MUMPY_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(wpk) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "conv(bf16): in, w, out must be 16-byte aligned");
  MUMPY_REQUIRE(2 * ph == kh - 1 && 2 * pw == kw - 1, "conv(bf16): only 'same' stride-1 convolutions");
  const int cblocks = (Cin + TC_BK - 1) / TC_BK;
  TcParams p = {};
  p.bias = bias;
  p.residual = residual;
  p.out = out;
  p.ldo = ldo;
  p.M = (long)B * H * W;
  p.N = Cout;
  p.K = kh * kw * cblocks;          // k-blocks
  TileChoice tc = pick_tile(p.M, Cout, p.K, residual != nullptr, act == MUMPY_ACT_GELU);
  p.BN = tc.bn;
  p.act = act;
  p.out_bf16 = (out_dtype != MUMPY_F32);
  p.f16 = in_dtype == MUMPY_F16;
  int splits = 1;
  if (splitk_ws && (reinterpret_cast<uintptr_t>(splitk_ws) & 15) == 0 && Cout % 4 == 0) {
    // widest tile (the A tile is fetched once per k-block), then as many k ranges as there are idle SMs, >= 8 k-blocks each
    const int bn = Cout <= 256 && Cout % 16 == 0 ? Cout : tc.bn;
    const long tiles = cdiv(p.M, TC_BM) * cdiv(Cout, bn);
    long s = g_num_sms / tiles;
    if (s > p.K / 8) s = p.K / 8;
    if (s > splitk_ws_bytes / (long)(p.M * Cout * sizeof(float))) s = splitk_ws_bytes / (long)(p.M * Cout * sizeof(float));
    if (s >= 2) {
      splits = (int)s;
      tc.bn = bn;
      tc.pair = false;
      p.BN = bn;
      p.bias = nullptr;
      p.residual = nullptr;
      p.act = MUMPY_ACT_NONE;
      p.out = splitk_ws;
      p.out_bf16 = 0;
      p.ldo = Cout;
      p.k_splits = (short)splits;
    }
  }
  p.conv = 1;
  p.Wout = W;
  p.Hout = H;
  p.lower_w = -pw;
  p.lower_h = -ph;
  p.kw = kw;
  p.cblocks = cblocks;
  CUtensorMap tmA, tmB;
  cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)ld_in * 2, (cuuint64_t)ld_in * 2 * W, (cuuint64_t)ld_in * 2 * W * H};
  int lower[2] = {-pw, -ph};
  int upper[2] = {pw - (kw - 1), ph - (kh - 1)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  MUMPY_REQUIRE(out_dtype == MUMPY_F32 || out_dtype == in_dtype, "conv(16-bit): the output is fp32 or the operand type");
  CUresult r = g_encode_im2col(&tmA, p.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(in), gdim, gstr, lower, upper, TC_BK, TC_BM,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (%d): B=%d H=%d W=%d Cin=%d ld=%ld k=%dx%d", (int)r, B, H, W, Cin, ld_in, kh, kw);
    return MUMPY_ERR_CUDA;
  }
  // same workaround CUTLASS applies (copy_traits_sm90_im2col.hpp) for small tensors on drivers <= 13.1
  if (g_driver_version <= 13010 && (size_t)B * H * W * ld_in * 2 < 131072) reinterpret_cast<uint64_t *>(&tmA)[1] &= ~(1ull << 21);
  rc = encode_2d_bf16(&tmB, wpk, p.f16 != 0, (uint64_t)p.K * TC_BK, (uint64_t)Cout, (uint64_t)p.K * TC_BK, TC_BK, (uint32_t)(tc.pair ? p.BN / 2 : p.BN));
  if (rc) return rc;
  rc = launch_tc(tmA, tmB, p, tc.pair, st);
  if (rc || splits == 1) return rc;
  const long total = p.M * (Cout / 4);
  const unsigned grid = (unsigned)(cdiv(total, 256) < 148l * 8 ? cdiv(total, 256) : 148l * 8);
  if (out_dtype == MUMPY_F32)
    launch_kernel(splitk_reduce_kernel<float>, grid, 256, 0, st, splitk_ws, splits, p.M * Cout, bias, residual, static_cast<float *>(out), ldo, p.M, Cout, act);
  else if (out_dtype == MUMPY_F16)
    launch_kernel(splitk_reduce_kernel<__half>, grid, 256, 0, st, splitk_ws, splits, p.M * Cout, bias, residual, static_cast<__half *>(out), ldo, p.M, Cout, act);
  else
    launch_kernel(splitk_reduce_kernel<__nv_bfloat16>, grid, 256, 0, st, splitk_ws, splits, p.M * Cout, bias, residual, static_cast<__nv_bfloat16 *>(out), ldo, p.M, Cout, act);
  return launch_status("splitk_reduce_kernel");
}

}  // namespace mumpy
