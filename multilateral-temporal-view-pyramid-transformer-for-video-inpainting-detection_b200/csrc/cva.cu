// Deformable cross-view attention (SwinDAttention, deformableAttention.py:324-405): offset network, bilinear
// sampling of the key/value windows and the raw-reshape residual.  All three are HBM/L2-bound gathers on the
// token canvases; the k/v/out projections run through mumpy_linear and the softmax core is in attention.cu.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace mumpy {

__device__ __forceinline__ int cva_query_window_s(int j, int r, int N1, int nW1, int per_clip) {
  if (!per_clip) return j % N1;
  const int i = j / r;
  const int clip = i / nW1;
  return clip * nW1 + ((i % nW1) * r + j % r) % nW1;
}

// One CTA per (query window, group); one warp per pixel; lanes span the group's channels.
// conv_offset = dw5x5(pad 2, within the window) -> LayerNorm(Cg) -> GELU -> 1x1 (Cg->2)   (:253-258)
// pix = ((tanh(o) * (1/ws) * 2 + ref) + 1) / 2 * (ws-1)                                    (:338-356)
template <int MAXC>   // channels per lane = Cg/32 <= MAXC
__global__ void __launch_bounds__(256) cva_offsets_kernel(const float *__restrict__ q, const float *__restrict__ dw_w,
                                                          const float *__restrict__ dw_b, const float *__restrict__ ln_g,
                                                          const float *__restrict__ ln_b, const float *__restrict__ pw,
                                                          float *__restrict__ pix, int N1, int TH1, int W, int C, int groups, int ws) {
  pdl_grid_sync();
  // [(ws+4)^2][Cg] query window with a zero border of 2 pixels (the depthwise conv's padding: no bounds tests in the tap loop),
  // then [25][Cg] depthwise taps (tap-major: conflict-free per lane), then [P][2] raw offsets
  extern __shared__ float tile[];
  const int P = ws * ws;
  const int Cg = C / groups;
  const int wp = ws + 4;
  float *wsm = tile + wp * wp * Cg;
  float *off = wsm + 25 * Cg;
  // Persistent over the windows of one group (blockIdx.y): the taps, the per-lane channel constants and the zero border of
  // the tile are set up once per CTA -- with a CTA per (window, group) that set-up cost more than the 49 x Cg x 25 FMAs.
  const int g = blockIdx.y;
  const int nW1 = (TH1 / ws) * (W / ws);
  const long L1 = (long)TH1 * W;
  const int cg4 = Cg >> 2;
  for (int e = threadIdx.x; e < wp * wp * cg4; e += blockDim.x) reinterpret_cast<float4 *>(tile)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = threadIdx.x; e < 25 * Cg; e += blockDim.x) {
    const int c = e / 25, tap = e - c * 25;
    wsm[tap * Cg + c] = __ldg(dw_w + e);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // per-lane channel constants (bias, LayerNorm affine, 1x1 weights); for narrow groups the 25 taps live in registers too
  float cb[MAXC], cgam[MAXC], cbet[MAXC], cwy[MAXC], cwx[MAXC];
  constexpr bool kTapRegs = MAXC <= 2;
  float wreg[kTapRegs ? MAXC : 1][25];
#pragma unroll
  for (int u = 0; u < MAXC; ++u) {
    const int c = lane + 32 * u;
    const bool ok = c < Cg;
    cb[u] = ok ? __ldg(dw_b + c) : 0.0f;
    cgam[u] = ok ? __ldg(ln_g + c) : 0.0f;
    cbet[u] = ok ? __ldg(ln_b + c) : 0.0f;
    cwy[u] = ok ? __ldg(pw + c) : 0.0f;
    cwx[u] = ok ? __ldg(pw + Cg + c) : 0.0f;
    if (kTapRegs) {
#pragma unroll
      for (int t = 0; t < 25; ++t) wreg[u][t] = ok ? wsm[t * Cg + c] : 0.0f;
    }
  }
  for (int win = blockIdx.x; win < N1; win += gridDim.x) {
  const int b = win / nW1, n = win % nW1;
  __syncthreads();                       // the previous window's readers are done with the tile and the offsets
  for (int e = threadIdx.x; e < P * cg4; e += blockDim.x) {      // interior pixels only: the border stays zero
    const int pp = e / cg4, c4 = e - pp * cg4;
    const long row = b * L1 + window_token_row(n, pp, TH1, W, ws, 0);
    reinterpret_cast<float4 *>(tile)[((pp / ws + 2) * wp + pp % ws + 2) * cg4 + c4] = __ldg(reinterpret_cast<const float4 *>(q + row * C + g * Cg) + c4);
  }
  __syncthreads();
  for (int p = warp; p < P; p += nwarps) {
    const int pi = p / ws, pj = p % ws;
    float val[MAXC];
    float s = 0.0f;
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int c = lane + 32 * u;
      float acc = 0.0f;
      if (c < Cg) {
        const float *t0 = tile + (pi * wp + pj) * Cg + c;      // top-left tap of the zero-padded window
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
          for (int bb = 0; bb < 5; ++bb)
            acc = fmaf(t0[(a * wp + bb) * Cg], kTapRegs ? wreg[u][a * 5 + bb] : wsm[(a * 5 + bb) * Cg + c], acc);
        acc += cb[u];
        s += acc;
      }
      val[u] = acc;
    }
    const float mean = warp_sum(s) / Cg;
    float v = 0.0f;
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      if (lane + 32 * u < Cg) {
        const float d = val[u] - mean;
        v = fmaf(d, d, v);
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(v) / Cg + 1e-5f);
    float oy = 0.0f, ox = 0.0f;
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      if (lane + 32 * u < Cg) {
        const float a = gelu_erf((val[u] - mean) * rstd * cgam[u] + cbet[u]);
        oy = fmaf(a, cwy[u], oy);
        ox = fmaf(a, cwx[u], ox);
      }
    }
    oy = warp_sum(oy);
    ox = warp_sum(ox);
    if (lane == 0) {
      off[2 * p] = oy;
      off[2 * p + 1] = ox;
    }
  }
  __syncthreads();
  // one thread per pixel: pix = ((tanh(o) * (1/ws) * 2 + ref) + 1) / 2 * (ws-1)
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int pi = p / ws, pj = p % ws;
    const float inv = 1.0f / (float)ws;
    const float ref_y = ((pi + 0.5f) / ws) * 2.0f - 1.0f;
    const float ref_x = ((pj + 0.5f) / ws) * 2.0f - 1.0f;
    const float pos_y = tanhf(off[2 * p]) * inv * 2.0f + ref_y;
    const float pos_x = tanhf(off[2 * p + 1]) * inv * 2.0f + ref_x;
    float *o = pix + (((long)win * groups + g) * P + p) * 2;
    o[0] = ((pos_y + 1.0f) / 2.0f) * (ws - 1);
    o[1] = ((pos_x + 1.0f) / 2.0f) * (ws - 1);
  }
  }
}

// Register-resident variant (the one that runs for window sizes 7 and 8): a WARP per (window, group), persistent.
//   phase A, lane = channel: the channel's WS x WS window lives in registers, the depthwise 5x5 is fully unrolled with the
//            out-of-window taps dropped at compile time (841 of 1225 FMAs for WS = 7, no shared-memory reads, no address
//            arithmetic); conv outputs go to a [pixel][channel] tile in shared memory;
//   phase B, lane = pixel: LayerNorm over the channels, GELU, the 1x1 (Cg -> 2), tanh and the pixel-coordinate map -- no
//            shuffles, two rounds for 49 pixels.
// Against the CTA-per-unit kernel above (a warp per pixel: 25 LDS + 25 FMA + ~50 integer instructions + 20 shuffles per
// pixel) this is ~4x fewer instructions; the op sits on the critical path at the start of every stage.
template <int WS>
__global__ void __launch_bounds__(128, 2) cva_offsets_reg_kernel(const float *__restrict__ q, const float *__restrict__ dw_w,
                                                              const float *__restrict__ dw_b, const float *__restrict__ ln_g,
                                                              const float *__restrict__ ln_b, const float *__restrict__ pw,
                                                              float *__restrict__ pix, int N1, int TH1, int W, int C, int groups) {
  constexpr int P = WS * WS;
  extern __shared__ float cr_smem[];
  const int Cg = C / groups;
  const int pitch = Cg + 1;                              // [pixel][channel] tile, odd pitch: conflict-free in both phases
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *prm = cr_smem;                                  // [4][Cg]: LayerNorm gamma, beta, 1x1 weights (y, x)
  float *tile = cr_smem + 4 * Cg + warp * (P * pitch);
  for (int i = threadIdx.x; i < Cg; i += blockDim.x) {
    prm[i] = __ldg(ln_g + i);
    prm[Cg + i] = __ldg(ln_b + i);
    prm[2 * Cg + i] = __ldg(pw + i);
    prm[3 * Cg + i] = __ldg(pw + Cg + i);
  }
  pdl_grid_sync();                                       // (the parameters above are constants: staged while the producer of q drains)
  __syncthreads();
  const int wpr = W / WS;
  const int nW1 = (TH1 / WS) * wpr;
  const long L1 = (long)TH1 * W;
  const long n_units = (long)N1 * groups;
  for (long unit = (long)blockIdx.x * 4 + warp; unit < n_units; unit += (long)gridDim.x * 4) {
    const int win = (int)(unit / groups), g = (int)(unit - (long)win * groups);
    const int b = win / nW1, n = win - b * nW1;
    const int wr = n / wpr, wc = n - wr * wpr;
    const float *qbase = q + (b * L1 + (long)(wr * WS) * W + wc * WS) * C + g * Cg;
    // ---- phase A
    for (int cc = 0; cc < Cg; cc += 32) {
      const int c = cc + lane;
      if (c < Cg) {
        float in[P], w[25];
#pragma unroll
        for (int t = 0; t < 25; ++t) w[t] = __ldg(dw_w + c * 25 + t);
#pragma unroll
        for (int p = 0; p < P; ++p) in[p] = __ldg(qbase + ((long)(p / WS) * W + p % WS) * C + c);
        const float bias = __ldg(dw_b + c);
#pragma unroll
        for (int pi = 0; pi < WS; ++pi)
#pragma unroll
          for (int pj = 0; pj < WS; ++pj) {
            float acc = 0.0f;
#pragma unroll
            for (int a = 0; a < 5; ++a)
#pragma unroll
              for (int bb = 0; bb < 5; ++bb) {
                const int y = pi + a - 2, x = pj + bb - 2;
                if (y >= 0 && y < WS && x >= 0 && x < WS) acc = fmaf(in[y * WS + x], w[a * 5 + bb], acc);
              }
            tile[(pi * WS + pj) * pitch + c] = acc + bias;
          }
      }
    }
    __syncwarp();
    // ---- phase B
    for (int p = lane; p < P; p += 32) {
      const float *v = tile + p * pitch;
      float s = 0.0f;
      for (int c = 0; c < Cg; ++c) s += v[c];
      const float mean = s / Cg;
      float var = 0.0f;
      for (int c = 0; c < Cg; ++c) {
        const float d = v[c] - mean;
        var = fmaf(d, d, var);
      }
      const float rstd = 1.0f / sqrtf(var / Cg + 1e-5f);
      float oy = 0.0f, ox = 0.0f;
      for (int c = 0; c < Cg; c += 2) {              // (Cg is a multiple of 4) two channels per step on the packed-fp32 GELU:
        // |error| <= 5.5e-7 against the erf form (tc_common.cuh), far below what the sampling positions can resolve
        const float2 a = gelu_fast2(make_float2((v[c] - mean) * rstd * prm[c] + prm[Cg + c], (v[c + 1] - mean) * rstd * prm[c + 1] + prm[Cg + c + 1]));
        oy = fmaf(a.y, prm[2 * Cg + c + 1], fmaf(a.x, prm[2 * Cg + c], oy));
        ox = fmaf(a.y, prm[3 * Cg + c + 1], fmaf(a.x, prm[3 * Cg + c], ox));
      }
      const int pi = p / WS, pj = p % WS;
      const float inv = 1.0f / (float)WS;
      const float ref_y = ((pi + 0.5f) / WS) * 2.0f - 1.0f;
      const float ref_x = ((pj + 0.5f) / WS) * 2.0f - 1.0f;
      const float pos_y = tanhf(oy) * inv * 2.0f + ref_y;
      const float pos_x = tanhf(ox) * inv * 2.0f + ref_x;
      float *o = pix + (unit * P + p) * 2;
      o[0] = ((pos_y + 1.0f) / 2.0f) * (WS - 1);
      o[1] = ((pos_x + 1.0f) / 2.0f) * (WS - 1);
    }
    __syncwarp();
  }
}

// F.grid_sample(bilinear, zeros padding, align_corners=True) of kv window j with the offsets of its query window.
template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) cva_sample_kernel(const InT *__restrict__ x2, const float *__restrict__ pix,
                                                         OutT *__restrict__ sampled, int N1, int TH1, int TH2, int W, int C,
                                                         int groups, int ws, int per_clip) {
  pdl_grid_sync();
  const int P = ws * ws;
  const int Cg = C / groups;
  const int j = blockIdx.x;
  const int r = TH2 / TH1;
  const int nW1 = (TH1 / ws) * (W / ws);
  const int nW2 = (TH2 / ws) * (W / ws);
  const int qw = cva_query_window_s(j, r, N1, nW1, per_clip);
  const int b2 = j / nW2, n2 = j % nW2;
  const long L2 = (long)TH2 * W;
  const int wpr = W / ws;
  const long base_row = b2 * L2 + (long)(n2 / wpr) * ws * W + (n2 % wpr) * ws;      // canvas row of the window's pixel (0,0)
  const int C4 = C >> 2, Cg4 = Cg >> 2;
  // blockIdx.y splits the window's P*C/4 work items when there are too few windows to fill the machine.  The kernel is
  // instruction-bound (ncu: 3 IPC), so the (pixel, channel quad) of an item is stepped incrementally -- no division in the
  // loop -- and all addresses are 32-bit offsets from the window's first pixel.
  const int items = P * C4, per = (items + gridDim.y - 1) / gridDim.y;
  const int e_end = min(items, (int)(blockIdx.y + 1) * per);
  const int e0 = blockIdx.y * per + threadIdx.x;
  int p = e0 / C4, c4 = e0 - p * C4;
  const int step_p = blockDim.x / C4, step_c = blockDim.x - step_p * C4;
  const InT *wbase = x2 + base_row * C;                                   // pixel (0,0) of the kv window
  const float *pbase = pix + (long)qw * groups * P * 2;
  OutT *obase = sampled + (long)j * P * C;
  for (int e = e0; e < e_end; e += blockDim.x) {
    const int c = c4 * 4;
    int g = 0;
    for (int t = Cg4; t <= c4; t += Cg4) ++g;                              // Cg % 4 == 0: the four channels share a group
    const float2 pp = __ldg(reinterpret_cast<const float2 *>(pbase + (g * P + p) * 2));
    const float py = pp.x, px = pp.y;
    const float fy = floorf(py), fx = floorf(px);
    const int y0 = (int)fy, x0 = (int)fx;
    const float wy1 = py - fy, wx1 = px - fx;
    const float wy0 = 1.0f - wy1, wx0 = 1.0f - wx1;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = y0 + dy, xx = x0 + dx;
        if (yy < 0 || yy >= ws || xx < 0 || xx >= ws) continue;
        const float wgt = (dy ? wy1 : wy0) * (dx ? wx1 : wx0);
        const InT *src = wbase + (yy * W + xx) * C + c;
        float4 v;
        if (sizeof(InT) == 4) {
          v = __ldg(reinterpret_cast<const float4 *>(src));
        } else {
          const uint2 u = __ldg(reinterpret_cast<const uint2 *>(src));
          if constexpr (sizeof(InT) == 2) {
            const float2 a = unpack2<InT>(u.x), b = unpack2<InT>(u.y);
            v = make_float4(a.x, a.y, b.x, b.y);
          }
        }
        acc.x = fmaf(v.x, wgt, acc.x); acc.y = fmaf(v.y, wgt, acc.y); acc.z = fmaf(v.z, wgt, acc.z); acc.w = fmaf(v.w, wgt, acc.w);
      }
    }
    OutT *dst = obase + p * C + c;
    if constexpr (sizeof(OutT) == 2) {
      uint2 pk;
      pk.x = pack2<OutT>(acc.x, acc.y);
      pk.y = pack2<OutT>(acc.z, acc.w);
      *reinterpret_cast<uint2 *>(dst) = pk;
    } else {
      *reinterpret_cast<float4 *>(dst) = acc;
    }
    p += step_p;
    c4 += step_c;
    if (c4 >= C4) {
      c4 -= C4;
      ++p;
    }
  }
}

// x_new = h + window_partition(h) (window-major, added at flat position) + reinterpret_(C,P)->(P,C)(y):
// flat element f = p*C + c of window (b, n)'s (P, C) block receives  h[canvas(n, p), c]  and  y[(win*P + f % P), f / P]  -- the
// raw reshape of a (C, P) tensor (deformableAttention.py:403) is a transposed read of the token-major y.  One CTA per
// window: y's (P, C) block is staged in shared memory (coalesced read) so that the transposed access never touches DRAM
// with a stride; everything else is 16-byte coalesced.
// WS > 0: compile-time window size (7 or 8) -- the loop body has a dozen divisions by P and ws, which as run-time values made
// the kernel instruction-bound; WS == 0: generic.
template <int WS>
__global__ void __launch_bounds__(256) cva_residual_kernel(const float *__restrict__ h, const float *__restrict__ y,
                                                           float *__restrict__ x_new, int TH1, int W, int C, int ws_rt) {
  pdl_grid_sync();
  extern __shared__ float ysm[];      // [P][C + 1]
  const int ws = WS > 0 ? WS : ws_rt;
  const int P = ws * ws;
  const long L1 = (long)TH1 * W;
  const int nW = (int)(L1 / P);
  const long win = blockIdx.x;
  const long b = win / nW;
  const int n = (int)(win - b * nW);
  const int C4 = C >> 2, ldy = C + 1;
  const float4 *y4 = reinterpret_cast<const float4 *>(y + win * P * C);
  for (int e = threadIdx.x; e < P * C4; e += blockDim.x) {
    const int p = e / C4, c = (e - p * C4) * 4;
    const float4 v = __ldg(y4 + e);
    float *d = ysm + p * ldy + c;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const long blk = (b * L1 + (long)n * P) * C;      // flat offset of this window's (P, C) block in window-major order
  for (int e = threadIdx.x; e < P * C4; e += blockDim.x) {
    const int p = e / C4, c = (e - p * C4) * 4;
    const long wrow = b * L1 + window_token_row(n, p, TH1, W, ws, 0);
    const float4 a = __ldg(reinterpret_cast<const float4 *>(h + blk) + e);
    const float4 hw = __ldg(reinterpret_cast<const float4 *>(h + wrow * C + c));
    const int f = p * C + c;
    float4 o;
    o.x = a.x + hw.x + ysm[(f % P) * ldy + f / P];
    o.y = a.y + hw.y + ysm[((f + 1) % P) * ldy + (f + 1) / P];
    o.z = a.z + hw.z + ysm[((f + 2) % P) * ldy + (f + 2) / P];
    o.w = a.w + hw.w + ysm[((f + 3) % P) * ldy + (f + 3) / P];
    reinterpret_cast<float4 *>(x_new + blk)[e] = o;
  }
}

}  // namespace mumpy

using namespace mumpy;

template <int MAXC>
static int launch_cva_offsets(const float *q, const float *dw_w, const float *dw_b, const float *ln_g, const float *ln_b, const float *pw,
                              float *pix, int N1, int TH1, int W, int C, int groups, int ws, size_t smem, cudaStream_t st) {
  static size_t granted = 0;
  if (smem > 48 * 1024 && smem > granted) {
    cudaError_t e = cudaFuncSetAttribute(cva_offsets_kernel<MAXC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cva_offsets: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    granted = smem;
  }
  // resident CTAs: 8 x 256 threads per SM, shared memory permitting
  const long per_sm = (220 * 1024) / (long)(smem + 1024) < 8 ? ((220 * 1024) / (long)(smem + 1024) < 1 ? 1 : (220 * 1024) / (long)(smem + 1024)) : 8;
  const long want = cdiv(148 * per_sm, groups);
  dim3 grid((unsigned)(N1 < want ? N1 : want), (unsigned)groups);
  launch_kernel(cva_offsets_kernel<MAXC>, grid, 256, smem, st, q, dw_w, dw_b, ln_g, ln_b, pw, pix, N1, TH1, W, C, groups, ws);
  return launch_status("cva_offsets");
}

extern "C" int mumpy_cva_offsets(const float *q, const float *dw_w, const float *dw_b, const float *ln_g, const float *ln_b,
                                 const float *pw, float *pix, int B, int TH1, int W, int C, int groups, int ws, void *stream) {
  MUMPY_REQUIRE(q && dw_w && dw_b && ln_g && ln_b && pw && pix && B > 0 && C % groups == 0, "cva_offsets: bad arguments");
  MUMPY_REQUIRE(TH1 % ws == 0 && W % ws == 0 && ws <= 8, "cva_offsets: bad window geometry");
  const int Cg = C / groups;
  MUMPY_REQUIRE(Cg <= 256 && Cg % 4 == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0, "cva_offsets: group width %d unsupported (<= 256, multiple of 4)", Cg);
  const int N1 = B * (TH1 / ws) * (W / ws);
  cudaStream_t st = as_stream(stream);
  static int use_reg = -1;                 // MUMPY_CVA_REG=0 forces the CTA-per-unit kernel (A/B measurements)
  if (use_reg < 0) {
    const char *e = getenv("MUMPY_CVA_REG");
    use_reg = (e && e[0] == '0') ? 0 : 1;
  }
  if (use_reg && (ws == 7 || ws == 8)) {
    const size_t rsmem = ((size_t)4 * Cg + 4 * (size_t)ws * ws * (Cg + 1)) * sizeof(float);
    if (rsmem <= 200 * 1024) {
      cudaError_t e = cudaSuccess;
      static size_t granted7 = 0, granted8 = 0;
      size_t &granted = ws == 7 ? granted7 : granted8;
      if (rsmem > 48 * 1024 && rsmem > granted) {
        e = ws == 7 ? cudaFuncSetAttribute(cva_offsets_reg_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem)
                    : cudaFuncSetAttribute(cva_offsets_reg_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
        if (e != cudaSuccess) {
          set_error("cva_offsets: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
          return MUMPY_ERR_CUDA;
        }
        granted = rsmem;
      }
      const long units = (long)N1 * groups;
      long per_sm = (220 * 1024) / (long)(rsmem + 1024);
      per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);             // 2 CTAs x 4 warps: ~250 registers per thread (window, taps, interleaved accumulators)
      const long ctas = cdiv(units, 4) < 148 * per_sm ? cdiv(units, 4) : 148 * per_sm;
      if (ws == 7) launch_kernel(cva_offsets_reg_kernel<7>, (unsigned)ctas, 128, rsmem, st, q, dw_w, dw_b, ln_g, ln_b, pw, pix, N1, TH1, W, C, groups);
      else launch_kernel(cva_offsets_reg_kernel<8>, (unsigned)ctas, 128, rsmem, st, q, dw_w, dw_b, ln_g, ln_b, pw, pix, N1, TH1, W, C, groups);
      return launch_status("cva_offsets");
    }
  }
  const size_t smem = ((size_t)((ws + 4) * (ws + 4) + 25) * Cg + 2 * ws * ws) * sizeof(float);
  if (Cg <= 32) return launch_cva_offsets<1>(q, dw_w, dw_b, ln_g, ln_b, pw, pix, N1, TH1, W, C, groups, ws, smem, st);
  if (Cg <= 64) return launch_cva_offsets<2>(q, dw_w, dw_b, ln_g, ln_b, pw, pix, N1, TH1, W, C, groups, ws, smem, st);
  if (Cg <= 128) return launch_cva_offsets<4>(q, dw_w, dw_b, ln_g, ln_b, pw, pix, N1, TH1, W, C, groups, ws, smem, st);
  return launch_cva_offsets<8>(q, dw_w, dw_b, ln_g, ln_b, pw, pix, N1, TH1, W, C, groups, ws, smem, st);
}

extern "C" int mumpy_cva_sample(const void *x2, int x2_dtype, const float *pix, void *sampled, int out_dtype, int B, int TH1, int TH2,
                                int W, int C, int groups, int ws, int per_clip_pairing, void *stream) {
  MUMPY_REQUIRE(x2 && pix && sampled && B > 0 && TH2 % TH1 == 0, "cva_sample: bad arguments");
  MUMPY_REQUIRE(C % groups == 0 && (C / groups) % 4 == 0, "cva_sample: channels per group must be a multiple of 4");
  MUMPY_REQUIRE(((reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(sampled) | reinterpret_cast<uintptr_t>(pix)) & 15) == 0,
                "cva_sample: buffers must be 16-byte aligned");
  MUMPY_REQUIRE(x2_dtype == MUMPY_F32 || x2_dtype == out_dtype, "cva_sample: a 16-bit input needs the same 16-bit output type");
  const int N1 = B * (TH1 / ws) * (W / ws);
  const int N2 = B * (TH2 / ws) * (W / ws);
  cudaStream_t st = as_stream(stream);
  int ysplit = (int)cdiv(148 * 4, N2);
  const int max_split = (int)cdiv((long)ws * ws * (C / 4), 256);
  if (ysplit > max_split) ysplit = max_split;
  const dim3 sgrid((unsigned)N2, (unsigned)(ysplit < 1 ? 1 : ysplit));
  if (is_16bit(x2_dtype))
    MUMPY_WITH_16(x2_dtype, T, launch_kernel(cva_sample_kernel<T, T>, sgrid, 256, 0, st, static_cast<const T *>(x2), pix, static_cast<T *>(sampled), N1, TH1, TH2, W, C, groups, ws, per_clip_pairing));
  else if (is_16bit(out_dtype))
    MUMPY_WITH_16(out_dtype, T, launch_kernel(cva_sample_kernel<float, T>, sgrid, 256, 0, st, static_cast<const float *>(x2), pix, static_cast<T *>(sampled), N1, TH1, TH2, W, C, groups, ws, per_clip_pairing));
  else
    launch_kernel(cva_sample_kernel<float, float>, sgrid, 256, 0, st, static_cast<const float *>(x2), pix, static_cast<float *>(sampled), N1, TH1, TH2, W, C, groups, ws, per_clip_pairing);
  return launch_status("cva_sample");
}

extern "C" int mumpy_cva_residual(const float *h, const float *y, float *x_new, int B, int TH1, int W, int C, int ws,
                                  void *stream) {
  MUMPY_REQUIRE(h && y && x_new && h != x_new && B > 0, "cva_residual: bad arguments (must be out of place)");
  MUMPY_REQUIRE(C % 4 == 0 && TH1 % ws == 0 && W % ws == 0 &&
                    ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(x_new)) & 15) == 0,
                "cva_residual: C %% 4 == 0, whole windows and 16-byte aligned buffers required");
  const int P = ws * ws;
  const size_t smem = (size_t)P * (C + 1) * sizeof(float);
  MUMPY_REQUIRE(smem <= 227 * 1024, "cva_residual: window block of %zu B does not fit shared memory", smem);
  static size_t granted_ws[3] = {0, 0, 0};
  size_t &granted = granted_ws[ws == 7 ? 0 : (ws == 8 ? 1 : 2)];
  if (smem > 48 * 1024 && smem > granted) {
    cudaError_t e = ws == 7 ? cudaFuncSetAttribute(cva_residual_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                  : ws == 8 ? cudaFuncSetAttribute(cva_residual_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                            : cudaFuncSetAttribute(cva_residual_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cva_residual: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    granted = smem;
  }
  const long wins = (long)B * (TH1 / ws) * (W / ws);
  if (ws == 7) launch_kernel(cva_residual_kernel<7>, (unsigned)wins, 256, smem, as_stream(stream), h, y, x_new, TH1, W, C, ws);
  else if (ws == 8) launch_kernel(cva_residual_kernel<8>, (unsigned)wins, 256, smem, as_stream(stream), h, y, x_new, TH1, W, C, ws);
  else launch_kernel(cva_residual_kernel<0>, (unsigned)wins, 256, smem, as_stream(stream), h, y, x_new, TH1, W, C, ws);
  return launch_status("cva_residual");
}
