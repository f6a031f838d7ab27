// LayerNorm (plain rows and the PatchMerging 2x2 gather) and NHWC GroupNorm + activation.
// Statistics are always fp32 two-pass (mean, then centred variance), like the reference's ATen kernels.
#include <stdlib.h>

#include "common.cuh"

namespace mumpy {

struct PlainRows {
  const float *x;
  int C;
  static constexpr int kSegs = 1;
  __device__ __forceinline__ const float *seg(long row, int) const { return x + row * C; }
  // float4 vector i of the (kSegs * C)-wide row
  __device__ __forceinline__ const float4 *vec(long row, int i) const { return reinterpret_cast<const float4 *>(x + row * C) + i; }
};

// PatchMerging: output token (b, r', c') = cat of canvas tokens (2r'+dy, 2c'+dx), segment q: dy = q&1, dx = q>>1
// (swinTransformer.py:357-361).
struct MergeRows {
  const float *x;
  int C, TH, W;
  static constexpr int kSegs = 4;
  __device__ __forceinline__ const float *seg(long row, int q) const {
    const int W2 = W / 2, H2 = TH / 2;
    const int c2 = (int)(row % W2);
    const long t = row / W2;
    const int r2 = (int)(t % H2);
    const long b = t / H2;
    return x + ((b * TH + 2 * r2 + (q & 1)) * W + 2 * c2 + (q >> 1)) * C;
  }
  __device__ __forceinline__ const float4 *vec(long row, int i) const {
    const int nvseg = C >> 2;
    const int q = i / nvseg;
    return reinterpret_cast<const float4 *>(seg(row, q)) + (i - q * nvseg);
  }
};

template <typename Rows, typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(Rows rows, const float *__restrict__ gamma, const float *__restrict__ beta,
                                                        OutT *__restrict__ out, long n_rows, int C, float eps) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int CT = C * Rows::kSegs;
  float s = 0.0f;
  for (int q = 0; q < Rows::kSegs; ++q) {
    const float *p = rows.seg(row, q);
    for (int i = lane; i < C; i += 32) s += p[i];
  }
  const float mean = warp_sum(s) / CT;
  float v = 0.0f;
  for (int q = 0; q < Rows::kSegs; ++q) {
    const float *p = rows.seg(row, q);
    for (int i = lane; i < C; i += 32) {
      const float d = p[i] - mean;
      v = fmaf(d, d, v);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / CT + eps);
  OutT *o = out + row * CT;
  float amax = 0.0f;
  for (int q = 0; q < Rows::kSegs; ++q) {
    const float *p = rows.seg(row, q);
    for (int i = lane; i < C; i += 32) {
      const int c = q * C + i;
      const float y = (p[i] - mean) * rstd * gamma[c] + beta[c];
      amax = fmaxf(amax, fabsf(y));
      o[c] = from_f32<OutT>(y);
    }
  }
  if (is_half_t<OutT>::value) f16_guard(amax);
}

// Plain rows, C % 4 == 0, C <= 1024: rows live in registers as float4.  LPR lanes share a row (8 / 16 / 32, so that a lane
// holds NV = C/4/LPR = 3..8 vectors and the statistics need only log2(LPR) shuffle steps -- with a full warp per 96..256-wide
// row the kernel was issue-bound on shuffles, ncu: 2.3 IPC at 37 % of HBM); a warp handles 32/LPR rows side by side and RI such
// groups per iteration to keep ~8 independent 16-byte loads in flight per lane.  Exact two-pass statistics, coalesced 128-byte
// (LPR = 8) or longer row segments, 8-byte (16-bit x 4) or 16-byte (fp32 x 4) stores.
template <typename Rows, typename OutT, int LPR, int NV, int RI>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const Rows rows, const float *__restrict__ gamma,
                                                            const float *__restrict__ beta, OutT *__restrict__ out, long n_rows, int C, float eps) {
  pdl_grid_sync();
  constexpr int RPW = (32 / LPR) * RI;                     // rows per warp iteration
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, grp = lane / LPR;
  const int nv = C >> 2;
  const long warp_stride = (long)gridDim.x * (blockDim.x >> 5) * RPW;
  [[maybe_unused]] float amax = 0.0f;
  for (long row0 = ((long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; row0 < n_rows; row0 += warp_stride) {
    float4 v[RI][NV];
    float s[RI], q[RI];
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      const long row = row0 + r * (32 / LPR) + grp;
      s[r] = 0.0f;
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int i = sub + LPR * u;
        v[r][u] = (row < n_rows && i < nv) ? *rows.vec(row, i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int r = 0; r < RI; ++r)
#pragma unroll
      for (int u = 0; u < NV; ++u) s[r] += (v[r][u].x + v[r][u].y) + (v[r][u].z + v[r][u].w);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < RI; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      s[r] = s[r] / C;                                     // mean
      q[r] = 0.0f;
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        if (sub + LPR * u < nv) {
          const float a = v[r][u].x - s[r], b = v[r][u].y - s[r], c = v[r][u].z - s[r], d = v[r][u].w - s[r];
          q[r] += (a * a + b * b) + (c * c + d * d);
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < RI; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
#pragma unroll
    for (int r = 0; r < RI; ++r) q[r] = 1.0f / sqrtf(q[r] / C + eps);      // rstd
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int i = sub + LPR * u;
      if (i >= nv) continue;
      const float4 g4 = __ldg(reinterpret_cast<const float4 *>(gamma) + i), b4 = __ldg(reinterpret_cast<const float4 *>(beta) + i);
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        const long row = row0 + r * (32 / LPR) + grp;
        if (row < n_rows) {
          const float rstd = q[r];
          const float o0 = (v[r][u].x - s[r]) * rstd * g4.x + b4.x, o1 = (v[r][u].y - s[r]) * rstd * g4.y + b4.y;
          const float o2 = (v[r][u].z - s[r]) * rstd * g4.z + b4.z, o3 = (v[r][u].w - s[r]) * rstd * g4.w + b4.w;
          if constexpr (sizeof(OutT) == 2) {
            if constexpr (is_half_t<OutT>::value) amax = fmaxf(fmaxf(amax, fabsf(o0)), fmaxf(fmaxf(fabsf(o1), fabsf(o2)), fabsf(o3)));
            uint2 pk;
            pk.x = pack2<OutT>(o0, o1);
            pk.y = pack2<OutT>(o2, o3);
            reinterpret_cast<uint2 *>(out + row * C)[i] = pk;
          } else {
            reinterpret_cast<float4 *>(out + row * C)[i] = make_float4(o0, o1, o2, o3);
          }
        }
      }
    }
  }
  if constexpr (is_half_t<OutT>::value) f16_guard(amax);
}

// `C` below is the full normalised width (4 x the canvas width for MergeRows)
template <typename Rows, typename OutT, int LPR, int NV, int RI>
static void launch_ln_vec(const Rows &x, const float *gamma, const float *beta, void *out, long rows, int C, float eps, cudaStream_t st) {
  const long warps = cdiv(rows, (32 / LPR) * RI);
  const long ctas = cdiv(warps, 8);
  launch_kernel(layernorm_vec_kernel<Rows, OutT, LPR, NV, RI>, (unsigned)(ctas < 148 * 64 ? ctas : 148 * 64), 256, 0, st, x, gamma, beta, static_cast<OutT *>(out), rows, C, eps);
}

template <typename Rows, typename OutT>
static void dispatch_ln_vec(const Rows &x, const float *gamma, const float *beta, void *out, long rows, int C, float eps, cudaStream_t st) {
  const int nv = C / 4;                 // float4 vectors per row
  if (nv <= 24) launch_ln_vec<Rows, OutT, 8, 3, 2>(x, gamma, beta, out, rows, C, eps, st);
  else if (nv <= 32) launch_ln_vec<Rows, OutT, 8, 4, 2>(x, gamma, beta, out, rows, C, eps, st);
  else if (nv <= 48) launch_ln_vec<Rows, OutT, 16, 3, 2>(x, gamma, beta, out, rows, C, eps, st);
  else if (nv <= 64) launch_ln_vec<Rows, OutT, 16, 4, 2>(x, gamma, beta, out, rows, C, eps, st);
  else if (nv <= 96) launch_ln_vec<Rows, OutT, 32, 3, 2>(x, gamma, beta, out, rows, C, eps, st);
  else if (nv <= 128) launch_ln_vec<Rows, OutT, 32, 4, 2>(x, gamma, beta, out, rows, C, eps, st);
  else if (nv <= 192) launch_ln_vec<Rows, OutT, 32, 6, 1>(x, gamma, beta, out, rows, C, eps, st);
  else launch_ln_vec<Rows, OutT, 32, 8, 1>(x, gamma, beta, out, rows, C, eps, st);
}

template <typename Rows>
static int launch_ln(Rows rows, const float *gamma, const float *beta, void *out, int out_dtype, long n_rows, int C, float eps,
                     cudaStream_t st) {
  const int warps = 8;
  dim3 grid((unsigned)cdiv(n_rows, warps));
  if (is_16bit(out_dtype))
    MUMPY_WITH_16(out_dtype, T, launch_kernel(layernorm_kernel<Rows, T>, grid, warps * 32, 0, st, rows, gamma, beta, static_cast<T *>(out), n_rows, C, eps));
  else
    launch_kernel(layernorm_kernel<Rows, float>, grid, warps * 32, 0, st, rows, gamma, beta, static_cast<float *>(out), n_rows, C, eps);
  return launch_status("layernorm");
}

// ---- GroupNorm on NHWC ----
// (1) gn_partial_kernel: one CTA per (image, pixel chunk) stages its chunk in shared memory and writes, per group, the
//     chunk's exact two-pass (mean, M2);  (2) gn_finalize_kernel combines the chunks with Chan's parallel-variance formula in
//     a fixed order (deterministic);  (3) groupnorm_apply_kernel normalises + activates with float4 accesses.
constexpr int GN_SMEM_FLOATS = 12 * 1024;      // 48 KB staging

// generic fallback (any C): one warp per group, scalar indexing
__global__ void __launch_bounds__(256) gn_partial_generic_kernel(const float *__restrict__ x, float *__restrict__ partial, int HW, int C, int groups,
                                                                 int pix, int nchunks) {
  pdl_grid_sync();
  extern __shared__ float tile[];   // [pix][C]
  const int b = blockIdx.x / nchunks, chunk = blockIdx.x % nchunks;
  const int p0 = chunk * pix;
  const int np = min(pix, HW - p0);
  const float4 *src = reinterpret_cast<const float4 *>(x + ((long)b * HW + p0) * C);
  float4 *dst = reinterpret_cast<float4 *>(tile);
  const int nvec = np * C / 4;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  const int cg = C / groups;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int n = np * cg;
  for (int g = warp; g < groups; g += nwarps) {
    float s = 0.0f;
    for (int e = lane; e < n; e += 32) s += tile[(e / cg) * C + g * cg + (e % cg)];
    const float mean = warp_sum(s) / n;
    float q = 0.0f;
    for (int e = lane; e < n; e += 32) {
      const float d = tile[(e / cg) * C + g * cg + (e % cg)] - mean;
      q = fmaf(d, d, q);
    }
    q = warp_sum(q);
    if (lane == 0) {
      float *o = partial + (((long)b * groups + g) * nchunks + chunk) * 2;
      o[0] = mean;
      o[1] = q;
    }
  }
}

// C/4 divides 256 (C = 32 ... 1024, powers of two): a thread owns one channel quad (one group, cg % 4 == 0) and every
// (256 / (C/4))-th pixel of the staged chunk: float4 shared-memory reads, no index arithmetic per element; the per-thread
// partials are combined per group in a fixed order (deterministic).  Exact two-pass (mean, then M2 about the chunk mean).
__global__ void __launch_bounds__(256) gn_partial_kernel(const float *__restrict__ x, float *__restrict__ partial, int HW, int C, int groups, int pix,
                                                         int nchunks) {
  pdl_grid_sync();
  extern __shared__ float tile[];   // [pix][C]
  __shared__ float red[256];
  __shared__ float gmean[64];
  const int b = blockIdx.x / nchunks, chunk = blockIdx.x % nchunks;
  const int p0 = chunk * pix;
  const int np = min(pix, HW - p0);
  const float4 *src = reinterpret_cast<const float4 *>(x + ((long)b * HW + p0) * C);
  float4 *t4 = reinterpret_cast<float4 *>(tile);
  const int C4 = C >> 2;
  const int nvec = np * C4;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) t4[i] = src[i];
  __syncthreads();
  const int cg = C / groups, cqg = cg >> 2;           // channel quads per group
  const int lanes_p = 256 / C4;
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cnt = cqg * lanes_p;                       // partials per group
  const float inv_n = 1.0f / (float)(np * cg);
  float s = 0.0f;
  for (int p = pl; p < np; p += lanes_p) {
    const float4 v = t4[p * C4 + c4];
    s += (v.x + v.y) + (v.z + v.w);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int g = warp; g < groups; g += 8) {
    float a = 0.0f;
    for (int k = lane; k < cnt; k += 32) a += red[(k / cqg) * C4 + g * cqg + (k % cqg)];
    a = warp_sum(a);
    if (lane == 0) gmean[g] = a * inv_n;
  }
  __syncthreads();
  const float mean = gmean[(c4 * 4) / cg];
  float q = 0.0f;
  for (int p = pl; p < np; p += lanes_p) {
    const float4 v = t4[p * C4 + c4];
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  red[threadIdx.x] = q;
  __syncthreads();
  for (int g = warp; g < groups; g += 8) {
    float a = 0.0f;
    for (int k = lane; k < cnt; k += 32) a += red[(k / cqg) * C4 + g * cqg + (k % cqg)];
    a = warp_sum(a);
    if (lane == 0) {
      float *o = partial + (((long)b * groups + g) * nchunks + chunk) * 2;
      o[0] = gmean[g];
      o[1] = a;
    }
  }
}

// one warp per (image, group): every lane folds its chunks (lane, lane+32, ...) with Chan's parallel-variance update, then the
// 32 lane results are merged pairwise in a fixed order (deterministic)
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float *__restrict__ partial, float *__restrict__ stats, int total_bg, int HW, int cg,
                                                          int pix, int nchunks, float eps) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int bg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bg >= total_bg) return;
  const float *p = partial + (long)bg * nchunks * 2;
  float n_a = 0.0f, mean_a = 0.0f, m2_a = 0.0f;
  for (int c = lane; c < nchunks; c += 32) {
    const float n_b = (float)(min(pix, HW - c * pix) * cg);
    const float mean_b = p[2 * c], m2_b = p[2 * c + 1];
    const float n_ab = n_a + n_b;
    const float delta = mean_b - mean_a;
    mean_a += delta * (n_b / n_ab);
    m2_a += m2_b + delta * delta * (n_a * n_b / n_ab);
    n_a = n_ab;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float n_b = __shfl_down_sync(0xffffffffu, n_a, o);
    const float mean_b = __shfl_down_sync(0xffffffffu, mean_a, o);
    const float m2_b = __shfl_down_sync(0xffffffffu, m2_a, o);
    const float n_ab = n_a + n_b;
    if (n_b > 0.0f) {
      const float delta = mean_b - mean_a;
      mean_a += delta * (n_b / n_ab);
      m2_a += m2_b + delta * delta * (n_a * n_b / n_ab);
      n_a = n_ab;
    }
  }
  if (lane == 0) {
    stats[2 * bg] = mean_a;
    stats[2 * bg + 1] = 1.0f / sqrtf(m2_a / n_a + eps);
  }
}

__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const float *__restrict__ x, const float *__restrict__ stats,
                                                              const float *__restrict__ gamma, const float *__restrict__ beta,
                                                              float *__restrict__ out, long ld_out, int out_col, long total4, int HW,
                                                              int C, int groups, int act, int quad_mean) {
  pdl_grid_sync();
  const int cg = C / groups;
  const int C4 = C / 4;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  const bool small = total4 < (1l << 31);
  for (; i < total4; i += stride) {
    int c, prem;
    long pix, b;
    divmod_idx(i, C4, small, pix, c);
    c *= 4;
    divmod_idx(pix, HW, small, b, prem);
    const float4 v = reinterpret_cast<const float4 *>(x)[i];
    const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + c)), bt = __ldg(reinterpret_cast<const float4 *>(beta + c));
    const float *st = stats + 2 * (b * groups + c / cg);          // cg % 4 == 0: the four channels share a group
    const float mean = st[0], rstd = st[1];
    float4 o;
    o.x = apply_act((v.x - mean) * rstd * g.x + bt.x, act);
    o.y = apply_act((v.y - mean) * rstd * g.y + bt.y, act);
    o.z = apply_act((v.z - mean) * rstd * g.z + bt.z, act);
    o.w = apply_act((v.w - mean) * rstd * g.w + bt.w, act);
    if (quad_mean) out[pix * ld_out + out_col + (c >> 2)] = (o.x + o.y + o.z + o.w) * 0.25f;      // DAP: mean of 4 consecutive channels
    else *reinterpret_cast<float4 *>(out + pix * ld_out + out_col + c) = o;
  }
}

// The same pass for C/4 dividing 256 (C = 32 ... 1024, powers of two -- every GroupNorm of the decoder): a thread keeps ONE
// channel quad (its gamma / beta / group statistics are loaded once) and walks the pixels of a contiguous range of one image.
// The flat kernel above spends ~40 instructions per float4 on index decomposition and per-element parameter loads (ncu: issue
// bound at 3.3 TB/s); here the loop body is load, 4 FMA-pairs, activation, store.
template <int ACT, bool QUAD>
__global__ void __launch_bounds__(256) groupnorm_apply_rows_kernel(const float *__restrict__ x, const float *__restrict__ stats,
                                                                   const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                   float *__restrict__ out, long ld_out, int out_col, int HW, int C4, int cqg,
                                                                   int groups, int chunks, int pix_per_cta) {
  pdl_grid_sync();
  const int b = blockIdx.x / chunks, chunk = blockIdx.x - b * chunks;
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4, lanes_p = 256 / C4;
  const int p1 = min(HW, (chunk + 1) * pix_per_cta);
  const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma) + c4), bt = __ldg(reinterpret_cast<const float4 *>(beta) + c4);
  const float *st = stats + 2 * ((long)b * groups + c4 / cqg);
  const float mean = st[0], rstd = st[1];
  const float4 *x4 = reinterpret_cast<const float4 *>(x) + (long)b * HW * C4 + c4;
  float *o = out + (long)b * HW * ld_out + out_col + (QUAD ? c4 : 4 * c4);
#pragma unroll 4
  for (int p = chunk * pix_per_cta + pl; p < p1; p += lanes_p) {
    const float4 v = x4[(long)p * C4];
    float4 r;
    r.x = (v.x - mean) * rstd * g.x + bt.x;
    r.y = (v.y - mean) * rstd * g.y + bt.y;
    r.z = (v.z - mean) * rstd * g.z + bt.z;
    r.w = (v.w - mean) * rstd * g.w + bt.w;
    if (ACT == MUMPY_ACT_RELU) {
      r.x = fmaxf(r.x, 0.0f); r.y = fmaxf(r.y, 0.0f); r.z = fmaxf(r.z, 0.0f); r.w = fmaxf(r.w, 0.0f);
    } else if (ACT != MUMPY_ACT_NONE) {
      r.x = apply_act(r.x, ACT); r.y = apply_act(r.y, ACT); r.z = apply_act(r.z, ACT); r.w = apply_act(r.w, ACT);
    }
    if (QUAD) o[(long)p * ld_out] = (r.x + r.y + r.z + r.w) * 0.25f;      // DAP: mean of 4 consecutive channels
    else *reinterpret_cast<float4 *>(o + (long)p * ld_out) = r;
  }
}

// Small maps (HW * C/groups <= 16 K values: the 7x7 ... 28x28 levels of the decoder pyramids): statistics and normalisation in ONE
// kernel, one CTA per (image, group) holding its slice in shared memory -- exact two-pass mean / variance over the whole slice,
// one read of the map instead of two and one launch instead of three on chains that are launch-latency bound.
constexpr int GN_SMALL_FLOATS = 16 * 1024;
__device__ __forceinline__ float gn_block_sum(float v, float *red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                       // (red may still be read from the previous reduction)
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = lane < 8 ? red[lane] : 0.0f;
  t = warp_sum(t);
  return t;                              // every thread holds the CTA total (256 threads = 8 warps)
}
__global__ void __launch_bounds__(256) gn_small_kernel(const float *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                                                       float *__restrict__ out, long ld_out, int out_col, int HW, int C, int groups, float eps, int act,
                                                       int quad_mean) {
  pdl_grid_sync();
  extern __shared__ float4 gtile[];      // [HW][cg / 4]
  __shared__ float red[8];
  const int b = blockIdx.x / groups, g = blockIdx.x - b * groups;
  const int cg = C / groups, cq = cg >> 2;
  const int n4 = HW * cq;
  const float *src = x + (long)b * HW * C + g * cg;
  float s = 0.0f;
  for (int i = threadIdx.x; i < n4; i += 256) {
    const int p = i / cq, q = i - p * cq;
    const float4 v = *reinterpret_cast<const float4 *>(src + (long)p * C + 4 * q);
    gtile[i] = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = gn_block_sum(s, red) / (float)(n4 * 4);
  float m2 = 0.0f;
  for (int i = threadIdx.x; i < n4; i += 256) {
    const float4 v = gtile[i];
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    m2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  const float rstd = 1.0f / sqrtf(gn_block_sum(m2, red) / (float)(n4 * 4) + eps);
  float *dst = out + (long)b * HW * ld_out + out_col + (quad_mean ? g * cq : g * cg);
  for (int i = threadIdx.x; i < n4; i += 256) {
    const int p = i / cq, q = i - p * cq;
    const float4 v = gtile[i];
    const float4 gm = __ldg(reinterpret_cast<const float4 *>(gamma + g * cg) + q), bt = __ldg(reinterpret_cast<const float4 *>(beta + g * cg) + q);
    float4 r;
    r.x = apply_act((v.x - mean) * rstd * gm.x + bt.x, act);
    r.y = apply_act((v.y - mean) * rstd * gm.y + bt.y, act);
    r.z = apply_act((v.z - mean) * rstd * gm.z + bt.z, act);
    r.w = apply_act((v.w - mean) * rstd * gm.w + bt.w, act);
    if (quad_mean) dst[(long)p * ld_out + q] = (r.x + r.y + r.z + r.w) * 0.25f;
    else *reinterpret_cast<float4 *>(dst + (long)p * ld_out + 4 * q) = r;
  }
}

template <int ACT>
static void launch_gn_apply_rows(const float *x, const float *stats, const float *gamma, const float *beta, float *out, long ld_out, int out_col, int B,
                                 int HW, int C, int groups, int quad_mean, cudaStream_t st) {
  const int C4 = C / 4, cqg = C / groups / 4;
  int pix = 128;
  while (pix > 16 && (long)B * cdiv(HW, pix) < 148 * 8) pix >>= 1;
  const int chunks = (int)cdiv(HW, pix);
  if (quad_mean)
    launch_kernel(groupnorm_apply_rows_kernel<ACT, true>, (unsigned)(B * chunks), 256, 0, st, x, stats, gamma, beta, out, ld_out, out_col, HW, C4, cqg, groups, chunks, pix);
  else
    launch_kernel(groupnorm_apply_rows_kernel<ACT, false>, (unsigned)(B * chunks), 256, 0, st, x, stats, gamma, beta, out, ld_out, out_col, HW, C4, cqg, groups, chunks, pix);
}

}  // namespace mumpy

using namespace mumpy;

extern "C" int mumpy_layernorm(const float *x, const float *gamma, const float *beta, void *out, int out_dtype, long rows, int C,
                               float eps, void *stream) {
  MUMPY_REQUIRE(x && gamma && beta && out && rows > 0 && C > 0, "layernorm: bad arguments");
  if (C % 4 == 0 && C <= 1024 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) |
                                   reinterpret_cast<uintptr_t>(beta)) & 15) == 0) {
    if (is_16bit(out_dtype))
      MUMPY_WITH_16(out_dtype, T, dispatch_ln_vec<PlainRows, T>(PlainRows{x, C}, gamma, beta, out, rows, C, eps, as_stream(stream)));
    else
      dispatch_ln_vec<PlainRows, float>(PlainRows{x, C}, gamma, beta, out, rows, C, eps, as_stream(stream));
    return launch_status("layernorm_vec");
  }
  return launch_ln(PlainRows{x, C}, gamma, beta, out, out_dtype, rows, C, eps, as_stream(stream));
}

extern "C" int mumpy_patch_merge_norm(const float *x, const float *gamma, const float *beta, void *out, int out_dtype, int B,
                                      int TH, int W, int C, float eps, void *stream) {
  MUMPY_REQUIRE(x && gamma && beta && out && B > 0 && TH % 2 == 0 && W % 2 == 0, "patch_merge_norm: bad arguments");
  const long rows = (long)B * (TH / 2) * (W / 2);
  const MergeRows mr{x, C, TH, W};
  if (C % 4 == 0 && 4 * C <= 1024 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) |
                                       reinterpret_cast<uintptr_t>(beta)) & 15) == 0) {
    if (is_16bit(out_dtype))
      MUMPY_WITH_16(out_dtype, T, dispatch_ln_vec<MergeRows, T>(mr, gamma, beta, out, rows, 4 * C, eps, as_stream(stream)));
    else
      dispatch_ln_vec<MergeRows, float>(mr, gamma, beta, out, rows, 4 * C, eps, as_stream(stream));
    return launch_status("patch_merge_norm_vec");
  }
  return launch_ln(mr, gamma, beta, out, out_dtype, rows, C, eps, as_stream(stream));
}

// floats of the caller-owned stats_ws of mumpy_groupnorm_nhwc: 2 B groups statistics + 2 B groups nchunks partial sums (the
// chunking below); callers size the buffer with this instead of repeating the formula
extern "C" long mumpy_groupnorm_workspace_floats(int B, int HW, int C, int groups) {
  if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C > GN_SMEM_FLOATS) return 0;
  int pix = GN_SMEM_FLOATS / C;
  if (pix > 64) pix = 64;
  if (pix > HW) pix = HW;
  const long nchunks = cdiv(HW, pix);
  return 2l * B * groups * (1 + nchunks);
}

extern "C" int mumpy_groupnorm_nhwc(const float *x, const float *gamma, const float *beta, float *stats_ws, float *out,
                                    long ld_out, int out_col, int B, int HW, int C, int groups, float eps, int act,
                                    int quad_mean, void *stream) {
  MUMPY_REQUIRE(x && gamma && beta && stats_ws && out && C % groups == 0, "groupnorm_nhwc: bad arguments");
  MUMPY_REQUIRE((C / groups) % 4 == 0 && (quad_mean || (ld_out % 4 == 0 && out_col % 4 == 0)),
                "groupnorm_nhwc: channels per group, ld_out, out_col must be multiples of 4");
  MUMPY_REQUIRE(C <= GN_SMEM_FLOATS, "groupnorm_nhwc: C too large");
  cudaStream_t st = as_stream(stream);
  static int gn_small = -1;
  if (gn_small < 0) {
    const char *v = getenv("MUMPY_GN_SMALL");
    gn_small = (v && v[0] == '0') ? 0 : 1;
  }
  if (gn_small && (long)HW * (C / groups) <= GN_SMALL_FLOATS && C % 4 == 0 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0) {
    static bool small_attr = false;
    if (!small_attr) {
      cudaFuncSetAttribute(gn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GN_SMALL_FLOATS * (int)sizeof(float));
      small_attr = true;
    }
    launch_kernel(gn_small_kernel, (unsigned)(B * groups), 256, (size_t)HW * (C / groups) * sizeof(float), st, x, gamma, beta, out, ld_out, out_col, HW, C,
                  groups, eps, act, quad_mean);
    return launch_status("gn_small");
  }
  int pix = GN_SMEM_FLOATS / C;
  if (pix > 64) pix = 64;
  if (pix > HW) pix = HW;
  const int nchunks = (int)cdiv(HW, pix);
  float *stats = stats_ws;                                  // 2 * B * groups
  float *partial = stats_ws + 2 * (long)B * groups;         // 2 * B * groups * nchunks
  const int C4 = C / 4;
  static bool gn_attr = false;
  if (!gn_attr) {        // 48 KB of dynamic staging + the static reduction arrays exceed the default 48 KB window
    cudaFuncSetAttribute(gn_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GN_SMEM_FLOATS * (int)sizeof(float));
    gn_attr = true;
  }
  if (C4 <= 256 && 256 % C4 == 0 && groups <= 64)
    launch_kernel(gn_partial_kernel, B * nchunks, 256, (size_t)pix * C * sizeof(float), st, x, partial, HW, C, groups, pix, nchunks);
  else
    launch_kernel(gn_partial_generic_kernel, B * nchunks, 256, (size_t)pix * C * sizeof(float), st, x, partial, HW, C, groups, pix, nchunks);
  int rc = launch_status("gn_partial");
  if (rc) return rc;
  launch_kernel(gn_finalize_kernel, (unsigned)cdiv(B * groups, 8), 256, 0, st, partial, stats, B * groups, HW, C / groups, pix, nchunks, eps);
  rc = launch_status("gn_finalize");
  if (rc) return rc;
  if (C4 <= 256 && 256 % C4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0) {
    switch (act) {
      case MUMPY_ACT_NONE: launch_gn_apply_rows<MUMPY_ACT_NONE>(x, stats, gamma, beta, out, ld_out, out_col, B, HW, C, groups, quad_mean, st); break;
      case MUMPY_ACT_RELU: launch_gn_apply_rows<MUMPY_ACT_RELU>(x, stats, gamma, beta, out, ld_out, out_col, B, HW, C, groups, quad_mean, st); break;
      case MUMPY_ACT_SIGMOID: launch_gn_apply_rows<MUMPY_ACT_SIGMOID>(x, stats, gamma, beta, out, ld_out, out_col, B, HW, C, groups, quad_mean, st); break;
      default: launch_gn_apply_rows<MUMPY_ACT_GELU>(x, stats, gamma, beta, out, ld_out, out_col, B, HW, C, groups, quad_mean, st); break;
    }
    return launch_status("groupnorm_apply_rows");
  }
  const long total4 = (long)B * HW * C / 4;
  const int blocks = (int)(cdiv(total4, 256) < 148 * 16 ? cdiv(total4, 256) : 148 * 16);
  launch_kernel(groupnorm_apply_kernel, blocks, 256, 0, st, x, stats, gamma, beta, out, ld_out, out_col, total4, HW, C, groups, act, quad_mean);
  return launch_status("groupnorm_apply");
}
