// HBM-bound data-movement kernels: view merges (row gathers), decoder resampling (bilinear up, avg-pool,
// pixel-shuffle) with fused gate/skip, layout changes, im2col for the tensor-core conv path, DAP and the
// mask + metric counts.  All are grid-stride, coalesced along the channel (innermost) dimension.
#include "common.cuh"

namespace mumpy {

static inline int flat_blocks(long total, int threads = 256) {
  const long b = cdiv(total, threads);
  return (int)(b < 148 * 16 ? b : 148 * 16);
}

template <typename OutT>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ src, int C, OutT *__restrict__ dst, long dst_ld,
                                                          int dst_col, long total, int rows_out, int rows_src, int div, int mul_hi,
                                                          int mul_lo, int add) {
  pdl_grid_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  float amax = 0.0f;
  for (; i < total; i += stride) {
    const int c = (int)(i % C);
    const long r = i / C;
    const long b = r / rows_out;
    const int q = (int)(r % rows_out);
    const long srow = b * rows_src + (q / div) * mul_hi + (q % div) * mul_lo + add;
    const float v = src[srow * C + c];
    amax = fmaxf(amax, fabsf(v));
    dst[r * dst_ld + dst_col + c] = from_f32<OutT>(v);
  }
  if (is_half_t<OutT>::value) f16_guard(amax);
}

__device__ __forceinline__ void store4(float *dst, float4 v) { *reinterpret_cast<float4 *>(dst) = v; }
template <typename T>
__device__ __forceinline__ void store4(T *dst, float4 v) {
  uint2 pk;
  pk.x = pack2<T>(v.x, v.y);
  pk.y = pack2<T>(v.z, v.w);
  *reinterpret_cast<uint2 *>(dst) = pk;
}

// 4 channels per thread (C, dst_ld, dst_col multiples of 4; 16-byte aligned bases)
template <typename OutT>
__global__ void __launch_bounds__(256) gather_rows_vec_kernel(const float *__restrict__ src, int C4, OutT *__restrict__ dst, long dst_ld,
                                                              int dst_col, long total4, int rows_out, int rows_src, int div, int mul_hi,
                                                              int mul_lo, int add) {
  pdl_grid_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  const bool small = total4 < (1l << 31);
  float amax = 0.0f;
  for (; i < total4; i += stride) {
    int c4, q;
    long r, b;
    divmod_idx(i, C4, small, r, c4);
    divmod_idx(r, rows_out, small, b, q);
    const long srow = b * rows_src + (q / div) * mul_hi + (q % div) * mul_lo + add;
    const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + srow * C4 + c4);
    amax = fmaxf(fmaxf(amax, fabsf(v.x)), fmaxf(fmaxf(fabsf(v.y), fabsf(v.z)), fabsf(v.w)));
    store4(dst + r * dst_ld + dst_col + 4 * c4, v);
  }
  if (is_half_t<OutT>::value) f16_guard(amax);
}

template <typename OutT>
__global__ void __launch_bounds__(256) im2col_kernel(const float *__restrict__ in, long ld_in, OutT *__restrict__ out, long total,
                                                     int H, int W, int Cin, int kh, int kw, int ph, int pw, int Kpad) {
  pdl_grid_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  const int K = kh * kw * Cin;
  float amax = 0.0f;
  for (; i < total; i += stride) {
    const int k = (int)(i % Kpad);
    const long m = i / Kpad;
    float v = 0.0f;
    if (k < K) {
      const int c = k % Cin;
      const int t = k / Cin;
      const int kx = t % kw, ky = t / kw;
      const int x = (int)(m % W);
      const long r = m / W;
      const int y = (int)(r % H);
      const long b = r / H;
      const int yy = y + ky - ph, xx = x + kx - pw;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = in[((b * H + yy) * W + xx) * ld_in + c];
    }
    amax = fmaxf(amax, fabsf(v));
    out[i] = from_f32<OutT>(v);
  }
  if (is_half_t<OutT>::value) f16_guard(amax);
}

// source index + weights of nn.Upsample(bilinear) along one axis (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void bilinear_axis(int o, int n_in, int n_out, int scale, bool aligned, int &i0, int &i1, float &w1) {
  float src;
  if (aligned) {
    const float s = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f;
    src = s * o;
  } else {
    src = (1.0f / scale) * (o + 0.5f) - 0.5f;
    if (src < 0.0f) src = 0.0f;
  }
  i0 = (int)src;
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  w1 = src - i0;
}

__global__ void __launch_bounds__(256) resample_kernel(const float *__restrict__ in, const float *__restrict__ mul,
                                                       const float *__restrict__ add, float *__restrict__ out, long ld_out, int out_col,
                                                       long total, int H, int W, int C, int Ho, int Wo, int Co, int mode, int scale) {
  pdl_grid_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int c = (int)(i % Co);
    long t = i / Co;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const long b = t / Ho;
    const float *img = in + b * H * W * C;
    float v;
    if (mode == MUMPY_RS_IDENTITY) {
      v = img[((long)ho * W + wo) * C + c];
    } else if (mode == MUMPY_RS_AVGPOOL2) {
      const float *p = img + ((long)(2 * ho) * W + 2 * wo) * C + c;
      v = (p[0] + p[C] + p[(long)W * C] + p[(long)W * C + C]) * 0.25f;
    } else if (mode == MUMPY_RS_PIXEL_SHUFFLE2) {
      v = img[((long)(ho >> 1) * W + (wo >> 1)) * C + c * 4 + (ho & 1) * 2 + (wo & 1)];
    } else {
      int y0, y1, x0, x1;
      float wy, wx;
      const bool aligned = mode == MUMPY_RS_UP_ALIGNED;
      bilinear_axis(ho, H, Ho, scale, aligned, y0, y1, wy);
      bilinear_axis(wo, W, Wo, scale, aligned, x0, x1, wx);
      const float v00 = img[((long)y0 * W + x0) * C + c], v01 = img[((long)y0 * W + x1) * C + c];
      const float v10 = img[((long)y1 * W + x0) * C + c], v11 = img[((long)y1 * W + x1) * C + c];
      // ATen upsample_bilinear2d: w0*(w0x*v00 + w1x*v01) + w1*(w0x*v10 + w1x*v11)
      v = (1.0f - wy) * ((1.0f - wx) * v00 + wx * v01) + wy * ((1.0f - wx) * v10 + wx * v11);
    }
    if (mul) v *= mul[i];
    if (add) v += add[i];
    out[(i / Co) * ld_out + out_col + c] = v;
  }
}

// 4 channels per thread for every mode but pixel-shuffle (C, ld_out, out_col multiples of 4).  Idx = uint32_t whenever the
// element count allows: the index decomposition is three divisions by run-time values per element, and 64-bit ones cost
// ~100 instructions each (the kernel was instruction-bound at a third of HBM bandwidth).
// OutT = float, or the 16-bit operand type when the only consumer is a tensor-core convolution (the separate cast pass -- one
// more read and write of the map -- disappears; same rounding as mumpy_cast16).
template <typename Idx, typename OutT>
__global__ void __launch_bounds__(256) resample_vec_kernel(const float *__restrict__ in, const float *__restrict__ mul,
                                                           const float *__restrict__ add, OutT *__restrict__ out, long ld_out, int out_col,
                                                           long total4, int H, int W, int C4, int Ho, int Wo, int mode, int scale) {
  pdl_grid_sync();
  [[maybe_unused]] float amax = 0.0f;
  Idx i = (Idx)blockIdx.x * blockDim.x + threadIdx.x;
  const Idx stride = (Idx)gridDim.x * blockDim.x;
  const float4 *in4 = reinterpret_cast<const float4 *>(in);
  for (; i < (Idx)total4; i += stride) {
    const Idx pixel = i / (Idx)C4;
    const int c4 = (int)(i - pixel * (Idx)C4);
    const Idx row = pixel / (Idx)Wo;
    const int wo = (int)(pixel - row * (Idx)Wo);
    const Idx b = row / (Idx)Ho;
    const int ho = (int)(row - b * (Idx)Ho);
    const float4 *img = in4 + (long)b * H * W * C4 + c4;
    float4 v;
    if (mode == MUMPY_RS_IDENTITY) {
      v = __ldg(img + ((long)ho * W + wo) * C4);
    } else if (mode == MUMPY_RS_AVGPOOL2) {
      const float4 *p = img + ((long)(2 * ho) * W + 2 * wo) * C4;
      const float4 a = __ldg(p), bb = __ldg(p + C4), c = __ldg(p + (long)W * C4), d = __ldg(p + (long)W * C4 + C4);
      v.x = (a.x + bb.x + c.x + d.x) * 0.25f; v.y = (a.y + bb.y + c.y + d.y) * 0.25f;
      v.z = (a.z + bb.z + c.z + d.z) * 0.25f; v.w = (a.w + bb.w + c.w + d.w) * 0.25f;
    } else {
      int y0, y1, x0, x1;
      float wy, wx;
      const bool aligned = mode == MUMPY_RS_UP_ALIGNED;
      bilinear_axis(ho, H, Ho, scale, aligned, y0, y1, wy);
      bilinear_axis(wo, W, Wo, scale, aligned, x0, x1, wx);
      const float4 v00 = __ldg(img + ((long)y0 * W + x0) * C4), v01 = __ldg(img + ((long)y0 * W + x1) * C4);
      const float4 v10 = __ldg(img + ((long)y1 * W + x0) * C4), v11 = __ldg(img + ((long)y1 * W + x1) * C4);
      // ATen upsample_bilinear2d: w0*(w0x*v00 + w1x*v01) + w1*(w0x*v10 + w1x*v11)
      const float ux = 1.0f - wx, uy = 1.0f - wy;
      v.x = uy * (ux * v00.x + wx * v01.x) + wy * (ux * v10.x + wx * v11.x);
      v.y = uy * (ux * v00.y + wx * v01.y) + wy * (ux * v10.y + wx * v11.y);
      v.z = uy * (ux * v00.z + wx * v01.z) + wy * (ux * v10.z + wx * v11.z);
      v.w = uy * (ux * v00.w + wx * v01.w) + wy * (ux * v10.w + wx * v11.w);
    }
    if (mul) {
      const float4 m = __ldg(reinterpret_cast<const float4 *>(mul) + i);
      v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
    }
    if (add) {
      const float4 m = __ldg(reinterpret_cast<const float4 *>(add) + i);
      v.x += m.x; v.y += m.y; v.z += m.z; v.w += m.w;
    }
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float4 *>(out + (long)pixel * ld_out + out_col + 4 * c4) = v;
    } else {
      if constexpr (is_half_t<OutT>::value) amax = fmaxf(fmaxf(amax, fabsf(v.x)), fmaxf(fmaxf(fabsf(v.y), fabsf(v.z)), fabsf(v.w)));
      uint2 pk;
      pk.x = pack2<OutT>(v.x, v.y);
      pk.y = pack2<OutT>(v.z, v.w);
      *reinterpret_cast<uint2 *>(out + (long)pixel * ld_out + out_col + 4 * c4) = pk;
    }
  }
  if constexpr (is_half_t<OutT>::value) f16_guard(amax);
}

// The same operation for C/4 dividing 256 (every decoder map): a CTA owns one output row, a thread one channel quad, and walks
// the row's pixels -- no index decomposition per element, the vertical weights once per CTA (the flat kernel above spends ~70
// instructions per float4 and is issue bound below 3 TB/s).
template <typename OutT>
__global__ void __launch_bounds__(256) resample_rows_kernel(const float *__restrict__ in, const float *__restrict__ mul, const float *__restrict__ add,
                                                            OutT *__restrict__ out, long ld_out, int out_col, int H, int W, int C4, int Ho, int Wo,
                                                            int mode, int scale) {
  pdl_grid_sync();
  const int b = blockIdx.x / Ho, ho = blockIdx.x - b * Ho;
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4, lanes_p = 256 / C4;
  const float4 *img = reinterpret_cast<const float4 *>(in) + (long)b * H * W * C4 + c4;
  const long orow = ((long)b * Ho + ho) * Wo;
  const bool aligned = mode == MUMPY_RS_UP_ALIGNED;
  int y0 = ho, y1 = ho;
  float wy = 0.0f;
  if (mode == MUMPY_RS_UP_ALIGNED || mode == MUMPY_RS_UP_HALFPIX) bilinear_axis(ho, H, Ho, scale, aligned, y0, y1, wy);
  const float uy = 1.0f - wy;
  const float4 *r0 = img + (long)(mode == MUMPY_RS_AVGPOOL2 ? 2 * ho : y0) * W * C4;
  const float4 *r1 = img + (long)(mode == MUMPY_RS_AVGPOOL2 ? 2 * ho + 1 : y1) * W * C4;
  [[maybe_unused]] float amax = 0.0f;
#pragma unroll 2
  for (int wo = pl; wo < Wo; wo += lanes_p) {
    float4 v;
    if (mode == MUMPY_RS_IDENTITY) {
      v = __ldg(r0 + (long)wo * C4);
    } else if (mode == MUMPY_RS_AVGPOOL2) {
      const float4 a = __ldg(r0 + (long)(2 * wo) * C4), bb = __ldg(r0 + (long)(2 * wo + 1) * C4);
      const float4 c = __ldg(r1 + (long)(2 * wo) * C4), d = __ldg(r1 + (long)(2 * wo + 1) * C4);
      v.x = (a.x + bb.x + c.x + d.x) * 0.25f; v.y = (a.y + bb.y + c.y + d.y) * 0.25f;
      v.z = (a.z + bb.z + c.z + d.z) * 0.25f; v.w = (a.w + bb.w + c.w + d.w) * 0.25f;
    } else {
      int x0, x1;
      float wx;
      bilinear_axis(wo, W, Wo, scale, aligned, x0, x1, wx);
      const float4 v00 = __ldg(r0 + (long)x0 * C4), v01 = __ldg(r0 + (long)x1 * C4);
      const float4 v10 = __ldg(r1 + (long)x0 * C4), v11 = __ldg(r1 + (long)x1 * C4);
      const float ux = 1.0f - wx;
      v.x = uy * (ux * v00.x + wx * v01.x) + wy * (ux * v10.x + wx * v11.x);
      v.y = uy * (ux * v00.y + wx * v01.y) + wy * (ux * v10.y + wx * v11.y);
      v.z = uy * (ux * v00.z + wx * v01.z) + wy * (ux * v10.z + wx * v11.z);
      v.w = uy * (ux * v00.w + wx * v01.w) + wy * (ux * v10.w + wx * v11.w);
    }
    const long i = (orow + wo) * C4 + c4;
    if (mul) {
      const float4 m = __ldg(reinterpret_cast<const float4 *>(mul) + i);
      v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
    }
    if (add) {
      const float4 m = __ldg(reinterpret_cast<const float4 *>(add) + i);
      v.x += m.x; v.y += m.y; v.z += m.z; v.w += m.w;
    }
    OutT *o = out + (orow + wo) * ld_out + out_col + 4 * c4;
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float4 *>(o) = v;
    } else {
      if constexpr (is_half_t<OutT>::value) amax = fmaxf(fmaxf(amax, fabsf(v.x)), fmaxf(fmaxf(fabsf(v.y), fabsf(v.z)), fabsf(v.w)));
      uint2 pk;
      pk.x = pack2<OutT>(v.x, v.y);
      pk.y = pack2<OutT>(v.z, v.w);
      *reinterpret_cast<uint2 *>(o) = pk;
    }
  }
  if constexpr (is_half_t<OutT>::value) f16_guard(amax);
}

template <typename OutT>
__global__ void __launch_bounds__(256) mul_add_vec_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b, const float4 *__restrict__ c,
                                                          OutT *__restrict__ out, long n4) {
  pdl_grid_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  [[maybe_unused]] float amax = 0.0f;
  for (; i < n4; i += stride) {
    const float4 x = __ldg(a + i), y = b ? __ldg(b + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 v = make_float4(x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w);
    if (c) {
      const float4 z = __ldg(c + i);
      v.x += z.x; v.y += z.y; v.z += z.z; v.w += z.w;
    }
    if constexpr (sizeof(OutT) == 4) {
      reinterpret_cast<float4 *>(out)[i] = v;
    } else {
      if constexpr (is_half_t<OutT>::value) amax = fmaxf(fmaxf(amax, fabsf(v.x)), fmaxf(fmaxf(fabsf(v.y), fabsf(v.z)), fabsf(v.w)));
      uint2 pk;
      pk.x = pack2<OutT>(v.x, v.y);
      pk.y = pack2<OutT>(v.z, v.w);
      reinterpret_cast<uint2 *>(out)[i] = pk;
    }
  }
  if constexpr (is_half_t<OutT>::value) f16_guard(amax);
}

__global__ void __launch_bounds__(256) mul_add_kernel(const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ c,
                                                      float *__restrict__ out, long n) {
  pdl_grid_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = a[i] * b[i];
    if (c) v += c[i];
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256) add_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, long n) {
  pdl_grid_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = a[i] + b[i];
}

// in (batch, R, Cc) with row stride ld_in -> out (batch, Cc, R) with row stride ld_out (+ column offset), tiled via smem
__global__ void __launch_bounds__(256) transpose_kernel(const float *__restrict__ in, long ld_in, long in_batch, float *__restrict__ out,
                                                        long ld_out, long out_batch, int out_col, int R, int Cc) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < R && c < Cc) tile[i][tx] = in[b * in_batch + (long)r * ld_in + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < R && c < Cc) out[b * out_batch + (long)c * ld_out + out_col + r] = tile[tx][i];
  }
}

// NCHW -> NHWC with a fused 2x2 average (ffinfo -> decoder_frequency_0 input)
__global__ void __launch_bounds__(256) nchw_pool2_to_nhwc_kernel(const float *__restrict__ in, float *__restrict__ out, long ld_out,
                                                                 int out_col, long total, int C, int H, int W) {
  pdl_grid_sync();
  const int Ho = H / 2, Wo = W / 2;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {       // i over (b, c, ho, wo): coalesced-ish reads
    const int wo = (int)(i % Wo);
    long t = i / Wo;
    const int ho = (int)(t % Ho);
    t /= Ho;
    const int c = (int)(t % C);
    const long b = t / C;
    const float *p = in + ((b * C + c) * H + 2 * ho) * W + 2 * wo;
    const float v = (p[0] + p[1] + p[W] + p[W + 1]) * 0.25f;
    out[((b * Ho + ho) * Wo + wo) * ld_out + out_col + c] = v;
  }
}

__global__ void __launch_bounds__(256) channel_group_mean_kernel(const float *__restrict__ in, float *__restrict__ out, long total, int C,
                                                                 int k) {
  pdl_grid_sync();
  const int Co = C / k;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int g = (int)(i % Co);
    const long p = i / Co;
    const float *src = in + p * C + g * k;
    float s = 0.0f;
    for (int j = 0; j < k; ++j) s += src[j];
    out[i] = s / (float)k;
  }
}

// Cout == 1 convolution (final_out, decoder.py:95): one thread per output pixel, float4 channel loads, weights in smem.
__global__ void __launch_bounds__(256) conv_cout1_kernel(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bias,
                                                         float *__restrict__ out, long pixels, int H, int W, int Cin, int kh, int kw, int ph,
                                                         int pw) {
  pdl_grid_sync();
  extern __shared__ float ws[];     // [kh*kw*Cin]
  for (int i = threadIdx.x; i < kh * kw * Cin; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= pixels) return;
  const int x = (int)(m % W);
  const long r = m / W;
  const int y = (int)(r % H);
  const long b = r / H;
  float acc = bias ? bias[0] : 0.0f;
  for (int ky = 0; ky < kh; ++ky) {
    const int yy = y + ky - ph;
    if (yy < 0 || yy >= H) continue;
    for (int kx = 0; kx < kw; ++kx) {
      const int xx = x + kx - pw;
      if (xx < 0 || xx >= W) continue;
      const float4 *src = reinterpret_cast<const float4 *>(in + ((b * H + yy) * W + xx) * Cin);
      const float *wt = ws + (ky * kw + kx) * Cin;
      for (int c = 0; c < Cin / 4; ++c) {
        const float4 v = src[c];
        acc = fmaf(v.x, wt[4 * c], acc);
        acc = fmaf(v.y, wt[4 * c + 1], acc);
        acc = fmaf(v.z, wt[4 * c + 2], acc);
        acc = fmaf(v.w, wt[4 * c + 3], acc);
      }
    }
  }
  out[m] = acc;
}

// Cout == 1, LPP = Cin/4 lanes per output pixel: every lane owns one float4 of channels over the kh*kw taps (coalesced
// 16-byte loads, neighbouring pixels re-hit L1), the partial sums are combined with log2(LPP) shuffles.
template <int LPP>
__global__ void __launch_bounds__(256) conv_cout1_vec_kernel(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bias,
                                                             float *__restrict__ out, long pixels, int H, int W, int kh, int kw, int ph, int pw) {
  pdl_grid_sync();
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long m = gid / LPP;
  const int sub = (int)(gid % LPP);
  const bool live = m < pixels;
  const long mm = live ? m : 0;
  const int x = (int)(mm % W);
  const long r = mm / W;
  const int y = (int)(r % H);
  const long b = r / H;
  const float4 *in4 = reinterpret_cast<const float4 *>(in);
  const float4 *w4 = reinterpret_cast<const float4 *>(w);
  float acc = 0.0f;
  for (int ky = 0; ky < kh; ++ky) {
    const int yy = y + ky - ph;
    if (yy < 0 || yy >= H) continue;
    for (int kx = 0; kx < kw; ++kx) {
      const int xx = x + kx - pw;
      if (xx < 0 || xx >= W) continue;
      const float4 v = __ldg(in4 + ((b * H + yy) * W + xx) * LPP + sub);
      const float4 wt = __ldg(w4 + (ky * kw + kx) * LPP + sub);
      acc = fmaf(v.x, wt.x, acc);
      acc = fmaf(v.y, wt.y, acc);
      acc = fmaf(v.z, wt.z, acc);
      acc = fmaf(v.w, wt.w, acc);
    }
  }
#pragma unroll
  for (int o = LPP / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (live && sub == 0) out[m] = acc + (bias ? bias[0] : 0.0f);
}

// Cout == 1, 3x3 'same': LPP lanes per pixel column, every thread produces COUT1_ROWS vertically adjacent outputs from one
// pass over the COUT1_ROWS + 2 input rows of its column (three float4 loads per row, each reused by up to three outputs; the
// nine weight float4 of its channel quad live in registers) -- half the loads and ~2.5x fewer instructions per output than
// one thread per (pixel, quad).  Four neighbouring columns per LPP = 8 warp: 512 contiguous bytes per load instruction.
constexpr int COUT1_ROWS = 4;
template <int LPP>
__global__ void __launch_bounds__(256, 4) conv_cout1_rows_kernel(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bias,
                                                              float *__restrict__ out, long units, int H, int W) {
  pdl_grid_sync();
  // 32-bit index decomposition (the host checks units * LPP < 2^31): five 64-bit divisions by run-time values were ~500 of the
  // ~740 instructions a thread executed (ncu: 74 M warp instructions for 14 M warp FMAs)
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned u = gid / LPP;                  // unit = (image, row tile, column)
  const int sub = (int)(gid % LPP);
  const bool live = u < (unsigned)units;
  const unsigned uu = live ? u : 0u;
  const unsigned r = uu / (unsigned)W;
  const int x = (int)(uu - r * (unsigned)W);
  const unsigned tiles = (unsigned)(H + COUT1_ROWS - 1) / COUT1_ROWS;
  const long b = r / tiles;
  const int y0 = (int)(r - (unsigned)b * tiles) * COUT1_ROWS;
  const float4 *in4 = reinterpret_cast<const float4 *>(in);
  float4 wt[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wt[t] = __ldg(reinterpret_cast<const float4 *>(w) + t * LPP + sub);
  float acc[COUT1_ROWS];
#pragma unroll
  for (int i = 0; i < COUT1_ROWS; ++i) acc[i] = 0.0f;
  // (rows outside the image are predicated loads, not branches: the compiler batches the loads of several rows -- the kernel was
  // latency-bound at 34 % occupancy with one row's loads in flight per thread, ncu: 72 % of the stall samples on long scoreboard)
#pragma unroll
  for (int ry = 0; ry < COUT1_ROWS + 2; ++ry) {
    const int yy = y0 + ry - 1;
    const bool in_y = yy >= 0 && yy < H;
    const float4 *row = in4 + ((b * H + (in_y ? yy : 0)) * W) * LPP + sub;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v[3];
    v[0] = (in_y && x > 0) ? __ldg(row + (long)(x - 1) * LPP) : zero;
    v[1] = in_y ? __ldg(row + (long)x * LPP) : zero;
    v[2] = (in_y && x + 1 < W) ? __ldg(row + (long)(x + 1) * LPP) : zero;
#pragma unroll
    for (int i = 0; i < COUT1_ROWS; ++i) {
      const int ky = ry - i;                     // input row ry feeds output row i through filter row ky
      if (ky < 0 || ky > 2) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 f = wt[ky * 3 + kx];
        acc[i] = fmaf(v[kx].x, f.x, acc[i]);
        acc[i] = fmaf(v[kx].y, f.y, acc[i]);
        acc[i] = fmaf(v[kx].z, f.z, acc[i]);
        acc[i] = fmaf(v[kx].w, f.w, acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < COUT1_ROWS; ++i) {
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  if (live && sub == 0) {
    const float bv = bias ? bias[0] : 0.0f;
#pragma unroll
    for (int i = 0; i < COUT1_ROWS; ++i)
      if (y0 + i < H) out[(b * H + y0 + i) * W + x] = acc[i] + bv;
  }
}

// Clip assembly (test.py:22-25 ToTensor + Normalize, universaldataloader.py:45-48 sliding window): frames are uint8 HWC
// images resident on the device, every clip names its T frames by index (consecutive clips share T-1 frames, so each
// frame is uploaded once); out[b,t,c,y,x] = (frame[idx[b,t]][y,x,c] / 255 - mean[c]) / std[c] with the same operation
// order as torchvision (true divisions), so the result is bit-identical to the reference's CPU transform.
// One thread per 4 horizontally adjacent pixels: 12 contiguous input bytes, one 16-byte store per channel.
__global__ void __launch_bounds__(256) assemble_clips_kernel(const unsigned char *__restrict__ frames, const int *__restrict__ clip_frames,
                                                             float *__restrict__ out, long total4, int T, int HW4, float m0, float m1,
                                                             float m2, float s0, float s1, float s2) {
  pdl_grid_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  for (; i < total4; i += stride) {
    const int p4 = (int)(i % HW4);
    const long bt = i / HW4;                                     // b*T + t
    const int f = __ldg(clip_frames + bt);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(frames + ((long)f * HW4 + p4) * 12);
    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
    const unsigned char px[12] = {(unsigned char)(w0), (unsigned char)(w0 >> 8), (unsigned char)(w0 >> 16), (unsigned char)(w0 >> 24),
                                  (unsigned char)(w1), (unsigned char)(w1 >> 8), (unsigned char)(w1 >> 16), (unsigned char)(w1 >> 24),
                                  (unsigned char)(w2), (unsigned char)(w2 >> 8), (unsigned char)(w2 >> 16), (unsigned char)(w2 >> 24)};
    float *dst = out + (bt * 3) * (long)HW4 * 4 + (long)p4 * 4;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float4 v;
      v.x = ((float)px[c] / 255.0f - mean[c]) / sd[c];
      v.y = ((float)px[3 + c] / 255.0f - mean[c]) / sd[c];
      v.z = ((float)px[6 + c] / 255.0f - mean[c]) / sd[c];
      v.w = ((float)px[9 + c] / 255.0f - mean[c]) / sd[c];
      *reinterpret_cast<float4 *>(dst + (long)c * HW4 * 4) = v;
    }
  }
}

__global__ void __launch_bounds__(256) mask_counts_kernel(const float *__restrict__ logits, const unsigned char *__restrict__ gt,
                                                          unsigned char *__restrict__ mask, unsigned long long *__restrict__ counts, int HW) {
  pdl_grid_sync();
  __shared__ unsigned int red[4][8];
  const int b = blockIdx.y;
  unsigned int tp = 0, np = 0, ng = 0, nu = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const bool p = logits[(long)b * HW + i] > 0.0f;      // sigmoid(x) > 0.5  <=>  x > 0   (test.py:100-106)
    if (mask) mask[(long)b * HW + i] = p ? 255 : 0;
    if (gt) {
      const bool g = gt[(long)b * HW + i] != 0;
      tp += (p && g);
      ng += g;
      nu += (p || g);
    }
    np += p;
  }
  unsigned int vals[4] = {tp, np, ng, nu};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    unsigned int v = vals[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[q][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4 && counts) {
    unsigned int s = 0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    if (s) atomicAdd(&counts[b * 4 + threadIdx.x], (unsigned long long)s);
  }
}

}  // namespace mumpy

using namespace mumpy;

extern "C" int mumpy_gather_rows(const float *src, int C, void *dst, int dst_dtype, long dst_ld, int dst_col, int B, int rows_out,
                                 int rows_src, int div, int mul_hi, int mul_lo, int add, void *stream) {
  MUMPY_REQUIRE(src && dst && C > 0 && B > 0 && rows_out > 0 && div > 0, "gather_rows: bad arguments");
  const long total = (long)B * rows_out * C;
  cudaStream_t st = as_stream(stream);
  if (C % 4 == 0 && dst_ld % 4 == 0 && dst_col % 4 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const long total4 = total / 4;
    if (is_16bit(dst_dtype))
      MUMPY_WITH_16(dst_dtype, T, launch_kernel(gather_rows_vec_kernel<T>, flat_blocks(total4), 256, 0, st, src, C / 4, static_cast<T *>(dst), dst_ld, dst_col, total4, rows_out, rows_src, div, mul_hi, mul_lo, add));
    else
      launch_kernel(gather_rows_vec_kernel<float>, flat_blocks(total4), 256, 0, st, src, C / 4, static_cast<float *>(dst), dst_ld, dst_col, total4, rows_out, rows_src, div, mul_hi, mul_lo, add);
    return launch_status("gather_rows_vec");
  }
  if (is_16bit(dst_dtype))
    MUMPY_WITH_16(dst_dtype, T, launch_kernel(gather_rows_kernel<T>, flat_blocks(total), 256, 0, st, src, C, static_cast<T *>(dst), dst_ld, dst_col, total, rows_out, rows_src, div, mul_hi, mul_lo, add));
  else
    launch_kernel(gather_rows_kernel<float>, flat_blocks(total), 256, 0, st, src, C, static_cast<float *>(dst), dst_ld, dst_col, total, rows_out, rows_src, div, mul_hi, mul_lo, add);
  return launch_status("gather_rows");
}

extern "C" int mumpy_im2col_nhwc(const float *in, long ld_in, void *out, int out_dtype, int B, int H, int W, int Cin, int kh, int kw,
                                 int ph, int pw, int Kpad, void *stream) {
  MUMPY_REQUIRE(in && out && Kpad >= kh * kw * Cin && is_16bit(out_dtype), "im2col_nhwc: bad arguments");
  const long total = (long)B * H * W * Kpad;
  MUMPY_WITH_16(out_dtype, T, launch_kernel(im2col_kernel<T>, flat_blocks(total), 256, 0, as_stream(stream), in, ld_in, static_cast<T *>(out), total, H, W, Cin, kh, kw, ph, pw, Kpad));
  return launch_status("im2col_nhwc");
}

template <typename OutT>
static void launch_resample_vec(const float *in, const float *mul, const float *add, void *out, long ld_out, int out_col, long total4, int H, int W, int C4,
                                int Ho, int Wo, int mode, int scale, cudaStream_t st) {
  if (total4 < (1l << 31) - (1l << 24))
    launch_kernel(resample_vec_kernel<uint32_t, OutT>, flat_blocks(total4), 256, 0, st, in, mul, add, static_cast<OutT *>(out), ld_out, out_col, total4, H, W, C4, Ho, Wo, mode, scale);
  else
    launch_kernel(resample_vec_kernel<long, OutT>, flat_blocks(total4), 256, 0, st, in, mul, add, static_cast<OutT *>(out), ld_out, out_col, total4, H, W, C4, Ho, Wo, mode, scale);
}

extern "C" int mumpy_resample_nhwc(const float *in, const float *mul, const float *add, void *out, int out_dtype, long ld_out, int out_col,
                                   int B, int H, int W, int C, int mode, int scale, void *stream) {
  MUMPY_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && C > 0, "resample_nhwc: bad arguments");
  int Ho = H, Wo = W, Co = C;
  switch (mode) {
    case MUMPY_RS_IDENTITY: break;
    case MUMPY_RS_UP_ALIGNED:
    case MUMPY_RS_UP_HALFPIX:
      MUMPY_REQUIRE(scale >= 1, "resample_nhwc: bad scale");
      Ho = H * scale; Wo = W * scale; break;
    case MUMPY_RS_AVGPOOL2:
      MUMPY_REQUIRE(H % 2 == 0 && W % 2 == 0, "resample_nhwc: odd map for avgpool2");
      Ho = H / 2; Wo = W / 2; break;
    case MUMPY_RS_PIXEL_SHUFFLE2:
      MUMPY_REQUIRE(C % 4 == 0, "resample_nhwc: pixel_shuffle needs C %% 4 == 0");
      Ho = 2 * H; Wo = 2 * W; Co = C / 4; break;
    default: set_error("resample_nhwc: unknown mode %d", mode); return MUMPY_ERR_ARG;
  }
  const long total = (long)B * Ho * Wo * Co;
  if (mode != MUMPY_RS_PIXEL_SHUFFLE2 && C % 4 == 0 && ld_out % 4 == 0 && out_col % 4 == 0 &&
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(mul) | reinterpret_cast<uintptr_t>(add)) & 15) == 0) {
    const int C4 = C / 4;
    if (C4 <= 256 && 256 % C4 == 0 && (long)B * Ho < (1l << 31)) {
      const unsigned rgrid = (unsigned)((long)B * Ho);
      if (out_dtype == MUMPY_F32) launch_kernel(resample_rows_kernel<float>, rgrid, 256, 0, as_stream(stream), in, mul, add, static_cast<float *>(out), ld_out, out_col, H, W, C4, Ho, Wo, mode, scale);
      else if (out_dtype == MUMPY_F16) launch_kernel(resample_rows_kernel<__half>, rgrid, 256, 0, as_stream(stream), in, mul, add, static_cast<__half *>(out), ld_out, out_col, H, W, C4, Ho, Wo, mode, scale);
      else launch_kernel(resample_rows_kernel<__nv_bfloat16>, rgrid, 256, 0, as_stream(stream), in, mul, add, static_cast<__nv_bfloat16 *>(out), ld_out, out_col, H, W, C4, Ho, Wo, mode, scale);
      return launch_status("resample_rows");
    }
    if (out_dtype == MUMPY_F32) launch_resample_vec<float>(in, mul, add, out, ld_out, out_col, total / 4, H, W, C / 4, Ho, Wo, mode, scale, as_stream(stream));
    else if (out_dtype == MUMPY_F16) launch_resample_vec<__half>(in, mul, add, out, ld_out, out_col, total / 4, H, W, C / 4, Ho, Wo, mode, scale, as_stream(stream));
    else launch_resample_vec<__nv_bfloat16>(in, mul, add, out, ld_out, out_col, total / 4, H, W, C / 4, Ho, Wo, mode, scale, as_stream(stream));
    return launch_status("resample_vec");
  }
  MUMPY_REQUIRE(out_dtype == MUMPY_F32, "resample_nhwc: 16-bit output needs C, ld_out, out_col multiples of 4, 16-byte aligned buffers and a mode other than pixel shuffle");
  launch_kernel(resample_kernel, flat_blocks(total), 256, 0, as_stream(stream), in, mul, add, static_cast<float *>(out), ld_out, out_col, total, H, W, C, Ho, Wo, Co, mode, scale);
  return launch_status("resample_nhwc");
}

extern "C" int mumpy_mul_add(const float *a, const float *b, const float *c, void *out, int out_dtype, long n, void *stream) {
  MUMPY_REQUIRE(a && b && out && n > 0, "mul_add: bad arguments");
  if (n % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const float4 *a4 = reinterpret_cast<const float4 *>(a), *b4 = reinterpret_cast<const float4 *>(b), *c4 = reinterpret_cast<const float4 *>(c);
    if (out_dtype == MUMPY_F32) launch_kernel(mul_add_vec_kernel<float>, flat_blocks(n / 4), 256, 0, as_stream(stream), a4, b4, c4, static_cast<float *>(out), n / 4);
    else if (out_dtype == MUMPY_F16) launch_kernel(mul_add_vec_kernel<__half>, flat_blocks(n / 4), 256, 0, as_stream(stream), a4, b4, c4, static_cast<__half *>(out), n / 4);
    else launch_kernel(mul_add_vec_kernel<__nv_bfloat16>, flat_blocks(n / 4), 256, 0, as_stream(stream), a4, b4, c4, static_cast<__nv_bfloat16 *>(out), n / 4);
    return launch_status("mul_add_vec");
  }
  MUMPY_REQUIRE(out_dtype == MUMPY_F32, "mul_add: 16-bit output needs n %% 4 == 0 and 16-byte aligned buffers");
  launch_kernel(mul_add_kernel, flat_blocks(n), 256, 0, as_stream(stream), a, b, c, static_cast<float *>(out), n);
  return launch_status("mul_add");
}

extern "C" int mumpy_add(const float *a, const float *b, float *out, long n, void *stream) {
  MUMPY_REQUIRE(a && b && out && n > 0, "add: bad arguments");
  if (n % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    // out = a * 1 + b through the fused kernel (b slot = nullptr means multiply by one)
    launch_kernel(mul_add_vec_kernel<float>, flat_blocks(n / 4), 256, 0, as_stream(stream), reinterpret_cast<const float4 *>(a), nullptr, reinterpret_cast<const float4 *>(b),
                                                                        out, n / 4);
    return launch_status("add_vec");
  }
  launch_kernel(add_kernel, flat_blocks(n), 256, 0, as_stream(stream), a, b, out, n);
  return launch_status("add");
}

extern "C" int mumpy_nchw_to_nhwc(const float *in, float *out, long ld_out, int out_col, int B, int C, int H, int W, int pool2,
                                  void *stream) {
  MUMPY_REQUIRE(in && out && B > 0, "nchw_to_nhwc: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (pool2) {
    MUMPY_REQUIRE(H % 2 == 0 && W % 2 == 0, "nchw_to_nhwc: odd map for pool2");
    const long total = (long)B * C * (H / 2) * (W / 2);
    launch_kernel(nchw_pool2_to_nhwc_kernel, flat_blocks(total), 256, 0, st, in, out, ld_out, out_col, total, C, H, W);
    return launch_status("nchw_pool2_to_nhwc");
  }
  const int P = H * W;
  dim3 grid((unsigned)cdiv(P, 32), (unsigned)cdiv(C, 32), (unsigned)B);   // in: (B, R=C, Cc=P)
  launch_kernel(transpose_kernel, grid, 256, 0, st, in, P, (long)C * P, out, ld_out, (long)P * ld_out, out_col, C, P);
  return launch_status("nchw_to_nhwc");
}

extern "C" int mumpy_nhwc_to_nchw(const float *in, long ld_in, float *out, int B, int C, int H, int W, void *stream) {
  MUMPY_REQUIRE(in && out && B > 0, "nhwc_to_nchw: bad arguments");
  const int P = H * W;
  dim3 grid((unsigned)cdiv(C, 32), (unsigned)cdiv(P, 32), (unsigned)B);   // in: (B, R=P, Cc=C)
  launch_kernel(transpose_kernel, grid, 256, 0, as_stream(stream), in, ld_in, (long)P * ld_in, out, P, (long)C * P, 0, P, C);
  return launch_status("nhwc_to_nchw");
}

extern "C" int mumpy_channel_group_mean(const float *in, float *out, long pixels, int C, int k, void *stream) {
  MUMPY_REQUIRE(in && out && pixels > 0 && k > 0 && C % k == 0, "channel_group_mean: bad arguments");
  const long total = pixels * (C / k);
  launch_kernel(channel_group_mean_kernel, flat_blocks(total), 256, 0, as_stream(stream), in, out, total, C, k);
  return launch_status("channel_group_mean");
}

extern "C" int mumpy_conv2d_nhwc_cout1(const float *in, const float *w, const float *bias, float *out, int B, int H, int W, int Cin,
                                       int kh, int kw, int ph, int pw, void *stream) {
  MUMPY_REQUIRE(in && w && out && B > 0 && Cin % 4 == 0 && kh * kw * Cin * 4 <= 48 * 1024, "conv2d_nhwc_cout1: bad arguments");
  const long pixels = (long)B * H * W;
  if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w)) & 15) == 0 && (Cin == 32 || Cin == 64 || Cin == 128 || Cin == 16)) {
    cudaStream_t st = as_stream(stream);
    const int lpp = Cin / 4;
    const long units_chk = (long)B * ((H + COUT1_ROWS - 1) / COUT1_ROWS) * W;
    if (kh == 3 && kw == 3 && ph == 1 && pw == 1 && (lpp == 8 || lpp == 4) && units_chk * lpp + 256 < (1l << 31)) {
      const long units = units_chk;
      const unsigned rgrid = (unsigned)cdiv(units * lpp, 256);
      if (lpp == 8) launch_kernel(conv_cout1_rows_kernel<8>, rgrid, 256, 0, st, in, w, bias, out, units, H, W);
      else launch_kernel(conv_cout1_rows_kernel<4>, rgrid, 256, 0, st, in, w, bias, out, units, H, W);
      return launch_status("conv2d_nhwc_cout1_rows");
    }
    const unsigned grid = (unsigned)cdiv(pixels * lpp, 256);
    if (lpp == 4) launch_kernel(conv_cout1_vec_kernel<4>, grid, 256, 0, st, in, w, bias, out, pixels, H, W, kh, kw, ph, pw);
    else if (lpp == 8) launch_kernel(conv_cout1_vec_kernel<8>, grid, 256, 0, st, in, w, bias, out, pixels, H, W, kh, kw, ph, pw);
    else if (lpp == 16) launch_kernel(conv_cout1_vec_kernel<16>, grid, 256, 0, st, in, w, bias, out, pixels, H, W, kh, kw, ph, pw);
    else launch_kernel(conv_cout1_vec_kernel<32>, grid, 256, 0, st, in, w, bias, out, pixels, H, W, kh, kw, ph, pw);
    return launch_status("conv2d_nhwc_cout1_vec");
  }
  launch_kernel(conv_cout1_kernel, (unsigned)cdiv(pixels, 256), 256, kh * kw * Cin * sizeof(float), as_stream(stream), in, w, bias, out, pixels, H, W,
                                                                                                           Cin, kh, kw, ph, pw);
  return launch_status("conv2d_nhwc_cout1");
}

extern "C" int mumpy_assemble_clips(const unsigned char *frames, const int *clip_frames, float *out, int B, int T, int H, int W,
                                    const float *mean3, const float *std3, void *stream) {
  MUMPY_REQUIRE(frames && clip_frames && out && mean3 && std3 && B > 0 && T > 0 && H > 0 && W > 0, "assemble_clips: bad arguments");
  MUMPY_REQUIRE((H * W) % 4 == 0 && ((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "assemble_clips: H*W must be a multiple of 4 and the buffers 16-byte aligned");
  const int HW4 = H * W / 4;
  const long total4 = (long)B * T * HW4;
  launch_kernel(assemble_clips_kernel, flat_blocks(total4), 256, 0, as_stream(stream), frames, clip_frames, out, total4, T, HW4, mean3[0], mean3[1],
                mean3[2], std3[0], std3[1], std3[2]);
  return launch_status("assemble_clips");
}

extern "C" int mumpy_mask_counts(const float *logits, const unsigned char *gt, unsigned char *mask, long long *counts, int B, int HW,
                                 void *stream) {
  MUMPY_REQUIRE(logits && B > 0 && HW > 0 && (mask || counts), "mask_counts: bad arguments");
  dim3 grid((unsigned)(cdiv(HW, 256) < 64 ? cdiv(HW, 256) : 64), (unsigned)B);
  launch_kernel(mask_counts_kernel, grid, 256, 0, as_stream(stream), logits, gt, mask, reinterpret_cast<unsigned long long *>(counts), HW);
  return launch_status("mask_counts");
}
