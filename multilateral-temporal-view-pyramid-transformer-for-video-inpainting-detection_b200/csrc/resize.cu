// Frame resize of the loader (dataloaders/universaldataset.py:68-79: PIL `img.resize(self.inputRes)` before ToTensor /
// Normalize) on the device, bit-identical to Pillow's 8-bit resampler so that uint8 frames can be uploaded at their native
// resolution and the CPU never touches pixels.
//
//   bicubic (Pillow >= 7 default): two separable passes, horizontal first, each output byte =
//       clip8((2^21 + sum_k src[first + k] * coef[k]) >> 22), the intermediate image rounded to uint8 like Pillow's.
//       Coefficients are 22-bit fixed point, built on the host in double precision exactly as Pillow's precompute_coeffs /
//       normalize_coeffs_8bpc do (mumpy_resize_taps); when down-scaling the cubic's support grows with the scale factor
//       (17 taps for 854 -> 224).
//   nearest (the default of the pillow==4.0.0 pinned by requirements.txt:9): index tables built on the host by repeated
//       double-precision addition like ImagingScaleAffine.
//
// Horizontal pass: persistent CTAs, coefficient table and one source row at a time in shared memory (16-byte loads), a thread
// per output pixel.  Vertical pass: four neighbouring bytes per thread (the taps of a row are uniform across the row).  Both
// read every source byte once from HBM; byte-granular fallbacks cover odd shapes.
#include <math.h>

#include "common.cuh"

namespace mumpy {

constexpr int RESIZE_PRECISION_BITS = 32 - 8 - 2;

// out[img][y][ox][c] = clip8(sum_k in[img][y][first(ox) + k][c] * coef[ox][k]);  rows = n * in_h
__global__ void resize_horizontal_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, const int *__restrict__ bounds,
                                         const int *__restrict__ coefs, int ksize, long rows, int in_w, int out_w, int C) {
  pdl_grid_sync();
  const long per_row = (long)out_w * C;
  const long total = rows * per_row;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long row = i / per_row;
    const int rem = (int)(i - row * per_row);
    const int ox = rem / C, c = rem - ox * C;
    const int first = __ldg(bounds + 2 * ox), n = __ldg(bounds + 2 * ox + 1);
    const uint8_t *src = in + (row * in_w + first) * C + c;
    const int *k = coefs + (long)ox * ksize;
    int acc = 1 << (RESIZE_PRECISION_BITS - 1);
    for (int t = 0; t < n; ++t) acc += (int)__ldg(src + (long)t * C) * __ldg(k + t);
    out[i] = (uint8_t)min(max(acc >> RESIZE_PRECISION_BITS, 0), 255);
  }
}

// Row-staged horizontal pass: persistent CTAs keep the whole coefficient table in shared memory and stage RESIZE_ROWS source
// rows at a time (they are contiguous in memory: one span copied with 16-byte loads starting at the enclosing 16-byte
// boundary, ~20 KB in flight per CTA), de-interleave them into channel planes (neighbouring pixels become neighbouring bytes:
// conflict-free tap reads) and produce the C channels of one output pixel per thread, every coefficient read once per pixel.
constexpr int RESIZE_ROWS = 2;      // (8 rows per iteration measured slower: fewer resident warps; the pass is instruction-bound)
template <int C>
__global__ void __launch_bounds__(256) resize_horizontal_rows_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                                                      const int *__restrict__ bounds, const int *__restrict__ coefs, int ksize,
                                                                      long rows, int in_w, int out_w, long in_bytes) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t rs_smem[];
  int *sc = reinterpret_cast<int *>(rs_smem);                    // [out_w][ksize]
  int *sb = sc + (long)out_w * ksize;                            // [out_w][2]
  // (offsets from the shared base in integer arithmetic, so that every access below stays a shared-memory instruction)
  const int raw_off = (out_w * (ksize + 2) * (int)sizeof(int) + 15) & ~15;
  uint8_t *sraw = rs_smem + raw_off;                              // staged source bytes, 16-byte aligned
  const int row_bytes = in_w * C;
  const int plane = (in_w + 3) & ~3;
  uint8_t *splanes = rs_smem + raw_off + ((RESIZE_ROWS * row_bytes + 15 + 16 + 15) & ~15);      // [row][channel][plane]
  for (int i = threadIdx.x; i < out_w * ksize; i += blockDim.x) sc[i] = __ldg(coefs + i);
  for (int i = threadIdx.x; i < 2 * out_w; i += blockDim.x) sb[i] = __ldg(bounds + i);
  const uintptr_t tensor_end = reinterpret_cast<uintptr_t>(in) + (uintptr_t)in_bytes;
  for (long row0 = (long)blockIdx.x * RESIZE_ROWS; row0 < rows; row0 += (long)gridDim.x * RESIZE_ROWS) {
    const int nr = (int)(rows - row0 < RESIZE_ROWS ? rows - row0 : RESIZE_ROWS);
    // absolute byte addresses: the copy starts at the 16-byte boundary at or below the span start (inside the same allocation:
    // either the previous row or, for a sliced tensor, the bytes before the slice) and never reads past the end of the tensor
    const uintptr_t start = reinterpret_cast<uintptr_t>(in) + (uintptr_t)(row0 * (long)row_bytes);
    const uintptr_t a0 = start & ~uintptr_t(15);
    const int off = (int)(start - a0);
    const uintptr_t end = start + (uintptr_t)((long)nr * row_bytes);
    const uintptr_t vec_end = (end + 15) & ~uintptr_t(15);
    const uintptr_t safe_end = vec_end <= tensor_end ? vec_end : (end & ~uintptr_t(15));
    __syncthreads();                                             // readers of the previous rows are done (and the tables are visible)
    const int n_vec = (int)((safe_end - a0) >> 4), n_bytes = (int)(end - a0);
    for (int v = threadIdx.x; v < n_vec; v += blockDim.x)
      reinterpret_cast<uint4 *>(sraw)[v] = __ldg(reinterpret_cast<const uint4 *>(a0) + v);
    for (int b = (n_vec << 4) + threadIdx.x; b < n_bytes; b += blockDim.x) sraw[b] = __ldg(reinterpret_cast<const uint8_t *>(a0) + b);
    __syncthreads();
    if (C > 1) {
      for (int r = 0; r < nr; ++r)
        for (int x = threadIdx.x; x < in_w; x += blockDim.x) {
#pragma unroll
          for (int c = 0; c < C; ++c) splanes[(r * C + c) * plane + x] = sraw[off + (r * in_w + x) * C + c];
        }
      __syncthreads();
    }
    for (int r = 0; r < nr; ++r)
      for (int ox = threadIdx.x; ox < out_w; ox += blockDim.x) {
        const int first = sb[2 * ox], n = sb[2 * ox + 1];
        const uint8_t *p = (C > 1 ? splanes + r * C * plane : sraw + off + r * in_w) + first;
        const int *k = sc + ox * ksize;
        int acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 1 << (RESIZE_PRECISION_BITS - 1);
#pragma unroll 4
        for (int t = 0; t < n; ++t) {
          const int w = k[t];
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] += (int)p[c * plane + t] * w;
        }
        uint8_t *dst = out + ((row0 + r) * out_w + ox) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) dst[c] = (uint8_t)min(max(acc[c] >> RESIZE_PRECISION_BITS, 0), 255);
      }
  }
}

// out[img][oy][x] = clip8(sum_k in[img][first(oy) + k][x] * coef[oy][k]);  x runs over the out_w * C bytes of a row, four per thread
__global__ void resize_vertical_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, const int *__restrict__ bounds,
                                       const int *__restrict__ coefs, int ksize, int n_img, int in_h, int out_h, long row_bytes) {
  pdl_grid_sync();
  const long row_words = row_bytes >> 2;
  const long total = (long)n_img * out_h * row_words;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long orow = i / row_words;
    const long x = (i - orow * row_words) << 2;
    const int img = (int)(orow / out_h), oy = (int)(orow - (long)img * out_h);
    const int first = __ldg(bounds + 2 * oy), n = __ldg(bounds + 2 * oy + 1);
    const uint8_t *src = in + ((long)img * in_h + first) * row_bytes + x;
    const int *k = coefs + (long)oy * ksize;
    int a0 = 1 << (RESIZE_PRECISION_BITS - 1), a1 = a0, a2 = a0, a3 = a0;
    for (int t = 0; t < n; ++t) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(src + (long)t * row_bytes));
      const int w = __ldg(k + t);
      a0 += (int)(v & 0xffu) * w;
      a1 += (int)((v >> 8) & 0xffu) * w;
      a2 += (int)((v >> 16) & 0xffu) * w;
      a3 += (int)(v >> 24) * w;
    }
    const uint32_t r = (uint32_t)min(max(a0 >> RESIZE_PRECISION_BITS, 0), 255) | ((uint32_t)min(max(a1 >> RESIZE_PRECISION_BITS, 0), 255) << 8) |
                       ((uint32_t)min(max(a2 >> RESIZE_PRECISION_BITS, 0), 255) << 16) | ((uint32_t)min(max(a3 >> RESIZE_PRECISION_BITS, 0), 255) << 24);
    *reinterpret_cast<uint32_t *>(out + orow * row_bytes + x) = r;
  }
}

// byte-granular variant for rows whose length is not a multiple of four
__global__ void resize_vertical_bytes_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, const int *__restrict__ bounds,
                                             const int *__restrict__ coefs, int ksize, int n_img, int in_h, int out_h, long row_bytes) {
  pdl_grid_sync();
  const long total = (long)n_img * out_h * row_bytes;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long orow = i / row_bytes;
    const long x = i - orow * row_bytes;
    const int img = (int)(orow / out_h), oy = (int)(orow - (long)img * out_h);
    const int first = __ldg(bounds + 2 * oy), n = __ldg(bounds + 2 * oy + 1);
    const uint8_t *src = in + ((long)img * in_h + first) * row_bytes + x;
    const int *k = coefs + (long)oy * ksize;
    int acc = 1 << (RESIZE_PRECISION_BITS - 1);
    for (int t = 0; t < n; ++t) acc += (int)__ldg(src + (long)t * row_bytes) * __ldg(k + t);
    out[i] = (uint8_t)min(max(acc >> RESIZE_PRECISION_BITS, 0), 255);
  }
}

__global__ void resize_nearest_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, const int *__restrict__ iy,
                                      const int *__restrict__ ix, int n_img, int in_h, int in_w, int out_h, int out_w, int C) {
  pdl_grid_sync();
  const long per_row = (long)out_w * C;
  const long total = (long)n_img * out_h * per_row;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long orow = i / per_row;
    const int rem = (int)(i - orow * per_row);
    const int ox = rem / C, c = rem - ox * C;
    const int img = (int)(orow / out_h), oy = (int)(orow - (long)img * out_h);
    out[i] = __ldg(in + (((long)img * in_h + __ldg(iy + oy)) * in_w + __ldg(ix + ox)) * C + c);
  }
}

static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

static unsigned grid_for(long total) {
  const long want = cdiv(total, 256);
  const long cap = 148l * 16;
  return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace mumpy

using namespace mumpy;

extern "C" int mumpy_resize_taps(int in_size, int out_size, int filter, int *bounds, int *coefs, int coef_capacity, int *ksize_out) {
  MUMPY_REQUIRE(in_size > 0 && out_size > 0 && bounds && ksize_out, "resize_taps: bad arguments");
  if (filter == MUMPY_RESIZE_NEAREST) {
    // ImagingScaleAffine: position accumulated by repeated addition; bounds[o] = source index (one int per output)
    const double scale = (double)in_size / out_size;
    double pos = 0.0 + scale * 0.5;
    for (int o = 0; o < out_size; ++o) {
      int i = pos < 0.0 ? -1 : (int)pos;
      if (i < 0) i = 0;
      if (i > in_size - 1) i = in_size - 1;
      bounds[o] = i;
      pos += scale;
    }
    *ksize_out = 1;
    return MUMPY_OK;
  }
  MUMPY_REQUIRE(filter == MUMPY_RESIZE_BICUBIC, "resize_taps: filter %d unsupported (0 nearest, 3 bicubic)", filter);
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  *ksize_out = ksize;
  if (!coefs) return MUMPY_OK;                 // size query
  MUMPY_REQUIRE(ksize <= 64, "resize_taps: %d taps (down-scaling by more than ~15x is not supported)", ksize);
  MUMPY_REQUIRE((long)coef_capacity >= (long)out_size * ksize, "resize_taps: coefficient buffer holds %d ints, %ld needed", coef_capacity, (long)out_size * ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int *k = coefs + (long)xx * ksize;
    double w[64];
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      k[x] = v < 0 ? (int)(-0.5 + v * (1 << RESIZE_PRECISION_BITS)) : (int)(0.5 + v * (1 << RESIZE_PRECISION_BITS));
    }
    for (int x = xmax; x < ksize; ++x) k[x] = 0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return MUMPY_OK;
}

extern "C" int mumpy_resize_u8(const unsigned char *in, unsigned char *out, unsigned char *tmp, int n, int in_h, int in_w, int out_h, int out_w,
                               int channels, int filter, const int *bounds_h, const int *coefs_h, int ksize_h, const int *bounds_v,
                               const int *coefs_v, int ksize_v, void *stream) {
  MUMPY_REQUIRE(in && out && n > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && channels >= 1 && channels <= 4, "resize_u8: bad arguments");
  MUMPY_REQUIRE((long)n * in_h * in_w * channels < (1l << 40), "resize_u8: input too large");
  cudaStream_t st = as_stream(stream);
  if (filter == MUMPY_RESIZE_NEAREST) {
    MUMPY_REQUIRE(bounds_h && bounds_v, "resize_u8(nearest): index tables missing");
    const long total = (long)n * out_h * out_w * channels;
    launch_kernel(resize_nearest_kernel, grid_for(total), 256, 0, st, in, out, bounds_v, bounds_h, n, in_h, in_w, out_h, out_w, channels);
    return launch_status("resize_nearest");
  }
  MUMPY_REQUIRE(filter == MUMPY_RESIZE_BICUBIC, "resize_u8: filter %d unsupported (0 nearest, 3 bicubic)", filter);
  const bool horiz = in_w != out_w, vert = in_h != out_h;      // Pillow skips a pass whose size does not change
  if (!horiz && !vert) {
    cudaError_t e = cudaMemcpyAsync(out, in, (size_t)n * in_h * in_w * channels, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) {
      set_error("resize_u8: copy: %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    return MUMPY_OK;
  }
  MUMPY_REQUIRE(!horiz || (bounds_h && coefs_h && ksize_h > 0), "resize_u8: horizontal taps missing");
  MUMPY_REQUIRE(!vert || (bounds_v && coefs_v && ksize_v > 0), "resize_u8: vertical taps missing");
  MUMPY_REQUIRE(!(horiz && vert) || tmp, "resize_u8: a two-pass resize needs the n*in_h*out_w*channels byte workspace");
  const unsigned char *vsrc = in;
  if (horiz) {
    unsigned char *hdst = vert ? tmp : out;
    const long rows = (long)n * in_h;
    const size_t smem = (size_t)out_w * (ksize_h + 2) * sizeof(int) + 16 + ((size_t)RESIZE_ROWS * in_w * channels + 64) +
                        (size_t)RESIZE_ROWS * channels * (in_w + 4);
    if (smem <= 200 * 1024 && (channels == 1 || channels == 3 || channels == 4)) {
      const long groups = cdiv(rows, RESIZE_ROWS);
      const long per_sm = (220 * 1024) / (long)(smem + 1024) < 1 ? 1 : (220 * 1024) / (long)(smem + 1024);
      const unsigned grid = (unsigned)(groups < 148l * per_sm ? groups : 148l * per_sm);
      const long in_bytes = rows * in_w * channels;
      cudaError_t e = cudaSuccess;
#define RESIZE_ROWS(C_)                                                                                                                       \
  {                                                                                                                                          \
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(resize_horizontal_rows_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess)                                                                                                                    \
      launch_kernel(resize_horizontal_rows_kernel<C_>, grid, 256, smem, st, in, hdst, bounds_h, coefs_h, ksize_h, rows, in_w, out_w, in_bytes); \
  }
      if (channels == 1) RESIZE_ROWS(1) else if (channels == 3) RESIZE_ROWS(3) else RESIZE_ROWS(4)
#undef RESIZE_ROWS
      if (e != cudaSuccess) {
        set_error("resize_u8: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return MUMPY_ERR_CUDA;
      }
    } else {
      const long total = rows * out_w * channels;
      launch_kernel(resize_horizontal_kernel, grid_for(total), 256, 0, st, in, hdst, bounds_h, coefs_h, ksize_h, rows, in_w, out_w, channels);
    }
    vsrc = hdst;
  }
  if (vert) {
    const long row_bytes = (long)out_w * channels;
    const bool words = row_bytes % 4 == 0 && ((reinterpret_cast<uintptr_t>(vsrc) | reinterpret_cast<uintptr_t>(out)) & 3) == 0;
    const long total = (long)n * out_h * (words ? row_bytes / 4 : row_bytes);
    if (words) launch_kernel(resize_vertical_kernel, grid_for(total), 256, 0, st, vsrc, out, bounds_v, coefs_v, ksize_v, n, in_h, out_h, row_bytes);
    else launch_kernel(resize_vertical_bytes_kernel, grid_for(total), 256, 0, st, vsrc, out, bounds_v, coefs_v, ksize_v, n, in_h, out_h, row_bytes);
  }
  return launch_status("resize_u8");
}
