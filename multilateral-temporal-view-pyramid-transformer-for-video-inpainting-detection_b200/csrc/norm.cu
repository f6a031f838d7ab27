// LayerNorm (plain rows and the PatchMerging 2x2 gather) and NHWC GroupNorm + activation.
// Statistics are always fp32 two-pass (mean, then centred variance), like the reference's ATen kernels.
#include "common.cuh"

namespace mumpy {

struct PlainRows {
  const float *x;
  int C;
  static constexpr int kSegs = 1;
  __device__ __forceinline__ const float *seg(long row, int) const { return x + row * C; }
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
};

template <typename Rows, typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(Rows rows, const float *__restrict__ gamma, const float *__restrict__ beta,
                                                        OutT *__restrict__ out, long n_rows, int C, float eps) {
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
  for (int q = 0; q < Rows::kSegs; ++q) {
    const float *p = rows.seg(row, q);
    for (int i = lane; i < C; i += 32) {
      const int c = q * C + i;
      o[c] = from_f32<OutT>((p[i] - mean) * rstd * gamma[c] + beta[c]);
    }
  }
}

template <typename Rows>
static int launch_ln(Rows rows, const float *gamma, const float *beta, void *out, int out_dtype, long n_rows, int C, float eps,
                     cudaStream_t st) {
  const int warps = 8;
  dim3 grid((unsigned)cdiv(n_rows, warps));
  if (out_dtype == MUMPY_BF16)
    layernorm_kernel<Rows, __nv_bfloat16><<<grid, warps * 32, 0, st>>>(rows, gamma, beta, static_cast<__nv_bfloat16 *>(out), n_rows, C, eps);
  else
    layernorm_kernel<Rows, float><<<grid, warps * 32, 0, st>>>(rows, gamma, beta, static_cast<float *>(out), n_rows, C, eps);
  return launch_status("layernorm");
}

// ---- GroupNorm on NHWC: one CTA per (b, group) for the statistics, then a flat apply pass ----
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const float *__restrict__ x, float *__restrict__ stats, int HW, int C,
                                                              int groups, float eps) {
  __shared__ float red[8];
  __shared__ float bcast;
  const int bg = blockIdx.x;
  const int b = bg / groups, g = bg % groups;
  const int cg = C / groups;
  const float *base = x + (long)b * HW * C + g * cg;
  const long n = (long)HW * cg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float s = 0.0f;
  for (long e = threadIdx.x; e < n; e += blockDim.x) s += base[(e / cg) * C + (e % cg)];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bcast = t / (float)n;
  }
  __syncthreads();
  const float mean = bcast;
  float v = 0.0f;
  for (long e = threadIdx.x; e < n; e += blockDim.x) {
    const float d = base[(e / cg) * C + (e % cg)] - mean;
    v = fmaf(d, d, v);
  }
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int i = 0; i < 8; ++i) t += red[i];
    stats[2 * bg] = mean;
    stats[2 * bg + 1] = 1.0f / sqrtf(t / (float)n + eps);
  }
}

__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const float *__restrict__ x, const float *__restrict__ stats,
                                                              const float *__restrict__ gamma, const float *__restrict__ beta,
                                                              float *__restrict__ out, long ld_out, int out_col, long total, int HW,
                                                              int C, int groups, int act) {
  const int cg = C / groups;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int c = (int)(i % C);
    const long pix = i / C;
    const long b = pix / HW;
    const float *st = stats + 2 * (b * groups + c / cg);
    const float v = (x[i] - st[0]) * st[1] * gamma[c] + beta[c];
    out[pix * ld_out + out_col + c] = apply_act(v, act);
  }
}

}  // namespace mumpy

using namespace mumpy;

extern "C" int mumpy_layernorm(const float *x, const float *gamma, const float *beta, void *out, int out_dtype, long rows, int C,
                               float eps, void *stream) {
  MUMPY_REQUIRE(x && gamma && beta && out && rows > 0 && C > 0, "layernorm: bad arguments");
  return launch_ln(PlainRows{x, C}, gamma, beta, out, out_dtype, rows, C, eps, as_stream(stream));
}

extern "C" int mumpy_patch_merge_norm(const float *x, const float *gamma, const float *beta, void *out, int out_dtype, int B,
                                      int TH, int W, int C, float eps, void *stream) {
  MUMPY_REQUIRE(x && gamma && beta && out && B > 0 && TH % 2 == 0 && W % 2 == 0, "patch_merge_norm: bad arguments");
  const long rows = (long)B * (TH / 2) * (W / 2);
  return launch_ln(MergeRows{x, C, TH, W}, gamma, beta, out, out_dtype, rows, C, eps, as_stream(stream));
}

extern "C" int mumpy_groupnorm_nhwc(const float *x, const float *gamma, const float *beta, float *stats_ws, float *out,
                                    long ld_out, int out_col, int B, int HW, int C, int groups, float eps, int act,
                                    void *stream) {
  MUMPY_REQUIRE(x && gamma && beta && stats_ws && out && C % groups == 0, "groupnorm_nhwc: bad arguments");
  groupnorm_stats_kernel<<<B * groups, 256, 0, as_stream(stream)>>>(x, stats_ws, HW, C, groups, eps);
  int rc = launch_status("groupnorm_stats");
  if (rc) return rc;
  const long total = (long)B * HW * C;
  const int blocks = (int)(cdiv(total, 256) < 148 * 16 ? cdiv(total, 256) : 148 * 16);
  groupnorm_apply_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, stats_ws, gamma, beta, out, ld_out, out_col, total, HW, C, groups, act);
  return launch_status("groupnorm_apply");
}
