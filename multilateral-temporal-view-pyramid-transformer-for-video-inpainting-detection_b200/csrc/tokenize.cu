// CrossThreeViewTokenize for one view (multiTemporalViewEncoder.py:605-618): Conv3d(3->C, kernel=stride=(kt,4,4))
// is a GEMM over 3*kt*16-element patches (K = 144/96/48), followed by LayerNorm(C).  fp32 throughout (the tokens seed the
// fp32 residual stream).  Persistent CTAs: the (K x C) filter matrix stays in shared memory for the CTA's lifetime, each
// iteration stages one row of S/4 patches (3*kt*4 contiguous image rows, 16-byte coalesced reads) as a K x S/4 tile; a warp
// owns 8 neighbouring tokens, a lane 4 output channels: per k one 16-byte filter read + two broadcast 16-byte patch reads
// feed 32 FMAs.  LayerNorm statistics are warp reductions over the channel lanes; stores are 16 bytes, 4C bytes per token
// contiguous.
#include "common.cuh"

namespace mumpy {

constexpr int TOK_PER_WARP = 8;

__global__ void __launch_bounds__(512) tokenize_kernel(const float *__restrict__ x, const float *__restrict__ w_kc,
                                                        const float *__restrict__ bias, const float *__restrict__ gamma,
                                                        const float *__restrict__ beta, float *__restrict__ out, int T, int S, int kt,
                                                        int C, float eps, long n_rows) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float sm[];
  const int K = 3 * kt * 16;
  const int Hp = S / 4;
  const int Hpp = (Hp + TOK_PER_WARP - 1) / TOK_PER_WARP * TOK_PER_WARP;      // patch-tile row length (tokens, padded)
  float *wsm = sm;                         // [K][C]
  float *patch = sm + (size_t)K * C;       // [K][Hpp]
  const int To = T / kt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C4 = C >> 2;
  for (int i = threadIdx.x; i < K * C4; i += blockDim.x) reinterpret_cast<float4 *>(wsm)[i] = __ldg(reinterpret_cast<const float4 *>(w_kc) + i);
  for (int i = threadIdx.x; i < K * (Hpp - Hp); i += blockDim.x) patch[(i / (Hpp - Hp)) * Hpp + Hp + i % (Hpp - Hp)] = 0.0f;   // padded tokens
  const bool live = lane < C4;
  float4 bo = make_float4(0.f, 0.f, 0.f, 0.f), g4 = bo, b4 = bo;
  if (live) {
    bo = __ldg(reinterpret_cast<const float4 *>(bias) + lane);
    g4 = __ldg(reinterpret_cast<const float4 *>(gamma) + lane);
    b4 = __ldg(reinterpret_cast<const float4 *>(beta) + lane);
  }
  const int S4 = S >> 2;
  for (long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int hp = (int)(row % Hp);
    const long t2 = row / Hp;
    const int to = (int)(t2 % To);
    const long b = t2 / To;
    __syncthreads();                       // previous iteration's readers are done with `patch` (and wsm is visible)
    // patch element k = ((c*kt + dt)*4 + dy)*4 + dx  (weight layout (C,3,kt,4,4)); image row (c, dt, dy) holds dx fastest
    for (int e = threadIdx.x; e < 3 * kt * 4 * S4; e += blockDim.x) {
      const int wp = e % S4;
      const int khi = e / S4;              // (c*kt + dt)*4 + dy
      const int dy = khi & 3;
      const int cdt = khi >> 2;
      const int dt = cdt % kt, c = cdt / kt;
      const float4 v = __ldg(reinterpret_cast<const float4 *>(x + ((((long)b * T + to * kt + dt) * 3 + c) * S + hp * 4 + dy) * S) + wp);
      float *dst = patch + (khi * 4) * Hpp + wp;
      dst[0] = v.x;
      dst[Hpp] = v.y;
      dst[2 * Hpp] = v.z;
      dst[3 * Hpp] = v.w;
    }
    __syncthreads();
    const int tok0 = warp * TOK_PER_WARP;
    if (tok0 >= Hp) continue;              // (whole warps only; they still take part in the barriers above)
    float acc[TOK_PER_WARP][4];
#pragma unroll
    for (int t = 0; t < TOK_PER_WARP; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.0f;
    if (live) {
      const float4 *w4 = reinterpret_cast<const float4 *>(wsm) + lane;
      const float4 *p4 = reinterpret_cast<const float4 *>(patch + tok0);
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        const float4 w = w4[(size_t)k * C4];
        const float4 pa = p4[(size_t)k * (Hpp >> 2)], pb = p4[(size_t)k * (Hpp >> 2) + 1];
        const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int t = 0; t < TOK_PER_WARP; ++t) {
          acc[t][0] = fmaf(pv[t], w.x, acc[t][0]);
          acc[t][1] = fmaf(pv[t], w.y, acc[t][1]);
          acc[t][2] = fmaf(pv[t], w.z, acc[t][2]);
          acc[t][3] = fmaf(pv[t], w.w, acc[t][3]);
        }
      }
    }
    float s[TOK_PER_WARP];
#pragma unroll
    for (int t = 0; t < TOK_PER_WARP; ++t) {
      acc[t][0] += bo.x; acc[t][1] += bo.y; acc[t][2] += bo.z; acc[t][3] += bo.w;
      s[t] = live ? (acc[t][0] + acc[t][1]) + (acc[t][2] + acc[t][3]) : 0.0f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int t = 0; t < TOK_PER_WARP; ++t) s[t] += __shfl_xor_sync(0xffffffffu, s[t], o);
    float q[TOK_PER_WARP];
#pragma unroll
    for (int t = 0; t < TOK_PER_WARP; ++t) {
      const float mean = s[t] / C;
      s[t] = mean;
      const float d0 = acc[t][0] - mean, d1 = acc[t][1] - mean, d2 = acc[t][2] - mean, d3 = acc[t][3] - mean;
      q[t] = live ? (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) : 0.0f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int t = 0; t < TOK_PER_WARP; ++t) q[t] += __shfl_xor_sync(0xffffffffu, q[t], o);
    if (live) {
      float *orow = out + (((long)b * To + to) * Hp * Hp + (long)hp * Hp + tok0) * C + 4 * lane;
#pragma unroll
      for (int t = 0; t < TOK_PER_WARP; ++t) {
        if (tok0 + t < Hp) {
          const float rstd = 1.0f / sqrtf(q[t] / C + eps);
          float4 y;
          y.x = (acc[t][0] - s[t]) * rstd * g4.x + b4.x;
          y.y = (acc[t][1] - s[t]) * rstd * g4.y + b4.y;
          y.z = (acc[t][2] - s[t]) * rstd * g4.z + b4.z;
          y.w = (acc[t][3] - s[t]) * rstd * g4.w + b4.w;
          *reinterpret_cast<float4 *>(orow + (long)t * C) = y;
        }
      }
    }
  }
}

// Tensor-core tokenizer, step 1: the patches of one view as a GEMM operand.  x (B,T,3,S,S) fp32 -> P (B*To*(S/4)^2, 3K) in the
// 16-bit operand type, K = 3*kt*16 in the order (c, dt, dy, dx) of Conv3d's weight.reshape(C, -1), each row = [hi | hi | lo]
// with v = hi + lo (the split of the FAF passes: against a weight row [w_hi | w_lo | w_hi] the three partial products give
// ~22 mantissa bits, so the tokens that seed the fp32 residual stream keep fp32 accuracy).  One CTA per row of S/4 patches:
// the 3*kt*4 image rows it covers are staged with coalesced 16-byte reads (row pitch S + 4 floats: conflict-free 16-byte reads
// across filter rows), then written token by token, 8 contiguous bytes per thread.
template <typename T16>
__global__ void __launch_bounds__(256) patchify16_kernel(const float *__restrict__ x, T16 *__restrict__ out, int T, int S, int kt, long n_rows) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float rows[];      // [3*kt*4][S + 4]
  const int G = 3 * kt * 4;                          // image rows (c, dt, dy) = groups of 4 consecutive K entries
  const int K = G * 4;
  const int Hp = S / 4, To = T / kt, pitch = S + 4, S4 = S >> 2;
  float amax = 0.0f;
  for (long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int hp = static_cast<int>(row % Hp);
    const long t2 = row / Hp;
    const int to = static_cast<int>(t2 % To);
    const long b = t2 / To;
    __syncthreads();
    for (int i = threadIdx.x; i < G * S4; i += blockDim.x) {
      const int g = i / S4, c4 = i - g * S4;
      const int c = g / (kt * 4), dt = (g / 4) % kt, dy = g & 3;
      const float4 v = __ldg(reinterpret_cast<const float4 *>(x + (((b * T + to * kt + dt) * 3 + c) * S + 4 * hp + dy) * (long)S) + c4);
      *reinterpret_cast<float4 *>(rows + g * pitch + 4 * c4) = v;
    }
    __syncthreads();
    T16 *o = out + row * Hp * 3 * K;
    for (int i = threadIdx.x; i < Hp * G; i += blockDim.x) {
      const int pw = i / G, g = i - pw * G;
      const float4 v = *reinterpret_cast<const float4 *>(rows + g * pitch + 4 * pw);
      if (is_half_t<T16>::value) amax = fmaxf(fmaxf(amax, fabsf(v.x)), fmaxf(fmaxf(fabsf(v.y), fabsf(v.z)), fabsf(v.w)));
      const T16 h0 = from_f32<T16>(v.x), h1 = from_f32<T16>(v.y), h2 = from_f32<T16>(v.z), h3 = from_f32<T16>(v.w);
      uint2 hi, lo;
      hi.x = pack2<T16>(to_f32(h0), to_f32(h1));
      hi.y = pack2<T16>(to_f32(h2), to_f32(h3));
      lo.x = pack2<T16>(v.x - to_f32(h0), v.y - to_f32(h1));
      lo.y = pack2<T16>(v.z - to_f32(h2), v.w - to_f32(h3));
      T16 *t = o + (long)pw * 3 * K + 4 * g;
      *reinterpret_cast<uint2 *>(t) = hi;
      *reinterpret_cast<uint2 *>(t + K) = hi;
      *reinterpret_cast<uint2 *>(t + 2 * K) = lo;
    }
  }
  if (is_half_t<T16>::value) f16_guard(amax);
}

}  // namespace mumpy

using namespace mumpy;

extern "C" int mumpy_patchify16(const float *x, void *out, int out_dtype, int B, int T, int S, int kt, void *stream) {
  MUMPY_REQUIRE(x && out && B > 0 && S % 4 == 0 && kt >= 1 && kt <= T && is_16bit(out_dtype), "patchify16: bad arguments (S must be a multiple of 4)");
  MUMPY_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "patchify16: buffers must be 16-byte aligned");
  const int Hp = S / 4, To = T / kt, G = 3 * kt * 4;
  const size_t smem = (size_t)G * (S + 4) * sizeof(float);
  MUMPY_REQUIRE(smem <= 200 * 1024, "patchify16: S=%d too large", S);
  static bool granted[2] = {false, false};
  const int gi = out_dtype == MUMPY_F16 ? 0 : 1;
  if (smem > 48 * 1024 && !granted[gi]) {
    cudaError_t e = gi == 0 ? cudaFuncSetAttribute(patchify16_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                            : cudaFuncSetAttribute(patchify16_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      set_error("patchify16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    granted[gi] = true;
  }
  const long n_rows = (long)B * To * Hp;
  const long ctas = n_rows < 148l * 8 ? n_rows : 148l * 8;
  MUMPY_WITH_16(out_dtype, T16, launch_kernel(patchify16_kernel<T16>, (unsigned)ctas, 256, smem, as_stream(stream), x, static_cast<T16 *>(out), T, S, kt, n_rows));
  return launch_status("patchify16");
}

extern "C" int mumpy_tokenize(const float *x, const float *w_kc, const float *bias, const float *gamma, const float *beta,
                              float *out, int B, int T, int S, int kt, int C, float eps, void *stream) {
  MUMPY_REQUIRE(x && w_kc && bias && gamma && beta && out && B > 0 && S % 4 == 0 && kt >= 1 && kt <= T, "tokenize: bad arguments (S must be a multiple of 4)");
  MUMPY_REQUIRE(C <= 128 && C % 4 == 0, "tokenize: C=%d unsupported (multiple of 4, <= 128)", C);
  MUMPY_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_kc) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(gamma) |
                  reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "tokenize: buffers must be 16-byte aligned");
  const int Hp = S / 4, To = T / kt, K = 3 * kt * 16;
  const int warps = (Hp + TOK_PER_WARP - 1) / TOK_PER_WARP;
  MUMPY_REQUIRE(warps <= 16, "tokenize: S=%d too large (<= 512)", S);
  const int Hpp = warps * TOK_PER_WARP;
  const size_t smem = (size_t)K * (C + Hpp) * sizeof(float);
  MUMPY_REQUIRE(smem <= 227 * 1024, "tokenize: tile of %zu B does not fit shared memory", smem);
  static size_t granted = 0;
  if (smem > 48 * 1024 && smem > granted) {
    cudaError_t e = cudaFuncSetAttribute(tokenize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("tokenize: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
    granted = smem;
  }
  const long n_rows = (long)B * To * Hp;
  const int per_sm = smem > 110 * 1024 ? 1 : 2;
  const long ctas = n_rows < 148l * per_sm ? n_rows : 148l * per_sm;
  launch_kernel(tokenize_kernel, (unsigned)ctas, warps * 32, smem, as_stream(stream), x, w_kc, bias, gamma, beta, out, T, S, kt, C, eps, n_rows);
  return launch_status("tokenize");
}
