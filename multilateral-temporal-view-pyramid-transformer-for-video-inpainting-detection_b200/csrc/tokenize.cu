// CrossThreeViewTokenize for one view (multiTemporalViewEncoder.py:605-618): Conv3d(3->C, kernel=stride=(kt,4,4))
// is a GEMM over 3*kt*16-element patches (K = 144/96/48), followed by LayerNorm(C).  HBM-bound: each CTA
// stages 8 horizontally adjacent patches through shared memory with coalesced 128 B row reads.
#include "common.cuh"

namespace mumpy {

constexpr int TOK_PER_CTA = 8;

__global__ void __launch_bounds__(128) tokenize_kernel(const float *__restrict__ x, const float *__restrict__ w_kc,
                                                       const float *__restrict__ bias, const float *__restrict__ gamma,
                                                       const float *__restrict__ beta, float *__restrict__ out, int T, int S, int kt,
                                                       int C, float eps) {
  pdl_grid_sync();
  extern __shared__ float sm[];
  const int K = 3 * kt * 16;
  float *patch = sm;                       // [TOK][K]
  float *vals = sm + TOK_PER_CTA * K;      // [TOK][C]
  const int Hp = S / 4;
  const int To = T / kt;
  const int groups_per_row = (Hp + TOK_PER_CTA - 1) / TOK_PER_CTA;
  int gidx = blockIdx.x;
  const int wg = gidx % groups_per_row; gidx /= groups_per_row;
  const int hp = gidx % Hp; gidx /= Hp;
  const int to = gidx % To;
  const int b = gidx / To;
  const int wp0 = wg * TOK_PER_CTA;
  const int ntok = min(TOK_PER_CTA, Hp - wp0);

  // patch element k = ((c*kt + dt)*4 + dy)*4 + dx  (weight layout (C,3,kt,4,4))
  for (int e = threadIdx.x; e < 3 * kt * 4 * TOK_PER_CTA * 4; e += blockDim.x) {
    const int dx = e % 4;
    const int tok = (e / 4) % TOK_PER_CTA;
    const int khi = e / (4 * TOK_PER_CTA);        // (c*kt + dt)*4 + dy
    const int dy = khi % 4;
    const int cdt = khi / 4;
    const int dt = cdt % kt, c = cdt / kt;
    float v = 0.0f;
    if (tok < ntok)
      v = x[((((long)b * T + to * kt + dt) * 3 + c) * S + hp * 4 + dy) * S + (wp0 + tok) * 4 + dx];
    patch[tok * K + khi * 4 + dx] = v;
  }
  __syncthreads();
  const int o = threadIdx.x;
  if (o < C) {
    float acc[TOK_PER_CTA];
#pragma unroll
    for (int t = 0; t < TOK_PER_CTA; ++t) acc[t] = 0.0f;
    for (int k = 0; k < K; ++k) {
      const float w = w_kc[(long)k * C + o];
#pragma unroll
      for (int t = 0; t < TOK_PER_CTA; ++t) acc[t] = fmaf(patch[t * K + k], w, acc[t]);
    }
    const float bo = bias[o];
#pragma unroll
    for (int t = 0; t < TOK_PER_CTA; ++t) vals[t * C + o] = acc[t] + bo;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < ntok; t += 4) {
    const float *v = vals + t * C;
    float s = 0.0f;
    for (int i = lane; i < C; i += 32) s += v[i];
    const float mean = warp_sum(s) / C;
    float q = 0.0f;
    for (int i = lane; i < C; i += 32) {
      const float d = v[i] - mean;
      q = fmaf(d, d, q);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
    float *dst = out + (((long)b * To + to) * Hp * Hp + (long)hp * Hp + wp0 + t) * C;
    for (int i = lane; i < C; i += 32) dst[i] = (v[i] - mean) * rstd * gamma[i] + beta[i];
  }
}

}  // namespace mumpy

using namespace mumpy;

extern "C" int mumpy_tokenize(const float *x, const float *w_kc, const float *bias, const float *gamma, const float *beta,
                              float *out, int B, int T, int S, int kt, int C, float eps, void *stream) {
  MUMPY_REQUIRE(x && w_kc && bias && gamma && beta && out && B > 0 && S % 4 == 0 && kt >= 1 && kt <= T, "tokenize: bad arguments");
  MUMPY_REQUIRE(C <= 128, "tokenize: C=%d > 128 unsupported", C);
  const int Hp = S / 4, To = T / kt, K = 3 * kt * 16;
  const int groups_per_row = (Hp + TOK_PER_CTA - 1) / TOK_PER_CTA;
  const long ctas = (long)B * To * Hp * groups_per_row;
  const size_t smem = (size_t)TOK_PER_CTA * (K + C) * sizeof(float);
  launch_kernel(tokenize_kernel, (unsigned)ctas, 128, smem, as_stream(stream), x, w_kc, bias, gamma, beta, out, T, S, kt, C, eps);
  return launch_status("tokenize");
}
