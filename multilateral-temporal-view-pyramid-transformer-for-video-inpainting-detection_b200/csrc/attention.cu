// Attention cores: Swin (shifted-)window attention with relative-position bias and mask, the short-sequence
// ViT attention of the temporal "global" encoder, and the deformable cross-view attention core.
// Window partition / reverse / cyclic shift are index maps on the token canvas (window_token_row), never copies.
// Scores, softmax and P.V are fp32; q/k/v are read as stored (fp32 or bf16).
#include "common.cuh"

namespace mumpy {

// smem layout for one (window, head): q,k,v [N][D+1] and scores [N][N+1]
template <int D>
__host__ __device__ constexpr int attn_smem_floats(int N) { return 3 * N * (D + 1) + N * (N + 1); }

// softmax over rows of S (N x N, row stride N+1); one warp per row
__device__ __forceinline__ void softmax_rows(float *S, int N, int warp, int nwarps, int lane) {
  for (int i = warp; i < N; i += nwarps) {
    float *row = S + i * (N + 1);
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float s = 0.0f;
    for (int j = lane; j < N; j += 32) {
      const float e = expf(row[j] - m);
      row[j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int j = lane; j < N; j += 32) row[j] *= inv;
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(128) window_attention_kernel(const T *__restrict__ qkv, const float *__restrict__ bias,
                                                               const float *__restrict__ mask, T *__restrict__ out, int TH, int W,
                                                               int C, int ws, int shift) {
  pdl_grid_sync();
  extern __shared__ float sm[];
  const int N = ws * ws;
  float *q = sm, *k = q + N * (D + 1), *v = k + N * (D + 1), *S = v + N * (D + 1);
  const int nW = (TH / ws) * (W / ws);
  const int win = blockIdx.x;
  const int b = win / nW, n = win % nW;
  const int h = blockIdx.y;
  const long L = (long)TH * W;
  const float qscale = (D == 64) ? 0.125f : 0.17677669529663687f;   // head_dim ** -0.5 rounded to fp32
  for (int e = threadIdx.x; e < N * D; e += blockDim.x) {
    const int p = e / D, d = e % D;
    const long row = b * L + window_token_row(n, p, TH, W, ws, shift);
    const T *src = qkv + row * 3 * C + h * D + d;
    q[p * (D + 1) + d] = to_f32(src[0]) * qscale;
    k[p * (D + 1) + d] = to_f32(src[C]);
    v[p * (D + 1) + d] = to_f32(src[2 * C]);
  }
  __syncthreads();
  const float *bh = bias + (long)h * N * N;
  const float *mw = mask ? mask + (long)n * N * N : nullptr;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e / N, j = e % N;
    float acc = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) acc = fmaf(q[i * (D + 1) + d], k[j * (D + 1) + d], acc);
    acc += bh[e];
    if (mw) acc += mw[e];
    S[i * (N + 1) + j] = acc;
  }
  __syncthreads();
  softmax_rows(S, N, threadIdx.x >> 5, blockDim.x >> 5, threadIdx.x & 31);
  __syncthreads();
  for (int e = threadIdx.x; e < N * D; e += blockDim.x) {
    const int i = e / D, d = e % D;
    float acc = 0.0f;
    for (int j = 0; j < N; ++j) acc = fmaf(S[i * (N + 1) + j], v[j * (D + 1) + d], acc);
    const long row = b * L + window_token_row(n, i, TH, W, ws, shift);
    out[row * C + h * D + d] = from_f32<T>(acc);
  }
}

// blocks.py:55-70 for N <= 8 tokens: one warp per (head, query); lanes span the head dim.
template <typename T>
__global__ void __launch_bounds__(128) mha_short_kernel(const T *__restrict__ qkv, T *__restrict__ out, int N, int C, int heads) {
  pdl_grid_sync();
  const long bn = blockIdx.x;
  const int d = C / heads;
  const float scale = 1.0f / sqrtf((float)d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const T *base = qkv + bn * N * 3 * C;
  for (int pair = warp; pair < heads * N; pair += nwarps) {
    const int h = pair / N, i = pair % N;
    float s[8];
    const T *qi = base + (long)i * 3 * C + h * d;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] = -INFINITY;
      if (j < N) {
        const T *kj = base + (long)j * 3 * C + C + h * d;
        float a = 0.0f;
        for (int dd = lane; dd < d; dd += 32) a = fmaf(to_f32(qi[dd]), to_f32(kj[dd]), a);
        s[j] = warp_sum(a) * scale;
      }
    }
    float m = s[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) m = fmaxf(m, s[j]);
    float den = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] = (j < N) ? expf(s[j] - m) : 0.0f;
      den += s[j];
    }
    const float inv = 1.0f / den;
    for (int dd = lane; dd < d; dd += 32) {
      float a = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < N) a = fmaf(s[j] * inv, to_f32(base[(long)j * 3 * C + 2 * C + h * d + dd]), a);
      out[(bn * N + i) * C + h * d + dd] = from_f32<T>(a);
    }
  }
}

// N == 3 tokens (the temporal "global" encoder): LP = d*sizeof(T)/16 lanes per (sequence, head); each lane owns one
// 16-byte chunk of the head dimension of all nine q/k/v slices (nine independent 16-byte loads), partial dot products are
// combined with log2(LP) shuffles, and the lane writes its chunk of the three output rows.
template <typename T, int LP>
__global__ void __launch_bounds__(256) mha3_kernel(const T *__restrict__ qkv, T *__restrict__ out, long n_pairs, int C, int heads, float scale) {
  pdl_grid_sync();
  constexpr int E = 16 / sizeof(T);                 // elements per 16-byte chunk
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long pair = gid / LP;
  const int sub = (int)(gid % LP);
  const bool live = pair < n_pairs;
  const long bn = live ? pair / heads : 0;
  const int h = live ? (int)(pair - bn * heads) : 0;
  const int d = C / heads;
  const T *base = qkv + bn * 3 * 3 * C + h * d + sub * E;
  float x[9][E];                                    // [token*3 + which][e]
#pragma unroll
  for (int s = 0; s < 9; ++s) {
    const uint4 u = live ? __ldg(reinterpret_cast<const uint4 *>(base + (long)(s / 3) * 3 * C + (s % 3) * C)) : make_uint4(0, 0, 0, 0);
    if (sizeof(T) == 4) {
      x[s][0] = __uint_as_float(u.x); x[s][1] = __uint_as_float(u.y); x[s][2 % E] = __uint_as_float(u.z); x[s][3 % E] = __uint_as_float(u.w);
    } else {
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if constexpr (sizeof(T) == 2) {
          const float2 f = unpack2<T>(w[e]);
          x[s][(2 * e) % E] = f.x;
          x[s][(2 * e + 1) % E] = f.y;
        }
      }
    }
  }
  float sc[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = 0.0f;
#pragma unroll
      for (int e = 0; e < E; ++e) a = fmaf(x[i * 3][e], x[j * 3 + 1][e], a);
      sc[i][j] = a;
    }
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) sc[i][j] += __shfl_xor_sync(0xffffffffu, sc[i][j], o);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float s0 = sc[i][0] * scale, s1 = sc[i][1] * scale, s2 = sc[i][2] * scale;
    const float m = fmaxf(s0, fmaxf(s1, s2));
    const float p0 = expf(s0 - m), p1 = expf(s1 - m), p2 = expf(s2 - m);
    const float inv = 1.0f / (p0 + p1 + p2);
    float o[E];
#pragma unroll
    for (int e = 0; e < E; ++e) o[e] = fmaf(p0 * inv, x[2][e], fmaf(p1 * inv, x[5][e], (p2 * inv) * x[8][e]));
    if (live) {
      T *dst = out + (bn * 3 + i) * C + h * d + sub * E;
      uint4 u;
      if (sizeof(T) == 4) {
        u = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2 % E]), __float_as_uint(o[3 % E]));
      } else {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if constexpr (sizeof(T) == 2) w[e] = pack2<T>(o[(2 * e) % E], o[(2 * e + 1) % E]);
          else w[e] = 0;
        }
        u = make_uint4(w[0], w[1], w[2], w[3]);
      }
      *reinterpret_cast<uint4 *>(dst) = u;
    }
  }
}

// query window paired with kv window j (deformableAttention.py:329-330 + :394)
__device__ __forceinline__ int cva_query_window(int j, int r, int N1, int nW1, int per_clip) {
  if (!per_clip) return j % N1;
  const int i = j / r;
  const int clip = i / nW1;
  return clip * nW1 + ((i % nW1) * r + j % r) % nW1;
}

template <typename TKV, typename TO, int D>
__global__ void __launch_bounds__(128) cva_attention_kernel(const float *__restrict__ q, const TKV *__restrict__ kv, TO *__restrict__ o,
                                                            int N1, int TH1, int W, int C, int ws, int r, int per_clip) {
  pdl_grid_sync();
  extern __shared__ float sm[];
  const int N = ws * ws;
  float *qs = sm, *ks = qs + N * (D + 1), *vs = ks + N * (D + 1), *S = vs + N * (D + 1);
  const int i = blockIdx.x;      // output window
  const int h = blockIdx.y;
  const int nW1 = (TH1 / ws) * (W / ws);
  const long L1 = (long)TH1 * W;
  const float scale = (D == 64) ? 0.125f : 0.17677669529663687f;
  // each thread owns the same (p,d) outputs across t: accumulate in registers
  constexpr int kPerThread = (64 * D + 127) / 128;
  float acc[kPerThread];
#pragma unroll
  for (int u = 0; u < kPerThread; ++u) acc[u] = 0.0f;
  for (int t = 0; t < r; ++t) {
    const int j = r * i + t;
    const int qw = cva_query_window(j, r, N1, nW1, per_clip);
    const int qb = qw / nW1, qn = qw % nW1;
    __syncthreads();
    for (int e = threadIdx.x; e < N * D; e += blockDim.x) {
      const int p = e / D, d = e % D;
      const long qrow = qb * L1 + window_token_row(qn, p, TH1, W, ws, 0);
      qs[p * (D + 1) + d] = q[qrow * C + h * D + d];
      const TKV *src = kv + ((long)j * N + p) * 2 * C + h * D + d;
      ks[p * (D + 1) + d] = to_f32(src[0]);
      vs[p * (D + 1) + d] = to_f32(src[C]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
      const int a = e / N, b = e % N;
      float s = 0.0f;
#pragma unroll
      for (int d = 0; d < D; ++d) s = fmaf(qs[a * (D + 1) + d], ks[b * (D + 1) + d], s);
      S[a * (N + 1) + b] = s * scale;
    }
    __syncthreads();
    softmax_rows(S, N, threadIdx.x >> 5, blockDim.x >> 5, threadIdx.x & 31);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
      const int e = threadIdx.x + u * 128;
      if (e < N * D) {
        const int a = e / D, d = e % D;
        float s = 0.0f;
        for (int b = 0; b < N; ++b) s = fmaf(S[a * (N + 1) + b], vs[b * (D + 1) + d], s);
        acc[u] += s;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kPerThread; ++u) {
    const int e = threadIdx.x + u * 128;
    if (e < N * D) {
      const int p = e / D, d = e % D;
      o[((long)i * N + p) * C + h * D + d] = from_f32<TO>(acc[u]);
    }
  }
}

template <typename K>
static int ensure_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return MUMPY_OK;
  static const void *seen[16];          // keyed by entry address: instantiations may share one function-pointer type
  static size_t granted[16];
  static int n_seen = 0;
  const void *key = reinterpret_cast<const void *>(kernel);
  int slot = -1;
  for (int i = 0; i < n_seen; ++i)
    if (seen[i] == key) slot = i;
  if (slot >= 0 && bytes <= granted[slot]) return MUMPY_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e));
    return MUMPY_ERR_CUDA;
  }
  if (slot < 0 && n_seen < 16) slot = n_seen++;
  if (slot >= 0) {
    seen[slot] = key;
    granted[slot] = bytes;
  }
  return MUMPY_OK;
}

int window_attention_mma(const void *qkv, const float *bias, const float *mask, const float *rel_table, int standard_mask, void *out, int dtype,
                         int B, int TH, int W, int C, int heads, int ws, int shift, cudaStream_t st);
int cva_attention_mma(const float *q, const void *kv, void *o, int dtype, int B, int TH1, int TH2, int W, int C, int heads, int ws, int per_clip,
                      cudaStream_t st);

}  // namespace mumpy

using namespace mumpy;

extern "C" int mumpy_window_attention(const void *qkv, const float *bias, const float *mask, const float *rel_table, int standard_mask,
                                      void *out, int dtype, int B, int TH, int W, int C, int heads, int ws, int shift, void *stream) {
  MUMPY_REQUIRE(qkv && bias && out && B > 0 && heads > 0 && C % heads == 0, "window_attention: bad arguments");
  MUMPY_REQUIRE(TH % ws == 0 && W % ws == 0 && ws * ws <= 64 && shift >= 0 && shift < ws, "window_attention: bad window geometry");
  const int D = C / heads;
  MUMPY_REQUIRE(D == 32 || D == 64, "window_attention: head dim %d unsupported (32 or 64)", D);
  const int N = ws * ws;
  dim3 grid((unsigned)(B * (TH / ws) * (W / ws)), (unsigned)heads);
  cudaStream_t st = as_stream(stream);
  if (is_16bit(dtype) && D == 32 && C % 8 == 0)      // tensor-pipe path (attention_mma.cu)
    return window_attention_mma(qkv, bias, mask, rel_table, standard_mask, out, dtype, B, TH, W, C, heads, ws, shift, st);
  int rc = MUMPY_OK;
#define LAUNCH(T, DD)                                                                                                     \
  {                                                                                                                       \
    const size_t smem = attn_smem_floats<DD>(N) * sizeof(float);                                                          \
    rc = ensure_smem(window_attention_kernel<T, DD>, smem);                                                               \
    if (rc) return rc;                                                                                                    \
    launch_kernel(window_attention_kernel<T, DD>, grid, 128, smem, st, static_cast<const T *>(qkv), bias, mask, static_cast<T *>(out), TH, W, C, ws, shift); \
  }
  if (dtype == MUMPY_BF16) {
    if (D == 32) LAUNCH(__nv_bfloat16, 32) else LAUNCH(__nv_bfloat16, 64)
  } else if (dtype == MUMPY_F16) {
    if (D == 32) LAUNCH(__half, 32) else LAUNCH(__half, 64)
  } else {
    if (D == 32) LAUNCH(float, 32) else LAUNCH(float, 64)
  }
#undef LAUNCH
  return launch_status("window_attention");
}

extern "C" int mumpy_mha_short(const void *qkv, void *out, int dtype, long Bn, int N, int C, int heads, void *stream) {
  MUMPY_REQUIRE(qkv && out && Bn > 0 && N > 0 && N <= 8 && C % heads == 0, "mha_short: bad arguments (N=%d)", N);
  cudaStream_t st = as_stream(stream);
  const int d = C / heads;
  const int esz = is_16bit(dtype) ? 2 : 4;
  const int lp = d * esz / 16;
  const bool aligned = ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 && (d * esz) % 16 == 0;
  if (N == 3 && aligned && (lp == 4 || lp == 8 || lp == 16)) {
    const long n_pairs = Bn * heads;
    const float scale = 1.0f / sqrtf((float)d);
    const unsigned grid = (unsigned)cdiv(n_pairs * lp, 256);
#define MHA3(T, LP) launch_kernel(mha3_kernel<T, LP>, grid, 256, 0, st, static_cast<const T *>(qkv), static_cast<T *>(out), n_pairs, C, heads, scale)
    if (dtype == MUMPY_BF16) {
      if (lp == 4) MHA3(__nv_bfloat16, 4); else if (lp == 8) MHA3(__nv_bfloat16, 8); else MHA3(__nv_bfloat16, 16);
    } else if (dtype == MUMPY_F16) {
      if (lp == 4) MHA3(__half, 4); else if (lp == 8) MHA3(__half, 8); else MHA3(__half, 16);
    } else {
      if (lp == 4) MHA3(float, 4); else if (lp == 8) MHA3(float, 8); else MHA3(float, 16);
    }
#undef MHA3
    return launch_status("mha3");
  }
  if (dtype == MUMPY_F16)
    launch_kernel(mha_short_kernel<__half>, (unsigned)Bn, 128, 0, st, static_cast<const __half *>(qkv), static_cast<__half *>(out), N, C, heads);
  else if (dtype == MUMPY_BF16)
    launch_kernel(mha_short_kernel<__nv_bfloat16>, (unsigned)Bn, 128, 0, st, static_cast<const __nv_bfloat16 *>(qkv), static_cast<__nv_bfloat16 *>(out), N, C, heads);
  else
    launch_kernel(mha_short_kernel<float>, (unsigned)Bn, 128, 0, st, static_cast<const float *>(qkv), static_cast<float *>(out), N, C, heads);
  return launch_status("mha_short");
}

// ---- attention maps (diagnostic outputs of the reference: blocks.py:66-68,88-89; deformableAttention.py:364,389,399; the hot path never
// materialises them).  Plain fp32 arithmetic on the same q / k operands the fused kernels read. -------------------------------------
template <typename T>
__device__ __forceinline__ float probs_load(const T *p) { return to_f32(*p); }
template <>
__device__ __forceinline__ float probs_load<float>(const float *p) { return *p; }

// softmax(q k^T * d^-1/2) of short sequences: probs (Bn, heads, N, N) fp32 from qkv (Bn, N, 3C); one thread per (sequence, head, row)
template <typename T>
__global__ void __launch_bounds__(128) mha_short_probs_kernel(const T *__restrict__ qkv, float *__restrict__ probs, long n_rows, int N, int C, int heads) {
  pdl_grid_sync();
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_rows) return;
  const int i = (int)(gid % N);
  const long bh = gid / N;
  const int h = (int)(bh % heads);
  const long bn = bh / heads;
  const int d = C / heads;
  const float scale = 1.0f / sqrtf((float)d);
  const T *q = qkv + (bn * N + i) * 3 * C + h * d;
  float sc[8];
  float m = -3.0e38f;
  for (int j = 0; j < N; ++j) {
    const T *k = qkv + (bn * N + j) * 3 * C + C + h * d;
    float a = 0.0f;
    for (int e = 0; e < d; ++e) a = fmaf(probs_load(q + e), probs_load(k + e), a);
    sc[j] = a * scale;
    m = fmaxf(m, sc[j]);
  }
  float sum = 0.0f;
  for (int j = 0; j < N; ++j) {
    sc[j] = expf(sc[j] - m);
    sum += sc[j];
  }
  float *o = probs + gid * N;
  for (int j = 0; j < N; ++j) o[j] = sc[j] / sum;
}

// deformable cross-view attention map: probs (N2, heads, P, P) fp32 (= the reference's (N1, r * heads, P, P) view) from the canvas-ordered
// fp32 queries and the window-major sampled k (kv rows of 2C, k first); one thread per (kv window, head, query pixel)
template <typename T>
__global__ void __launch_bounds__(64) cva_attention_probs_kernel(const float *__restrict__ q, const T *__restrict__ kv, float *__restrict__ probs, int N1, int TH1,
                                                                int W, int C, int heads, int ws, int r, int per_clip) {
  pdl_grid_sync();
  const int P = ws * ws;
  const int j = blockIdx.x, h = blockIdx.y, i = threadIdx.x;
  if (i >= P) return;
  const int nW1 = (TH1 / ws) * (W / ws);
  const int qw = cva_query_window(j, r, N1, nW1, per_clip);
  const long qb = qw / nW1;
  const int qn = qw - (int)qb * nW1;
  const long L1 = (long)TH1 * W;
  const float *qrow = q + (qb * L1 + window_token_row(qn, i, TH1, W, ws, 0)) * C + h * 32;
  float qv[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) qv[e] = qrow[e];
  float *o = probs + (((long)j * heads + h) * P + i) * P;
  float m = -3.0e38f;
  for (int p = 0; p < P; ++p) {
    const T *k = kv + ((long)j * P + p) * 2 * C + h * 32;
    float a = 0.0f;
#pragma unroll
    for (int e = 0; e < 32; ++e) a = fmaf(qv[e], probs_load(k + e), a);
    a *= 0.17677669529663687f;
    o[p] = a;
    m = fmaxf(m, a);
  }
  float sum = 0.0f;
  for (int p = 0; p < P; ++p) {
    const float e = expf(o[p] - m);
    o[p] = e;
    sum += e;
  }
  const float inv = 1.0f / sum;
  for (int p = 0; p < P; ++p) o[p] *= inv;
}

extern "C" int mumpy_mha_short_probs(const void *qkv, float *probs, int dtype, long Bn, int N, int C, int heads, void *stream) {
  MUMPY_REQUIRE(qkv && probs && Bn > 0 && N > 0 && N <= 8 && heads > 0 && C % heads == 0, "mha_short_probs: bad arguments (N=%d)", N);
  cudaStream_t st = as_stream(stream);
  const long n_rows = Bn * heads * N;
  const unsigned grid = (unsigned)cdiv(n_rows, 128);
  if (dtype == MUMPY_F16) launch_kernel(mha_short_probs_kernel<__half>, grid, 128, 0, st, static_cast<const __half *>(qkv), probs, n_rows, N, C, heads);
  else if (dtype == MUMPY_BF16) launch_kernel(mha_short_probs_kernel<__nv_bfloat16>, grid, 128, 0, st, static_cast<const __nv_bfloat16 *>(qkv), probs, n_rows, N, C, heads);
  else launch_kernel(mha_short_probs_kernel<float>, grid, 128, 0, st, static_cast<const float *>(qkv), probs, n_rows, N, C, heads);
  return launch_status("mha_short_probs");
}

extern "C" int mumpy_cva_attention_probs(const float *q, const void *kv, int kv_dtype, float *probs, int B, int TH1, int TH2, int W, int C, int heads, int ws,
                                         int per_clip_pairing, void *stream) {
  MUMPY_REQUIRE(q && kv && probs && B > 0 && heads > 0 && C % heads == 0 && C / heads == 32, "cva_attention_probs: bad arguments (head dim must be 32)");
  MUMPY_REQUIRE(TH1 % ws == 0 && TH2 % TH1 == 0 && W % ws == 0 && ws * ws <= 64, "cva_attention_probs: bad window geometry");
  const int N1 = B * (TH1 / ws) * (W / ws);
  const int r = TH2 / TH1;
  dim3 grid((unsigned)(N1 * r), (unsigned)heads);
  cudaStream_t st = as_stream(stream);
  if (kv_dtype == MUMPY_F16) launch_kernel(cva_attention_probs_kernel<__half>, grid, 64, 0, st, q, static_cast<const __half *>(kv), probs, N1, TH1, W, C, heads, ws, r, per_clip_pairing);
  else if (kv_dtype == MUMPY_BF16)
    launch_kernel(cva_attention_probs_kernel<__nv_bfloat16>, grid, 64, 0, st, q, static_cast<const __nv_bfloat16 *>(kv), probs, N1, TH1, W, C, heads, ws, r, per_clip_pairing);
  else launch_kernel(cva_attention_probs_kernel<float>, grid, 64, 0, st, q, static_cast<const float *>(kv), probs, N1, TH1, W, C, heads, ws, r, per_clip_pairing);
  return launch_status("cva_attention_probs");
}

extern "C" int mumpy_cva_attention(const float *q, const void *kv, int kv_dtype, void *o, int out_dtype, int B, int TH1, int TH2,
                                   int W, int C, int heads, int ws, int per_clip_pairing, void *stream) {
  MUMPY_REQUIRE(q && kv && o && B > 0 && C % heads == 0 && C / heads == 32, "cva_attention: bad arguments (head dim must be 32)");
  MUMPY_REQUIRE(TH1 % ws == 0 && TH2 % TH1 == 0 && W % ws == 0 && ws * ws <= 64, "cva_attention: bad window geometry");
  MUMPY_REQUIRE(kv_dtype == out_dtype, "cva_attention: kv and output dtypes must match");
  const int N = ws * ws;
  const int N1 = B * (TH1 / ws) * (W / ws);
  const int r = TH2 / TH1;
  dim3 grid((unsigned)N1, (unsigned)heads);
  const size_t smem = attn_smem_floats<32>(N) * sizeof(float);
  cudaStream_t st = as_stream(stream);
  if (is_16bit(kv_dtype) && C % 8 == 0) return cva_attention_mma(q, kv, o, kv_dtype, B, TH1, TH2, W, C, heads, ws, per_clip_pairing, st);
  if (kv_dtype == MUMPY_F16)
    launch_kernel(cva_attention_kernel<__half, __half, 32>, grid, 128, smem, st, q, static_cast<const __half *>(kv), static_cast<__half *>(o), N1, TH1, W, C, ws, r, per_clip_pairing);
  else if (kv_dtype == MUMPY_BF16)
    launch_kernel(cva_attention_kernel<__nv_bfloat16, __nv_bfloat16, 32>, grid, 128, smem, st, q, static_cast<const __nv_bfloat16 *>(kv), static_cast<__nv_bfloat16 *>(o), N1, TH1, W, C, ws, r, per_clip_pairing);
  else
    launch_kernel(cva_attention_kernel<float, float, 32>, grid, 128, smem, st, q, static_cast<const float *>(kv), static_cast<float *>(o), N1, TH1, W, C, ws, r, per_clip_pairing);
  return launch_status("cva_attention");
}
