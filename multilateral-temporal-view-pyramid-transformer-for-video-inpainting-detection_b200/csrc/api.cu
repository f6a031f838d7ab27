// Error plumbing, init and the dtype dispatch of mumpy_linear.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace mumpy {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_pdl = -1;      // -1: read MUMPY_PDL (default on) at first use
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char *v = getenv("MUMPY_PDL");
    g_pdl = (v && v[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}

int launch_status(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MUMPY_ERR_CUDA;
  }
  return MUMPY_OK;
}

int resolve_driver_entry_points();
void set_gemm_pair_mode(int mode);
void set_attention_tc(int enabled);
void set_gemm_tile_override(int bn);
int linear_f32(const float *A, long lda, const float *W, const float *bias, const float *residual, float *out, long ldo,
               long M, int N, int K, int act, cudaStream_t st);
int linear_bf16(const void *A, long lda, const void *W, const float *bias, const float *residual, void *out, void *aux, long ldo,
                long M, int N, int K, int ab_dtype, int out_dtype, int act, cudaStream_t st);

bool mlp_fused_supported(int C);
void set_mlp_fused_shape(int shape);
int mlp_fused_16(const float *x, const float *gamma, const float *beta, float eps, const void *W1, const float *b1, const void *W2, const float *b2,
                 float *out, long M, int C, int w_dtype, cudaStream_t st);
bool ln_linear_supported(int N, int K);
void set_ln_linear_pair_mode(int mode);
int ln_linear_16(const float *x, const float *gamma, const float *beta, float eps, const void *W, const float *bias, void *out, long ldo,
                 long M, int N, int K, int w_dtype, int act, cudaStream_t st);

int conv_bf16(const void *in, long ld_in, const void *wpk, const float *bias, const float *residual, void *out, long ldo, int B,
              int H, int W, int Cin, int Cout, int kh, int kw, int ph, int pw, int in_dtype, int out_dtype, int act, float *splitk_ws,
              long splitk_ws_bytes, cudaStream_t st);

// setters of the per-translation-unit fp16 overflow flag pointers (common.cuh); filled by static initialisers
static F16FlagSetter *flag_setters(int **count) {
  static F16FlagSetter fns[32];
  static int n = 0;
  *count = &n;
  return fns;
}
void register_f16_flag_setter(F16FlagSetter fn) {
  int *n;
  F16FlagSetter *fns = flag_setters(&n);
  if (*n < 32) fns[(*n)++] = fn;
}

template <typename T>
__global__ void cast16_kernel(const float *__restrict__ in, T *__restrict__ out, long n) {
  pdl_grid_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  float amax = 0.0f;
  for (; i < n; i += stride) {
    const float v = in[i];
    amax = fmaxf(amax, fabsf(v));
    out[i] = from_f32<T>(v);
  }
  if (is_half_t<T>::value) f16_guard(amax);
}

template <typename T>
__global__ void cast16_vec_kernel(const float4 *__restrict__ in, uint2 *__restrict__ out, long n4) {
  pdl_grid_sync();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  float amax = 0.0f;
  for (; i < n4; i += stride) {
    const float4 v = __ldg(in + i);
    amax = fmaxf(fmaxf(amax, fabsf(v.x)), fmaxf(fmaxf(fabsf(v.y), fabsf(v.z)), fabsf(v.w)));
    uint2 pk;
    pk.x = pack2<T>(v.x, v.y);
    pk.y = pack2<T>(v.z, v.w);
    out[i] = pk;
  }
  if (is_half_t<T>::value) f16_guard(amax);
}

}  // namespace mumpy

using namespace mumpy;

extern "C" int mumpy_abi_version(void) { return 1; }

extern "C" const char *mumpy_last_error(void) { return g_err; }

extern "C" int mumpy_set_gemm_pair_mode(int mode) {
  set_gemm_pair_mode(mode);
  return MUMPY_OK;
}

extern "C" int mumpy_set_attention_tc(int enabled) {
  set_attention_tc(enabled);
  return MUMPY_OK;
}

extern "C" int mumpy_set_gemm_tile(int bn) {
  set_gemm_tile_override(bn);
  return MUMPY_OK;
}

extern "C" int mumpy_set_f16_overflow_flag(unsigned int *flag_dev) {
  int *n;
  F16FlagSetter *fns = flag_setters(&n);
  for (int i = 0; i < *n; ++i) {
    cudaError_t e = fns[i](flag_dev);
    if (e != cudaSuccess) {
      set_error("mumpy_set_f16_overflow_flag: %s", cudaGetErrorString(e));
      return MUMPY_ERR_CUDA;
    }
  }
  return MUMPY_OK;
}

extern "C" int mumpy_set_pdl(int enabled) {
  g_pdl = enabled ? 1 : 0;
  return MUMPY_OK;
}

extern "C" int mumpy_init(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    return MUMPY_ERR_CUDA;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return MUMPY_ERR_CUDA;
  }
  if (prop.major != 10) {
    set_error("libmumpy_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return MUMPY_ERR_UNSUPPORTED;
  }
  return resolve_driver_entry_points();
}

extern "C" int mumpy_linear(const void *A, long lda, const void *W, const float *bias, const float *residual, void *out,
                            long ldo, long M, int N, int K, int ab_dtype, int out_dtype, int act, void *stream) {
  MUMPY_REQUIRE(A && W && out && M > 0 && N > 0 && K > 0, "linear: bad arguments (M=%ld N=%d K=%d)", M, N, K);
  if (ab_dtype == MUMPY_F32) {
    MUMPY_REQUIRE(out_dtype == MUMPY_F32, "linear(fp32): output must be fp32");
    return linear_f32(static_cast<const float *>(A), lda, static_cast<const float *>(W), bias, residual,
                      static_cast<float *>(out), ldo, M, N, K, act, as_stream(stream));
  }
  if (is_16bit(ab_dtype)) return linear_bf16(A, lda, W, bias, residual, out, nullptr, ldo, M, N, K, ab_dtype, out_dtype, act, as_stream(stream));
  set_error("linear: unknown dtype %d", ab_dtype);
  return MUMPY_ERR_ARG;
}

extern "C" int mumpy_mlp_fused_supported(int C) { return mlp_fused_supported(C) ? 1 : 0; }
extern "C" int mumpy_set_mlp_fused_shape(int shape) {
  set_mlp_fused_shape(shape);
  return MUMPY_OK;
}

extern "C" int mumpy_mlp_fused(const float *x, const float *gamma, const float *beta, float eps, const void *W1, const float *b1, const void *W2,
                               const float *b2, float *out, long M, int C, int w_dtype, void *stream) {
  MUMPY_REQUIRE(x && gamma && beta && W1 && b1 && W2 && b2 && out && M > 0 && C > 0 && is_16bit(w_dtype), "mlp_fused: bad arguments (M=%ld C=%d)", M, C);
  return mlp_fused_16(x, gamma, beta, eps, W1, b1, W2, b2, out, M, C, w_dtype, as_stream(stream));
}

extern "C" int mumpy_set_ln_linear_pair_mode(int mode) {
  set_ln_linear_pair_mode(mode);
  return MUMPY_OK;
}

extern "C" int mumpy_ln_linear_supported(int N, int K) { return ln_linear_supported(N, K) ? 1 : 0; }

extern "C" int mumpy_ln_linear(const float *x, const float *gamma, const float *beta, float eps, const void *W, const float *bias, void *out,
                               long ldo, long M, int N, int K, int w_dtype, int act, void *stream) {
  MUMPY_REQUIRE(x && gamma && beta && W && out && M > 0 && N > 0 && K > 0 && is_16bit(w_dtype), "ln_linear: bad arguments (M=%ld N=%d K=%d)", M, N, K);
  return ln_linear_16(x, gamma, beta, eps, W, bias, out, ldo, M, N, K, w_dtype, act, as_stream(stream));
}

extern "C" int mumpy_linear_dual(const void *A, long lda, const void *W, const float *bias, const float *residual, float *out,
                                 void *aux16, long ldo, long M, int N, int K, int ab_dtype, int act, void *stream) {
  MUMPY_REQUIRE(A && W && out && aux16 && M > 0 && N > 0 && K > 0 && is_16bit(ab_dtype), "linear_dual: bad arguments (M=%ld N=%d K=%d)", M, N, K);
  return linear_bf16(A, lda, W, bias, residual, out, aux16, ldo, M, N, K, ab_dtype, MUMPY_F32, act, as_stream(stream));
}

extern "C" int mumpy_conv2d_nhwc_bf16(const void *in, long ld_in, const void *w_packed, const float *bias, const float *residual,
                                      void *out, long ld_out, int B, int H, int W, int Cin, int Cout, int kh, int kw, int ph, int pw,
                                      int in_dtype, int out_dtype, int act, float *splitk_ws, long splitk_ws_bytes, void *stream) {
  MUMPY_REQUIRE(in && w_packed && out && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && is_16bit(in_dtype), "conv2d_nhwc_bf16: bad arguments");
  return conv_bf16(in, ld_in, w_packed, bias, residual, out, ld_out, B, H, W, Cin, Cout, kh, kw, ph, pw, in_dtype, out_dtype, act, splitk_ws,
                   splitk_ws_bytes, as_stream(stream));
}

extern "C" int mumpy_cast16(const float *in, void *out, int out_dtype, long n, void *stream) {
  MUMPY_REQUIRE(in && out && n > 0 && is_16bit(out_dtype), "cast16: bad arguments");
  if (n % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const long n4 = n / 4;
    const int vb = (int)(cdiv(n4, 256) < 148 * 16 ? cdiv(n4, 256) : 148 * 16);
    MUMPY_WITH_16(out_dtype, T, launch_kernel(cast16_vec_kernel<T>, vb, 256, 0, as_stream(stream), reinterpret_cast<const float4 *>(in), reinterpret_cast<uint2 *>(out), n4));
    return launch_status("cast16_vec");
  }
  int blocks = (int)(cdiv(n, 256) < 148 * 8 ? cdiv(n, 256) : 148 * 8);
  MUMPY_WITH_16(out_dtype, T, launch_kernel(cast16_kernel<T>, blocks, 256, 0, as_stream(stream), in, static_cast<T *>(out), n));
  return launch_status("cast16");
}
