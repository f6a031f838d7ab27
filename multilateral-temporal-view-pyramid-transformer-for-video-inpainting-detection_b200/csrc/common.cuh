// Shared helpers for libmumpy_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mumpy_b200.h"

namespace mumpy {

void set_error(const char *fmt, ...);
int launch_status(const char *what);   // cudaGetLastError -> MUMPY_OK / MUMPY_ERR_CUDA (+ message)

#define MUMPY_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::mumpy::set_error(__VA_ARGS__);      \
      return MUMPY_ERR_ARG;                 \
    }                                       \
  } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: every kernel of the library is launched with the programmatic-stream-serialisation
// attribute and starts with pdl_grid_sync(): griddepcontrol.wait blocks until the preceding kernel of the stream has
// completed and its writes are visible (a no-op without the attribute), launch_dependents lets the next kernel's CTAs be
// scheduled -- and run their own prologue up to their wait -- while this one is still running.  Inside a CUDA graph
// this removes the drain-then-launch bubble between the ~750 dependent kernels of one forward.
bool pdl_enabled();
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  if (pdl_enabled()) {
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
static inline long cdiv(long a, long b) { return (a + b - 1) / b; }

// q = i / d, r = i % d for a flat element index: 32-bit arithmetic when the caller knows the index fits (`small`, uniform
// across the grid) -- a 64-bit division by a run-time value is a ~100-instruction routine, which made several flat kernels
// instruction-bound.
__device__ __forceinline__ void divmod_idx(long i, int d, bool small, long &q, int &r) {
  if (small) {
    const unsigned qq = (unsigned)i / (unsigned)d;
    q = qq;
    r = (int)((unsigned)i - qq * (unsigned)d);
  } else {
    q = i / d;
    r = (int)(i - q * d);
  }
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
// fp32 -> IEEE half always SATURATES to +-65504 (cvt.satfinite) instead of producing inf: a value beyond the half range can then
// never turn into inf - inf = NaN further down the forward.  Saturation alone would be silent, so the kernels that create new
// magnitudes in operand precision (GEMM / conv epilogues, casts, gathers, LayerNorm, the DCT repack) also track the largest
// |value| they converted and report it through f16_guard() below.
template <> __device__ __forceinline__ __half from_f32<__half>(float v) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return __ushort_as_half(h);
}

// two fp32 -> one 32-bit word of two 16-bit operands (round to nearest even) and back, for either operand type
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  uint32_t u;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(hi), "f"(lo));
  return u;
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) {
  const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&u);
  return make_float2(__low2float(h), __high2float(h));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) {
  const __half2 h = *reinterpret_cast<const __half2 *>(&u);
  return __half22float2(h);
}
// ---- fp16 range guard ("fp16" precision mode) ----
// A caller-owned device word (mumpy_set_f16_overflow_flag) that kernels OR 1 into when they had to saturate: the host reads
// it with the step's results (ops.f16_overflowed / evaluate.py raise), so leaving the half range is loud, never silent.
// The pointer lives in one __device__ variable per translation unit (the library is built without relocatable device code);
// every unit that includes this header registers its setter with api.cu.
static __device__ unsigned int *tu_f16_flag = nullptr;
typedef cudaError_t (*F16FlagSetter)(unsigned int *);
void register_f16_flag_setter(F16FlagSetter fn);
static cudaError_t tu_set_f16_flag(unsigned int *p) { return cudaMemcpyToSymbol(tu_f16_flag, &p, sizeof(p)); }
namespace {
struct F16FlagRegistrar {
  F16FlagRegistrar() { register_f16_flag_setter(&tu_set_f16_flag); }
};
static F16FlagRegistrar f16_flag_registrar_;
}  // namespace
constexpr float kHalfMax = 65504.0f;
// amax = largest |fp32 value| this thread converted to IEEE half (track with fmaxf(amax, fabsf(v)); NaN-transparent)
__device__ __forceinline__ void f16_guard(float amax) {
  if (amax > kHalfMax) {
    unsigned int *f = tu_f16_flag;
    if (f) atomicOr(f, 1u);
  }
}
template <typename T> struct is_half_t { static constexpr bool value = false; };
template <> struct is_half_t<__half> { static constexpr bool value = true; };

static inline bool is_16bit(int dtype) { return dtype == MUMPY_BF16 || dtype == MUMPY_F16; }
// runs `...` with T bound to the 16-bit operand type that `dtype` (MUMPY_BF16 / MUMPY_F16) names
#define MUMPY_WITH_16(dtype, T, ...)       \
  do {                                      \
    if ((dtype) == MUMPY_F16) {             \
      using T = __half;                     \
      __VA_ARGS__;                          \
    } else {                                \
      using T = __nv_bfloat16;              \
      __VA_ARGS__;                          \
    }                                       \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case MUMPY_ACT_GELU: return gelu_erf(v);
    case MUMPY_ACT_RELU: return fmaxf(v, 0.0f);
    case MUMPY_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default: return v;
  }
}

// canvas row of token p (0..ws*ws) of window n (0..nW) on a (TH x W) canvas, optionally cyclically
// shifted: window (wr,wc) position (pr,pc) reads canvas ((7wr+pr+shift) mod TH, (7wc+pc+shift) mod W)
// (swinTransformer.py:63-66,273,295).
__device__ __forceinline__ int window_token_row(int n, int p, int TH, int W, int ws, int shift) {
  const int wpr = W / ws;
  int r = (n / wpr) * ws + p / ws + shift;
  int c = (n % wpr) * ws + p % ws + shift;
  if (r >= TH) r -= TH;
  if (c >= W) c -= W;
  return r * W + c;
}

}  // namespace mumpy
