// fp32-mode nn.Linear and decoder convolutions on the exact-fp32 GEMM skeleton (gemm_simt.cuh).
#include "gemm_simt.cuh"

namespace mumpy {

struct RowMajorA {
  const float *a;
  long lda;
  __device__ __forceinline__ float operator()(int, long m, int k) const { return a[m * lda + k]; }
};

struct WeightNK {   // nn.Linear weight (N,K) row-major
  const float *w;
  int K;
  __device__ __forceinline__ float operator()(int, int n, int k) const { return w[(long)n * K + k]; }
};

// A(m,k) of a stride-1 convolution on an NHWC map: m = (b,y,x), k = (ky,kx,c)
struct ConvA {
  const float *in;
  long ld;
  int H, W, Cin, kw, ph, pw;
  __device__ __forceinline__ float operator()(int, long m, int k) const {
    const int c = k % Cin;
    const int t = k / Cin;
    const int kx = t % kw, ky = t / kw;
    const int x = (int)(m % W);
    const long r = m / W;
    const int y = (int)(r % H);
    const long b = r / H;
    const int yy = y + ky - ph, xx = x + kx - pw;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) return 0.0f;
    return in[((b * H + yy) * W + xx) * ld + c];
  }
};

struct BiasActResidual {
  const float *bias;
  const float *residual;
  float *out;
  long ldo;
  int act;
  __device__ __forceinline__ void operator()(int, long m, int n, float v) const {
    if (bias) v += bias[n];
    v = apply_act(v, act);
    if (residual) v += residual[m * ldo + n];
    out[m * ldo + n] = v;
  }
};

int linear_f32(const float *A, long lda, const float *W, const float *bias, const float *residual, float *out, long ldo,
               long M, int N, int K, int act, cudaStream_t st) {
  return launch_gemm_simt(RowMajorA{A, lda}, WeightNK{W, K}, BiasActResidual{bias, residual, out, ldo, act}, M, N, K, 1, st,
                          "gemm_simt(linear)");
}

}  // namespace mumpy

using namespace mumpy;

extern "C" int mumpy_conv2d_nhwc(const float *in, long ld_in, const float *w, const float *bias, float *out, long ld_out,
                                 int B, int H, int W, int Cin, int Cout, int kh, int kw, int ph, int pw, void *stream) {
  MUMPY_REQUIRE(in && w && out && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv2d_nhwc: bad arguments");
  const long M = (long)B * H * W;
  const int K = kh * kw * Cin;
  return launch_gemm_simt(ConvA{in, ld_in, H, W, Cin, kw, ph, pw}, WeightNK{w, K},
                          BiasActResidual{bias, nullptr, out, ld_out, MUMPY_ACT_NONE}, M, Cout, K, 1, as_stream(stream),
                          "gemm_simt(conv2d)");
}
