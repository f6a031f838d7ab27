// FAF frequency branch (dct.py:71-79) on the frame the encoder keeps (multiTemporalViewEncoder.py:734):
//   Xf = D X D^T ;  Y_band = D^T (F_band o Xf) D ,  F_band keeps lo <= i+j <= hi.
// Exact fp32 (bf16 operands would leak the large low-frequency coefficients into the small high band).
// Four batched GEMM passes on the fp32 skeleton; the band mask is applied in the A loader of pass 3.
#include "gemm_simt.cuh"

namespace mumpy {

// generic strided operand: value(z, r, c) = base[z*batch_stride + r*rs + c*cs]
struct Strided {
  const float *base;
  long batch_stride;
  long rs, cs;
  __device__ __forceinline__ float operator()(int z, long r, int c) const { return base[z * batch_stride + r * rs + c * cs]; }
};

// pass 3 A operand: Xf of image z % n_img masked by band z / n_img
struct BandMaskedA {
  const float *xf;
  long img_stride;
  int S, n_img;
  int lo0, hi0, lo1, hi1, lo2, hi2;
  __device__ __forceinline__ float operator()(int z, long m, int k) const {
    const int band = z / n_img, img = z % n_img;
    const int lo = band == 0 ? lo0 : (band == 1 ? lo1 : lo2);
    const int hi = band == 0 ? hi0 : (band == 1 ? hi1 : hi2);
    const int s = (int)m + k;
    if (s < lo || s > hi) return 0.0f;
    return xf[img * img_stride + m * S + k];
  }
};

struct StoreStrided {
  float *base;
  long batch_stride;
  int ld;
  __device__ __forceinline__ void operator()(int z, long m, int n, float v) const { base[z * batch_stride + m * ld + n] = v; }
};

// pass 1 A operand: frame `frame` of x (B,T,3,S,S): image z = b*3 + c
struct FrameA {
  const float *x;
  int T, frame, S;
  __device__ __forceinline__ float operator()(int z, long m, int k) const {
    const int b = z / 3, c = z % 3;
    return x[(((long)b * T + frame) * 3 + c) * S * S + m * S + k];
  }
};

// pass 4 output: out (B,9,S,S), channel = band*3 + rgb; z = band*n_img + (b*3 + rgb)
struct StoreBands {
  float *out;
  int S, n_img;
  __device__ __forceinline__ void operator()(int z, long m, int n, float v) const {
    const int band = z / n_img, img = z % n_img;
    const int b = img / 3, c = img % 3;
    out[(((long)b * 9 + band * 3 + c) * S + m) * S + n] = v;
  }
};

}  // namespace mumpy

using namespace mumpy;

extern "C" int mumpy_faf(const float *x, const float *dct, float *ws, float *out, int B, int T, int frame, int S,
                         const int *band_lo_hi6, void *stream) {
  MUMPY_REQUIRE(x && dct && ws && out && band_lo_hi6 && B > 0 && frame >= 0 && frame < T, "faf: bad arguments");
  cudaStream_t st = as_stream(stream);
  const int n_img = B * 3;
  const long SS = (long)S * S;
  float *t1 = ws;                 // X D^T          (n_img, S, S)
  float *xf = ws + n_img * SS;    // D X D^T        (n_img, S, S)
  float *u = ws + 2 * n_img * SS; // (F o Xf) D     (3*n_img, S, S)
  // 1. T1[m,n] = sum_k X[m,k] D[n,k]
  int rc = launch_gemm_simt(FrameA{x, T, frame, S}, Strided{dct, 0, S, 1}, StoreStrided{t1, SS, S}, S, S, S, n_img, st, "faf pass 1");
  if (rc) return rc;
  // 2. Xf[m,n] = sum_k D[m,k] T1[k,n]          (B operand indexed (n,k) -> T1[k*S + n])
  rc = launch_gemm_simt(Strided{dct, 0, S, 1}, Strided{t1, SS, 1, S}, StoreStrided{xf, SS, S}, S, S, S, n_img, st, "faf pass 2");
  if (rc) return rc;
  // 3. U[m,n] = sum_k (F o Xf)[m,k] D[k,n]
  BandMaskedA a3{xf, SS, S, n_img, band_lo_hi6[0], band_lo_hi6[1], band_lo_hi6[2], band_lo_hi6[3], band_lo_hi6[4], band_lo_hi6[5]};
  rc = launch_gemm_simt(a3, Strided{dct, 0, 1, S}, StoreStrided{u, SS, S}, S, S, S, 3 * n_img, st, "faf pass 3");
  if (rc) return rc;
  // 4. Y[m,n] = sum_k D[k,m] U[k,n]
  return launch_gemm_simt(Strided{dct, 0, 1, S}, Strided{u, SS, 1, S}, StoreBands{out, S, n_img}, S, S, S, 3 * n_img, st, "faf pass 4");
}
