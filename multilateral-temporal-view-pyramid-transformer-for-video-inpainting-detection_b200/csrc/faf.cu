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

// ---- tensor-core path (16-bit modes): the four DCT passes as tcgen05 GEMMs on split operands -------------------------
// a = a_hi + a_lo with both halves in the 16-bit operand type; a.b ~ a_hi.b_hi + a_hi.b_lo + a_lo.b_hi accumulated in fp32 is
// one GEMM over the K-concatenation  A' = [a_hi | a_hi | a_lo],  W' = [w_hi | w_lo | w_hi]  (K' = 3S): max-abs error vs the
// fp64 transform 2.9e-5 (bf16 halves) / 2.5e-6 (f16 halves, = fp32's own rounding) on outputs of magnitude ~3, far below
// the operand rounding of the convolutions that consume the result.  Between the GEMMs a repack kernel transposes each
// S x S image (the next pass contracts over the other index), applies the band masks, splits into halves and writes the
// K-concatenated operand; the last repack writes the (B,9,S,S) result.
struct RepackParams {
  const float *in;          // fp32 images, row-major R x C
  long group_stride;        // image z lives at in + (z / group) * group_stride + (z % group) * R * C
  int group;
  int R, C;
  int transpose;            // output element (orow, ocol) = transpose ? (c, r) : (r, c)
  int n_img;                // input images
  int n_bands;              // 1, or 3: output image band*n_img + z keeps lo[band] <= r + c <= hi[band]
  int lo[3], hi[3];
  void *out16;              // split mode: rows of 3*OC 16-bit values [hi | hi | lo]
  float *out32;             // final mode: fp32, output image (z % n_img3 ... ) remapped to (b, band*3 + c)
  int final_imgs;           // final mode: images per band (= 3B); 0 otherwise
};

template <typename T16>
__global__ void __launch_bounds__(256) faf_repack_kernel(const RepackParams p) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int z = blockIdx.z;
  const int r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float *img = p.in + (long)(z / p.group) * p.group_stride + (long)(z % p.group) * p.R * p.C;
  const int OR = p.transpose ? p.C : p.R, OC = p.transpose ? p.R : p.C;
  // one CTA per 32-row band of an image, walking the column tiles (a CTA per 32 x 32 tile was launch-bound: 14 k CTAs of
  // 1 k elements each)
  for (int c0 = blockIdx.x * 32; c0 < p.C; c0 += gridDim.x * 32) {
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < p.R && c < p.C) ? img[(long)r * p.C + c] : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    // output row index runs over i, output column over tx (coalesced along the output row)
    const int r = p.transpose ? r0 + tx : r0 + i;
    const int c = p.transpose ? c0 + i : c0 + tx;
    if (r >= p.R || c >= p.C) continue;
    const float v = p.transpose ? tile[tx][i] : tile[i][tx];
    const int orow = p.transpose ? c : r, ocol = p.transpose ? r : c;
    if (p.out32) {
      const int band = z / p.final_imgs, im = z % p.final_imgs;
      const long oimg = (long)(im / 3) * 9 + band * 3 + im % 3;
      p.out32[(oimg * OR + orow) * OC + ocol] = v;
    } else {
      T16 *o = static_cast<T16 *>(p.out16);
      for (int b = 0; b < p.n_bands; ++b) {
        float m = v;
        if (p.n_bands > 1 && (r + c < p.lo[b] || r + c > p.hi[b])) m = 0.0f;
        if (is_half_t<T16>::value && fabsf(m) > kHalfMax) f16_guard(fabsf(m));      // DCT coefficient beyond the half range
        const T16 h = from_f32<T16>(m);
        const T16 l = from_f32<T16>(m - to_f32(h));
        T16 *row = o + (((long)b * p.n_img + z) * OR + orow) * 3 * OC;
        row[ocol] = h;
        row[OC + ocol] = h;
        row[2 * OC + ocol] = l;
      }
    }
  }
  }
}

int linear_faf_pass(const void *A, const void *W, void *out, long M, int K, int S, int imgs_per_band, int bands, const int *lo_hi6, int final_pass,
                    int ab_dtype, cudaStream_t st, int in_sparse = 0);

static int faf_repack(const RepackParams &p, int n_img, int dtype, cudaStream_t st) {
  dim3 grid(1, (unsigned)cdiv(p.R, 32), (unsigned)n_img);
  if (dtype == MUMPY_F16)
    launch_kernel(faf_repack_kernel<__half>, grid, 256, 0, st, p);
  else
    launch_kernel(faf_repack_kernel<__nv_bfloat16>, grid, 256, 0, st, p);
  return launch_status("faf_repack");
}

}  // namespace mumpy

using namespace mumpy;

// workspace sizes of the two FAF entry points (floats of `ws` / bytes of `ws16`), so that callers do not repeat the layout
extern "C" long mumpy_faf_workspace_floats(int B, int S) { return B > 0 && S > 0 ? 5l * B * 3 * S * S : 0; }
extern "C" long mumpy_faf16_workspace_bytes(int B, int S) { return B > 0 && S > 0 ? 2l * 9 * B * S * 3 * S * 2 : 0; }

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

// Tensor-core FAF.  dcat (S, 3S) = [D_hi | D_lo | D_hi], dtcat (S, 3S) = the same for D^T, both of `dtype`;
// ws16: 2 x (9*B*S rows x 3S) 16-bit values (the passes ping-pong between the halves); ws32 is no longer used (may be NULL).
extern "C" int mumpy_faf16(const float *x, const void *dcat, const void *dtcat, void *ws16, float *ws32, float *out, int B, int T,
                           int frame, int S, const int *band_lo_hi6, int dtype, void *stream) {
  MUMPY_REQUIRE(x && dcat && dtcat && ws16 && out && band_lo_hi6 && B > 0 && frame >= 0 && frame < T, "faf16: bad arguments");
  MUMPY_REQUIRE(is_16bit(dtype) && S % 8 == 0, "faf16: 16-bit operand type and S %% 8 == 0 required");
  cudaStream_t st = as_stream(stream);
  const int n_img = B * 3;
  const long SS = (long)S * S;
  RepackParams p = {};
  p.R = S;
  p.C = S;
  p.out16 = ws16;
  // P0: frame `frame` of x -> [x_hi | x_hi | x_lo], rows (img, i)
  p.in = x;
  p.group = 3;
  p.group_stride = (long)T * 3 * SS;
  p.in += (long)frame * 3 * SS;
  p.transpose = 0;
  p.n_img = n_img;
  p.n_bands = 1;
  int rc = faf_repack(p, n_img, dtype, st);
  if (rc) return rc;
  // The four GEMMs write their results already transposed per image, band masked and split (faf_ts_epilogue, gemm_tcgen05.cu)
  // into the other half of ws16, so no repack kernel runs between them.
  char *ws_a = static_cast<char *>(ws16);
  char *ws_b = ws_a + (size_t)9 * B * S * 3 * S * 2;
  int lohi[6];
  for (int i = 0; i < 6; ++i) lohi[i] = band_lo_hi6[i];
  for (int b = 0; b < 3; ++b) MUMPY_REQUIRE(lohi[2 * b] >= 0 && lohi[2 * b + 1] < 65536, "faf16: band limits out of range");
  // G1: T1[(img,i), k] = sum_j x[i,j] D[k,j]                      -> rows (img, k), columns i
  rc = linear_faf_pass(ws_a, dcat, ws_b, (long)n_img * S, 3 * S, S, n_img, 1, nullptr, 0, dtype, st);
  if (rc) return rc;
  // G2: Xf^T[(img,k), k'] = sum_i T1[i,k] D[k',i]                 -> rows (band, img, k'), columns k of F_band o Xf
  rc = linear_faf_pass(ws_b, dcat, ws_a, (long)n_img * S, 3 * S, S, n_img, 3, lohi, 0, dtype, st);
  if (rc) return rc;
  // G3: U[(band,img,k'), j] = sum_k (F o Xf)[k',k] D[k,j]          -> rows (band, img, j), columns k'
  // (passes 3 and 4 skip the k-blocks that the band masks zeroed: 6 / 7 / 11 of 11 blocks for the low / middle / high band at S = 224;
  //  MUMPY_FAF_SPARSE=0 multiplies them all)
  static int faf_sparse = -1;
  if (faf_sparse < 0) {
    const char *v = getenv("MUMPY_FAF_SPARSE");
    faf_sparse = (v && v[0] == '0') ? 0 : 1;
  }
  rc = linear_faf_pass(ws_a, dtcat, ws_b, 3l * n_img * S, 3 * S, S, 3 * n_img, 1, lohi, 0, dtype, st, faf_sparse);
  if (rc) return rc;
  // G4: Y^T[(band,img,j), i] = sum_k' U[k',j] D[k',i]              -> out (B, 9, S, S), channel = band*3 + rgb, element (i, j)
  return linear_faf_pass(ws_b, dtcat, out, 3l * n_img * S, 3 * S, S, n_img, 1, lohi, 1, dtype, st, faf_sparse);
}
