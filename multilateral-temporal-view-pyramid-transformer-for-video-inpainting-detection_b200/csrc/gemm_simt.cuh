// Exact-fp32 tiled GEMM skeleton on the FMA pipe, parameterised by operand loaders and an epilogue:
//   acc(m,n) = sum_k A(z,m,k) * B(z,n,k)   ->   E(z,m,n,acc)          z = blockIdx.z (batch)
// Used for the fp32 parity mode of every nn.Linear, the decoder convolutions (implicit GEMM: the im2col
// gather lives in the A loader) and the DCT band split.
#pragma once
#include "common.cuh"

namespace mumpy {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_TM = 4, SG_TN = 4;

template <typename ALoader, typename BLoader, typename Epilogue>
__global__ void __launch_bounds__(256) gemm_simt_kernel(ALoader A, BLoader Bm, Epilogue E, long M, int N, int K) {
  pdl_grid_sync();
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int z = blockIdx.z;
  const long m0 = (long)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  float acc[SG_TM][SG_TN];
#pragma unroll
  for (int i = 0; i < SG_TM; ++i)
#pragma unroll
    for (int j = 0; j < SG_TN; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < (SG_BM * SG_BK) / 256; ++i) {
      const int e = tid + i * 256;
      const int mm = e / SG_BK, kk = e % SG_BK;
      const long m = m0 + mm;
      const int k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? A(z, m, k) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < (SG_BN * SG_BK) / 256; ++i) {
      const int e = tid + i * 256;
      const int nn = e / SG_BK, kk = e % SG_BK;
      const int n = n0 + nn;
      const int k = k0 + kk;
      Bs[kk][nn] = (n < N && k < K) ? Bm(z, n, k) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[SG_TM], b[SG_TN];
#pragma unroll
      for (int i = 0; i < SG_TM; ++i) a[i] = As[kk][ty * SG_TM + i];
#pragma unroll
      for (int j = 0; j < SG_TN; ++j) b[j] = Bs[kk][tx * SG_TN + j];
#pragma unroll
      for (int i = 0; i < SG_TM; ++i)
#pragma unroll
        for (int j = 0; j < SG_TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < SG_TM; ++i) {
    const long m = m0 + ty * SG_TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < SG_TN; ++j) {
      const int n = n0 + tx * SG_TN + j;
      if (n < N) E(z, m, n, acc[i][j]);
    }
  }
}

template <typename ALoader, typename BLoader, typename Epilogue>
static inline int launch_gemm_simt(ALoader A, BLoader Bm, Epilogue E, long M, int N, int K, int batch, cudaStream_t st,
                                   const char *what) {
  dim3 grid((unsigned)cdiv(M, SG_BM), (unsigned)cdiv(N, SG_BN), (unsigned)batch);
  launch_kernel(gemm_simt_kernel<ALoader, BLoader, Epilogue>, grid, 256, 0, st, A, Bm, E, M, N, K);
  return launch_status(what);
}

}  // namespace mumpy
