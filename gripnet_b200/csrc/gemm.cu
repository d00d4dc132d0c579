// K2/K6 (round 1): fp32 dense transforms on the CUDA cores.
//
// Strided / batched / batch-reduce SGEMM with fused epilogue (accumulate, addend,
// ReLU-mask) and a deterministic split-K for the weight-gradient products whose
// reduction dimension is the node count.  fp32 FFMA keeps the 1e-5 parity bar of
// north_star without error compensation; the tcgen05 3xTF32 path for the large
// X.W_r transforms is the next step for this file (DESIGN.md "next").
#include "common.cuh"

namespace gn {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int kGemmThreads = 256;

struct GemmParams {
  int M, N, K;
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int batch; int64_t sA, sB, sC;
  int batch_reduce;
  float alpha; int accumulate;
  const float* addend; int64_t ldd;
  const float* mask; int64_t ldm;
  const int64_t* a_rows;
  int split_k; int k_per_split;
  float* ws;
};

__device__ __forceinline__ void gemm_epilogue(const GemmParams& p, float* C, int m, int n, float v) {
  v *= p.alpha;
  if (p.accumulate) v += C[int64_t(m) * p.ldc + n];
  if (p.addend) v += p.addend[int64_t(m) * p.ldd + n];
  if (p.mask && !(p.mask[int64_t(m) * p.ldm + n] > 0.f)) v = 0.f;
  C[int64_t(m) * p.ldc + n] = v;
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(kGemmThreads) sgemm_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;  // M tiles on grid.x (2^31 limit)
  const int z = blockIdx.z;

  int k_begin = 0, k_end = p.K;
  int b_begin = 0, b_end = 1;
  if (p.split_k > 1) {
    k_begin = z * p.k_per_split;
    k_end = min(p.K, k_begin + p.k_per_split);
  } else if (p.batch_reduce) {
    b_end = p.batch;
  } else {
    b_begin = z;
    b_end = z + 1;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int b = b_begin; b < b_end; ++b) {
    const float* A = p.A + int64_t(b) * p.sA;
    const float* B = p.B + int64_t(b) * p.sB;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
      // ---- stage A tile: As[k][m] = op(A)[m0+m, k0+k]
#pragma unroll
      for (int i = 0; i < (BM * BK) / kGemmThreads; ++i) {
        const int e = tid + i * kGemmThreads;
        int m, k;
        if (!TA) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
        const int gm = m0 + m, gk = k0 + k;
        float v = 0.f;
        if (gm < p.M && gk < k_end) {
          if (!TA) {
            const int64_t r = p.a_rows ? p.a_rows[gm] : gm;
            v = __ldg(A + r * p.lda + gk);
          } else {
            const int64_t r = p.a_rows ? p.a_rows[gk] : gk;
            v = __ldg(A + r * p.lda + gm);
          }
        }
        As[k][m] = v;
      }
      // ---- stage B tile: Bs[k][n] = op(B)[k0+k, n0+n]
#pragma unroll
      for (int i = 0; i < (BN * BK) / kGemmThreads; ++i) {
        const int e = tid + i * kGemmThreads;
        int n, k;
        if (!TB) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
        const int gn_ = n0 + n, gk = k0 + k;
        float v = 0.f;
        if (gn_ < p.N && gk < k_end) v = !TB ? __ldg(B + int64_t(gk) * p.ldb + gn_) : __ldg(B + int64_t(gn_) * p.ldb + gk);
        Bs[k][n] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
        const float a[TM] = {a4.x, a4.y, a4.z, a4.w};
        const float bb[TN] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  if (p.split_k > 1) {
    float* W = p.ws + int64_t(z) * p.M * p.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n < p.N) W[int64_t(m) * p.N + n] = acc[i][j];
      }
    }
    return;
  }
  float* C = p.C + (p.batch_reduce ? 0 : int64_t(z) * p.sC);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < p.N) gemm_epilogue(p, C, m, n, acc[i][j]);
    }
  }
}

// sum the split-K slices in slice order, then the epilogue
__global__ void splitk_reduce_kernel(const GemmParams p) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = int64_t(p.M) * p.N;
  if (idx >= total) return;
  float s = 0.f;
  for (int z = 0; z < p.split_k; ++z) s += p.ws[int64_t(z) * total + idx];
  gemm_epilogue(p, p.C, int(idx / p.N), int(idx % p.N), s);
}

}  // namespace gn

using namespace gn;

extern "C" int gn_sgemm(int transA, int transB, int32_t M, int32_t N, int32_t K, const float* A, int64_t lda,
                        const float* B, int64_t ldb, float* C, int64_t ldc, int32_t batch, int64_t strideA,
                        int64_t strideB, int64_t strideC, int batch_reduce, float alpha, int accumulate,
                        const float* addend, int64_t ld_addend, const float* relu_mask, int64_t ld_mask,
                        const int64_t* a_rows, int32_t split_k, float* ws, size_t ws_bytes, void* stream) {
  if (M < 0 || N < 0 || K < 0 || batch < 1 || !C) return GN_ERR_ARG;
  if (M == 0 || N == 0) return GN_OK;
  if (K > 0 && (!A || !B)) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.batch = batch; p.sA = strideA; p.sB = strideB; p.sC = strideC;
  p.batch_reduce = batch_reduce ? 1 : 0;
  p.alpha = alpha; p.accumulate = accumulate ? 1 : 0;
  p.addend = addend; p.ldd = ld_addend; p.mask = relu_mask; p.ldm = ld_mask;
  p.a_rows = a_rows;
  p.split_k = 1; p.k_per_split = K; p.ws = ws;
  if (split_k > 1 && batch == 1 && K > BK) {
    int kps = int(ceil_div(ceil_div(K, split_k), BK) * BK);
    int splits = int(ceil_div(K, kps));
    if (splits > 1) {
      if (!ws || ws_bytes < size_t(splits) * M * N * sizeof(float)) return GN_ERR_WORKSPACE;
      p.split_k = splits;
      p.k_per_split = kps;
    }
  }
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN),
            (unsigned)(p.split_k > 1 ? p.split_k : (p.batch_reduce ? 1 : batch)));
  if (grid.y > 65535 || grid.z > 65535) return GN_ERR_RANGE;
  if (!transA && !transB) { GN_LAUNCH((sgemm_kernel<false, false>), grid, kGemmThreads, 0, st, p); }
  else if (!transA && transB) { GN_LAUNCH((sgemm_kernel<false, true>), grid, kGemmThreads, 0, st, p); }
  else if (transA && !transB) { GN_LAUNCH((sgemm_kernel<true, false>), grid, kGemmThreads, 0, st, p); }
  else { GN_LAUNCH((sgemm_kernel<true, true>), grid, kGemmThreads, 0, st, p); }
  if (p.split_k > 1) {
    GN_LAUNCH(splitk_reduce_kernel, (unsigned)ceil_div(int64_t(M) * N, 256), 256, 0, st, p);
  }
  return GN_OK;
}
