// K2/K6 (CUDA-core path): fp32 dense transforms with exact FFMA arithmetic.
//
// The GEMMs on the GripNet path are all "skinny": either a huge node dimension M
// against tiny feature widths (X.W, dY.W^T), or tiny outputs with a huge reduction
// (dW = X^T dY, datt, the batch-reduced dX of the relational conv).  One generic
// 64x64 tile leaves those latency-bound on a handful of CTAs, so this file
//   * picks the CTA tile from the output shape (128x16 / 128x32 / 64x64 for tall
//     outputs, 32x32 for small ones),
//   * flattens (batch-to-reduce, K) into ONE iteration space and splits it over
//     CTAs whenever the output tiles alone cannot fill 148 SMs,
//   * sums the split slices in slice order in a second tiny kernel, so every
//     result is bit-reproducible (no atomics).
// Fused epilogue: alpha, accumulate, addend, ReLU mask.  fp32 FFMA keeps the 1e-5
// parity bar without error compensation; the tcgen05 3xTF32 kernel (tc_gemm.cu)
// takes the large X.W / dY.W^T transforms.
#include "common.cuh"

namespace gn {

constexpr int kBK = 16;         // default k-tile depth
constexpr int kGemmThreads = 256;
constexpr int kTargetCtas = 148 * 2;
constexpr int kMaxSplits = 3 * 148;

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
  int kt;              // k-tiles per batch entry: ceil(K / bk)
  int iters;           // flattened reduction length: (batch_reduce ? batch : 1) * kt
  int splits;          // CTAs along the reduction
  int iters_per_split;
  int vec_ok;          // C / addend / mask rows are 16-byte aligned (float4 epilogue allowed)
  float* ws;           // [batch_indep][splits][M][N] when splits > 1
};

struct GemmPlan {
  int cfg;      // 0: 64x64  1: 128x32  2: 128x16  3: 32x32  4: 32x32 with a 64-deep k-tile
  int bm, bn, bk;
  int splits, iters_per_split, iters, kt;
  int batch_indep;
};

static GemmPlan make_plan(int M, int N, int K, int batch, int batch_reduce, int allow_split) {
  GemmPlan pl;
  if (M >= 1024) {
    if (N <= 16) { pl.cfg = 2; pl.bm = 128; pl.bn = 16; }
    else if (N <= 32) { pl.cfg = 1; pl.bm = 128; pl.bn = 32; }
    else { pl.cfg = 0; pl.bm = 64; pl.bn = 64; }
  } else if (M <= 32 || N <= 32 || int64_t(M) * N <= 4096 * 4) {
    pl.cfg = 3; pl.bm = 32; pl.bn = 32;
  } else {
    pl.cfg = 0; pl.bm = 64; pl.bn = 64;
  }
  pl.bk = kBK;
  if (pl.cfg == 3 && K >= 256) { pl.cfg = 4; pl.bk = 64; }   // long reductions: fewer, deeper iterations
  pl.kt = int(ceil_div(K, pl.bk));
  pl.batch_indep = batch_reduce ? 1 : batch;
  pl.iters = (batch_reduce ? batch : 1) * pl.kt;
  const int64_t tiles = ceil_div(M, pl.bm) * ceil_div(N, pl.bn) * pl.batch_indep;
  int splits = 1;
  if (allow_split && pl.iters >= 4 && tiles < kTargetCtas) {
    int64_t want = (pl.bk >= 64 ? 3 * 148 : kTargetCtas) / tiles;
    // at least two 16-deep k-tiles per slice; ONE 64-deep tile is enough (weight gradients: 19 081 rows over
    // ~296 CTAs of one tile each instead of 100 CTAs of three — the kernel is latency-bound, not bandwidth-bound)
    int64_t cap = pl.bk >= 64 ? pl.iters : pl.iters / 2;
    splits = int(want < cap ? want : cap);
    if (splits > kMaxSplits) splits = kMaxSplits;
    if (splits < 1) splits = 1;
  }
  pl.iters_per_split = int(ceil_div(pl.iters > 0 ? pl.iters : 1, splits));
  pl.splits = int(ceil_div(pl.iters > 0 ? pl.iters : 1, pl.iters_per_split));
  return pl;
}

template <int TN>
__device__ __forceinline__ void epilogue_row(const GemmParams& p, float* C, int m, int n0, const float (&v)[TN]) {
  float r[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) r[j] = v[j] * p.alpha;
  float* crow = C + int64_t(m) * p.ldc + n0;
  if (TN == 4 && p.vec_ok && n0 + 3 < p.N) {
    if (p.accumulate) {
      const float4 c = *reinterpret_cast<const float4*>(crow);
      r[0] += c.x; r[1] += c.y; r[2] += c.z; r[3] += c.w;
    }
    if (p.addend) {
      const float4 a = *reinterpret_cast<const float4*>(p.addend + int64_t(m) * p.ldd + n0);
      r[0] += a.x; r[1] += a.y; r[2] += a.z; r[3] += a.w;
    }
    if (p.mask) {
      const float4 k = *reinterpret_cast<const float4*>(p.mask + int64_t(m) * p.ldm + n0);
      if (!(k.x > 0.f)) r[0] = 0.f;
      if (!(k.y > 0.f)) r[1] = 0.f;
      if (!(k.z > 0.f)) r[2] = 0.f;
      if (!(k.w > 0.f)) r[3] = 0.f;
    }
    *reinterpret_cast<float4*>(crow) = make_float4(r[0], r[1], r[2], r[3]);
    return;
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = n0 + j;
    if (n >= p.N) break;
    float x = r[j];
    if (p.accumulate) x += crow[j];
    if (p.addend) x += p.addend[int64_t(m) * p.ldd + n];
    if (p.mask && !(p.mask[int64_t(m) * p.ldm + n] > 0.f)) x = 0.f;
    crow[j] = x;
  }
}

template <bool TA, bool TB, int BM, int BN, int TM, int TN, int BK>
__global__ void __launch_bounds__(kGemmThreads) sgemm_kernel(const GemmParams p) {
  constexpr int NTX = BN / TN;
  static_assert((BM / TM) * NTX == kGemmThreads, "tile / micro-tile mismatch");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % NTX, ty = tid / NTX;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;  // M tiles on grid.x (2^31 limit)
  const int zb = blockIdx.z / p.splits;                   // independent batch entry
  const int zs = blockIdx.z - zb * p.splits;              // reduction slice

  const int it_begin = zs * p.iters_per_split;
  const int it_end = min(p.iters, it_begin + p.iters_per_split);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int it = it_begin; it < it_end; ++it) {
    int b, k0;
    if (p.batch_reduce) { b = it / p.kt; k0 = (it - b * p.kt) * BK; } else { b = zb; k0 = it * BK; }
    const float* A = p.A + int64_t(b) * p.sA;
    const float* B = p.B + int64_t(b) * p.sB;
    // ---- stage A tile: As[k][m] = op(A)[m0+m, k0+k]
#pragma unroll
    for (int i = 0; i < (BM * BK) / kGemmThreads; ++i) {
      const int e = tid + i * kGemmThreads;
      int m, k;
      if (!TA) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < p.M && gk < p.K) {
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
      if (gn_ < p.N && gk < p.K) v = !TB ? __ldg(B + int64_t(gk) * p.ldb + gn_) : __ldg(B + int64_t(gn_) * p.ldb + gk);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], bb[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bb[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (p.splits > 1) {
    float* W = p.ws + (int64_t(zb) * p.splits + zs) * p.M * p.N;
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
  float* C = p.C + (p.batch_reduce ? 0 : int64_t(zb) * p.sC);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m < p.M) epilogue_row<TN>(p, C, m, n0 + tx * TN, acc[i]);
  }
}

// sum the reduction slices in a fixed order, then the epilogue.
// LANES == 1: one thread per output, slices added in slice order.
// LANES == 32: one warp per output; lane l adds slices l, l+32, ... in order, then a fixed
// shuffle tree combines the 32 lane sums (deterministic, and short for ~100 slices).
template <int LANES>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const GemmParams p, int batch_indep) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t idx = t / LANES;
  const int lane = int(t % LANES);
  const int64_t mn = int64_t(p.M) * p.N;
  if (idx >= mn * batch_indep) return;
  const int zb = int(idx / mn);
  const int64_t o = idx - int64_t(zb) * mn;
  const float* w = p.ws + int64_t(zb) * p.splits * mn + o;
  float s = 0.f;
  int z = lane;
  for (; z + 3 * LANES < p.splits; z += 4 * LANES) {   // loads issued together, adds kept in order
    const float a = __ldcg(w + int64_t(z) * mn), b = __ldcg(w + int64_t(z + LANES) * mn);
    const float c = __ldcg(w + int64_t(z + 2 * LANES) * mn), d = __ldcg(w + int64_t(z + 3 * LANES) * mn);
    s += a; s += b; s += c; s += d;
  }
  for (; z < p.splits; z += LANES) s += __ldcg(w + int64_t(z) * mn);
  if (LANES == 32) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(kFull, s, off);
    if (lane != 0) return;
  }
  const float v[1] = {s};
  epilogue_row<1>(p, p.C + (p.batch_reduce ? 0 : int64_t(zb) * p.sC), int(o / p.N), int(o % p.N), v);
}

template <bool TA, bool TB>
static int launch_cfg(const GemmPlan& pl, const GemmParams& p, dim3 grid, cudaStream_t st) {
  switch (pl.cfg) {
    case 0: GN_LAUNCH((sgemm_kernel<TA, TB, 64, 64, 4, 4, 16>), grid, kGemmThreads, 0, st, p); break;
    case 1: GN_LAUNCH((sgemm_kernel<TA, TB, 128, 32, 4, 4, 16>), grid, kGemmThreads, 0, st, p); break;
    case 2: GN_LAUNCH((sgemm_kernel<TA, TB, 128, 16, 2, 4, 16>), grid, kGemmThreads, 0, st, p); break;
    case 3: GN_LAUNCH((sgemm_kernel<TA, TB, 32, 32, 2, 2, 16>), grid, kGemmThreads, 0, st, p); break;
    default: GN_LAUNCH((sgemm_kernel<TA, TB, 32, 32, 2, 2, 64>), grid, kGemmThreads, 0, st, p); break;
  }
  return GN_OK;
}

}  // namespace gn

using namespace gn;

extern "C" size_t gn_sgemm_workspace_bytes(int32_t M, int32_t N, int32_t K, int32_t batch, int batch_reduce) {
  if (M <= 0 || N <= 0 || K <= 0 || batch < 1) return 0;
  const GemmPlan pl = make_plan(M, N, K, batch, batch_reduce, 1);
  if (pl.splits <= 1) return 0;
  return size_t(pl.batch_indep) * pl.splits * M * N * sizeof(float);
}

extern "C" int gn_sgemm(int transA, int transB, int32_t M, int32_t N, int32_t K, const float* A, int64_t lda,
                        const float* B, int64_t ldb, float* C, int64_t ldc, int32_t batch, int64_t strideA,
                        int64_t strideB, int64_t strideC, int batch_reduce, float alpha, int accumulate,
                        const float* addend, int64_t ld_addend, const float* relu_mask, int64_t ld_mask,
                        const int64_t* a_rows, int32_t split_k, float* ws, size_t ws_bytes, void* stream) {
  if (M < 0 || N < 0 || K < 0 || batch < 1 || !C) return GN_ERR_ARG;
  if (M == 0 || N == 0) return GN_OK;
  if (K > 0 && (!A || !B)) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const GemmPlan pl = make_plan(M, N, K, batch, batch_reduce, split_k > 0 && ws != nullptr);
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.batch = batch; p.sA = strideA; p.sB = strideB; p.sC = strideC;
  p.batch_reduce = batch_reduce ? 1 : 0;
  p.alpha = alpha; p.accumulate = accumulate ? 1 : 0;
  p.addend = addend; p.ldd = ld_addend; p.mask = relu_mask; p.ldm = ld_mask;
  p.a_rows = a_rows;
  p.kt = pl.kt; p.iters = pl.iters; p.splits = pl.splits; p.iters_per_split = pl.iters_per_split;
  p.ws = ws;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  p.vec_ok = al16(C) && ldc % 4 == 0 && (batch_reduce || batch == 1 || strideC % 4 == 0) &&
             (!addend || (al16(addend) && ld_addend % 4 == 0)) && (!relu_mask || (al16(relu_mask) && ld_mask % 4 == 0));
  if (pl.splits > 1 && ws_bytes < size_t(pl.batch_indep) * pl.splits * M * N * sizeof(float)) return GN_ERR_WORKSPACE;
  const int64_t gz = int64_t(pl.batch_indep) * pl.splits;
  dim3 grid((unsigned)ceil_div(M, pl.bm), (unsigned)ceil_div(N, pl.bn), (unsigned)gz);
  if (grid.y > 65535 || gz > 65535) return GN_ERR_RANGE;
  if (!transA && !transB) GN_CHECK((launch_cfg<false, false>(pl, p, grid, st)));
  else if (!transA && transB) GN_CHECK((launch_cfg<false, true>(pl, p, grid, st)));
  else if (transA && !transB) GN_CHECK((launch_cfg<true, false>(pl, p, grid, st)));
  else GN_CHECK((launch_cfg<true, true>(pl, p, grid, st)));
  if (pl.splits > 1) {
    const int64_t total = int64_t(M) * N * pl.batch_indep;
    // a warp per output (lane = slice) only pays when there are few outputs and many slices (weight gradients:
    // 32 x 16 outputs, ~100 slices); with many outputs a thread per output reads every slice COALESCED across the
    // warp (the relation-reduced dX of the RGCN: 645 x 48 outputs, 16 slices — 11.6 -> ~3 us on the backward's chain)
    if (pl.splits >= 16 && total <= 8192) {
      GN_LAUNCH(splitk_reduce_kernel<32>, (unsigned)ceil_div(total * 32, 256), 256, 0, st, p, pl.batch_indep);
    } else {
      GN_LAUNCH(splitk_reduce_kernel<1>, (unsigned)ceil_div(total, 256), 256, 0, st, p, pl.batch_indep);
    }
  }
  return GN_OK;
}
