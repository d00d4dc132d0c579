// Weight-gradient products on the 5th-generation tensor cores:  C[Mo, No] = A^T B,  A: [n, Mo], B: [n, No],
// both row-major with the FEATURES contiguous and the reduction running over the n node rows —
// dW = H_{l-1}^T dY of every GCN layer (autograd of gripnet/layers.py:73) and dW_r = X^T dY[:, r, :] of the
// relational layer (layers.py:171-189, all relations at once: B = dY viewed as [n, R*f]).
//
// Same 3xTF32 error-compensated scheme and stage layout as tc_gemm.cu, but BOTH operands arrive "MN-major"
// (the reduction index is the slow one in memory), so the eight loader warps transpose while they stage:
//   * a thread owns a 4-feature x 4-row micro-block: four coalesced 128-bit loads (a warp reads 512 contiguous
//     bytes of each row), a 4x4 transpose in registers, hi/lo split, four 128-bit stores into the canonical
//     K-major core-matrix layout (chunk = 4 consecutive reduction rows of one feature);
//   * global loads run one k-block (32 rows) ahead of the staging.
// Long reductions (n is the number of nodes: millions) cannot sit in one TMEM accumulator — the tensor core
// truncates when it adds — so the MMA warp works in runs of kGroup k-blocks (128 rows) that ping-pong between
// two TMEM accumulator pairs (hi*hi | cross terms); the loader warps drain a finished run with tcgen05.ld and
// keep the running sum in fp32 REGISTERS (round-to-nearest adds) while the next run is being multiplied.
// The grid splits the rows over CTAs; partial products go to `ws` and a second kernel adds them in split
// order (deterministic) and writes C, optionally in the [R][k][f] layout of the relational weights.
#include "tc_common.cuh"

namespace gn {
namespace tc {

constexpr int kGroup = 4;             // k-blocks per accumulator run
constexpr int kTnThreads = (kLoaderWarps + 1) * 32;

struct TnParams {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  int64_t n;
  int Mo, No;
  int nt;                  // columns of B per N tile: 32, 64 or 128
  int kb_total, kb_per_cta;
  int stages, tmem_cols;
  float* part; int64_t part_ld, part_split_stride;
};

__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// 4 rows x 4 features in registers -> for each feature the 4 consecutive reduction rows as one hi / lo chunk
__device__ __forceinline__ void stage_block(const float4 (&r)[4], unsigned char* hi_base, unsigned char* lo_base,
                                            int plane, int chunk, int row4) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = make_float4(comp(r[0], i), comp(r[1], i), comp(r[2], i), comp(r[3], i));
    float4 hi, lo;
    split4(v, hi, lo);
    const int off = chunk * plane + (row4 + i) * 16;
    *reinterpret_cast<float4*>(hi_base + off) = hi;
    *reinterpret_cast<float4*>(lo_base + off) = lo;
  }
}

template <int NG>   // NG = nt / 32: 16-column groups each epilogue thread owns
__global__ void __launch_bounds__(kTnThreads, 1) tc_tn_kernel(const TnParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[3], empty_bar[3], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = p.nt, S = p.stages;
  const int stage_sz = stage_bytes(nt);
  const int kb0 = blockIdx.x * p.kb_per_cta;
  const int kb1 = kb0 + p.kb_per_cta < p.kb_total ? kb0 + p.kb_per_cta : p.kb_total;
  const int nkb = kb1 - kb0;
  const int m0 = blockIdx.z * BM;
  const int n_tile0 = blockIdx.y * nt;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], kLoaderThreads);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kLoaderThreads);
    }
    fence_barrier_init();
  }
  if (warp == kLoaderWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(uint32_t(p.tmem_cols)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;

  if (warp < kLoaderWarps) {
    // ============================ loaders + running sums ============================
    const int fq = threadIdx.x & 31;          // A: features m0 + 4*fq .. +3
    const int c = threadIdx.x >> 5;           //    rows 4*c .. 4*c+3 of the k-block
    const int nbq = nt / 4;                   // B: feature quads in the tile
    const bool has_b = int(threadIdx.x) < nbq * CHUNKS;
    const int bq = int(threadIdx.x) % nbq, bc = int(threadIdx.x) / nbq;
    const bool a_ok = m0 + 4 * fq < p.Mo;
    const bool b_ok = has_b && n_tile0 + 4 * bq < p.No;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ca[4], cb[4], na[4], nb[4];
    auto issue = [&](int kb, float4(&a)[4], float4(&b)[4]) {
      const int64_t row_base = int64_t(kb0 + kb) * BK;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t ra = row_base + 4 * c + j;
        a[j] = (a_ok && ra < p.n) ? ldg4(p.A + ra * p.lda + m0 + 4 * fq) : zero4;
        const int64_t rb = row_base + 4 * bc + j;
        b[j] = (b_ok && rb < p.n) ? ldg4(p.B + rb * p.ldb + n_tile0 + 4 * bq) : zero4;
      }
    };
    float creg[NG * 16];
#pragma unroll
    for (int i = 0; i < NG * 16; ++i) creg[i] = 0.f;
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;

    issue(0, ca, cb);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S, use = kb / S;
      if (kb + 1 < nkb) issue(kb + 1, na, nb);
      if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1);
      unsigned char* a_hi = smem + s * stage_sz;
      unsigned char* a_lo = a_hi + part_bytes(BM);
      unsigned char* b_hi = a_lo + part_bytes(BM);
      unsigned char* b_lo = b_hi + part_bytes(nt);
      stage_block(ca, a_hi, a_lo, plane_bytes(BM), c, 4 * fq);
      if (has_b) stage_block(cb, b_hi, b_lo, plane_bytes(nt), bc, 4 * bq);
      fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core
      mbar_arrive(&full_bar[s]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ca[j] = na[j];
        cb[j] = nb[j];
      }
      if ((kb % kGroup) == kGroup - 1 || kb == nkb - 1) {
        // the run that just received its last k-block: wait for its MMAs, fold it into the fp32 running sum
        const int g = kb / kGroup, slot = g & 1;
        mbar_wait(&acc_full[slot], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t base = tmem_acc + lane_addr + uint32_t(slot * 2 * nt + half * (nt / 2));
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          uint32_t r[16], r2[16];
          tmem_ld16(base + uint32_t(nt + gi * 16), r2);      // cross terms (small) first
          tmem_ld16(base + uint32_t(gi * 16), r);
#pragma unroll
          for (int i = 0; i < 16; ++i) creg[gi * 16 + i] += __uint_as_float(r2[i]) + __uint_as_float(r[i]);
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[slot]);
      }
    }
    // ---- this split's partial tile
    const int row = m0 + quarter * 32 + lane;
    float* dst = p.part + int64_t(blockIdx.x) * p.part_split_stride + int64_t(row) * p.part_ld + n_tile0 +
                 half * (nt / 2);
#pragma unroll
    for (int gi = 0; gi < NG; ++gi)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(dst + gi * 16 + q * 4) =
            make_float4(creg[gi * 16 + q * 4], creg[gi * 16 + q * 4 + 1], creg[gi * 16 + q * 4 + 2],
                        creg[gi * 16 + q * 4 + 3]);
  } else {
    // ================================= MMA issuer ==================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, nt);
      const uint32_t lbo_a = plane_bytes(BM), lbo_b = plane_bytes(nt);
      uint32_t accumulate = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S, use = kb / S;
        const int g = kb / kGroup, slot = g & 1;
        if (kb % kGroup == 0) {
          if (g >= 2) {                    // the run two back used this accumulator pair: wait for its drain
            mbar_wait(&acc_empty[slot], ((g >> 1) - 1) & 1);
            tc_fence_after();
          }
          accumulate = 0;
        }
        const uint32_t tmem_d = tmem_acc + uint32_t(slot * 2 * nt);        // hi*hi
        const uint32_t tmem_x = tmem_d + uint32_t(nt);                     // lo*hi + hi*lo
        mbar_wait(&full_bar[s], use & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * stage_sz);
        const uint32_t a_lo = a_hi + part_bytes(BM);
        const uint32_t b_hi = a_lo + part_bytes(BM);
        const uint32_t b_lo = b_hi + part_bytes(nt);
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          const uint32_t ao = 2 * j * lbo_a, bo = 2 * j * lbo_b;
          const uint64_t dah = make_desc(a_hi + ao, lbo_a, 128), dal = make_desc(a_lo + ao, lbo_a, 128);
          const uint64_t dbh = make_desc(b_hi + bo, lbo_b, 128), dbl = make_desc(b_lo + bo, lbo_b, 128);
          umma_tf32(tmem_x, dal, dbh, idesc, accumulate);
          umma_tf32(tmem_x, dah, dbl, idesc, 1);
          umma_tf32(tmem_d, dah, dbh, idesc, accumulate);
          accumulate = 1;
        }
        umma_commit(&empty_bar[s]);
        if ((kb % kGroup) == kGroup - 1 || kb == nkb - 1) umma_commit(&acc_full[slot]);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kLoaderWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(uint32_t(p.tmem_cols)));
  }
}

// C(m, n) = sum over the splits, in split order; c_inner > 0: C is [No / c_inner][Mo][c_inner] (the [R][k][f]
// layout of the relational weights), else row-major [Mo, No] with leading dimension ldc
__global__ void __launch_bounds__(256) tn_reduce_kernel(const float* __restrict__ part, int n_splits,
                                                        int64_t split_stride, int64_t part_ld, int Mo, int No,
                                                        float* __restrict__ C, int64_t ldc, int c_inner,
                                                        int64_t c_stride) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(Mo) * No) return;
  const int m = int(idx / No), n = int(idx - int64_t(m) * No);
  const float* src = part + int64_t(m) * part_ld + n;
  float s = 0.f;
  int k = 0;
  for (; k + 7 < n_splits; k += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(src + int64_t(k + u) * split_stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  for (; k < n_splits; ++k) s += __ldg(src + int64_t(k) * split_stride);
  const int64_t dst = c_inner > 0 ? int64_t(n / c_inner) * c_stride + int64_t(m) * c_inner + (n % c_inner)
                                  : int64_t(m) * ldc + n;
  C[dst] = s;
}

struct TnPlan {
  bool ok;
  int nt, n_tiles, m_tiles, kb_total, kb_per_cta, n_splits, stages, tmem_cols;
  size_t smem_bytes, ws_bytes;
  int64_t part_ld, split_stride;
};

static TnPlan tn_plan(int64_t n, int Mo, int No) {
  TnPlan pl{};
  pl.ok = false;
  if (n <= 0 || Mo <= 0 || No <= 0 || n >= (int64_t(1) << 36)) return pl;
  pl.nt = No <= 32 ? 32 : (No <= 64 ? 64 : 128);
  pl.n_tiles = int(ceil_div(No, pl.nt));
  pl.m_tiles = int(ceil_div(Mo, BM));
  pl.kb_total = int(ceil_div(n, BK));
  const int64_t tiles = int64_t(pl.n_tiles) * pl.m_tiles;
  int64_t splits = 296 / tiles;                            // ~2 CTAs per SM's worth of loads in flight
  const int64_t max_splits = ceil_div(pl.kb_total, 2 * kGroup);   // at least two runs per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  pl.kb_per_cta = int(ceil_div(pl.kb_total, splits));
  pl.n_splits = int(ceil_div(pl.kb_total, pl.kb_per_cta));
  pl.stages = 3 * stage_bytes(pl.nt) <= 200 * 1024 ? 3 : 2;
  pl.tmem_cols = 4 * pl.nt;
  pl.smem_bytes = size_t(pl.stages) * stage_bytes(pl.nt);
  pl.part_ld = int64_t(pl.n_tiles) * pl.nt;
  pl.split_stride = int64_t(pl.m_tiles) * BM * pl.part_ld;
  pl.ws_bytes = size_t(pl.n_splits) * size_t(pl.split_stride) * 4;
  pl.ok = true;
  return pl;
}

}  // namespace tc
}  // namespace gn

using namespace gn;

extern "C" size_t gn_tc_tn_workspace_bytes(int64_t n, int32_t Mo, int32_t No) {
  const tc::TnPlan pl = tc::tn_plan(n, Mo, No);
  return pl.ok ? align_up(pl.ws_bytes) : 0;
}

extern "C" int gn_tc_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t n, int32_t Mo, int32_t No,
                        float* C, int64_t ldc, int32_t c_inner, int64_t c_stride, void* ws, size_t ws_bytes,
                        void* stream) {
  if (n < 0 || Mo <= 0 || No <= 0 || !C) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (n == 0) return GN_ERR_ARG;                      // callers route empty reductions elsewhere
  if (!A || !B || !al16(A) || !al16(B) || lda % 4 || ldb % 4 || Mo % 4 || No % 4) return GN_ERR_ARG;
  if (c_inner < 0 || (c_inner > 0 && No % c_inner != 0)) return GN_ERR_ARG;
  const tc::TnPlan pl = tc::tn_plan(n, Mo, No);
  if (!pl.ok) return GN_ERR_ARG;
  if (!ws || ws_bytes < pl.ws_bytes || !al16(ws)) return GN_ERR_WORKSPACE;
  tc::TnParams p;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.n = n; p.Mo = Mo; p.No = No;
  p.nt = pl.nt; p.kb_total = pl.kb_total; p.kb_per_cta = pl.kb_per_cta; p.stages = pl.stages;
  p.tmem_cols = pl.tmem_cols;
  p.part = static_cast<float*>(ws); p.part_ld = pl.part_ld; p.part_split_stride = pl.split_stride;
  static std::atomic<int> attr_set{0};
  if (!attr_set.load(std::memory_order_acquire)) {
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::tc_tn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::tc_tn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc::tc_tn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      g_last_cuda_error.store(int(e), std::memory_order_relaxed);
      return GN_ERR_CUDA;
    }
    attr_set.store(1, std::memory_order_release);
  }
  dim3 grid((unsigned)pl.n_splits, (unsigned)pl.n_tiles, (unsigned)pl.m_tiles);
  switch (pl.nt) {
    case 32: GN_LAUNCH(tc::tc_tn_kernel<1>, grid, tc::kTnThreads, pl.smem_bytes, st, p); break;
    case 64: GN_LAUNCH(tc::tc_tn_kernel<2>, grid, tc::kTnThreads, pl.smem_bytes, st, p); break;
    default: GN_LAUNCH(tc::tc_tn_kernel<4>, grid, tc::kTnThreads, pl.smem_bytes, st, p); break;
  }
  const int64_t outs = int64_t(Mo) * No;
  GN_LAUNCH(tc::tn_reduce_kernel, (unsigned)ceil_div(outs, 256), 256, 0, st, (const float*)p.part, pl.n_splits,
            pl.split_stride, pl.part_ld, Mo, No, C, ldc, c_inner, c_stride);
  return GN_OK;
}
