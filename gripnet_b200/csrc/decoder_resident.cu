// K9/K10 with the embedding table resident in shared memory.
//
// The link decoder (gripnet/decoder.py:19-23) scores edges between the nodes of the TASK supervertex, whose
// embedding table is small: 645 drugs x 80 floats = 206 KB on every pose dataset (GripNet-pose.py:86-99) —
// it fits the 227 KB of shared memory of one B200 SM.  The global-memory kernels of decoder.cu gather 64 B
// pieces of that table through L1 (8 distinct cache lines per 128-bit warp load: the L1 tag stage, not HBM
// and not L2, bounds them).  Here every persistent CTA (one per SM) copies the table into shared memory once
// and all row gathers are shared-memory reads: 8 lanes read one 128-byte piece of a row (one conflict-free
// wavefront per quarter warp), 4 entries per warp step.  Same arithmetic per element as decoder.cu
// (fmaf(a*b, c, s) forward; fmaf(g, a, run) / fmaf(run, w, acc) / fmaf(g*a, b, acc) backward); only the
// association of the partial sums over lanes differs (8 lanes x 4 floats instead of 4 x 4), fixed per build.
// Row-split, atomic-free reductions as everywhere else (rowsplit.cuh).
#include "rowsplit.cuh"

namespace gn {

constexpr int kResThreads = 512;
constexpr int kResWarps = kResThreads / 32;
constexpr int kResLPE = 8;                 // lanes per entry
constexpr int kResEPI = 32 / kResLPE;      // entries per warp step
constexpr int kResMaxSmem = 227 * 1024;    // opt-in dynamic shared memory per CTA on sm_100

// z [n_nodes, D] (leading dimension ldz) -> zs [n_nodes, D] packed; D % 4 == 0
__device__ __forceinline__ void stage_table(const float* __restrict__ z, int64_t ldz, int n_nodes, int D, float* zs) {
  const int d4 = D >> 2;
  const int total = n_nodes * d4;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = i / d4, c = i - r * d4;
    reinterpret_cast<float4*>(zs)[i] = ldg4(z + int64_t(r) * ldz + 4 * c);
  }
  __syncthreads();
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ---------------------------------------------------------------------------------------------
// forward: 32 consecutive edges per warp batch (coalesced index loads and score stores)
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kResThreads, 1) distmult_fwd_res_kernel(
    const float* __restrict__ z, int64_t ldz, int n_nodes, int D, const float* __restrict__ w,
    const int64_t* __restrict__ src, const int64_t* __restrict__ dst, const int64_t* __restrict__ etype,
    int64_t n_edges, int sigmoid, float* __restrict__ out) {
  extern __shared__ __align__(16) float zs[];
  stage_table(z, ldz, n_nodes, D, zs);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane / kResLPE, fl = lane % kResLPE;
  const int64_t n_batches = (n_edges + 31) >> 5;
  for (int64_t b = int64_t(blockIdx.x) * kResWarps + warp; b < n_batches; b += int64_t(gridDim.x) * kResWarps) {
    const int64_t e = (b << 5) + lane;
    int s = 0, d = 0, r = 0;
    if (e < n_edges) {
      s = int(src[e]);
      d = int(dst[e]);
      r = int(etype[e]);
    }
    float mine = 0.f;
#pragma unroll
    for (int t = 0; t < 32 / kResEPI; ++t) {
      const int from = t * kResEPI + slot;
      const int ss = __shfl_sync(kFull, s, from), dd = __shfl_sync(kFull, d, from), rr = __shfl_sync(kFull, r, from);
      const float* za = zs + ss * D;
      const float* zb = zs + dd * D;
      const float* wr = w + int64_t(rr) * D;
      float acc = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int f = (v * kResLPE + fl) * 4;
        if (f < D) {
          const float4 a = lds4(za + f), bb = lds4(zb + f), c = ldg4(wr + f);
          acc = fmaf(a.x * bb.x, c.x, acc);
          acc = fmaf(a.y * bb.y, c.y, acc);
          acc = fmaf(a.z * bb.z, c.z, acc);
          acc = fmaf(a.w * bb.w, c.w, acc);
        }
      }
#pragma unroll
      for (int o = kResLPE / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
      // lane L keeps the score of edge L of the batch: edge t*EPI + j was computed by slot j
      const float got = __shfl_sync(kFull, acc, (lane % kResEPI) * kResLPE);
      if (lane / kResEPI == t) mine = got;
    }
    if (e < n_edges) out[e] = sigmoid ? 1.0f / (1.0f + expf(-mine)) : mine;
  }
}

// ---------------------------------------------------------------------------------------------
// backward, persistent warps over the row-split work list
//   MODE 0 (dz): rows = nodes, entry (other, rel, e): run += coef[e] z[other]; acc += run .* w[rel] per relation run
//   MODE 1 (dw): rows = relations, entry e: acc += (coef[e] z[src_e]) .* z[dst_e]
// ---------------------------------------------------------------------------------------------
template <int NV, int MODE>
__global__ void __launch_bounds__(kResThreads, 1) distmult_bwd_res_kernel(
    const gn_csr csr, const int32_t* __restrict__ ent_a, const int32_t* __restrict__ ent_b,
    const int32_t* __restrict__ ent_eid, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
    const float* __restrict__ coef, const float* __restrict__ z, int64_t ldz, int n_nodes, int D,
    const float* __restrict__ w, float* __restrict__ outp, int64_t ldo, float* __restrict__ partial) {
  extern __shared__ __align__(16) float zs[];
  stage_table(z, ldz, n_nodes, D, zs);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane / kResLPE, fl = lane % kResLPE;
  const int total_warps = gridDim.x * kResWarps;
  for (int chunk = blockIdx.x * kResWarps + warp;; chunk += total_warps) {
    ChunkInfo ci;
    if (!chunk_info_at(csr, chunk, ci)) break;
    Vec<4> acc[NV], run[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[v].v[i] = run[v].v[i] = 0.f;
    int cur_rel = -1;

    auto flush = [&](int rel) {
      const float* wr = w + int64_t(rel) * D;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int f = (v * kResLPE + fl) * 4;
        if (f < D) {
          const float4 c = ldg4(wr + f);
          acc[v].v[0] = fmaf(run[v].v[0], c.x, acc[v].v[0]);
          acc[v].v[1] = fmaf(run[v].v[1], c.y, acc[v].v[1]);
          acc[v].v[2] = fmaf(run[v].v[2], c.z, acc[v].v[2]);
          acc[v].v[3] = fmaf(run[v].v[3], c.w, acc[v].v[3]);
#pragma unroll
          for (int i = 0; i < 4; ++i) run[v].v[i] = 0.f;
        }
      }
    };

    for (int base = ci.beg; base < ci.end; base += 32) {
      const int mine = base + lane;
      int ia = 0, ib = 0;        // MODE 0: other, rel;  MODE 1: src, dst
      float g = 0.f;
      if (mine < ci.end) {
        const int e = ent_eid ? __ldg(ent_eid + mine) : mine;
        g = __ldg(coef + e);
        if (MODE == 0) {
          ia = __ldg(ent_a + mine);
          ib = __ldg(ent_b + mine);
        } else {
          ia = int(src[e]);
          ib = int(dst[e]);
        }
      }
      const int n_here = min(32, ci.end - base);
#pragma unroll
      for (int t = 0; t < 32 / kResEPI; ++t) {
        const int from = t * kResEPI + slot;
        const int a_i = __shfl_sync(kFull, ia, from), b_i = __shfl_sync(kFull, ib, from);
        const float gg = __shfl_sync(kFull, g, from);
        if (from < n_here) {
          const float* za = zs + a_i * D;
          if (MODE == 0) {
            if (b_i != cur_rel) {
              if (cur_rel >= 0) flush(cur_rel);
              cur_rel = b_i;
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              const int f = (v * kResLPE + fl) * 4;
              if (f < D) {
                const float4 a = lds4(za + f);
                run[v].v[0] = fmaf(gg, a.x, run[v].v[0]);
                run[v].v[1] = fmaf(gg, a.y, run[v].v[1]);
                run[v].v[2] = fmaf(gg, a.z, run[v].v[2]);
                run[v].v[3] = fmaf(gg, a.w, run[v].v[3]);
              }
            }
          } else {
            const float* zb = zs + b_i * D;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              const int f = (v * kResLPE + fl) * 4;
              if (f < D) {
                const float4 a = lds4(za + f), bb = lds4(zb + f);
                acc[v].v[0] = fmaf(gg * a.x, bb.x, acc[v].v[0]);
                acc[v].v[1] = fmaf(gg * a.y, bb.y, acc[v].v[1]);
                acc[v].v[2] = fmaf(gg * a.z, bb.z, acc[v].v[2]);
                acc[v].v[3] = fmaf(gg * a.w, bb.w, acc[v].v[3]);
              }
            }
          }
        }
      }
    }
    if (MODE == 0 && cur_rel >= 0) flush(cur_rel);
#pragma unroll
    for (int v = 0; v < NV; ++v) reduce_slots<kResLPE, 4>(acc[v]);
    const int row = ci.row;
    auto emit = [&](int, int f, const Vec<4>& sum) { store_vec<4>(outp + int64_t(row) * ldo + f, sum); };
    finish_row<kResLPE, 4, NV, 2>(csr, ci, acc, D, partial, emit);
  }
}

static bool resident_fits(int64_t n_nodes, int D) {
  return n_nodes > 0 && D > 0 && D % 4 == 0 && D <= 4 * kResLPE * 4 && n_nodes * int64_t(D) * 4 <= kResMaxSmem;
}

static int resident_grid(int64_t work_warps) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
  }
  const int64_t ctas = ceil_div(work_warps, kResWarps);
  return int(ctas < sms ? (ctas > 0 ? ctas : 1) : sms);
}

template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)) == cudaSuccess
             ? GN_OK
             : GN_ERR_CUDA;
}

}  // namespace gn

using namespace gn;

extern "C" {

int gn_distmult_resident_ok(int64_t n_nodes, int32_t D, int64_t ldz) {
  return resident_fits(n_nodes, D) && ldz % 4 == 0 ? 1 : 0;
}

int gn_distmult_fwd_resident(const float* z, int64_t ldz, int32_t n_nodes, int32_t D, const float* w,
                             const int64_t* src, const int64_t* dst, const int64_t* etype, int64_t n_edges,
                             int sigmoid, float* out, void* stream) {
  if (n_edges < 0 || D <= 0 || n_nodes <= 0) return GN_ERR_ARG;
  if (n_edges == 0) return GN_OK;
  if (!z || !w || !src || !dst || !etype || !out) return GN_ERR_ARG;
  if (!resident_fits(n_nodes, D) || ldz % 4 || !aligned16(z) || !aligned16(w)) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const size_t smem = size_t(n_nodes) * D * 4;
  const int grid = resident_grid(ceil_div(n_edges, 32));
  const int nv = (D + 4 * kResLPE - 1) / (4 * kResLPE);
#define GN_CASE(N)                                                                                             \
  if (nv == N) {                                                                                               \
    GN_CHECK(opt_in_smem(distmult_fwd_res_kernel<N>, smem));                                                   \
    GN_LAUNCH((distmult_fwd_res_kernel<N>), grid, kResThreads, smem, st, z, ldz, (int)n_nodes, (int)D, w, src, \
              dst, etype, n_edges, sigmoid, out);                                                              \
    return GN_OK;                                                                                              \
  }
  GN_CASE(1) GN_CASE(2) GN_CASE(3) GN_CASE(4)
#undef GN_CASE
  return GN_ERR_ARG;
}

static int launch_bwd_res(int mode, const gn_csr& csr, const int32_t* ent_a, const int32_t* ent_b,
                          const int32_t* ent_eid, const int64_t* src, const int64_t* dst, const float* coef,
                          const float* z, int64_t ldz, int n_nodes, int D, const float* w, float* outp, int64_t ldo,
                          float* partial, cudaStream_t st) {
  if (csr.n_rows == 0 || csr.n_chunks == 0) return GN_OK;
  if (csr.n_chunks > csr.n_rows && !partial) return GN_ERR_ARG;
  if (!resident_fits(n_nodes, D) || ldz % 4 || ldo % 4 || !aligned16(z) || !aligned16(outp) || (w && !aligned16(w)) ||
      (partial && !aligned16(partial)))
    return GN_ERR_ARG;
  const size_t smem = size_t(n_nodes) * D * 4;
  const int grid = resident_grid(csr.n_chunks);
  const int nv = (D + 4 * kResLPE - 1) / (4 * kResLPE);
#define GN_CASE(N, M)                                                                                              \
  if (nv == N && mode == M) {                                                                                      \
    GN_CHECK(opt_in_smem(distmult_bwd_res_kernel<N, M>, smem));                                                    \
    GN_LAUNCH((distmult_bwd_res_kernel<N, M>), grid, kResThreads, smem, st, csr, ent_a, ent_b, ent_eid, src, dst,  \
              coef, z, ldz, n_nodes, D, w, outp, ldo, partial);                                                    \
    return GN_OK;                                                                                                  \
  }
  GN_CASE(1, 0) GN_CASE(2, 0) GN_CASE(3, 0) GN_CASE(4, 0) GN_CASE(1, 1) GN_CASE(2, 1) GN_CASE(3, 1) GN_CASE(4, 1)
#undef GN_CASE
  return GN_ERR_ARG;
}

int gn_distmult_bwd_z_resident(const gn_csr* node_csr, const int32_t* ent_other, const int32_t* ent_rel,
                               const int32_t* ent_eid, const float* coef, const float* z, int64_t ldz, int32_t D,
                               const float* w, float* dz, int64_t lddz, float* partial, void* stream) {
  if (!node_csr || !z || !w || !dz || D <= 0) return GN_ERR_ARG;
  if (node_csr->nnz > 0 && (!ent_other || !ent_rel || !ent_eid || !coef)) return GN_ERR_ARG;
  return launch_bwd_res(0, *node_csr, ent_other, ent_rel, ent_eid, nullptr, nullptr, coef, z, ldz, node_csr->n_rows,
                        D, w, dz, lddz, partial, as_stream(stream));
}

int gn_distmult_bwd_w_resident(const gn_csr* rel_csr, const int32_t* rel_eid, const int64_t* src, const int64_t* dst,
                               const float* coef, const float* z, int64_t ldz, int32_t n_nodes, int32_t D, float* dw,
                               float* partial, void* stream) {
  if (!rel_csr || !z || !dw || D <= 0 || n_nodes <= 0) return GN_ERR_ARG;
  if (rel_csr->nnz > 0 && (!src || !dst || !coef)) return GN_ERR_ARG;      // rel_eid == NULL: identity order
  return launch_bwd_res(1, *rel_csr, nullptr, nullptr, rel_eid, src, dst, coef, z, ldz, n_nodes, D, nullptr, dw, D,
                        partial, as_stream(stream));
}

}  // extern "C"
