// K9/K10, dense-relation form — EXPERIMENTAL: parity-checked on a B200 at the end of round 1 (3 shapes incl.
// pose-0 size, profiles/r01_v16_dense_decoder_pytest.txt) but NOT YET TIMED, so off by default
// (ops.DECODER_PATH == "dense" selects it).  DESIGN.md §9a has the reasoning.
//
// The link decoder of the pose family scores edges inside a SMALL task supervertex (645 drugs) whose relation
// slices are dense: 25 k edges per relation = 6 % of the 645^2 node pairs.  Then, per relation r,
//     S_r = (z .* w_r) z^T                      score_e = S_r[src_e, dst_e]                (forward)
//     C_r[n, m] = sum of coef_e over edges of relation r joining n and m (both directions)
//     T_r = C_r z        dz = sum_r T_r .* w_r        dw_r = 1/2 sum_n z_n .* T_r[n]       (backward)
// (identities checked in float64 by scratch/dense_decoder_identity.py) turn the two 320-byte row gathers per edge
// and the three gather walks of the backward into R batched dense products — 16x the arithmetic, no row gathers —
// plus one 4-byte gather per edge.  The products go through the library's GEMM entry points (the host side issues
// them); this file holds the glue kernels.  Determinism as everywhere else: C_r is accumulated by ONE warp per
// node row walking the endpoint CSR in entry order (duplicates inside a 32-entry batch are combined in lane
// order), no floating-point atomics.
#include "common.cuh"

namespace gn {

__global__ void dense_scale_kernel(const float* __restrict__ z, int64_t ldz, int n, int D, const float* __restrict__ w,
                                   int R, float* __restrict__ zw) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t per_rel = int64_t(n) * D;
  if (idx >= per_rel * R) return;
  const int r = int(idx / per_rel);
  const int64_t rem = idx - int64_t(r) * per_rel;
  const int row = int(rem / D), k = int(rem - int64_t(row) * D);
  zw[idx] = z[int64_t(row) * ldz + k] * w[int64_t(r) * D + k];
}

__global__ void dense_score_kernel(const float* __restrict__ S, int n, const int64_t* __restrict__ src,
                                   const int64_t* __restrict__ dst, const int64_t* __restrict__ etype, int64_t n_edges,
                                   int sigmoid, float* __restrict__ out) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const float s = __ldg(S + (etype[e] * n + src[e]) * int64_t(n) + dst[e]);
  out[e] = sigmoid ? 1.0f / (1.0f + expf(-s)) : s;
}

// one warp per node row of the endpoint CSR; C is [R][n][n], this warp is the only writer of C[*][row][*]
__global__ void __launch_bounds__(256) dense_coef_kernel(const int32_t* __restrict__ rowptr,
                                                         const int32_t* __restrict__ ent_other,
                                                         const int32_t* __restrict__ ent_rel,
                                                         const int32_t* __restrict__ ent_eid,
                                                         const float* __restrict__ coef, int n, float* C) {
  const int row = int((int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  for (int base = beg; base < end; base += 32) {
    const int i = base + lane;
    const bool valid = i < end;
    int other = 0, rel = 0;
    float g = 0.f;
    if (valid) {
      other = __ldg(ent_other + i);
      rel = __ldg(ent_rel + i);
      g = __ldg(coef + __ldg(ent_eid + i));
    }
    const int key = valid ? rel * n + other : -1 - lane;        // invalid lanes never match anybody
    const unsigned peers = __match_any_sync(kFull, key);
    float sum = 0.f;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) {                              // lane order: a fixed summation order
      const float v = __shfl_sync(kFull, g, l);
      if ((peers >> l) & 1u) sum += v;
    }
    if (valid && lane == __ffs(peers) - 1) {
      float* p = C + (int64_t(rel) * n + row) * int64_t(n) + other;
      *p += sum;
    }
    __syncwarp();                                               // the next batch may touch the same element
  }
}

__global__ void dense_dz_kernel(const float* __restrict__ T, int n, int D, int R, const float* __restrict__ w,
                                float* __restrict__ dz, int64_t lddz) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= int64_t(n) * D) return;
  const int row = int(idx / D), k = int(idx - int64_t(row) * D);
  float s = 0.f;
  for (int r = 0; r < R; ++r) s = fmaf(T[(int64_t(r) * n + row) * D + k], w[int64_t(r) * D + k], s);
  dz[int64_t(row) * lddz + k] = s;
}

// one thread per (relation, feature): rows added in order
__global__ void dense_dw_kernel(const float* __restrict__ T, int n, int D, int R, const float* __restrict__ z,
                                int64_t ldz, float* __restrict__ dw) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * D) return;
  const int r = idx / D, k = idx - r * D;
  float s = 0.f;
  for (int row = 0; row < n; ++row) s = fmaf(z[int64_t(row) * ldz + k], T[(int64_t(r) * n + row) * D + k], s);
  dw[idx] = 0.5f * s;
}

}  // namespace gn

using namespace gn;

extern "C" {

int gn_distmult_dense_scale(const float* z, int64_t ldz, int32_t n_nodes, int32_t D, const float* w, int32_t n_rel,
                            float* zw, void* stream) {
  if (!z || !w || !zw || n_nodes <= 0 || D <= 0 || n_rel <= 0) return GN_ERR_ARG;
  const int64_t total = int64_t(n_nodes) * D * n_rel;
  GN_LAUNCH(dense_scale_kernel, (unsigned)ceil_div(total, 256), 256, 0, as_stream(stream), z, ldz, (int)n_nodes, (int)D,
            w, (int)n_rel, zw);
  return GN_OK;
}

int gn_distmult_dense_scores(const float* S, int32_t n_nodes, const int64_t* src, const int64_t* dst,
                             const int64_t* etype, int64_t n_edges, int sigmoid, float* out, void* stream) {
  if (n_edges < 0 || n_nodes <= 0) return GN_ERR_ARG;
  if (n_edges == 0) return GN_OK;
  if (!S || !src || !dst || !etype || !out) return GN_ERR_ARG;
  GN_LAUNCH(dense_score_kernel, (unsigned)ceil_div(n_edges, 256), 256, 0, as_stream(stream), S, (int)n_nodes, src, dst,
            etype, n_edges, sigmoid, out);
  return GN_OK;
}

int gn_distmult_dense_coef(const int32_t* node_rowptr, const int32_t* ent_other, const int32_t* ent_rel,
                           const int32_t* ent_eid, const float* coef, int32_t n_nodes, int32_t n_rel, int zero_first,
                           float* C, void* stream) {
  if (!node_rowptr || !C || n_nodes <= 0 || n_rel <= 0) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  if (zero_first &&
      cudaMemsetAsync(C, 0, size_t(n_rel) * size_t(n_nodes) * size_t(n_nodes) * sizeof(float), st) != cudaSuccess)
    return GN_ERR_CUDA;
  if (!ent_other || !ent_rel || !ent_eid || !coef) return GN_ERR_ARG;
  GN_LAUNCH(dense_coef_kernel, (unsigned)ceil_div(int64_t(n_nodes) * 32, 256), 256, 0, st, node_rowptr, ent_other,
            ent_rel, ent_eid, coef, (int)n_nodes, C);
  return GN_OK;
}

int gn_distmult_dense_grads(const float* T, int32_t n_nodes, int32_t D, int32_t n_rel, const float* z, int64_t ldz,
                            const float* w, float* dz, int64_t lddz, float* dw, void* stream) {
  if (!T || !z || !w || n_nodes <= 0 || D <= 0 || n_rel <= 0) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  if (dz) GN_LAUNCH(dense_dz_kernel, (unsigned)ceil_div(int64_t(n_nodes) * D, 256), 256, 0, st, T, (int)n_nodes, (int)D,
                    (int)n_rel, w, dz, lddz);
  if (dw) GN_LAUNCH(dense_dw_kernel, (unsigned)ceil_div(int64_t(n_rel) * D, 128), 128, 0, st, T, (int)n_nodes, (int)D,
                    (int)n_rel, z, ldz, dw);
  return GN_OK;
}

}  // extern "C"
