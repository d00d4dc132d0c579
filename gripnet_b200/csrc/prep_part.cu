// K1 for destination-partitioned graphs (new design, the reference is single-device):
// every rank builds ONLY its own rows of the dst-sorted CSR (forward) and of the src-sorted transpose CSR
// (backward) from the edges that touch its block — no rank ever sorts or stores the global CSR.
//
//   gn_edge_filter         order-preserving selection of the edges whose destination (or source) lies in
//                          [lo, hi): flag -> exclusive scan -> scatter.  Applied to the global list (or to
//                          the chunks of a streamed generator) it yields the rank's shard.
//   gn_gcn_part_structure  rows [row0, row0 + n_rows) of the CSR keyed by `key_end`: self-loop rewrite of
//                          gripnet/layers.py:52-60 (PyG add_remaining_self_loops) restricted to the block,
//                          stable counting sort by local row, loop entry appended at the end of each row,
//                          per-entry weight left in `val`, weighted in-degree / deg^-1/2 of the block.
//   gn_gcn_part_values     val <- (dis[source] * w) * dis[target] (layers.py:66-69) once the blocks'
//                          deg^-1/2 have been all-gathered: the only global quantity the rows need.
//
// Because the filter keeps the original relative order and the sort is stable, rows, columns and values
// are BIT-IDENTICAL to rows [row0, row0 + n_rows) of the global CSR built by gn_gcn_prep
// (tests/test_gpu_prep.py::test_partitioned_prep_equals_slices_of_the_global_csr).
#include "sortkit.cuh"

namespace gn {

__global__ void filter_flag_kernel(const int64_t* __restrict__ key, int64_t n, int64_t lo, int64_t hi,
                                   int32_t* __restrict__ flag) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int64_t k = key[e];
  flag[e] = (k >= lo && k < hi) ? 1 : 0;
}

__global__ void filter_scatter_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                      const float* __restrict__ weight, int64_t n, const int32_t* __restrict__ flag,
                                      const int32_t* __restrict__ pos, int64_t pos_base, int64_t* __restrict__ out_src,
                                      int64_t* __restrict__ out_dst, float* __restrict__ out_w,
                                      int64_t* __restrict__ out_idx, int64_t idx_base) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n || !flag[e]) return;
  const int64_t k = pos_base + pos[e];
  out_src[k] = src[e];
  out_dst[k] = dst[e];
  if (out_w) out_w[k] = weight ? weight[e] : 1.0f;
  if (out_idx) out_idx[k] = idx_base + e;
}

// keep = not a self-loop (square graphs); dropped loops remember the LAST one listed per node
__global__ void part_mark_kernel(const int64_t* __restrict__ key_end, const int64_t* __restrict__ other_end,
                                 int64_t n_edges, int with_loops, int32_t row0, int32_t n_rows,
                                 int32_t* __restrict__ key, int32_t* __restrict__ last_loop) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int r = int32_t(key_end[e] - row0);
  const bool keep = !with_loops || key_end[e] != other_end[e];
  key[e] = keep ? r : n_rows;                       // sentinel row: sorts behind every real row
  if (!keep) atomicMax(&last_loop[r], int32_t(e));  // integer max: order independent
}

__global__ void part_rowptr_kernel(const int32_t* __restrict__ rp_nl, int32_t n_rows, int with_loops, int32_t row0,
                                   const float* __restrict__ weight, const int32_t* __restrict__ last_loop,
                                   float fill_value, int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                   float* __restrict__ val, int32_t* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_rows) return;
  rowptr[i] = rp_nl[i] + (with_loops ? i : 0);
  if (i < n_rows) {
    if (with_loops) {
      const int s = rp_nl[i + 1] + i;               // last slot of row i: its self-loop
      col[s] = row0 + i;
      const int l = last_loop[i];
      val[s] = (l >= 0) ? (weight ? weight[l] : 1.0f) : fill_value;
    }
  } else if (counts) {
    counts[0] = rp_nl[n_rows] + (with_loops ? n_rows : 0);
  }
}

__global__ void part_assemble_kernel(const int32_t* __restrict__ sorted_key, const int32_t* __restrict__ sorted_pos,
                                     int64_t n_edges, int32_t n_rows, int with_loops,
                                     const int64_t* __restrict__ other_end, const float* __restrict__ weight,
                                     int32_t* __restrict__ col, float* __restrict__ val) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_edges) return;
  const int r = sorted_key[j];
  if (r >= n_rows) return;                           // dropped self-loop
  const int64_t s = j + (with_loops ? r : 0);
  const int k = sorted_pos[j];
  col[s] = int32_t(other_end[k]);
  val[s] = weight ? weight[k] : 1.0f;
}

// weighted in-degree by TARGET of the block's rows, summed in list order exactly as gcn_degree_kernel
// (prep.cu) does for the global graph, then deg^-1/2 with inf -> 0
__global__ void part_degree_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ val, int32_t n_rows,
                                   int with_loops, int unit_weight, float* __restrict__ deg, float* __restrict__ dis) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int b = rowptr[i], e = rowptr[i + 1];
  float d;
  if (!with_loops) {                                 // bipartite: the stacked graph's own loop on the target
    if (unit_weight) {
      d = float(e - b) + 1.0f;
    } else {
      d = 0.f;
      for (int s = b; s < e; ++s) d = __fadd_rn(d, val[s]);
      d = __fadd_rn(d, 1.0f);
    }
  } else {
    if (unit_weight) {
      d = float(e - b - 1) + val[e - 1];
    } else {
      d = 0.f;
      for (int s = b; s < e - 1; ++s) d = __fadd_rn(d, val[s]);
      d = __fadd_rn(d, val[e - 1]);
    }
  }
  if (deg) deg[i] = d;
  float r = float(pow(double(d), -0.5));
  if (isinf(r)) r = 0.f;
  dis[i] = r;
}

// val[s] <- (dis_src[source] * w) * dis_dst[target], warp per row
__global__ void part_values_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int32_t n_rows,
                                   int32_t row0, const float* __restrict__ dis_src, const float* __restrict__ dis_dst,
                                   int transpose, float* __restrict__ val) {
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  const int row = row0 + int(warp);
  const int b = rowptr[warp], e = rowptr[warp + 1];
  for (int s = b + lane; s < e; s += 32) {
    const int c = col[s];
    const int s_node = transpose ? row : c;
    const int d_node = transpose ? c : row;
    const float ds = dis_src ? dis_src[s_node] : 1.0f;
    val[s] = __fmul_rn(__fmul_rn(ds, val[s]), dis_dst[d_node]);
  }
}

inline unsigned grid1(int64_t n, int block = 256) { return (unsigned)(n > 0 ? ceil_div(n, block) : 1); }

}  // namespace gn

using namespace gn;

extern "C" {

size_t gn_edge_filter_workspace_bytes(int64_t n_edges) {
  const size_t E = size_t(n_edges > 0 ? n_edges : 1);
  return 2 * align_up((E + 1) * 4) + scan_ws_bytes(n_edges + 1) + 1024;
}

int gn_edge_filter(const int64_t* src, const int64_t* dst, const float* weight, int64_t n_edges, int by_src,
                   int64_t lo, int64_t hi, int64_t* out_src, int64_t* out_dst, float* out_weight, int64_t* out_index,
                   int64_t index_base, int32_t* count, void* ws, size_t ws_bytes, void* stream) {
  if (n_edges < 0 || !count || n_edges >= (int64_t(1) << 31)) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  if (n_edges == 0) {
    if (cudaMemsetAsync(count, 0, sizeof(int32_t), st) != cudaSuccess) return GN_ERR_CUDA;
    return GN_OK;
  }
  if (!src || !dst) return GN_ERR_ARG;
  Arena a(ws, ws_bytes);
  int32_t* flag = a.take<int32_t>(size_t(n_edges));
  int32_t* pos = a.take<int32_t>(size_t(n_edges) + 1);
  if (!a.ok()) return GN_ERR_WORKSPACE;
  GN_LAUNCH(filter_flag_kernel, grid1(n_edges), 256, 0, st, by_src ? src : dst, n_edges, lo, hi, flag);
  GN_CHECK(exclusive_scan_i32(flag, pos, n_edges, count, a.base + a.off, a.cap - a.off, st));
  if (out_src && out_dst) {
    GN_LAUNCH(filter_scatter_kernel, grid1(n_edges), 256, 0, st, src, dst, weight, n_edges, (const int32_t*)flag,
              (const int32_t*)pos, int64_t(0), out_src, out_dst, out_weight, out_index, index_base);
  }
  return GN_OK;
}

size_t gn_gcn_part_workspace_bytes(int64_t n_edges, int32_t n_rows) {
  const size_t E = size_t(n_edges > 0 ? n_edges : 1);
  const size_t N = size_t(n_rows) + 2;
  return 3 * align_up(E * 4) + 2 * align_up(N * 4) + sort_ws_bytes(n_edges) + 4096;
}

int gn_gcn_part_structure(const int64_t* key_end, const int64_t* other_end, const float* weight, int64_t n_edges,
                          int32_t row0, int32_t n_rows, int with_loops, float fill_value, int32_t* rowptr,
                          int32_t* col, float* val, float* deg, float* dis, int32_t* counts, void* ws,
                          size_t ws_bytes, void* stream) {
  if (n_edges < 0 || n_rows <= 0 || row0 < 0 || !rowptr || !col || !val || !counts) return GN_ERR_ARG;
  if (n_edges > 0 && (!key_end || !other_end)) return GN_ERR_ARG;
  if (n_edges + int64_t(n_rows) >= (int64_t(1) << 31)) return GN_ERR_RANGE;
  cudaStream_t st = as_stream(stream);
  const int64_t E = n_edges;
  const size_t Ea = size_t(E > 0 ? E : 1);
  Arena a(ws, ws_bytes);
  int32_t* key = a.take<int32_t>(Ea);
  int32_t* sorted_key = a.take<int32_t>(Ea);
  int32_t* sorted_pos = a.take<int32_t>(Ea);
  int32_t* last_loop = a.take<int32_t>(size_t(n_rows));
  int32_t* rp_nl = a.take<int32_t>(size_t(n_rows) + 2);
  if (!a.ok()) return GN_ERR_WORKSPACE;
  void* sub_ws = a.base + a.off;
  const size_t sub_bytes = a.cap - a.off;
  if (cudaMemsetAsync(last_loop, 0xFF, size_t(n_rows) * 4, st) != cudaSuccess) return GN_ERR_CUDA;
  if (cudaMemsetAsync(counts, 0, 4 * sizeof(int32_t), st) != cudaSuccess) return GN_ERR_CUDA;
  if (E > 0) {
    GN_LAUNCH(part_mark_kernel, grid1(E), 256, 0, st, key_end, other_end, E, with_loops, row0, n_rows, key, last_loop);
  }
  GN_CHECK(sort_pairs(key, nullptr, sorted_key, sorted_pos, E, bits_for(int64_t(n_rows) + 1), sub_ws, sub_bytes, st));
  GN_CHECK(rowptr_from_sorted(sorted_key, E, n_rows, rp_nl, st));
  GN_LAUNCH(part_rowptr_kernel, grid1(int64_t(n_rows) + 1), 256, 0, st, (const int32_t*)rp_nl, n_rows, with_loops,
            row0, weight, (const int32_t*)last_loop, fill_value, rowptr, col, val, counts);
  if (E > 0) {
    GN_LAUNCH(part_assemble_kernel, grid1(E), 256, 0, st, (const int32_t*)sorted_key, (const int32_t*)sorted_pos, E,
              n_rows, with_loops, other_end, weight, col, val);
  }
  if (dis) {
    GN_LAUNCH(part_degree_kernel, grid1(n_rows), 256, 0, st, (const int32_t*)rowptr, (const float*)val, n_rows,
              with_loops, weight == nullptr ? 1 : 0, deg, dis);
  }
  return GN_OK;
}

int gn_gcn_part_values(const int32_t* rowptr, const int32_t* col, int32_t n_rows, int32_t row0, const float* dis_src,
                       const float* dis_dst, int transpose, float* val, void* stream) {
  if (n_rows <= 0 || !rowptr || !col || !val || !dis_dst) return GN_ERR_ARG;
  GN_LAUNCH(part_values_kernel, grid1(int64_t(n_rows) * 32), 256, 0, as_stream(stream), rowptr, col, n_rows, row0,
            dis_src, dis_dst, transpose, val);
  return GN_OK;
}

}  // extern "C"
