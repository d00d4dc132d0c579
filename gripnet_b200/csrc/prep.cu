// K1 graph preprocessing: int64 edge lists -> dst-sorted CSR + transpose CSR,
// GCN symmetric normalisation, multi-relational and decoder index structures.
// Reference semantics restated (not translated): gripnet/layers.py:52-69 + PyG
// add_remaining_self_loops, layers.py:165-189, :363-368.
#include "sortkit.cuh"

namespace gn {

// ---------------------------------------------------------------------------
// GCN prep
// ---------------------------------------------------------------------------
// pass 1: classify edges.  flag[e] = 1 for kept (non-loop) edges; for each node
// remember the LAST self-loop listed on it (its weight becomes the loop weight).
__global__ void gcn_mark_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t n_edges,
                                int bipartite, int32_t* __restrict__ flag, int32_t* __restrict__ last_loop) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const bool keep = bipartite || (src[e] != dst[e]);
  flag[e] = keep ? 1 : 0;
  if (!keep) atomicMax(&last_loop[src[e]], int32_t(e));  // integer max: order independent
}

// pass 2: compact kept edges to their augmented-list position; build sort keys.
__global__ void gcn_compact_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                   const float* __restrict__ weight, int64_t n_edges, const int32_t* __restrict__ flag,
                                   const int32_t* __restrict__ pos, int32_t n_src, int32_t n_dst,
                                   int32_t* __restrict__ c_src, int32_t* __restrict__ c_dst, float* __restrict__ c_w,
                                   int32_t* __restrict__ key_dst, int32_t* __restrict__ key_src,
                                   int32_t* __restrict__ val_pos, int64_t* __restrict__ aug_src,
                                   int64_t* __restrict__ aug_dst) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int k = pos[e];
  if (flag[e]) {
    const int s = int32_t(src[e]), d = int32_t(dst[e]);
    c_src[k] = s;
    c_dst[k] = d;
    c_w[k] = weight ? weight[e] : 1.0f;
    key_dst[e] = d;
    key_src[e] = s;
    if (aug_src) {
      aug_src[k] = s;
      aug_dst[k] = d;
    }
  } else {  // dropped self-loop: sentinel key sorts it behind every real row
    key_dst[e] = n_dst;
    key_src[e] = n_src;
  }
  val_pos[e] = k;
}

__global__ void gcn_loop_kernel(const float* __restrict__ weight, const int32_t* __restrict__ last_loop,
                                int32_t n_nodes, float fill_value, const int32_t* __restrict__ n_kept,
                                float* __restrict__ loop_w, int64_t* __restrict__ aug_src,
                                int64_t* __restrict__ aug_dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const int l = last_loop[i];
  loop_w[i] = (l >= 0) ? (weight ? weight[l] : 1.0f) : fill_value;
  if (aug_src) {
    const int64_t k = int64_t(*n_kept) + i;
    aug_src[k] = i;
    aug_dst[k] = i;
  }
}

// Assemble one CSR from the sorted kept edges.  `rp_nl` are row pointers over kept
// edges only; with loops every row gets one extra slot at its end.
__global__ void gcn_assemble_kernel(const int32_t* __restrict__ sorted_key, const int32_t* __restrict__ sorted_pos,
                                    int64_t n_edges, int32_t n_rows, int with_loops,
                                    const int32_t* __restrict__ other_end /* c_src for dst-CSR, c_dst for transpose */,
                                    int32_t* __restrict__ col, int32_t* __restrict__ perm) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_edges) return;
  const int r = sorted_key[j];
  if (r >= n_rows) return;  // dropped self-loop
  const int64_t s = j + (with_loops ? r : 0);
  const int k = sorted_pos[j];
  col[s] = other_end[k];
  perm[s] = k;
}

__global__ void gcn_rowptr_kernel(const int32_t* __restrict__ rp_nl, int32_t n_rows, int with_loops,
                                  int32_t* __restrict__ rowptr, int32_t* __restrict__ col, int32_t* __restrict__ perm,
                                  int32_t* __restrict__ indeg, int32_t* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_rows) return;
  const int shift = with_loops ? i : 0;
  rowptr[i] = rp_nl[i] + shift;
  if (i < n_rows) {
    if (with_loops) {
      const int s = rp_nl[i + 1] + i;  // last slot of row i
      col[s] = i;
      perm[s] = rp_nl[n_rows] + i;     // position of loop i in the augmented list
    }
    if (indeg) indeg[i] = rp_nl[i + 1] - rp_nl[i] + (with_loops ? 1 : 0);
  } else if (counts) {
    counts[0] = rp_nl[n_rows] + (with_loops ? n_rows : 0);
  }
}

// weighted in-degree by target, summed in augmented-list order (the order a
// sequential CPU index_add visits), then deg^-1/2 with inf -> 0.
__global__ void gcn_degree_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                                  const float* __restrict__ c_w, const float* __restrict__ loop_w,
                                  const int32_t* __restrict__ n_kept_ptr, int32_t n_rows, int bipartite,
                                  int unit_weight, float* __restrict__ deg, float* __restrict__ dis) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int b = rowptr[i], e = rowptr[i + 1];
  float d;
  if (bipartite) {
    if (unit_weight) {
      d = float(e - b) + 1.0f;
    } else {
      d = 0.f;
      for (int s = b; s < e; ++s) d = __fadd_rn(d, c_w[perm[s]]);
      d = __fadd_rn(d, 1.0f);  // the stacked graph's own loop on the target (layers.py:367-368)
    }
  } else {
    if (unit_weight) {
      d = float(e - b - 1) + loop_w[i];
    } else {
      d = 0.f;
      for (int s = b; s < e - 1; ++s) d = __fadd_rn(d, c_w[perm[s]]);
      d = __fadd_rn(d, loop_w[i]);
    }
  }
  deg[i] = d;
  float r = float(pow(double(d), -0.5));
  if (isinf(r)) r = 0.f;
  dis[i] = r;
}

// per-slot coefficient  (dis[row]*w)*dis[col]  in the reference's association order
__global__ void gcn_val_kernel(const int32_t* __restrict__ rowptr_unused, const int32_t* __restrict__ col,
                               const int32_t* __restrict__ perm, const int32_t* __restrict__ row_of_slot_key,
                               int64_t nnz_cap, const int32_t* __restrict__ nnz_ptr, const float* __restrict__ c_w,
                               const float* __restrict__ loop_w, const float* __restrict__ dis_src,
                               const float* __restrict__ dis_dst, const int32_t* __restrict__ n_kept_ptr,
                               int transpose, float* __restrict__ val) {
  (void)rowptr_unused;
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= nnz_cap || s >= *nnz_ptr) return;
  const int k = perm[s];
  const int n_kept = *n_kept_ptr;
  const int row = row_of_slot_key[s];
  const float w = (k < n_kept) ? c_w[k] : loop_w[k - n_kept];
  // dst-CSR: row = target, col = source;  transpose: row = source, col = target
  const int s_node = transpose ? row : col[s];
  const int d_node = transpose ? col[s] : row;
  const float ds = dis_src ? dis_src[s_node] : 1.0f;
  val[s] = __fmul_rn(__fmul_rn(ds, w), dis_dst[d_node]);
}

// row id of every CSR slot (expand rowptr), one thread per row walking its slots
__global__ void expand_rows_kernel(const int32_t* __restrict__ rowptr, int32_t n_rows, int32_t* __restrict__ row_of) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  const int b = rowptr[warp], e = rowptr[warp + 1];
  for (int s = b + lane; s < e; s += 32) row_of[s] = warp;
}

__global__ void gcn_aug_norm_kernel(const int32_t* __restrict__ c_src, const int32_t* __restrict__ c_dst,
                                    const float* __restrict__ c_w, const float* __restrict__ loop_w,
                                    const float* __restrict__ dis, const int32_t* __restrict__ n_kept_ptr,
                                    int32_t n_nodes, int64_t cap, float* __restrict__ aug_norm) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n_kept = *n_kept_ptr;
  if (k >= cap || k >= int64_t(n_kept) + n_nodes) return;
  if (k < n_kept) {
    aug_norm[k] = __fmul_rn(__fmul_rn(dis[c_src[k]], c_w[k]), dis[c_dst[k]]);
  } else {
    const int i = int(k - n_kept);
    aug_norm[k] = __fmul_rn(__fmul_rn(dis[i], loop_w[i]), dis[i]);
  }
}

// ---------------------------------------------------------------------------
// RGCN prep
// ---------------------------------------------------------------------------
__global__ void rgcn_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t n_edges,
                                 const int64_t* __restrict__ range_list, int32_t n_rel, int32_t* __restrict__ key_dst,
                                 int32_t* __restrict__ key_srcrel) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  // relation = last r with range_list[r][0] <= e   (ranges are an ascending partition of [0,E))
  int lo = 0, hi = n_rel - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (range_list[2 * mid] <= e) lo = mid; else hi = mid - 1;
  }
  // skip over empty ranges that share the same start: take the one whose end is > e
  while (lo > 0 && range_list[2 * lo + 1] <= e) --lo;
  key_dst[e] = int32_t(dst[e]);
  key_srcrel[e] = int32_t(src[e]) * n_rel + lo;
}

__global__ void rgcn_fwd_fill_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ key_srcrel,
                                     int64_t n_edges, int32_t* __restrict__ col) {
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s < n_edges) col[s] = key_srcrel[perm[s]];
}

__global__ void rgcn_invcnt_kernel(const int32_t* __restrict__ rowptr, int32_t n_nodes, float* __restrict__ inv_cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const int c = rowptr[i + 1] - rowptr[i];
  inv_cnt[i] = 1.0f / float(c > 1 ? c : 1);
}

__global__ void rgcn_bwd_fill_kernel(const int32_t* __restrict__ perm_t, const int32_t* __restrict__ key_dst,
                                     const float* __restrict__ inv_cnt, int64_t n_edges, int32_t* __restrict__ col_t,
                                     float* __restrict__ val_t) {
  const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= n_edges) return;
  const int d = key_dst[perm_t[s]];
  col_t[s] = d;
  val_t[s] = inv_cnt[d];
}

// ---------------------------------------------------------------------------
// index-list prep
// ---------------------------------------------------------------------------
__global__ void narrow_kernel(const int64_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = int32_t(in[i]);
}

inline unsigned grid1d(int64_t n, int block = 256) { return (unsigned)(n > 0 ? ceil_div(n, block) : 1); }

}  // namespace gn

using namespace gn;

extern "C" {

size_t gn_gcn_prep_workspace_bytes(int64_t n_edges, int32_t n_src, int32_t n_dst) {
  const size_t E = size_t(n_edges > 0 ? n_edges : 1);
  const size_t N = size_t(n_src > n_dst ? n_src : n_dst) + 2;
  const size_t cap = E + N;
  return 10 * align_up(E * 4) + 8 * align_up(N * 4) + 2 * align_up(cap * 4) + scan_ws_bytes(n_edges + 1) +
         sort_ws_bytes(n_edges) + 4096;
}

int gn_gcn_prep(const int64_t* src, const int64_t* dst, const float* weight, int64_t n_edges, int32_t n_src,
                int32_t n_dst, int bipartite, float fill_value, int64_t* aug_src, int64_t* aug_dst, float* aug_norm,
                int32_t* rowptr, int32_t* col, float* val, int32_t* perm, int32_t* rowptr_t, int32_t* col_t,
                float* val_t, int32_t* perm_t, float* deg, int32_t* indeg, int32_t* counts, void* ws, size_t ws_bytes,
                void* stream) {
  if (n_edges < 0 || n_src <= 0 || n_dst <= 0) return GN_ERR_ARG;
  if (!bipartite && n_src != n_dst) return GN_ERR_ARG;
  if (n_edges > 0 && (src == nullptr || dst == nullptr)) return GN_ERR_ARG;
  if (!rowptr || !col || !val || !perm || !rowptr_t || !col_t || !val_t || !perm_t || !deg || !counts)
    return GN_ERR_ARG;
  if (n_edges + int64_t(n_dst) >= (int64_t(1) << 31)) return GN_ERR_RANGE;
  cudaStream_t st = as_stream(stream);
  const int with_loops = bipartite ? 0 : 1;
  const int64_t E = n_edges;
  const size_t Ea = size_t(E > 0 ? E : 1);
  const int64_t cap = E + (with_loops ? n_dst : 0);

  Arena a(ws, ws_bytes);
  int32_t* flag = a.take<int32_t>(Ea);
  int32_t* pos = a.take<int32_t>(Ea + 1);
  int32_t* c_src = a.take<int32_t>(Ea);
  int32_t* c_dst = a.take<int32_t>(Ea);
  float* c_w = a.take<float>(Ea);
  int32_t* key_dst = a.take<int32_t>(Ea);
  int32_t* key_src = a.take<int32_t>(Ea);
  int32_t* val_pos = a.take<int32_t>(Ea);
  int32_t* sorted_key = a.take<int32_t>(Ea);
  int32_t* sorted_pos = a.take<int32_t>(Ea);
  int32_t* last_loop = a.take<int32_t>(size_t(n_dst));
  float* loop_w = a.take<float>(size_t(n_dst));
  float* dis_dst = a.take<float>(size_t(n_dst));
  int32_t* rp_nl = a.take<int32_t>(size_t(n_dst) + 2);
  int32_t* rp_nl_t = a.take<int32_t>(size_t(n_src) + 2);
  int32_t* n_kept = a.take<int32_t>(64);
  int32_t* row_of = a.take<int32_t>(size_t(cap > 0 ? cap : 1));
  if (!a.ok()) return GN_ERR_WORKSPACE;
  void* sub_ws = a.base + a.off;
  const size_t sub_bytes = a.cap - a.off;

  if (cudaMemsetAsync(last_loop, 0xFF, size_t(n_dst) * 4, st) != cudaSuccess) return GN_ERR_CUDA;
  if (cudaMemsetAsync(counts, 0, 4 * sizeof(int32_t), st) != cudaSuccess) return GN_ERR_CUDA;
  if (E > 0) {
    GN_LAUNCH(gcn_mark_kernel, grid1d(E), 256, 0, st, src, dst, E, bipartite, flag, last_loop);
  }
  GN_CHECK(exclusive_scan_i32(flag, pos, E, n_kept, sub_ws, sub_bytes, st));
  if (E > 0) {
    GN_LAUNCH(gcn_compact_kernel, grid1d(E), 256, 0, st, src, dst, weight, E, (const int32_t*)flag,
              (const int32_t*)pos, n_src, n_dst, c_src, c_dst, c_w, key_dst, key_src, val_pos, aug_src, aug_dst);
  }
  if (with_loops) {
    GN_LAUNCH(gcn_loop_kernel, grid1d(n_dst), 256, 0, st, weight, (const int32_t*)last_loop, n_dst, fill_value,
              (const int32_t*)n_kept, loop_w, aug_src, aug_dst);
  }
  // ---- dst-sorted CSR
  GN_CHECK(sort_pairs(key_dst, val_pos, sorted_key, sorted_pos, E, bits_for(int64_t(n_dst) + 1), sub_ws, sub_bytes, st));
  GN_CHECK(rowptr_from_sorted(sorted_key, E, n_dst, rp_nl, st));
  GN_LAUNCH(gcn_rowptr_kernel, grid1d(int64_t(n_dst) + 1), 256, 0, st, (const int32_t*)rp_nl, n_dst, with_loops,
            rowptr, col, perm, indeg, counts);
  if (E > 0) {
    GN_LAUNCH(gcn_assemble_kernel, grid1d(E), 256, 0, st, (const int32_t*)sorted_key, (const int32_t*)sorted_pos, E,
              n_dst, with_loops, (const int32_t*)c_src, col, perm);
  }
  const int unit_weight = weight == nullptr ? 1 : 0;
  GN_LAUNCH(gcn_degree_kernel, grid1d(n_dst), 256, 0, st, (const int32_t*)rowptr, (const int32_t*)perm,
            (const float*)c_w, (const float*)loop_w, (const int32_t*)n_kept, n_dst, bipartite, unit_weight, deg,
            dis_dst);
  const float* dis_src = bipartite ? nullptr : dis_dst;
  if (cap > 0) {
    GN_LAUNCH(expand_rows_kernel, grid1d(int64_t(n_dst) * 32), 256, 0, st, (const int32_t*)rowptr, n_dst, row_of);
    GN_LAUNCH(gcn_val_kernel, grid1d(cap), 256, 0, st, (const int32_t*)rowptr, (const int32_t*)col,
              (const int32_t*)perm, (const int32_t*)row_of, cap, (const int32_t*)counts, (const float*)c_w,
              (const float*)loop_w, dis_src, (const float*)dis_dst, (const int32_t*)n_kept, 0, val);
  }
  if (aug_norm != nullptr && with_loops && cap > 0) {
    GN_LAUNCH(gcn_aug_norm_kernel, grid1d(cap), 256, 0, st, (const int32_t*)c_src, (const int32_t*)c_dst,
              (const float*)c_w, (const float*)loop_w, (const float*)dis_dst, (const int32_t*)n_kept, n_dst, cap,
              aug_norm);
  }
  // ---- src-sorted (transpose) CSR
  GN_CHECK(sort_pairs(key_src, val_pos, sorted_key, sorted_pos, E, bits_for(int64_t(n_src) + 1), sub_ws, sub_bytes, st));
  GN_CHECK(rowptr_from_sorted(sorted_key, E, n_src, rp_nl_t, st));
  GN_LAUNCH(gcn_rowptr_kernel, grid1d(int64_t(n_src) + 1), 256, 0, st, (const int32_t*)rp_nl_t, n_src, with_loops,
            rowptr_t, col_t, perm_t, (int32_t*)nullptr, (int32_t*)nullptr);
  if (E > 0) {
    GN_LAUNCH(gcn_assemble_kernel, grid1d(E), 256, 0, st, (const int32_t*)sorted_key, (const int32_t*)sorted_pos, E,
              n_src, with_loops, (const int32_t*)c_dst, col_t, perm_t);
  }
  if (cap > 0) {
    GN_LAUNCH(expand_rows_kernel, grid1d(int64_t(n_src) * 32), 256, 0, st, (const int32_t*)rowptr_t, n_src, row_of);
    GN_LAUNCH(gcn_val_kernel, grid1d(cap), 256, 0, st, (const int32_t*)rowptr_t, (const int32_t*)col_t,
              (const int32_t*)perm_t, (const int32_t*)row_of, cap, (const int32_t*)counts, (const float*)c_w,
              (const float*)loop_w, dis_src, (const float*)dis_dst, (const int32_t*)n_kept, 1, val_t);
  }
  // counts[1] = number of self-loops removed = E - n_kept (computed lazily on host from counts[0])
  return GN_OK;
}

size_t gn_rgcn_prep_workspace_bytes(int64_t n_edges, int32_t n_nodes, int32_t n_rel) {
  (void)n_nodes;
  (void)n_rel;
  const size_t E = size_t(n_edges > 0 ? n_edges : 1);
  return 3 * align_up(E * 4) + sort_ws_bytes(n_edges) + 4096;
}

int gn_rgcn_prep(const int64_t* src, const int64_t* dst, int64_t n_edges, const int64_t* range_list, int32_t n_nodes,
                 int32_t n_rel, int32_t* rowptr, int32_t* col, int32_t* perm, float* inv_cnt, int32_t* rowptr_t,
                 int32_t* col_t, float* val_t, int32_t* perm_t, void* ws, size_t ws_bytes, void* stream) {
  if (n_edges < 0 || n_nodes <= 0 || n_rel <= 0 || !rowptr || !inv_cnt || !rowptr_t) return GN_ERR_ARG;
  if (n_edges > 0 && (!src || !dst || !range_list || !col || !perm || !col_t || !val_t || !perm_t)) return GN_ERR_ARG;
  if (int64_t(n_nodes) * n_rel >= (int64_t(1) << 31) || n_edges >= (int64_t(1) << 31)) return GN_ERR_RANGE;
  cudaStream_t st = as_stream(stream);
  const int64_t E = n_edges;
  const size_t Ea = size_t(E > 0 ? E : 1);
  Arena a(ws, ws_bytes);
  int32_t* key_dst = a.take<int32_t>(Ea);
  int32_t* key_sr = a.take<int32_t>(Ea);
  int32_t* sorted = a.take<int32_t>(Ea);
  if (!a.ok()) return GN_ERR_WORKSPACE;
  void* sub_ws = a.base + a.off;
  const size_t sub_bytes = a.cap - a.off;
  const int32_t n_rows_t = n_nodes * n_rel;
  if (E > 0) {
    GN_LAUNCH(rgcn_keys_kernel, grid1d(E), 256, 0, st, src, dst, E, range_list, n_rel, key_dst, key_sr);
  }
  GN_CHECK(sort_pairs(key_dst, nullptr, sorted, perm, E, bits_for(n_nodes > 1 ? n_nodes : 2), sub_ws, sub_bytes, st));
  GN_CHECK(rowptr_from_sorted(sorted, E, n_nodes, rowptr, st));
  GN_LAUNCH(rgcn_invcnt_kernel, grid1d(n_nodes), 256, 0, st, (const int32_t*)rowptr, n_nodes, inv_cnt);
  if (E > 0) {
    GN_LAUNCH(rgcn_fwd_fill_kernel, grid1d(E), 256, 0, st, (const int32_t*)perm, (const int32_t*)key_sr, E, col);
  }
  GN_CHECK(sort_pairs(key_sr, nullptr, sorted, perm_t, E, bits_for(n_rows_t > 1 ? n_rows_t : 2), sub_ws, sub_bytes, st));
  GN_CHECK(rowptr_from_sorted(sorted, E, n_rows_t, rowptr_t, st));
  if (E > 0) {
    GN_LAUNCH(rgcn_bwd_fill_kernel, grid1d(E), 256, 0, st, (const int32_t*)perm_t, (const int32_t*)key_dst,
              (const float*)inv_cnt, E, col_t, val_t);
  }
  return GN_OK;
}

size_t gn_index_prep_workspace_bytes(int64_t n, int32_t n_nodes) {
  const size_t na = size_t(n > 0 ? n : 1);
  return align_up(na * 4) + gn_csr_from_keys_workspace_bytes(n, n_nodes) + 1024;
}

int gn_index_prep(const int64_t* index, int64_t n, int32_t n_nodes, int32_t* rowptr, int32_t* perm, void* ws,
                  size_t ws_bytes, void* stream) {
  if (n < 0 || n_nodes <= 0 || !rowptr) return GN_ERR_ARG;
  if (n > 0 && (!index || !perm)) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const int bits = bits_for(n_nodes > 1 ? n_nodes : 2);
  if (bits <= 10) {  // one counting pass straight from the int64 list
    const LongKeys ks{index};
    const PermSink sink{perm};
    if (bits <= 8) return rs_single_pass<8>(ks, sink, n, rowptr, n_nodes, ws, ws_bytes, st);
    return rs_single_pass<10>(ks, sink, n, rowptr, n_nodes, ws, ws_bytes, st);
  }
  Arena a(ws, ws_bytes);
  int32_t* keys = a.take<int32_t>(size_t(n > 0 ? n : 1));
  if (!a.ok()) return GN_ERR_WORKSPACE;
  if (n > 0) {
    GN_LAUNCH(narrow_kernel, grid1d(n), 256, 0, st, index, n, keys);
  }
  return gn_csr_from_keys(keys, n, n_nodes, rowptr, perm, a.base + a.off, a.cap - a.off, stream);
}

}  // extern "C"
