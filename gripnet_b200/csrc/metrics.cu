// K15: evaluation metrics on the device.
//
// Link prediction — replaces the per-relation host loop of GripNet-pose.py:148-164 / :188-199, which
// copies every relation's scores to the host and calls gripnet/utils.py:28-35 (auprc_auroc_ap: sklearn
// roc_auc_score, average_precision_score, auc(precision_recall_curve)) R times.  Here all relations are
// ranked by ONE stable radix sort (key 1: score descending, key 2: relation) and one CTA per relation
// walks its ranked slice once:
//   tp_k, fp_k  = positives / negatives with score >= the k-th DISTINCT threshold (ties form one point)
//   AUROC = sum_k (fp_k - fp_{k-1}) (tp_k + tp_{k-1}) / (2 P N)                [trapezoid from (0,0)]
//   AP    = sum_k (tp_k - tp_{k-1}) / P * tp_k / (tp_k + fp_k)                 [step-wise, no interpolation]
//   AUPRC = sum_k (tp_k - tp_{k-1}) / P * (prec_k + prec_{k-1}) / 2, prec_0=1  [trapezoid to (recall 0, precision 1)]
// Counts are exact integers; the three sums are float64 (what sklearn computes in) in a fixed order.
//
// Node classification — replaces torch.argmax + gripnet/utils.py:38-46 (micro_macro: sklearn f1_score
// micro / macro) and :49-52 (acc): one pass builds the C x C confusion matrix with integer atomics
// (order-independent), a single block turns it into micro-F1, macro-F1 and accuracy.
#include <cooperative_groups.h>

#include "common.cuh"

namespace gn {

// ---------------------------------------------------------------------------------------------
// link prediction
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int range_lookup(const int64_t* __restrict__ range, int n_rel, int64_t i) {
  int lo = 0, hi = n_rel;                       // last slice whose start <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (range[2 * mid] <= i) lo = mid; else hi = mid;
  }
  return (n_rel > 0 && range[2 * lo] <= i && i < range[2 * lo + 1]) ? lo : n_rel;   // n_rel: in no slice
}

__global__ void lp_rank_keys_kernel(const float* __restrict__ pos, int64_t n_pos, const float* __restrict__ neg,
                                    int64_t n_neg, const int64_t* __restrict__ pos_range,
                                    const int64_t* __restrict__ neg_range, int n_rel, int32_t* __restrict__ key,
                                    int32_t* __restrict__ rel) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_pos + n_neg) return;
  const bool is_pos = i < n_pos;
  const float s = is_pos ? pos[i] : neg[i - n_pos];
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // ascending u <=> ascending score
  key[i] = int32_t(~u);                               // ascending key <=> DESCENDING score
  rel[i] = is_pos ? range_lookup(pos_range, n_rel, i) : range_lookup(neg_range, n_rel, i - n_pos);
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int64_t n,
                                  int32_t* __restrict__ dst) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

constexpr int kMtThreads = 1024;   // wide tiles keep the walk of a ranked slice short
constexpr int kMtItems = 8;
constexpr int kMtTile = kMtThreads * kMtItems;
constexpr int kMtCluster = 8;      // CTAs (one thread-block cluster) per relation

// Running aggregate of a prefix of a relation's ranked slice: positives so far, the last threshold-end position
// (relative to the slice, -1 = none) and the positives up to and including that position.
struct ScanTriple {
  int sum, last, tp_last;
};

__device__ __forceinline__ ScanTriple combine(const ScanTriple a, const ScanTriple b) {   // a precedes b
  ScanTriple r;
  r.sum = a.sum + b.sum;
  if (b.last >= 0) { r.last = b.last; r.tp_last = a.sum + b.tp_last; }
  else { r.last = a.last; r.tp_last = a.tp_last; }
  return r;
}

__device__ __forceinline__ ScanTriple shfl_up(const ScanTriple v, int d) {
  return ScanTriple{__shfl_up_sync(kFull, v.sum, d), __shfl_up_sync(kFull, v.last, d),
                    __shfl_up_sync(kFull, v.tp_last, d)};
}

// One thread-block CLUSTER of kMtCluster CTAs per relation (round 1: one CTA per relation — 16 CTAs on 148 SMs at
// pose-0 size, 94 us, the tail of the training epoch).  CTA `rank` owns the rank-th contiguous part of the
// relation's ranked slice:
//   pass 1  aggregate of the part (positives, last threshold end, positives at that end) -> shared memory;
//   cluster barrier; the carry of a part = the aggregates of the parts before it, read over distributed shared memory;
//   pass 2  the trapezoid / step sums of the part's threshold ends, starting from the carry (a part of one tile
//           keeps its items in registers between the passes);
//   cluster barrier; CTA 0 adds the kMtCluster partial sums in part order and writes the record.
// Counts are exact integers; the float64 sums have a fixed order (thread items, shuffle tree, warps, parts).
__global__ void __cluster_dims__(kMtCluster, 1, 1) __launch_bounds__(kMtThreads, 1)
    lp_metrics_kernel(const float* __restrict__ pos, int64_t n_pos, const float* __restrict__ neg,
                      const int32_t* __restrict__ perm, const int32_t* __restrict__ rowptr, int n_rel,
                      double* __restrict__ record) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int r = int(blockIdx.x) / kMtCluster;
  const int rank = int(cluster.block_rank());
  const int beg = rowptr[r], end = rowptr[r + 1];
  const int len = end - beg;
  const int part = (len + kMtCluster - 1) / kMtCluster;
  const int c0 = min(len, rank * part), c1 = min(len, c0 + part);      // this CTA's positions [c0, c1)
  const int n_tiles = (c1 - c0 + kMtTile - 1) / kMtTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ ScanTriple s_warp[kMtThreads / 32];
  __shared__ ScanTriple s_total;           // running aggregate (carry + tiles so far) of the current pass
  __shared__ ScanTriple s_agg;             // pass-1 aggregate of this part: read by the other CTAs of the cluster
  __shared__ double s_red[2][kMtThreads / 32];
  __shared__ long long s_roc[kMtThreads / 32];
  __shared__ double s_part[2];             // this part's (ap, prc): read by CTA 0 of the cluster
  __shared__ long long s_part_roc;

  auto score_at = [&](int i) -> float {
    const int idx = perm[i];
    return idx < n_pos ? pos[idx] : neg[idx - n_pos];
  };

  int lab[kMtItems];
  bool bnd[kMtItems];
  ScanTriple mine;
  // items [j0, j0 + kMtItems) of the slice: labels, threshold ends, the thread's own aggregate
  auto load_tile = [&](int j0) {
    float sc[kMtItems + 1];
#pragma unroll
    for (int q = 0; q <= kMtItems; ++q) sc[q] = (j0 + q < len && j0 + q <= c1) ? score_at(beg + j0 + q) : 0.f;
    mine = ScanTriple{0, -1, 0};
#pragma unroll
    for (int q = 0; q < kMtItems; ++q) {
      const int j = j0 + q;
      const bool valid = j < c1;
      lab[q] = (valid && perm[beg + j] < n_pos) ? 1 : 0;
      bnd[q] = valid && (j + 1 == len || sc[q + 1] != sc[q]);
      mine.sum += lab[q];
      if (bnd[q]) { mine.last = j; mine.tp_last = mine.sum; }
    }
  };
  // exclusive prefix of `mine` over the CTA, on top of `carry`; s_total <- carry + the whole tile
  auto scan_tile = [&](const ScanTriple carry) -> ScanTriple {
    ScanTriple inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const ScanTriple o = shfl_up(inc, d);
      if (lane >= d) inc = combine(o, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    ScanTriple pre = carry;
    for (int w = 0; w < warp; ++w) pre = combine(pre, s_warp[w]);
    const ScanTriple e = shfl_up(inc, 1);
    if (lane > 0) pre = combine(pre, e);
    if (threadIdx.x == kMtThreads - 1) s_total = combine(pre, mine);
    __syncthreads();                                     // s_total visible, s_warp free for the next tile
    return pre;
  };

  // ---- pass 1: aggregate of this part
  ScanTriple carry{0, -1, 0};
  for (int t = 0; t < n_tiles; ++t) {
    load_tile(c0 + t * kMtTile + int(threadIdx.x) * kMtItems);
    (void)scan_tile(carry);
    carry = s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) s_agg = carry;
  cluster.sync();
  ScanTriple cin{0, -1, 0};
  for (int q = 0; q < rank; ++q) cin = combine(cin, *cluster.map_shared_rank(&s_agg, q));

  // ---- pass 2: sums over this part's threshold ends
  long long roc = 0;          // sum (fp_k - fp_{k-1}) (tp_k + tp_{k-1}); <= 2 P N < 2^63
  double ap = 0.0, prc = 0.0;
  carry = cin;
  for (int t = 0; t < n_tiles; ++t) {
    const int j0 = c0 + t * kMtTile + int(threadIdx.x) * kMtItems;
    if (n_tiles > 1) load_tile(j0);                      // a single tile is still in registers
    const ScanTriple pre = scan_tile(carry);
    carry = s_total;
    int tp = pre.sum;
    int prev = pre.last;
    long long tp_p = pre.last >= 0 ? pre.tp_last : 0;
#pragma unroll
    for (int q = 0; q < kMtItems; ++q) {
      tp += lab[q];
      if (!bnd[q]) continue;
      const int j = j0 + q;
      const long long tp_k = tp, fp_k = (long long)(j + 1) - tp_k;
      const long long fp_p = (long long)(prev + 1) - tp_p;
      roc += (fp_k - fp_p) * (tp_k + tp_p);
      const double prec_k = double(tp_k) / double(tp_k + fp_k);
      const double prec_p = prev >= 0 ? double(tp_p) / double(tp_p + fp_p) : 1.0;
      const double d_tp = double(tp_k - tp_p);
      ap += d_tp * prec_k;
      prc += d_tp * (prec_k + prec_p) * 0.5;
      prev = j;
      tp_p = tp_k;
    }
    __syncthreads();
  }
  // fixed-order block reduction
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    roc += __shfl_down_sync(kFull, roc, d);
    ap += __shfl_down_sync(kFull, ap, d);
    prc += __shfl_down_sync(kFull, prc, d);
  }
  if (lane == 0) {
    s_roc[warp] = roc;
    s_red[0][warp] = ap;
    s_red[1][warp] = prc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long roc_t = 0;
    double ap_t = 0.0, prc_t = 0.0;
    for (int w = 0; w < kMtThreads / 32; ++w) {
      roc_t += s_roc[w];
      ap_t += s_red[0][w];
      prc_t += s_red[1][w];
    }
    s_part_roc = roc_t;
    s_part[0] = ap_t;
    s_part[1] = prc_t;
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x == 0) {
    long long roc_t = 0;
    double ap_t = 0.0, prc_t = 0.0;
    int positives = 0;
    for (int q = 0; q < kMtCluster; ++q) {
      roc_t += *cluster.map_shared_rank(&s_part_roc, q);
      ap_t += cluster.map_shared_rank(s_part, q)[0];
      prc_t += cluster.map_shared_rank(s_part, q)[1];
      positives += cluster.map_shared_rank(&s_agg, q)->sum;
    }
    const double P = double(positives), N = double(len) - P;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    record[0 * n_rel + r] = P > 0 ? prc_t / P : nan;                         // auprc
    record[1 * n_rel + r] = (P > 0 && N > 0) ? double(roc_t) / (2.0 * P * N) : nan;   // auroc
    record[2 * n_rel + r] = P > 0 ? ap_t / P : nan;                          // ap
  }
  cluster.sync();            // no CTA leaves while CTA 0 still reads its shared memory
}

// ---------------------------------------------------------------------------------------------
// node classification
// ---------------------------------------------------------------------------------------------
__global__ void argmax_rows_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int C, int64_t* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* row = x + i * ldx;
  float best = row[0];
  int arg = 0;
  for (int c = 1; c < C; ++c) {
    const float v = row[c];
    if (v > best || (v != v && best == best)) {          // first maximum; NaN wins (torch.argmax)
      best = v;
      arg = c;
    }
  }
  out[i] = arg;
}

__global__ void confusion_kernel(const int64_t* __restrict__ target, const int64_t* __restrict__ pred, int64_t n,
                                 int C, int32_t* __restrict__ conf, int32_t* __restrict__ bad) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t t = target[i], p = pred[i];
  if (t < 0 || t >= C || p < 0 || p >= C) {
    atomicAdd(bad, 1);
    return;
  }
  atomicAdd(&conf[t * C + p], 1);                          // integer: order independent
}

// out[0] = micro-F1, out[1] = macro-F1 (mean over the classes PRESENT in target or pred, F1 := 0 where
// undefined — sklearn's zero_division="warn"), out[2] = accuracy
__global__ void nc_f1_kernel(const int32_t* __restrict__ conf, const int32_t* __restrict__ bad, int C, int64_t n,
                             double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long correct = 0;
  double macro = 0.0;
  int present = 0;
  for (int c = 0; c < C; ++c) {
    long long tp = conf[c * C + c], row = 0, col = 0;
    for (int k = 0; k < C; ++k) {
      row += conf[c * C + k];
      col += conf[k * C + c];
    }
    correct += tp;
    if (row + col == 0) continue;
    ++present;
    macro += (2.0 * double(tp)) / double(row + col);       // 2tp / (2tp + fp + fn)
  }
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  const bool ok = bad[0] == 0 && n > 0;
  const double acc = ok ? double(correct) / double(n) : nan;
  out[0] = acc;                                            // single-label multiclass: micro-F1 == accuracy
  out[1] = (ok && present > 0) ? macro / double(present) : nan;
  out[2] = acc;
}

}  // namespace gn

using namespace gn;

extern "C" {

size_t gn_lp_metrics_workspace_bytes(int64_t n_pos, int64_t n_neg, int32_t n_rel) {
  const int64_t n = n_pos + n_neg > 0 ? n_pos + n_neg : 1;
  return 5 * align_up(size_t(n) * 4) + align_up(size_t(n_rel + 2) * 4) + sort_ws_bytes(n) + 1024;
}

int gn_lp_metrics(const float* pos_score, int64_t n_pos, const float* neg_score, int64_t n_neg,
                  const int64_t* pos_range, const int64_t* neg_range, int32_t n_rel, double* record, void* ws,
                  size_t ws_bytes, void* stream) {
  if (n_pos < 0 || n_neg < 0 || n_rel <= 0 || pos_range == nullptr || record == nullptr) return GN_ERR_ARG;
  if ((n_pos > 0 && pos_score == nullptr) || (n_neg > 0 && neg_score == nullptr)) return GN_ERR_ARG;
  const int64_t n = n_pos + n_neg;
  if (n >= (int64_t(1) << 31)) return GN_ERR_RANGE;
  if (neg_range == nullptr) neg_range = pos_range;
  cudaStream_t st = as_stream(stream);
  Arena a(ws, ws_bytes);
  const size_t cap = size_t(n > 0 ? n : 1);
  int32_t* key = a.take<int32_t>(cap);      // rank keys, then the relation keys
  int32_t* rel = a.take<int32_t>(cap);
  int32_t* ksorted = a.take<int32_t>(cap);
  int32_t* perm1 = a.take<int32_t>(cap);
  int32_t* perm2 = a.take<int32_t>(cap);
  int32_t* rowptr = a.take<int32_t>(size_t(n_rel) + 2);
  if (!a.ok()) return GN_ERR_WORKSPACE;
  void* sws = a.base + a.off;
  const size_t sws_bytes = a.cap - a.off;
  if (n > 0) {
    const unsigned grid = (unsigned)ceil_div(n, 256);
    GN_LAUNCH(lp_rank_keys_kernel, grid, 256, 0, st, pos_score, n_pos, neg_score, n_neg, pos_range, neg_range,
              (int)n_rel, key, rel);
    GN_CHECK(sort_pairs(key, nullptr, ksorted, perm1, n, 32, sws, sws_bytes, st));
    GN_LAUNCH(gather_i32_kernel, grid, 256, 0, st, (const int32_t*)rel, (const int32_t*)perm1, n, key);
    GN_CHECK(sort_pairs(key, perm1, ksorted, perm2, n, bits_for(int64_t(n_rel) + 1), sws, sws_bytes, st));
  }
  GN_CHECK(rowptr_from_sorted(ksorted, n, n_rel + 1, rowptr, st));
  GN_LAUNCH(lp_metrics_kernel, (unsigned)n_rel * kMtCluster, kMtThreads, 0, st, pos_score, n_pos, neg_score,
            (const int32_t*)perm2, (const int32_t*)rowptr, (int)n_rel, record);
  return GN_OK;
}

int gn_argmax_rows(const float* x, int64_t ldx, int64_t n, int32_t C, int64_t* out, void* stream) {
  if (n < 0 || C <= 0 || (n > 0 && (x == nullptr || out == nullptr))) return GN_ERR_ARG;
  if (n == 0) return GN_OK;
  GN_LAUNCH(argmax_rows_kernel, (unsigned)ceil_div(n, 256), 256, 0, as_stream(stream), x, ldx, n, (int)C, out);
  return GN_OK;
}

size_t gn_nc_metrics_workspace_bytes(int32_t C) { return align_up(size_t(C) * size_t(C) * 4 + 4); }

int gn_nc_metrics(const int64_t* target, const int64_t* pred, int64_t n, int32_t C, double* out, void* ws,
                  size_t ws_bytes, void* stream) {
  if (n < 0 || C <= 0 || out == nullptr || (n > 0 && (target == nullptr || pred == nullptr))) return GN_ERR_ARG;
  if (ws == nullptr || ws_bytes < gn_nc_metrics_workspace_bytes(C)) return GN_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  int32_t* conf = static_cast<int32_t*>(ws);
  int32_t* bad = conf + size_t(C) * size_t(C);
  if (cudaMemsetAsync(conf, 0, size_t(C) * size_t(C) * 4 + 4, st) != cudaSuccess) return GN_ERR_CUDA;
  if (n > 0) GN_LAUNCH(confusion_kernel, (unsigned)ceil_div(n, 256), 256, 0, st, target, pred, n, (int)C, conf, bad);
  GN_LAUNCH(nc_f1_kernel, 1, 32, 0, st, (const int32_t*)conf, (const int32_t*)bad, (int)C, n, out);
  return GN_OK;
}

}  // extern "C"
