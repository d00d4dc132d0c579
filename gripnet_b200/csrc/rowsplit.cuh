// Row-split, atomic-free row reductions shared by the SpMM and decoder-backward
// kernels.
//
// Work unit = one warp per CHUNK (<= chunk_len consecutive entries of one CSR
// row).  A warp is tiled as EPI = 32/LPE entry slots x LPE feature lanes; every
// feature lane owns NV vectors of VEC floats.  Rows that fit one chunk are
// finished in registers.  Rows split over k chunks write their partial sums to
// `partial[chunk]`; the LAST warp to arrive (a per-row counter, integer atomics
// only) re-reads the k partials in chunk order and finishes the row, so the
// floating-point summation order is fixed and results are bit-reproducible
// run to run (north_star: deterministic, atomic-free backward).
#pragma once
#include "common.cuh"

namespace gn {

template <int VEC>
struct Vec {
  float v[VEC];
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> load_vec(const float* p) {
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}

// coherent (L2) load: data written by other SMs during this kernel
template <int VEC>
__device__ __forceinline__ Vec<VEC> load_vec_cg(const float* p) {
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = __ldcg(p);
  }
  return r;
}

template <int VEC>
__device__ __forceinline__ Vec<VEC> load_vec_plain(const float* p) {
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = *p;
  }
  return r;
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const Vec<VEC>& a) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
  } else {
    *p = a.v[0];
  }
}

// sum over the entry slots of a warp (lanes that share the same feature lane)
template <int LPE, int VEC>
__device__ __forceinline__ void reduce_slots(Vec<VEC>& a) {
#pragma unroll
  for (int o = LPE; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) a.v[i] += __shfl_xor_sync(kFull, a.v[i], o);
  }
}

struct ChunkInfo {
  int chunk, row, beg, end, first_chunk, n_chunks_of_row;
};

// chunk `warp` of the work list (false when it lies beyond the list)
__device__ __forceinline__ bool chunk_info_at(const gn_csr& csr, int warp, ChunkInfo& ci) {
  if (warp >= csr.n_chunks) return false;
  if (csr.flags & GN_CSR_ROW_IS_CHUNK) {   // chunk i is row i: no chunk tables, one round trip
    ci.chunk = ci.row = ci.first_chunk = warp;
    ci.beg = __ldg(csr.rowptr + warp);
    ci.end = __ldg(csr.rowptr + warp + 1);
    ci.n_chunks_of_row = 1;
    return true;
  }
  if (warp >= __ldg(csr.chunk_ptr + csr.n_rows)) return false;
  ci.chunk = warp;
  ci.row = __ldg(csr.chunk_row + warp);
  ci.beg = __ldg(csr.chunk_beg + warp);
  const int row_end = __ldg(csr.rowptr + ci.row + 1);
  ci.end = min(ci.beg + csr.chunk_len, row_end);
  ci.first_chunk = __ldg(csr.chunk_ptr + ci.row);
  ci.n_chunks_of_row = __ldg(csr.chunk_ptr + ci.row + 1) - ci.first_chunk;
  return true;
}

// one warp per chunk, chunk = global warp index
__device__ __forceinline__ bool chunk_info(const gn_csr& csr, ChunkInfo& ci) {
  return chunk_info_at(csr, int((int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5), ci);
}

// Sum the k partial rows of a split row in chunk order (slot s takes chunks s, s+EPI, ...; the slots are
// combined afterwards in the fixed shuffle order).  Rows of a few hundred chunks (relation rows of the
// decoder backward, hub rows of power-law graphs) make this the tail of the kernel, so DEPTH partial rows x
// NV vectors are in flight per step instead of one dependent load at a time; the order of the additions is
// that of a plain loop.  DEPTH 1 costs no registers beyond the accumulators it replaces.
template <int LPE, int VEC, int NV, int DEPTH>
__device__ __forceinline__ void sum_partials(const float* __restrict__ rows, int k, int width, int slot, int fl,
                                          Vec<VEC> (&s)[NV]) {
  constexpr int EPI = 32 / LPE;
  constexpr int kTailDepth = DEPTH;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[v].v[i] = 0.f;
  for (int j0 = slot; j0 < k; j0 += kTailDepth * EPI) {
    Vec<VEC> p[kTailDepth][NV];
#pragma unroll
    for (int u = 0; u < kTailDepth; ++u) {
      const int j = j0 + u * EPI;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int f = (v * LPE + fl) * VEC;
        if (j < k && f < width) p[u][v] = load_vec_cg<VEC>(rows + int64_t(j) * width + f);
      }
    }
#pragma unroll
    for (int u = 0; u < kTailDepth; ++u) {
      const int j = j0 + u * EPI;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int f = (v * LPE + fl) * VEC;
        if (j < k && f < width) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) s[v].v[i] += p[u][v].v[i];
        }
      }
    }
  }
}

// Finish a row.  `acc[NV]` holds this warp's slot-reduced sums (valid on slot 0);
// `width` = number of floats per row; feature lane `fl` owns floats
// [ (v*LPE + fl)*VEC , +VEC ) for v < NV.  `emit(v, f, vec)` writes the result.
// Returns true (warp-uniform) on the warp that emitted the row: the only chunk of the row, or the last of its
// chunks to arrive.
template <int LPE, int VEC, int NV, int TAIL_DEPTH = 1, typename Emit>
__device__ __forceinline__ bool finish_row(const gn_csr& csr, const ChunkInfo& ci, Vec<VEC> (&acc)[NV], int width,
                                           float* __restrict__ partial, Emit emit) {
  constexpr int EPI = 32 / LPE;
  const int lane = threadIdx.x & 31;
  const int slot = lane / LPE, fl = lane % LPE;
  if (ci.n_chunks_of_row == 1) {
    if (slot == 0) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int f = (v * LPE + fl) * VEC;
        if (f < width) emit(v, f, acc[v]);
      }
    }
    return true;
  }
  // Partial-sum slots: a row with k > 1 chunks owns slots [2*(first_chunk-row), +k).
  // first_chunk-row = number of EXTRA chunks in earlier rows and k <= 2*(k-1), so the
  // ranges never overlap and 2*(n_chunks-n_rows) slots suffice.
  const int64_t pbase = 2 * int64_t(ci.first_chunk - ci.row);
  if (slot == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int f = (v * LPE + fl) * VEC;
      if (f < width) store_vec<VEC>(partial + (pbase + (ci.chunk - ci.first_chunk)) * width + f, acc[v]);
    }
  }
  __threadfence();
  __syncwarp();
  int last = 0;
  if (lane == 0) last = (atomicAdd(csr.row_counter + ci.row, 1) == ci.n_chunks_of_row - 1) ? 1 : 0;
  last = __shfl_sync(kFull, last, 0);
  if (!last) return false;
  __threadfence();
  Vec<VEC> s[NV];
  sum_partials<LPE, VEC, NV, TAIL_DEPTH>(partial + pbase * width, ci.n_chunks_of_row, width, slot, fl, s);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int f = (v * LPE + fl) * VEC;
    reduce_slots<LPE, VEC>(s[v]);
    if (slot == 0 && f < width) emit(v, f, s[v]);
  }
  if (lane == 0) csr.row_counter[ci.row] = 0;  // every warp of this row has arrived: safe to re-arm
  return true;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int pow2_ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace gn
