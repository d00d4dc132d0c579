// K13: negative sampling on the device.
//
// Reference: gripnet/utils.py:98-112 (`negative_sampling`) and :115-119 (`typed_negative_sampling`),
// called once per epoch (GripNet-pose.py:131): a device->host copy of the positives, numpy draws over
// N^2 pair codes, `np.isin` rejection against the positives, host->device copy of the result — ~95 ms
// per epoch at pose size, two orders of magnitude more than the whole training step takes here.
//
// Here: the positives are hashed once into an open-addressing set of (relation, pair-code) keys; every
// epoch one thread per edge draws counter-based Philox4x32-10 pair codes (counter = edge id, attempt,
// epoch; key = seed) until one misses the set.  Rejection per element == the reference's "redraw the
// rejected ones" loop: i.i.d. uniform over the non-positive pairs.  The draw is a pure function of
// (seed, epoch, edge, attempt), so the result does not depend on scheduling, and oracle/negsample.py
// reproduces it bit for bit.  The epoch counter lives in device memory and is advanced by the kernel,
// so the sampler can sit inside a captured CUDA graph and still produce fresh negatives per replay.
#include "common.cuh"

namespace gn {

constexpr unsigned long long kEmpty = ~0ull;

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {   // splitmix64 finaliser
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

// relation of edge e = index of the range_list slice holding it (ascending contiguous slices)
__device__ __forceinline__ int relation_of(const int64_t* __restrict__ range_list, int n_rel, int64_t e) {
  int lo = 0, hi = n_rel - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (e < __ldg(range_list + 2 * mid + 1)) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__global__ void neg_table_fill(unsigned long long* table, int64_t cap) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < cap) table[i] = kEmpty;
}

__global__ void neg_table_insert(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t n_edges,
                                 int64_t n_nodes, const int64_t* __restrict__ range_list, int n_rel,
                                 unsigned long long* table, unsigned long long mask) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const unsigned long long n2 = (unsigned long long)n_nodes * (unsigned long long)n_nodes;
  const unsigned long long rel = n_rel > 0 ? (unsigned long long)relation_of(range_list, n_rel, e) : 0ull;
  const unsigned long long key = rel * n2 + (unsigned long long)src[e] * (unsigned long long)n_nodes +
                                 (unsigned long long)dst[e];
  unsigned long long h = mix64(key) & mask;
  while (true) {   // set semantics: the final table content does not depend on insertion order
    const unsigned long long old = atomicCAS(table + h, kEmpty, key);
    if (old == kEmpty || old == key) return;
    h = (h + 1) & mask;
  }
}

__device__ __forceinline__ bool in_table(const unsigned long long* __restrict__ table, unsigned long long mask,
                                         unsigned long long key) {
  unsigned long long h = mix64(key) & mask;
  while (true) {
    const unsigned long long v = __ldg(table + h);
    if (v == key) return true;
    if (v == kEmpty) return false;
    h = (h + 1) & mask;
  }
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t& r0, uint32_t& r1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  r0 = c0; r1 = c1;
}

constexpr int kMaxAttempts = 1 << 16;   // the reference would loop forever if every pair were positive

__global__ void __launch_bounds__(256) neg_draw_kernel(const unsigned long long* __restrict__ table,
                                                       unsigned long long mask, int64_t n_edges, int64_t n_nodes,
                                                       const int64_t* __restrict__ range_list, int n_rel,
                                                       unsigned long long seed, unsigned long long* state,
                                                       int64_t* __restrict__ neg_src, int64_t* __restrict__ neg_dst) {
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(state);
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < n_edges) {
    const unsigned long long n2 = (unsigned long long)n_nodes * (unsigned long long)n_nodes;
    const unsigned long long rel = n_rel > 0 ? (unsigned long long)relation_of(range_list, n_rel, e) : 0ull;
    unsigned long long code = 0;
    for (int attempt = 0; attempt < kMaxAttempts; ++attempt) {
      uint32_t x0, x1;
      philox4x32_10(uint32_t(e), uint32_t((unsigned long long)e >> 32), uint32_t(attempt), uint32_t(epoch), uint32_t(seed),
                    uint32_t(seed >> 32), x0, x1);
      const unsigned long long u = (unsigned long long)x0 | ((unsigned long long)x1 << 32);
      code = __umul64hi(u, n2);
      if (!in_table(table, mask, rel * n2 + code)) break;
    }
    neg_src[e] = int64_t(code / (unsigned long long)n_nodes);
    neg_dst[e] = int64_t(code % (unsigned long long)n_nodes);
  }
  // the last block to finish advances the epoch (every block has read it by then) and re-arms the counter
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long prev = atomicAdd(state + 1, 1ull);
    if (prev == gridDim.x - 1) {
      state[1] = 0;
      state[0] = epoch + 1;
    }
  }
}

inline int64_t table_capacity(int64_t n_edges) {
  int64_t cap = 64;
  while (cap < 2 * n_edges + 1) cap <<= 1;
  return cap;
}

}  // namespace gn

using namespace gn;

extern "C" size_t gn_negsample_table_bytes(int64_t n_edges) {
  return n_edges < 0 ? 0 : size_t(table_capacity(n_edges)) * sizeof(unsigned long long);
}

static int check_common(int64_t n_edges, int64_t n_nodes, const int64_t* range_list, int32_t n_rel, const void* table,
                        size_t table_bytes) {
  if (n_edges < 0 || n_nodes <= 0 || n_rel < 0 || !table) return GN_ERR_ARG;
  if (n_rel > 0 && !range_list) return GN_ERR_ARG;
  if (table_bytes < gn_negsample_table_bytes(n_edges)) return GN_ERR_WORKSPACE;
  // keys are rel * N^2 + code and must stay below the EMPTY marker
  const long double top = (long double)(n_rel > 0 ? n_rel : 1) * (long double)n_nodes * (long double)n_nodes;
  if (top >= 9.0e18L) return GN_ERR_RANGE;
  return GN_OK;
}

extern "C" int gn_negsample_build(const int64_t* src, const int64_t* dst, int64_t n_edges, int64_t n_nodes,
                                  const int64_t* range_list, int32_t n_rel, void* table, size_t table_bytes,
                                  void* stream) {
  GN_CHECK(check_common(n_edges, n_nodes, range_list, n_rel, table, table_bytes));
  if (n_edges > 0 && (!src || !dst)) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const int64_t cap = table_capacity(n_edges);
  unsigned long long* t = static_cast<unsigned long long*>(table);
  GN_LAUNCH(neg_table_fill, (unsigned)ceil_div(cap, 256), 256, 0, st, t, cap);
  if (n_edges > 0)
    GN_LAUNCH(neg_table_insert, (unsigned)ceil_div(n_edges, 256), 256, 0, st, src, dst, n_edges, n_nodes, range_list,
              n_rel, t, (unsigned long long)(cap - 1));
  return GN_OK;
}

extern "C" int gn_negsample_draw(const void* table, size_t table_bytes, int64_t n_edges, int64_t n_nodes,
                                 const int64_t* range_list, int32_t n_rel, uint64_t seed, uint64_t* state,
                                 int64_t* neg_src, int64_t* neg_dst, void* stream) {
  GN_CHECK(check_common(n_edges, n_nodes, range_list, n_rel, table, table_bytes));
  if (!state) return GN_ERR_ARG;
  if (n_edges > 0 && (!neg_src || !neg_dst)) return GN_ERR_ARG;
  const int64_t cap = table_capacity(n_edges);
  const unsigned grid = (unsigned)ceil_div(n_edges > 0 ? n_edges : 1, 256);
  GN_LAUNCH(neg_draw_kernel, grid, 256, 0, as_stream(stream), static_cast<const unsigned long long*>(table),
            (unsigned long long)(cap - 1), n_edges, n_nodes, range_list, n_rel, (unsigned long long)seed,
            reinterpret_cast<unsigned long long*>(state), neg_src, neg_dst);
  return GN_OK;
}
