// Templated building blocks of the stable LSD counting sort (K1): per-tile digit
// histogram and stable scatter, parameterised on
//   RB      radix bits per pass (8 -> 256 bins, 10 -> 1024 bins),
//   KeySrc  where key i comes from (an int32 array, or the two int64 endpoint
//           arrays of an edge list read in place),
//   Sink    what is written at the sorted position (key/value pair, or the
//           endpoint-CSR entry of the decoder backward, fused).
// One pass = histogram -> exclusive scan (digit-major) -> scatter.  Stable, integer
// only, no data-dependent atomics on the output: bit-reproducible.
#pragma once
#include "common.cuh"

namespace gn {

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;                       // per thread
constexpr int kRsTile = kRsThreads * kRsItems;     // 4096 keys per block
constexpr int kRsWarpSpan = kRsItems * 32;         // contiguous keys owned by one warp

struct PlainKeys {
  const int32_t* k;
  __device__ __forceinline__ int key(int64_t i) const { return k[i]; }
};

struct LongKeys {
  const int64_t* k;
  __device__ __forceinline__ int key(int64_t i) const { return int(k[i]); }
};

struct PairSink {
  int32_t* keys_out;
  int32_t* vals_out;
  const int32_t* vals_in;   // nullptr: iota
  __device__ __forceinline__ void put(int pos, int key, int64_t idx) const {
    keys_out[pos] = key;
    vals_out[pos] = vals_in ? vals_in[idx] : int32_t(idx);
  }
};

// value only (single-pass sorts whose row pointers come from the histogram)
struct PermSink {
  int32_t* perm;
  __device__ __forceinline__ void put(int pos, int, int64_t idx) const { perm[pos] = int32_t(idx); }
};

template <int RB, typename KeySrc>
__global__ void __launch_bounds__(kRsThreads) rs_histogram(const KeySrc ks, int64_t n, int shift,
                                                           int32_t* __restrict__ hist, int n_tiles) {
  constexpr int RADIX = 1 << RB;
  __shared__ int h[RADIX];
  for (int i = threadIdx.x; i < RADIX; i += kRsThreads) h[i] = 0;
  __syncthreads();
  const int64_t base = int64_t(blockIdx.x) * kRsTile;
#pragma unroll
  for (int i = 0; i < kRsItems; ++i) {
    const int64_t idx = base + int64_t(i) * kRsThreads + threadIdx.x;
    if (idx < n) atomicAdd(&h[(ks.key(idx) >> shift) & (RADIX - 1)], 1);   // integer: order independent
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RADIX; d += kRsThreads) hist[int64_t(d) * n_tiles + blockIdx.x] = h[d];
}

// offsets: exclusive scan of hist (digit-major).  rowptr_out (optional, single-pass
// sorts with shift == 0): rowptr[r] = first output slot of key r, r in [0, n_rows].
template <int RB, typename KeySrc, typename Sink>
__global__ void __launch_bounds__(kRsThreads) rs_scatter(const KeySrc ks, const Sink sink, int64_t n, int shift,
                                                         const int32_t* __restrict__ offsets, int n_tiles,
                                                         int32_t* __restrict__ rowptr_out, int32_t n_rows) {
  constexpr int RADIX = 1 << RB;
  // cnt[w][d]: first the number of digit-d keys in warp w's span, then (after the
  // fix-up) the global output position of that warp's first digit-d key.
  __shared__ int cnt[kRsWarps][RADIX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kRsWarps * RADIX; i += kRsThreads) (&cnt[0][0])[i] = 0;
  if (rowptr_out != nullptr && blockIdx.x == 0) {
    for (int r = threadIdx.x; r <= n_rows; r += kRsThreads)
      rowptr_out[r] = (r < RADIX && r < n_rows) ? offsets[int64_t(r) * n_tiles] : int32_t(n);
  }
  __syncthreads();

  const int64_t warp_base = int64_t(blockIdx.x) * kRsTile + int64_t(warp) * kRsWarpSpan;
  int32_t key[kRsItems];
  int32_t rank[kRsItems];
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int64_t idx = warp_base + r * 32 + lane;      // warp owns a contiguous span, visited in order
    const bool valid = idx < n;
    key[r] = valid ? ks.key(idx) : 0;
    const int digit = valid ? ((key[r] >> shift) & (RADIX - 1)) : RADIX;  // sentinel groups the tail lanes
    const unsigned peers = __match_any_sync(kFull, digit);
    int prior = 0;
    if (valid) prior = cnt[warp][digit];
    __syncwarp();
    rank[r] = prior + __popc(peers & lt_mask);
    if (valid && (peers & lt_mask) == 0) cnt[warp][digit] = prior + __popc(peers);  // lowest peer updates
    __syncwarp();
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RADIX; d += kRsThreads) {
    int running = offsets[int64_t(d) * n_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      const int c = cnt[w][d];
      cnt[w][d] = running;
      running += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int64_t idx = warp_base + r * 32 + lane;
    if (idx < n) {
      const int digit = (key[r] >> shift) & (RADIX - 1);
      sink.put(cnt[warp][digit] + rank[r], key[r], idx);
    }
  }
}

inline int64_t rs_tiles(int64_t n) { return ceil_div(n > 0 ? n : 1, kRsTile); }

// radix bits per pass for keys of `key_bits` bits: fewest passes, then the narrowest digit
inline void rs_plan(int key_bits, int* passes, int* rb) {
  if (key_bits < 1) key_bits = 1;
  int p = (key_bits + 9) / 10;
  const int digit = (key_bits + p - 1) / p;
  *passes = p;
  *rb = digit <= 8 ? 8 : 10;
}

// One complete single-pass sort (keys < 2^RB): histogram, scan, scatter (+ rowptr).
// ws: hist (tiles << RB ints) + scan scratch.
template <int RB, typename KeySrc, typename Sink>
int rs_single_pass(const KeySrc& ks, const Sink& sink, int64_t n, int32_t* rowptr_out, int32_t n_rows, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  const int64_t tiles = rs_tiles(n);
  Arena a(ws, ws_bytes);
  int32_t* hist = a.take<int32_t>(size_t(tiles) << RB);
  if (!a.ok()) return GN_ERR_WORKSPACE;
  GN_LAUNCH((rs_histogram<RB, KeySrc>), (unsigned)tiles, kRsThreads, 0, st, ks, n, 0, hist, (int)tiles);
  GN_CHECK(exclusive_scan_i32(hist, hist, tiles << RB, nullptr, a.base + a.off, a.cap - a.off, st));
  GN_LAUNCH((rs_scatter<RB, KeySrc, Sink>), (unsigned)tiles, kRsThreads, 0, st, ks, sink, n, 0,
            (const int32_t*)hist, (int)tiles, rowptr_out, n_rows);
  return GN_OK;
}

inline size_t rs_single_pass_ws_bytes(int64_t n, int rb) {
  const int64_t hist = rs_tiles(n) << rb;
  return align_up(size_t(hist) * 4) + scan_ws_bytes(hist) + 512;
}

}  // namespace gn
