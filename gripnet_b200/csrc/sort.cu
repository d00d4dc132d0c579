// K1 building blocks: exclusive scan, stable LSD radix sort (counting sort per 8- or
// 10-bit digit), CSR row pointers from sorted keys, row-split chunk lists.
//
// All of it is integer work and fully deterministic: the sort is stable, so slot k
// of a CSR row holds the row's entries in their original relative order — the
// bit-exact contract of SURVEY.md §8 a1 (checked against numpy argsort(kind="stable")).
#include "sortkit.cuh"

namespace gn {

std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_last_cuda_error{0};

int bits_for(int64_t n_values) {
  int b = 1;
  while ((int64_t(1) << b) < n_values) ++b;
  return b;
}

// ----------------------------------------------------------------------------
// exclusive scan: tile reduce -> recursive scan of tile sums -> tile scan
// ----------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem /*[kScanThreads/32]*/, int* block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < kScanThreads / 32) ? smem[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(kFull, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < kScanThreads / 32) smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == kScanThreads / 32 - 1) *block_total = wi;
  }
  __syncthreads();
  return incl - v + smem[warp];
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int32_t* __restrict__ in, int64_t n,
                                                               int32_t* __restrict__ tile_sums) {
  __shared__ int red[kScanThreads / 32];
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += red[w];
    tile_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_tiles(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                           int64_t n, const int32_t* __restrict__ tile_offsets,
                                                           int32_t* __restrict__ total_out) {
  __shared__ int sm[kScanThreads / 32];
  __shared__ int block_total;
  const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int off = block_exclusive_scan(s, sm, &block_total) + (tile_offsets ? tile_offsets[blockIdx.x] : 0);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = off;
    off += v[i];
  }
  if (total_out != nullptr && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *total_out = off;
}

size_t scan_ws_bytes(int64_t n) {
  size_t bytes = 0;
  int64_t tiles = ceil_div(n > 0 ? n : 1, kScanTile);
  while (tiles > 1) {
    bytes += align_up(size_t(tiles) * sizeof(int32_t));
    tiles = ceil_div(tiles, kScanTile);
  }
  return bytes + 256;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* total_out, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  if (n <= 0) {
    if (total_out) {
      if (cudaMemsetAsync(total_out, 0, sizeof(int32_t), st) != cudaSuccess) return GN_ERR_CUDA;
    }
    return GN_OK;
  }
  const int64_t tiles = ceil_div(n, kScanTile);
  if (tiles == 1) {
    GN_LAUNCH(scan_tiles, 1, kScanThreads, 0, st, in, out, n, (const int32_t*)nullptr, total_out);
    return GN_OK;
  }
  Arena a(ws, ws_bytes);
  int32_t* sums = a.take<int32_t>(size_t(tiles));
  if (!a.ok()) return GN_ERR_WORKSPACE;
  GN_LAUNCH(scan_tile_sums, (unsigned)tiles, kScanThreads, 0, st, in, n, sums);
  GN_CHECK(exclusive_scan_i32(sums, sums, tiles, nullptr, a.base + a.off, a.cap - a.off, st));
  GN_LAUNCH(scan_tiles, (unsigned)tiles, kScanThreads, 0, st, in, out, n, (const int32_t*)sums, total_out);
  return GN_OK;
}

// ----------------------------------------------------------------------------
// stable LSD radix sort (kernels in sortkit.cuh), 8 or 10 bits per pass
// ----------------------------------------------------------------------------
size_t sort_ws_bytes(int64_t n) {
  if (n <= 0) return 256;
  const int64_t hist = rs_tiles(n) << 10;
  return align_up(size_t(hist) * 4) + scan_ws_bytes(hist) + 2 * align_up(size_t(n) * 4) + 1024;
}

template <int RB>
static int sort_pass(const int32_t* src_k, const int32_t* src_v, int32_t* dst_k, int32_t* dst_v, int64_t n, int shift,
                     int32_t* hist, int64_t tiles, void* scan_ws, size_t scan_bytes, cudaStream_t st) {
  const PlainKeys ks{src_k};
  const PairSink sink{dst_k, dst_v, src_v};
  GN_LAUNCH((rs_histogram<RB, PlainKeys>), (unsigned)tiles, kRsThreads, 0, st, ks, n, shift, hist, (int)tiles);
  GN_CHECK(exclusive_scan_i32(hist, hist, tiles << RB, nullptr, scan_ws, scan_bytes, st));
  GN_LAUNCH((rs_scatter<RB, PlainKeys, PairSink>), (unsigned)tiles, kRsThreads, 0, st, ks, sink, n, shift,
            (const int32_t*)hist, (int)tiles, (int32_t*)nullptr, 0);
  return GN_OK;
}

int sort_pairs(const int32_t* keys_in, const int32_t* vals_in, int32_t* keys_out, int32_t* vals_out, int64_t n,
               int key_bits, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n <= 0) return GN_OK;
  if (n >= (int64_t(1) << 31)) return GN_ERR_RANGE;
  int passes, rb;
  rs_plan(key_bits, &passes, &rb);
  const int64_t tiles = rs_tiles(n);
  Arena a(ws, ws_bytes);
  int32_t* hist = a.take<int32_t>(size_t(tiles) << rb);
  int32_t* tk = passes > 1 ? a.take<int32_t>(size_t(n)) : nullptr;
  int32_t* tv = passes > 1 ? a.take<int32_t>(size_t(n)) : nullptr;
  if (!a.ok()) return GN_ERR_WORKSPACE;
  void* scan_ws = a.base + a.off;
  const size_t scan_bytes = a.cap - a.off;
  // ping-pong so that the LAST pass lands in (keys_out, vals_out)
  const int32_t* src_k = keys_in;
  const int32_t* src_v = vals_in;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    int32_t* dst_k = to_out ? keys_out : tk;
    int32_t* dst_v = to_out ? vals_out : tv;
    if (rb == 8) GN_CHECK(sort_pass<8>(src_k, src_v, dst_k, dst_v, n, p * 8, hist, tiles, scan_ws, scan_bytes, st));
    else GN_CHECK(sort_pass<10>(src_k, src_v, dst_k, dst_v, n, p * 10, hist, tiles, scan_ws, scan_bytes, st));
    src_k = dst_k;
    src_v = dst_v;
  }
  return GN_OK;
}

// ----------------------------------------------------------------------------
// rowptr from sorted keys: rowptr[r] = first position whose key >= r
// ----------------------------------------------------------------------------
__global__ void rowptr_kernel(const int32_t* __restrict__ keys, int64_t n, int32_t n_rows, int32_t* __restrict__ rowptr) {
  const int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n == 0) {
    if (k <= n_rows) rowptr[k] = 0;
    return;
  }
  if (k >= n) return;
  const int cur = keys[k];
  const int prev = (k == 0) ? -1 : keys[k - 1];
  for (int r = prev + 1; r <= cur; ++r) rowptr[r] = int32_t(k);
  if (k == n - 1)
    for (int r = cur + 1; r <= n_rows; ++r) rowptr[r] = int32_t(n);
}

int rowptr_from_sorted(const int32_t* sorted_keys, int64_t n, int32_t n_rows, int32_t* rowptr, cudaStream_t st) {
  const int64_t work = n > 0 ? n : int64_t(n_rows) + 1;
  GN_LAUNCH(rowptr_kernel, (unsigned)ceil_div(work, 256), 256, 0, st, sorted_keys, n, n_rows, rowptr);
  return GN_OK;
}

// rows [r0, r0+n) of a CSR: rebased row pointers (destination-partitioned graphs)
__global__ void rowptr_slice_kernel(const int32_t* __restrict__ rowptr, int32_t r0, int32_t n_rows,
                                    int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n_rows) out[i] = rowptr[r0 + i] - rowptr[r0];
}

// ----------------------------------------------------------------------------
// chunk lists
// ----------------------------------------------------------------------------
__global__ void chunk_count_kernel(const int32_t* __restrict__ rowptr, int32_t n_rows, int32_t chunk_len,
                                   int32_t* __restrict__ cnt, int32_t* __restrict__ row_counter) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int len = rowptr[r + 1] - rowptr[r];
  const int c = (len + chunk_len - 1) / chunk_len;
  cnt[r] = c > 0 ? c : 1;
  if (row_counter) row_counter[r] = 0;
}

__global__ void chunk_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ chunk_ptr,
                                  int32_t n_rows, int32_t chunk_len, int32_t* __restrict__ chunk_row,
                                  int32_t* __restrict__ chunk_beg, int64_t capacity) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int c0 = chunk_ptr[r], c1 = chunk_ptr[r + 1];
  int beg = rowptr[r];
  for (int c = c0; c < c1; ++c) {
    if (c < capacity) {
      chunk_row[c] = r;
      chunk_beg[c] = beg;
    }
    beg += chunk_len;
  }
}

// small CSRs (decoder structures rebuilt every epoch): count + scan + fill in ONE block
constexpr int kChunkOneBlockRows = 32768;
__global__ void __launch_bounds__(1024) chunk_one_block_kernel(const int32_t* __restrict__ rowptr, int32_t n_rows,
                                                               int32_t chunk_len, int32_t* __restrict__ chunk_ptr,
                                                               int32_t* __restrict__ chunk_row,
                                                               int32_t* __restrict__ chunk_beg, int64_t capacity,
                                                               int32_t* __restrict__ row_counter) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  constexpr int kItems = 4;                         // consecutive rows per thread: a third of the block-wide scans
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_rows; base += 1024 * kItems) {
    const int r0 = base + int(threadIdx.x) * kItems;
    int ptr[kItems + 1];
#pragma unroll
    for (int i = 0; i <= kItems; ++i) ptr[i] = (r0 + i <= n_rows) ? rowptr[r0 + i] : 0;
    int c[kItems], tsum = 0;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      c[i] = 0;
      if (r0 + i < n_rows) {
        const int len = ptr[i + 1] - ptr[i];
        c[i] = (len + chunk_len - 1) / chunk_len;
        if (c[i] < 1) c[i] = 1;
        if (row_counter) row_counter[r0 + i] = 0;
      }
      tsum += c[i];
    }
    int incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = warp_sums[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, wi, o);
        if (lane >= o) wi += t;
      }
      warp_sums[lane] = wi - w;
    }
    __syncthreads();
    int c0 = carry + warp_sums[warp] + incl - tsum;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
      if (r0 + i < n_rows) {
        chunk_ptr[r0 + i] = c0;
        for (int k = 0; k < c[i]; ++k) {
          if (c0 + k < capacity) {
            chunk_row[c0 + k] = r0 + i;
            chunk_beg[c0 + k] = ptr[i] + k * chunk_len;
          }
        }
        c0 += c[i];
      }
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = c0;            // rows beyond n_rows add nothing: the last thread ends the tile
    __syncthreads();
  }
  if (threadIdx.x == 0) chunk_ptr[n_rows] = carry;
}

}  // namespace gn

using namespace gn;

extern "C" {

int gn_version(void) { return 100; }

uint64_t gn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int gn_last_cuda_error(void) { return g_last_cuda_error.load(std::memory_order_relaxed); }

const char* gn_last_cuda_error_string(void) {
  return cudaGetErrorString(cudaError_t(g_last_cuda_error.load(std::memory_order_relaxed)));
}

const char* gn_error_string(int status) {
  switch (status) {
    case GN_OK: return "ok";
    case GN_ERR_ARG: return "invalid argument (size, null pointer or alignment)";
    case GN_ERR_RANGE: return "size exceeds the int32 index space";
    case GN_ERR_WORKSPACE: return "workspace too small";
    case GN_ERR_CUDA: return "CUDA runtime / launch failure (is a B200 visible?)";
    default: return "unknown status";
  }
}

size_t gn_csr_from_keys_workspace_bytes(int64_t n, int32_t n_rows) {
  (void)n_rows;
  const size_t multi = sort_ws_bytes(n) + align_up(size_t(n > 0 ? n : 1) * 4) + 256;
  const size_t single = rs_single_pass_ws_bytes(n, 10);
  return multi > single ? multi : single;
}

int gn_csr_from_keys(const int32_t* keys, int64_t n, int32_t n_rows, int32_t* rowptr, int32_t* perm, void* ws,
                     size_t ws_bytes, void* stream) {
  if (n < 0 || n_rows < 0 || rowptr == nullptr || (n > 0 && (keys == nullptr || perm == nullptr))) return GN_ERR_ARG;
  if (n >= (int64_t(1) << 31)) return GN_ERR_RANGE;
  cudaStream_t st = as_stream(stream);
  const int bits = bits_for(n_rows > 1 ? n_rows : 2);
  if (bits <= 10) {  // one pass; row pointers fall out of the scanned histogram
    const PlainKeys ks{keys};
    const PermSink sink{perm};
    if (bits <= 8) return rs_single_pass<8>(ks, sink, n, rowptr, n_rows, ws, ws_bytes, st);
    return rs_single_pass<10>(ks, sink, n, rowptr, n_rows, ws, ws_bytes, st);
  }
  Arena a(ws, ws_bytes);
  int32_t* sorted = a.take<int32_t>(size_t(n > 0 ? n : 1));
  if (!a.ok()) return GN_ERR_WORKSPACE;
  GN_CHECK(sort_pairs(keys, nullptr, sorted, perm, n, bits, a.base + a.off, a.cap - a.off, st));
  return rowptr_from_sorted(sorted, n, n_rows, rowptr, st);
}

int gn_rowptr_slice(const int32_t* rowptr, int32_t r0, int32_t n_rows, int32_t* out, void* stream) {
  if (!rowptr || !out || r0 < 0 || n_rows < 0) return GN_ERR_ARG;
  GN_LAUNCH(rowptr_slice_kernel, (unsigned)ceil_div(int64_t(n_rows) + 1, 256), 256, 0, as_stream(stream), rowptr, r0,
            n_rows, out);
  return GN_OK;
}

size_t gn_build_chunks_workspace_bytes(int32_t n_rows) { return scan_ws_bytes(int64_t(n_rows) + 1) + 256; }

int gn_build_chunks(const int32_t* rowptr, int32_t n_rows, int32_t chunk_len, int32_t* chunk_ptr, int32_t* chunk_row,
                    int32_t* chunk_beg, int64_t chunk_capacity, int32_t* row_counter, void* ws, size_t ws_bytes,
                    void* stream) {
  if (n_rows < 0 || chunk_len <= 0 || rowptr == nullptr || chunk_ptr == nullptr) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  if (n_rows == 0) {
    if (cudaMemsetAsync(chunk_ptr, 0, sizeof(int32_t), st) != cudaSuccess) return GN_ERR_CUDA;
    return GN_OK;
  }
  if (n_rows <= kChunkOneBlockRows) {
    GN_LAUNCH(chunk_one_block_kernel, 1, 1024, 0, st, rowptr, n_rows, chunk_len, chunk_ptr, chunk_row, chunk_beg,
              chunk_capacity, row_counter);
    return GN_OK;
  }
  const unsigned grid = (unsigned)ceil_div(n_rows, 256);
  GN_LAUNCH(chunk_count_kernel, grid, 256, 0, st, rowptr, n_rows, chunk_len, chunk_ptr, row_counter);
  GN_CHECK(exclusive_scan_i32(chunk_ptr, chunk_ptr, n_rows, chunk_ptr + n_rows, ws, ws_bytes, st));
  GN_LAUNCH(chunk_fill_kernel, grid, 256, 0, st, rowptr, (const int32_t*)chunk_ptr, n_rows, chunk_len, chunk_row,
            chunk_beg, chunk_capacity);
  return GN_OK;
}

}  // extern "C"
