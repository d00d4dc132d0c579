// Shared helpers for the gripnet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/gripnet_b200.h"

namespace gn {

extern std::atomic<uint64_t> g_launches;
// cudaError_t of the most recent failed launch / runtime call made by this library (0 = none); reported by
// gn_last_cuda_error() so the host side can show the real CUDA error string next to GN_ERR_CUDA
extern std::atomic<int> g_last_cuda_error;

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Every kernel launch goes through this so gn_launch_count() is exact.
#define GN_LAUNCH(kernel, grid, block, smem, stream, ...)                       \
  do {                                                                          \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                 \
    gn::g_launches.fetch_add(1, std::memory_order_relaxed);                     \
    const cudaError_t _gn_e = cudaGetLastError(); /* status of THIS launch */   \
    if (_gn_e != cudaSuccess) {                                                 \
      gn::g_last_cuda_error.store(int(_gn_e), std::memory_order_relaxed);       \
      return GN_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)

#define GN_CHECK(expr)                 \
  do {                                 \
    int _s = (expr);                   \
    if (_s != GN_OK) return _s;        \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  Arena(void* p, size_t bytes) : base(static_cast<char*>(p)), cap(bytes), off(0) {}
  template <typename T>
  T* take(size_t n) {
    size_t bytes = align_up(n * sizeof(T));
    if (base == nullptr || off + bytes > cap) {
      off = cap + 1;
      return nullptr;
    }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= cap; }
};

// ---- device-side primitives shared by several translation units ------------
// exclusive scan of n int32 values; out may alias in.  ws from scan_ws_bytes(n).
size_t scan_ws_bytes(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* total_out /*nullable*/,
                       void* ws, size_t ws_bytes, cudaStream_t st);

// stable LSD radix sort of (key, value) int32 pairs.  vals_in == nullptr means iota.
// Result ends in (keys_out, vals_out); keys_in/vals_in are preserved (tmp buffers from ws).
size_t sort_ws_bytes(int64_t n);
int sort_pairs(const int32_t* keys_in, const int32_t* vals_in, int32_t* keys_out, int32_t* vals_out,
               int64_t n, int key_bits, void* ws, size_t ws_bytes, cudaStream_t st);

// rowptr[0..n_rows] from SORTED keys (boundary search), no atomics.
int rowptr_from_sorted(const int32_t* sorted_keys, int64_t n, int32_t n_rows, int32_t* rowptr,
                       cudaStream_t st);

int bits_for(int64_t n_values);  // smallest b with 2^b >= n_values (>= 1)

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

}  // namespace gn
