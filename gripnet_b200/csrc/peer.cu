// K12: slot all-gather over NVLink peer memory (the exchange step before a SpMM).
//
// Every rank holds a [world*B, F] gather buffer at the SAME offset of a symmetric arena that all
// ranks of the node have mapped (torch symmetric memory: CUDA VMM + peer access through NVSwitch).
// The producing kernel wrote this rank's slot in place; this kernel
//   1. PUSHES the slot into every peer's buffer with 128-bit stores over NVLink (one local read,
//      world-1 remote writes per vector, peers visited in a rank-staggered order),
//   2. the last CTA to finish publishes "rank r has delivered buffer b, use e" in every peer's
//      flag block (fence.sys + st.release.sys), and
//   3. waits (ld.acquire.sys on its OWN flag block) until every peer has delivered, so the SpMM
//      that follows on the stream sees the complete operand.
// No host synchronisation, no NCCL, CUDA-graph capturable: the use counter `e` lives in device
// memory, so a replayed graph keeps counting.  Deadlock-free: every rank publishes before it
// waits, and only one thread per rank spins.
//
// Buffer reuse: a buffer (arena offset) is used once per step.  A peer can only be writing use
// e+1 of buffer b into this rank while this rank still reads use e if it ran a whole step ahead,
// which the other gathers of the step make impossible when a step holds >= 2 gathers (the host
// side falls back to NCCL otherwise).
#include "common.cuh"

namespace gn {

constexpr int kMaxPeers = 8;

struct PeerArgs {
  char* buf[kMaxPeers];                  // every rank's gather buffer (peer-mapped device pointers)
  unsigned long long* flags[kMaxPeers];  // every rank's flag block: [n_buffers][kMaxPeers]
  int world, rank;
  int64_t slot_bytes;
  int flag_index;
  unsigned long long* seq;               // local: use counter of this buffer
  unsigned int* done;                    // local: CTA arrival counter of this buffer (left zero)
  unsigned int* abort_flag;              // local: set when a wait timed out; later gathers do not wait
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) peer_allgather_kernel(const PeerArgs a) {
  const int64_t n16 = a.slot_bytes >> 4;
  const int64_t slot_off = int64_t(a.rank) * a.slot_bytes;
  const int4* __restrict__ src = reinterpret_cast<const int4*>(a.buf[a.rank] + slot_off);
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const int4 v = src[i];
#pragma unroll 1
    for (int s = 1; s < a.world; ++s) {
      const int peer = (a.rank + s) % a.world;
      reinterpret_cast<int4*>(a.buf[peer] + slot_off)[i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x != 0) return;
  const unsigned int prev = atomicAdd(a.done, 1u);
  if (prev != gridDim.x - 1) return;
  // last CTA: every CTA's pushes are ordered before this point (fence + atomic on each side)
  __threadfence_system();
  *a.done = 0;
  const unsigned long long e = *a.seq + 1;
  *a.seq = e;
  const int slot = a.flag_index * kMaxPeers;
  for (int s = 1; s < a.world; ++s) {
    const int peer = (a.rank + s) % a.world;
    st_release_sys(a.flags[peer] + slot + a.rank, e);
  }
  for (int s = 1; s < a.world; ++s) {
    const int peer = (a.rank + s) % a.world;
    const unsigned long long* f = a.flags[a.rank] + slot + peer;
    // bounded spin (~10 s): a rank that died must not hang this GPU.  A timeout raises the abort
    // flag: results are then visibly wrong (the host checks the flag) instead of the GPU being stuck.
    unsigned int spins = 0;
    while (ld_acquire_sys(f) < e) {
      if (*reinterpret_cast<volatile unsigned int*>(a.abort_flag) != 0) return;
      __nanosleep(256);
      if (++spins > (1u << 23)) {
        *reinterpret_cast<volatile unsigned int*>(a.abort_flag) = 1;
        return;
      }
    }
  }
}

}  // namespace gn

using namespace gn;

extern "C" int gn_peer_max_world(void) { return kMaxPeers; }

extern "C" int gn_peer_allgather(const uint64_t* arena_base /*host*/, int32_t world, int32_t rank, int64_t buf_offset,
                                 int64_t slot_bytes, int64_t flag_offset, int32_t flag_index, uint64_t* seq,
                                 uint32_t* done, uint32_t* abort_flag, void* stream) {
  if (!arena_base || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || slot_bytes < 0 || flag_index < 0 ||
      !seq || !done || !abort_flag)
    return GN_ERR_ARG;
  if (slot_bytes % 16 != 0 || buf_offset % 16 != 0 || flag_offset % 8 != 0) return GN_ERR_ARG;
  if (world == 1) return GN_OK;
  PeerArgs a;
  for (int p = 0; p < kMaxPeers; ++p) {
    const uint64_t base = p < world ? arena_base[p] : 0;
    if (p < world && base == 0) return GN_ERR_ARG;
    a.buf[p] = reinterpret_cast<char*>(base + (p < world ? uint64_t(buf_offset) : 0));
    a.flags[p] = reinterpret_cast<unsigned long long*>(base + (p < world ? uint64_t(flag_offset) : 0));
  }
  a.world = world; a.rank = rank; a.slot_bytes = slot_bytes; a.flag_index = flag_index;
  a.seq = reinterpret_cast<unsigned long long*>(seq);
  a.done = done;
  a.abort_flag = abort_flag;
  // enough CTAs to keep the NVLink ports busy, few enough that the arrival counter is cheap
  int64_t ctas = ceil_div((slot_bytes >> 4) > 0 ? (slot_bytes >> 4) : 1, 256 * 4);
  if (ctas > 148 * 2) ctas = 148 * 2;
  if (ctas < 1) ctas = 1;
  GN_LAUNCH(peer_allgather_kernel, (unsigned)ctas, 256, 0, as_stream(stream), a);
  return GN_OK;
}
