// K12: slot all-gather over NVLink peer memory (the exchange step before a SpMM).
//
// Every rank holds a [world*B, F] gather buffer at the SAME offset of a symmetric arena that all
// ranks of the node have mapped (torch symmetric memory: CUDA VMM + peer access through NVSwitch).
// The producing kernel wrote this rank's slot in place; this kernel
//   1. PUSHES the slot into every peer's buffer with 128-bit stores over NVLink (one local read,
//      world-1 remote writes per vector, peers visited in a rank-staggered order),
//   2. every CTA, once its stores are issued, adds 1 to "CTAs of rank r that delivered buffer b" in every
//      peer's flag block (fence.sys + remote red.release.sys), and
//   3. CTA 0 waits (ld.acquire.sys on its OWN flag block) until every peer's count is complete, so the
//      SpMM that follows on the stream sees the complete operand.
// No host synchronisation, no NCCL, CUDA-graph capturable: the use counter `e` lives in device
// memory, so a replayed graph keeps counting.  Deadlock-free: signalling never waits, and only
// one CTA per rank spins.
//
// Buffer reuse: a buffer (arena offset) is used once per step.  A peer can only be writing use
// e+1 of buffer b into this rank while this rank still reads use e if it ran a whole step ahead,
// which the other gathers of the step make impossible when a step holds >= 2 gathers (the host
// side falls back to NCCL otherwise).
#include "common.cuh"

namespace gn {

constexpr int kMaxPeers = 8;
constexpr int64_t kMulticastMaxSlotBytes = int64_t(4) << 20;

struct PeerArgs {
  char* buf[kMaxPeers];                  // every rank's gather buffer (peer-mapped device pointers)
  unsigned long long* flags[kMaxPeers];  // every rank's flag block: [n_buffers][kMaxPeers]
  char* mc;                              // multicast (NVLS) mapping of the same buffer on ALL ranks, or NULL
  unsigned long long* mc_flags;          // multicast mapping of the flag block, or NULL
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

// one store, delivered by the NVSwitch to the same offset of every rank's buffer (this rank's included)
__device__ __forceinline__ void multimem_st16(void* p, const int4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__int_as_float(v.x)),
               "f"(__int_as_float(v.y)), "f"(__int_as_float(v.z)), "f"(__int_as_float(v.w))
               : "memory");
}

__device__ __forceinline__ void red_relaxed_sys_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// one reduction, applied by the NVSwitch to the same word of EVERY rank's flag block
__device__ __forceinline__ void multimem_red_add(unsigned long long* p, unsigned long long v) {
  asm volatile("multimem.red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Publish / wait round of one exchange.  Flag (buffer b, rank r) in a rank's flag block COUNTS the CTAs of rank
// r that have delivered their part of buffer b, over all uses of the buffer (u64, never reset).
//   * every CTA: once its own stores are issued (barrier), ONE thread orders them system-wide (fence.sys, cumulative
//     over the CTA's stores through the barrier) and adds 1 to its flag in every peer's block — one
//     multimem.red through the NVSwitch when the arena has a multicast mapping, else world-1 relaxed remote reds —
//     no intra-GPU arrival counter, no serial "last CTA" tail: a CTA's signal leaves as soon as that CTA is done;
//   * CTA 0 alone waits: thread p spins (ld.acquire.sys on this rank's OWN memory) until peer p's count reaches
//     the running total of expected CTAs (`seq` + this launch's gridDim.x) — every rank launches the same grid
//     for the same use of a buffer, because buffer shapes are symmetric.  The kernel therefore ends only when every peer's data has arrived, and the kernels behind it
//     on the stream see the complete buffer.
// Deadlock-free: signalling never waits, and only one CTA per rank spins (bounded, see below).
__device__ __forceinline__ void signal_and_wait(const PeerArgs& a) {
  __shared__ unsigned long long s_use;
  __syncthreads();
  const int slot = a.flag_index * kMaxPeers;
  if (threadIdx.x == 0) {
    __threadfence_system();              // ONE system-scope fence orders the CTA's stores; the signals are relaxed
    if (a.mc_flags != nullptr) {
      multimem_red_add(a.mc_flags + slot + a.rank, 1ull);     // one instruction signals every rank
    } else {
      for (int s = 1; s < a.world; ++s) {
        const int peer = (a.rank + s) % a.world;
        red_relaxed_sys_add(a.flags[peer] + slot + a.rank, 1ull);
      }
    }
  }
  if (blockIdx.x != 0) return;
  // `seq` = CTAs every peer has been expected to deliver to this buffer so far (uses may differ in grid size)
  if (threadIdx.x == 0) s_use = *a.seq + gridDim.x;
  __syncthreads();
  const unsigned long long target = s_use;
  if (threadIdx.x >= 1 && threadIdx.x < a.world) {
    const int peer = (a.rank + int(threadIdx.x)) % a.world;
    const unsigned long long* f = a.flags[a.rank] + slot + peer;
    // bounded spin (~10 s): a rank that died must not hang this GPU.  A timeout raises the abort
    // flag: results are then visibly wrong (the host checks the flag) instead of the GPU being stuck.
    unsigned int spins = 0;
    while (ld_acquire_sys(f) < target) {
      if (*reinterpret_cast<volatile unsigned int*>(a.abort_flag) != 0) break;
      __nanosleep(64);
      if (++spins > (1u << 25)) {
        *reinterpret_cast<volatile unsigned int*>(a.abort_flag) = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *a.seq = s_use;
}

__global__ void __launch_bounds__(256) peer_allgather_kernel(const PeerArgs a) {
  const int64_t n16 = a.slot_bytes >> 4;
  const int64_t slot_off = int64_t(a.rank) * a.slot_bytes;
  const int4* __restrict__ src = reinterpret_cast<const int4*>(a.buf[a.rank] + slot_off);
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  // NVLS multicast for SMALL slots only (latency-bound: one store instruction instead of world-1).  The switch also
  // delivers a multicast store back to the sender, so a rank RECEIVES world slots instead of world-1: for large
  // slots the exchange is ingress-bound and unicast stores win — at world 2 by 2x (profiles/r02_v22_scaled_n2_*:
  // 512 MB slots moved at 353 GB/s of egress = 706 GB/s of ingress through the multicast mapping)
  const bool mc_data = a.mc != nullptr && (a.world > 2 && a.slot_bytes <= kMulticastMaxSlotBytes);
  if (mc_data) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride)
      multimem_st16(a.mc + slot_off + i * 16, src[i]);   // NVLS: one store leaves the GPU, the switch replicates it
  } else {
    // unicast: four independent 128-bit loads per thread in flight, then the remote stores (peers visited in a
    // rank-staggered order so the ranks do not all write into the same GPU at once)
    constexpr int kU = 4;
    for (int64_t i0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i0 < n16; i0 += stride * kU) {
      int4 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < n16) v[u] = src[i];
      }
#pragma unroll 1
      for (int s = 1; s < a.world; ++s) {
        int4* dst = reinterpret_cast<int4*>(a.buf[(a.rank + s) % a.world] + slot_off);
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int64_t i = i0 + u * stride;
          if (i < n16) dst[i] = v[u];
        }
      }
    }
  }
  signal_and_wait(a);
}

// ---------------------------------------------------------------------------------------------
// Segmented push + rank-ordered sum: the small reductions of a partitioned step without NCCL.
//   gn_peer_push : every segment (src pointer, bytes, offset inside a slot, target peer or ALL) is stored
//                  into slot `rank` of the exchange buffer of its target rank(s) — own arena included —
//                  then the same publish / wait protocol as the slot all-gather.  With one segment per
//                  peer this is an all-to-all (the reduce-scatter of the decoder's dz); with every
//                  segment sent to ALL it is the all-gather phase of an all-reduce (bucketed weight
//                  gradients + the loss share).
//   gn_slot_sum  : dst[i] = slot_0[i] + slot_1[i] + ... + slot_{world-1}[i], in RANK ORDER: every rank
//                  adds the same numbers in the same order, so replicated results are bit-identical on all
//                  ranks and run to run (NCCL's ring / tree order is neither specified nor rank-symmetric).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxSegments = 40;
constexpr int kPushBlockBytes = 256 * 16 * 4;   // 16 KB of a segment per CTA

struct PushBatch {
  const char* src[kMaxSegments];
  int64_t bytes[kMaxSegments];
  int64_t slot_off[kMaxSegments];
  int32_t peer[kMaxSegments];                 // -1: every rank
  int32_t block_first[kMaxSegments + 1];
  int32_t n_segs;
};

__global__ void __launch_bounds__(256) peer_push_kernel(const PeerArgs a, const PushBatch b) {
  __shared__ int s_seg;
  if (threadIdx.x == 0) {
    int s = 0;
    while (s + 1 < b.n_segs && int(blockIdx.x) >= b.block_first[s + 1]) ++s;
    s_seg = s;
  }
  __syncthreads();
  const int s = s_seg;
  const int64_t base = int64_t(int(blockIdx.x) - b.block_first[s]) * kPushBlockBytes;
  const int64_t bytes = b.bytes[s];
  const char* __restrict__ src = b.src[s];
  const int64_t dst_off = int64_t(a.rank) * a.slot_bytes + b.slot_off[s];
  const int p_first = b.peer[s] < 0 ? 0 : b.peer[s];
  const int p_last = b.peer[s] < 0 ? a.world - 1 : b.peer[s];
  const bool v16 = ((reinterpret_cast<uintptr_t>(src) | uintptr_t(bytes) | uintptr_t(dst_off)) & 15) == 0;
  if (v16) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int64_t i = base + (int64_t(it) * 256 + threadIdx.x) * 16;
      if (i >= bytes) break;
      const int4 v = *reinterpret_cast<const int4*>(src + i);
      if (b.peer[s] < 0 && a.mc != nullptr) {
        multimem_st16(a.mc + dst_off + i, v);
        continue;
      }
      for (int q = p_first; q <= p_last; ++q) {
        const int peer = (b.peer[s] < 0) ? (a.rank + q) % a.world : q;     // staggered start: not everyone hits rank 0 first
        *reinterpret_cast<int4*>(a.buf[peer] + dst_off + i) = v;
      }
    }
  } else {   // 4-byte granularity (the scalar loss share, odd-sized tensors)
    const int64_t end = base + kPushBlockBytes < bytes ? base + kPushBlockBytes : bytes;
    for (int64_t i = base + int64_t(threadIdx.x) * 4; i < end; i += 256 * 4) {
      const int v = *reinterpret_cast<const int*>(src + i);
      for (int q = p_first; q <= p_last; ++q) {
        const int peer = (b.peer[s] < 0) ? (a.rank + q) % a.world : q;
        *reinterpret_cast<int*>(a.buf[peer] + dst_off + i) = v;
      }
    }
  }
  signal_and_wait(a);
}

// ---------------------------------------------------------------------------------------------
// Halo-packed exchange (north_star: "all-gather of HALO source-feature rows").  Instead of its whole block a rank
// sends to peer p only the rows p's CSR actually references (index lists exchanged once at graph-build time),
// packed contiguously into p's operand buffer behind p's own rows; p's column indices were remapped to the packed
// positions.  On power-law graphs a rank references 20-50 % of a remote block, so the bytes ENTERING a GPU drop
// 2-5x against the full slot all-gather (which is ingress-bound: (P-1) blocks in, one block out through NVLS).
// Fixed grid (the same on every rank), grid-stride over (peer, row) items, then the usual publish / wait round.
// ---------------------------------------------------------------------------------------------
struct HaloArgs {
  const int32_t* idx[kMaxPeers];     // local row ids to send to peer p (device), NULL / count 0 for p == rank
  int64_t count[kMaxPeers];
  int64_t dst_row[kMaxPeers];        // first packed row of this rank's region in peer p's buffer
  const float* src;                  // this rank's rows (the head of its own buffer)
  int row_f4;                        // float4 per row
};

__global__ void __launch_bounds__(256) peer_halo_push_kernel(const PeerArgs a, const HaloArgs h) {
  const int lanes = h.row_f4 < 32 ? h.row_f4 : 32;            // threads that cooperate on one row
  const int rows_per_cta = 256 / lanes;
  const int sub = int(threadIdx.x) / lanes, l = int(threadIdx.x) % lanes;
  for (int s = 1; s < a.world; ++s) {
    const int peer = (a.rank + s) % a.world;
    const int32_t* __restrict__ idx = h.idx[peer];
    const int64_t cnt = h.count[peer];
    float4* __restrict__ dst = reinterpret_cast<float4*>(a.buf[peer]) + h.dst_row[peer] * h.row_f4;
    // kHU rows per thread group in flight: index loads, then the row loads, then the remote stores (one dependent
    // index -> row -> store chain per thread keeps only ~1 MB in flight per GPU, short of the NVLink latency-bandwidth
    // product)
    constexpr int kHU = 4;
    const int64_t step = int64_t(gridDim.x) * rows_per_cta;
    if (sub >= rows_per_cta) continue;
    for (int64_t j0 = int64_t(blockIdx.x) * rows_per_cta + sub; j0 < cnt; j0 += step * kHU) {
      int32_t r[kHU];
#pragma unroll
      for (int u = 0; u < kHU; ++u) {
        const int64_t j = j0 + u * step;
        r[u] = j < cnt ? __ldg(idx + j) : 0;
      }
      if (h.row_f4 <= 32) {                      // one float4 per lane and row (rows of up to 128 floats)
        float4 v[kHU];
#pragma unroll
        for (int u = 0; u < kHU; ++u)
          if (l < h.row_f4) v[u] = __ldg(reinterpret_cast<const float4*>(h.src) + int64_t(r[u]) * h.row_f4 + l);
#pragma unroll
        for (int u = 0; u < kHU; ++u) {
          const int64_t j = j0 + u * step;
          if (j < cnt && l < h.row_f4) dst[j * h.row_f4 + l] = v[u];
        }
      } else {
#pragma unroll 1
        for (int u = 0; u < kHU; ++u) {
          const int64_t j = j0 + u * step;
          if (j >= cnt) break;
          const float4* __restrict__ srow = reinterpret_cast<const float4*>(h.src) + int64_t(r[u]) * h.row_f4;
          for (int c = l; c < h.row_f4; c += lanes) dst[j * h.row_f4 + c] = srow[c];
        }
      }
    }
  }
  signal_and_wait(a);
}

struct SumBatch {
  float* dst[kMaxSegments];
  int64_t n[kMaxSegments];
  int64_t slot_off[kMaxSegments];             // bytes
  int32_t block_first[kMaxSegments + 1];
  int32_t n_segs;
};
constexpr int kSumBlockElems = 256 * 4 * 4;   // 4096 floats per CTA

__global__ void __launch_bounds__(256) slot_sum_kernel(const char* __restrict__ buf, int world, int64_t slot_bytes,
                                                       const SumBatch b) {
  __shared__ int s_seg;
  if (threadIdx.x == 0) {
    int s = 0;
    while (s + 1 < b.n_segs && int(blockIdx.x) >= b.block_first[s + 1]) ++s;
    s_seg = s;
  }
  __syncthreads();
  const int s = s_seg;
  const int64_t n = b.n[s];
  float* __restrict__ dst = b.dst[s];
  const char* __restrict__ src0 = buf + b.slot_off[s];
  const int64_t base = int64_t(int(blockIdx.x) - b.block_first[s]) * kSumBlockElems;
  const bool v16 = ((reinterpret_cast<uintptr_t>(dst) | uintptr_t(b.slot_off[s]) | uintptr_t(slot_bytes)) & 15) == 0;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int64_t i = base + (int64_t(it) * 256 + threadIdx.x) * 4;
    if (i >= n) break;
    if (v16 && i + 4 <= n) {
      float4 acc = *reinterpret_cast<const float4*>(src0 + i * 4);
      for (int p = 1; p < world; ++p) {
        const float4 v = *reinterpret_cast<const float4*>(src0 + int64_t(p) * slot_bytes + i * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(dst + i) = acc;
    } else {
      for (int64_t j = i; j < n && j < i + 4; ++j) {
        float acc = *reinterpret_cast<const float*>(src0 + j * 4);
        for (int p = 1; p < world; ++p) acc += *reinterpret_cast<const float*>(src0 + int64_t(p) * slot_bytes + j * 4);
        dst[j] = acc;
      }
    }
  }
}

}  // namespace gn

using namespace gn;

extern "C" int gn_peer_max_world(void) { return kMaxPeers; }

static int fill_peer_args(PeerArgs& a, const uint64_t* arena_base, int32_t world, int32_t rank, int64_t buf_offset,
                          int64_t slot_bytes, int64_t flag_offset, int32_t flag_index, uint64_t* seq, uint32_t* done,
                          uint32_t* abort_flag);

extern "C" int gn_peer_allgather(const uint64_t* arena_base /*host*/, int32_t world, int32_t rank, int64_t buf_offset,
                                 int64_t slot_bytes, int64_t flag_offset, int32_t flag_index, uint64_t* seq,
                                 uint32_t* done, uint32_t* abort_flag, void* stream) {
  if (world == 1) return GN_OK;
  PeerArgs a;
  GN_CHECK(fill_peer_args(a, arena_base, world, rank, buf_offset, slot_bytes, flag_offset, flag_index, seq, done,
                          abort_flag));
  // enough CTAs to keep the NVLink ports busy, few enough that the arrival counter is cheap
  int64_t ctas = ceil_div((slot_bytes >> 4) > 0 ? (slot_bytes >> 4) : 1, 256 * 4);
  if (ctas > 148 * 2) ctas = 148 * 2;
  if (ctas < 1) ctas = 1;
  GN_LAUNCH(peer_allgather_kernel, (unsigned)ctas, 256, 0, as_stream(stream), a);
  return GN_OK;
}

static int fill_peer_args(PeerArgs& a, const uint64_t* arena_base, int32_t world, int32_t rank, int64_t buf_offset,
                          int64_t slot_bytes, int64_t flag_offset, int32_t flag_index, uint64_t* seq, uint32_t* done,
                          uint32_t* abort_flag) {
  if (!arena_base || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || slot_bytes < 0 || flag_index < 0 ||
      !seq || !done || !abort_flag)
    return GN_ERR_ARG;
  if (slot_bytes % 16 != 0 || buf_offset % 16 != 0 || flag_offset % 8 != 0) return GN_ERR_ARG;
  for (int p = 0; p < kMaxPeers; ++p) {
    const uint64_t base = p < world ? arena_base[p] : 0;
    if (p < world && base == 0) return GN_ERR_ARG;
    a.buf[p] = reinterpret_cast<char*>(base + (p < world ? uint64_t(buf_offset) : 0));
    a.flags[p] = reinterpret_cast<unsigned long long*>(base + (p < world ? uint64_t(flag_offset) : 0));
  }
  a.mc = arena_base[world] ? reinterpret_cast<char*>(arena_base[world] + uint64_t(buf_offset)) : nullptr;
  a.mc_flags = arena_base[world] ? reinterpret_cast<unsigned long long*>(arena_base[world] + uint64_t(flag_offset))
                                 : nullptr;
  a.world = world; a.rank = rank; a.slot_bytes = slot_bytes; a.flag_index = flag_index;
  a.seq = reinterpret_cast<unsigned long long*>(seq);
  a.done = done;
  a.abort_flag = abort_flag;
  return GN_OK;
}

extern "C" int gn_peer_max_segments(void) { return kMaxSegments; }

extern "C" int gn_peer_push(const uint64_t* arena_base /*host*/, int32_t world, int32_t rank, int64_t buf_offset,
                            int64_t slot_bytes, const gn_peer_segment* segs /*host*/, int32_t n_segs,
                            int64_t flag_offset, int32_t flag_index, uint64_t* seq, uint32_t* done,
                            uint32_t* abort_flag, void* stream) {
  PeerArgs a;
  GN_CHECK(fill_peer_args(a, arena_base, world, rank, buf_offset, slot_bytes, flag_offset, flag_index, seq, done,
                          abort_flag));
  if (n_segs < 1 || n_segs > kMaxSegments || !segs) return GN_ERR_ARG;
  PushBatch b;
  int64_t blocks = 0;
  b.n_segs = 0;
  for (int i = 0; i < n_segs; ++i) {
    const gn_peer_segment& g = segs[i];
    if (g.bytes < 0 || g.bytes % 4 != 0 || g.slot_offset < 0 || g.slot_offset % 4 != 0 ||
        g.slot_offset + g.bytes > slot_bytes || g.peer >= world)
      return GN_ERR_ARG;
    if (g.bytes == 0) continue;
    if (!g.src) return GN_ERR_ARG;
    const int k = b.n_segs++;
    b.src[k] = static_cast<const char*>(g.src);
    b.bytes[k] = g.bytes;
    b.slot_off[k] = g.slot_offset;
    b.peer[k] = g.peer < 0 ? -1 : g.peer;
    b.block_first[k] = int32_t(blocks);
    blocks += ceil_div(g.bytes, kPushBlockBytes);
    if (blocks >= (int64_t(1) << 30)) return GN_ERR_RANGE;
  }
  if (b.n_segs == 0) {                      // nothing to send: still take part in the publish / wait round
    b.n_segs = 1;
    b.src[0] = nullptr; b.bytes[0] = 0; b.slot_off[0] = 0; b.peer[0] = rank; b.block_first[0] = 0;
    blocks = 1;
  }
  b.block_first[b.n_segs] = int32_t(blocks);
  GN_LAUNCH(peer_push_kernel, (unsigned)blocks, 256, 0, as_stream(stream), a, b);
  return GN_OK;
}

extern "C" int gn_slot_sum(const void* buf, int32_t world, int64_t slot_bytes, const gn_sum_segment* segs /*host*/,
                           int32_t n_segs, void* stream) {
  if (!buf || world < 1 || slot_bytes < 0 || slot_bytes % 4 != 0 || n_segs < 0 || n_segs > kMaxSegments ||
      (n_segs > 0 && !segs))
    return GN_ERR_ARG;
  SumBatch b;
  int64_t blocks = 0;
  b.n_segs = 0;
  for (int i = 0; i < n_segs; ++i) {
    const gn_sum_segment& g = segs[i];
    if (g.n < 0 || g.slot_offset < 0 || g.slot_offset % 4 != 0 || g.slot_offset + g.n * 4 > slot_bytes) return GN_ERR_ARG;
    if (g.n == 0) continue;
    if (!g.dst) return GN_ERR_ARG;
    const int k = b.n_segs++;
    b.dst[k] = g.dst;
    b.n[k] = g.n;
    b.slot_off[k] = g.slot_offset;
    b.block_first[k] = int32_t(blocks);
    blocks += ceil_div(g.n, kSumBlockElems);
    if (blocks >= (int64_t(1) << 30)) return GN_ERR_RANGE;
  }
  if (b.n_segs == 0) return GN_OK;
  b.block_first[b.n_segs] = int32_t(blocks);
  GN_LAUNCH(slot_sum_kernel, (unsigned)blocks, 256, 0, as_stream(stream), static_cast<const char*>(buf), world,
            slot_bytes, b);
  return GN_OK;
}

extern "C" int gn_peer_halo_grid(void) { return 148 * 2; }

extern "C" int gn_peer_halo_push(const uint64_t* arena_base /*host*/, int32_t world, int32_t rank, int64_t buf_offset,
                                 int64_t buf_bytes, const float* src_rows, int32_t row_floats,
                                 const gn_halo_peer* peers /*host, world entries*/, int64_t flag_offset,
                                 int32_t flag_index, uint64_t* seq, uint32_t* done, uint32_t* abort_flag, void* stream) {
  if (world == 1) return GN_OK;
  if (!peers || !src_rows || row_floats <= 0 || row_floats % 4 != 0 ||
      (reinterpret_cast<uintptr_t>(src_rows) & 15) != 0)
    return GN_ERR_ARG;
  PeerArgs a;
  GN_CHECK(fill_peer_args(a, arena_base, world, rank, buf_offset, (buf_bytes + 15) / 16 * 16, flag_offset, flag_index,
                          seq, done, abort_flag));
  a.mc = nullptr;                       // every peer receives a different subset: unicast stores
  HaloArgs h;
  for (int p = 0; p < kMaxPeers; ++p) {
    h.idx[p] = nullptr; h.count[p] = 0; h.dst_row[p] = 0;
    if (p < world && p != rank) {
      if (peers[p].count < 0 || peers[p].dst_row < 0 || (peers[p].count > 0 && !peers[p].idx)) return GN_ERR_ARG;
      if ((peers[p].dst_row + peers[p].count) * int64_t(row_floats) * 4 > buf_bytes) return GN_ERR_ARG;
      h.idx[p] = peers[p].idx; h.count[p] = peers[p].count; h.dst_row[p] = peers[p].dst_row;
    }
  }
  h.src = src_rows;
  h.row_f4 = row_floats / 4;
  GN_LAUNCH(peer_halo_push_kernel, (unsigned)gn_peer_halo_grid(), 256, 0, as_stream(stream), a, h);
  return GN_OK;
}
