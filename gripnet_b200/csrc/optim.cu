// K14: fused multi-tensor Adam — the optimiser step of the training scripts
// (torch.optim.Adam(model.parameters(), lr) at GripNet-pose.py:104 + optimizer.step() at :146,
// GripNet-aminer.py:113/:135).  One launch updates up to kAdamSlots parameter tensors: their
// (param, grad, exp_avg, exp_avg_sq, n) descriptors travel in the kernel-parameter block, so a
// captured CUDA graph keeps them by value and no descriptor table lives in device memory.  The step
// counter is DEVICE memory, read by the kernel and advanced after it: replays of one captured graph
// are successive optimiser steps with the right bias corrections.
//
// Arithmetic follows torch's single-tensor Adam (amsgrad = False, maximize = False):
//   g' = g + wd * p                     (weight_decay != 0 only)
//   m  = m + (g' - m) * (1 - beta1)     (lerp)
//   v  = v * beta2 + (1 - beta2) * g' * g'
//   p  = p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// with the hyper-parameters taken as doubles: the bias corrections and (1 - beta) are evaluated in double and
// only then cast to fp32 (as torch does on the host; 1 - float(0.999) would be off by 1.3e-5 relative), and
// everything elementwise in fp32.  HBM-bound: 28 B per element (p, m, v read+write, g read).
#include "common.cuh"

namespace gn {

constexpr int kAdamSlots = 24;
constexpr int kAdamThreads = 256;
constexpr int kAdamVecPerThread = 4;                                  // float4 per thread
constexpr int kAdamBlockElems = kAdamThreads * kAdamVecPerThread * 4;  // 4096 elements per block

struct AdamBatch {
  float* p[kAdamSlots];
  const float* g[kAdamSlots];
  float* m[kAdamSlots];
  float* v[kAdamSlots];
  int64_t n[kAdamSlots];
  int32_t block_first[kAdamSlots + 1];   // first block of each tensor
  int32_t n_tensors;
};

struct AdamHyper {
  double lr, beta1, beta2;                                   // bias corrections are evaluated in double
  float beta2f, one_minus_beta1, one_minus_beta2, eps, weight_decay;   // float(double expression), as torch casts them
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamHyper& h, float step_size,
                                            float bc2_sqrt) {
  if (h.weight_decay != 0.f) g = __fmaf_rn(h.weight_decay, p, g);
  m = m + (g - m) * h.one_minus_beta1;
  v = v * h.beta2f + h.one_minus_beta2 * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + h.eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(kAdamThreads) adam_kernel(const AdamBatch b, const AdamHyper h,
                                                            const uint64_t* __restrict__ step) {
  __shared__ float s_step_size, s_bc2_sqrt;
  __shared__ int s_slot;
  if (threadIdx.x == 0) {
    const double t = double(step[0] + 1);
    const double bc1 = 1.0 - pow(h.beta1, t);
    const double bc2 = 1.0 - pow(h.beta2, t);
    s_step_size = float(h.lr / bc1);
    s_bc2_sqrt = float(sqrt(bc2));
    int s = 0;
    while (s + 1 < b.n_tensors && int(blockIdx.x) >= b.block_first[s + 1]) ++s;
    s_slot = s;
  }
  __syncthreads();
  const int s = s_slot;
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  float* __restrict__ p = b.p[s];
  const float* __restrict__ g = b.g[s];
  float* __restrict__ m = b.m[s];
  float* __restrict__ v = b.v[s];
  const int64_t n = b.n[s];
  const int64_t base = int64_t(int(blockIdx.x) - b.block_first[s]) * kAdamBlockElems;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                         reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
#pragma unroll
  for (int it = 0; it < kAdamVecPerThread; ++it) {
    const int64_t i = base + (int64_t(it) * kAdamThreads + threadIdx.x) * 4;
    if (i >= n) break;
    if (aligned && i + 4 <= n) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      adam_update(pv.x, gv.x, mv.x, vv.x, h, step_size, bc2_sqrt);
      adam_update(pv.y, gv.y, mv.y, vv.y, h, step_size, bc2_sqrt);
      adam_update(pv.z, gv.z, mv.z, vv.z, h, step_size, bc2_sqrt);
      adam_update(pv.w, gv.w, mv.w, vv.w, h, step_size, bc2_sqrt);
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (int64_t j = i; j < n && j < i + 4; ++j) {
        float pj = p[j], mj = m[j], vj = v[j];
        adam_update(pj, g[j], mj, vj, h, step_size, bc2_sqrt);
        p[j] = pj;
        m[j] = mj;
        v[j] = vj;
      }
    }
  }
}

__global__ void adam_advance_kernel(uint64_t* step) { step[0] += 1; }

}  // namespace gn

using namespace gn;

extern "C" {

int gn_adam_max_tensors_per_launch(void) { return kAdamSlots; }

int gn_adam_step(const gn_adam_tensor* tensors, int32_t n_tensors, double lr, double beta1, double beta2, double eps,
                 double weight_decay, uint64_t* step, void* stream) {
  if (n_tensors < 0 || (n_tensors > 0 && tensors == nullptr) || step == nullptr) return GN_ERR_ARG;
  if (!(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0) || !(lr >= 0.0))
    return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const AdamHyper h{lr, beta1, beta2, float(beta2), float(1.0 - beta1), float(1.0 - beta2), float(eps),
                    float(weight_decay)};
  int32_t i = 0;
  while (i < n_tensors) {
    AdamBatch b;
    b.n_tensors = 0;
    int64_t blocks = 0;
    while (i < n_tensors && b.n_tensors < kAdamSlots) {
      const gn_adam_tensor& t = tensors[i++];
      if (t.n < 0) return GN_ERR_ARG;
      if (t.n == 0) continue;
      if (t.param == nullptr || t.grad == nullptr || t.exp_avg == nullptr || t.exp_avg_sq == nullptr)
        return GN_ERR_ARG;
      const int64_t nb = ceil_div(t.n, kAdamBlockElems);
      if (blocks + nb >= (int64_t(1) << 31)) return GN_ERR_RANGE;
      const int s = b.n_tensors++;
      b.p[s] = t.param;
      b.g[s] = t.grad;
      b.m[s] = t.exp_avg;
      b.v[s] = t.exp_avg_sq;
      b.n[s] = t.n;
      b.block_first[s] = int32_t(blocks);
      blocks += nb;
    }
    if (b.n_tensors == 0) continue;
    b.block_first[b.n_tensors] = int32_t(blocks);
    GN_LAUNCH(adam_kernel, (unsigned)blocks, kAdamThreads, 0, st, b, h, (const uint64_t*)step);
  }
  GN_LAUNCH(adam_advance_kernel, 1, 1, 0, st, step);
  return GN_OK;
}

}  // extern "C"
