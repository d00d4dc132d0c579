// Elementwise / reduction glue of the layer stacks: strided 2-D maps (the
// torch.cat / abs / relu of gripnet/layers.py:279-309, :369-384), ReLU and abs
// backward, deterministic column sums (bias gradients) and the fused
// link-prediction / node-classification losses (GripNet-pose.py:140-142,
// GripNet-aminer.py:133).
#include "common.cuh"

namespace gn {

template <int OP>
__device__ __forceinline__ float ew_apply(float s, float d) {
  if (OP == GN_EW_COPY) return s;
  if (OP == GN_EW_ABS) return fabsf(s);
  if (OP == GN_EW_RELU) return fmaxf(s, 0.f);
  return d + s;  // GN_EW_ADD
}

template <int OP>
__global__ void map2d_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd,
                             int64_t n, int F) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * F) return;
  const int64_t r = idx / F;
  const int c = int(idx - r * F);
  float* d = dst + r * ldd + c;
  *d = ew_apply<OP>(src[r * lds + c], OP == GN_EW_ADD ? *d : 0.f);
}

// 128-bit variant: F, lds, ldd multiples of 4 and 16-byte aligned bases
template <int OP>
__global__ void map2d_vec_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd,
                                 int64_t n, int F4) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * F4) return;
  const int64_t r = idx / F4;
  const int c = int(idx - r * F4) * 4;
  const float4 s = *reinterpret_cast<const float4*>(src + r * lds + c);
  float4* dp = reinterpret_cast<float4*>(dst + r * ldd + c);
  float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
  if (OP == GN_EW_ADD) d = *dp;
  *dp = make_float4(ew_apply<OP>(s.x, d.x), ew_apply<OP>(s.y, d.y), ew_apply<OP>(s.z, d.z), ew_apply<OP>(s.w, d.w));
}

__global__ void relu_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ y, int64_t ldy,
                                float* __restrict__ dst, int64_t ldd, int64_t n, int F) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * F) return;
  const int64_t r = idx / F;
  const int c = int(idx - r * F);
  dst[r * ldd + c] = (y[r * ldy + c] > 0.f) ? g[r * ldg + c] : 0.f;
}

// 128-bit variants (F and the leading dimensions multiples of 4, 16-byte aligned bases): one float4 per operand
__global__ void relu_bwd_vec_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ y, int64_t ldy,
                                    float* __restrict__ dst, int64_t ldd, int64_t n, int F4) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * F4) return;
  const int64_t r = idx / F4;
  const int c = int(idx - r * F4) * 4;
  const float4 gv = *reinterpret_cast<const float4*>(g + r * ldg + c);
  const float4 yv = *reinterpret_cast<const float4*>(y + r * ldy + c);
  *reinterpret_cast<float4*>(dst + r * ldd + c) = make_float4(yv.x > 0.f ? gv.x : 0.f, yv.y > 0.f ? gv.y : 0.f,
                                                              yv.z > 0.f ? gv.z : 0.f, yv.w > 0.f ? gv.w : 0.f);
}

__device__ __forceinline__ float sgn_scaled(float t, float g, float scale) {
  return g * ((t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f)) * scale;
}

__global__ void abs_bwd_vec_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ t, int64_t ldt,
                                   float* __restrict__ dst, int64_t ldd, int64_t n, int F4, float scale) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * F4) return;
  const int64_t r = idx / F4;
  const int c = int(idx - r * F4) * 4;
  const float4 gv = *reinterpret_cast<const float4*>(g + r * ldg + c);
  const float4 tv = *reinterpret_cast<const float4*>(t + r * ldt + c);
  *reinterpret_cast<float4*>(dst + r * ldd + c) =
      make_float4(sgn_scaled(tv.x, gv.x, scale), sgn_scaled(tv.y, gv.y, scale), sgn_scaled(tv.z, gv.z, scale),
                  sgn_scaled(tv.w, gv.w, scale));
}

__global__ void abs_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ t, int64_t ldt,
                               float* __restrict__ dst, int64_t ldd, int64_t n, int F, float scale) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * F) return;
  const int64_t r = idx / F;
  const int c = int(idx - r * F);
  const float tv = t[r * ldt + c];
  const float sgn = (tv > 0.f) ? 1.f : ((tv < 0.f) ? -1.f : 0.f);
  dst[r * ldd + c] = scale * g[r * ldg + c] * sgn;
}

__global__ void axpby_kernel(const float* __restrict__ a, int64_t lda, float alpha, const float* b, int64_t ldb,
                             float beta, float* dst, int64_t ldd, int64_t n, int F) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * F) return;
  const int64_t r = idx / F;
  const int c = int(idx - r * F);
  float v = alpha * a[r * lda + c];
  if (b) v += beta * b[r * ldb + c];
  dst[r * ldd + c] = v;
}

// ((a + b) + c) / 3 in the reference's evaluation order and with a true division (GripNet-freebase-d.py:160-161)
__global__ void mean3_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                             const float* __restrict__ c, int64_t ldc, float* __restrict__ dst, int64_t ldd, int64_t n,
                             int F) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * F) return;
  const int64_t r = idx / F;
  const int col = int(idx - r * F);
  const float v = __fadd_rn(__fadd_rn(a[r * lda + col], b[r * ldb + col]), c[r * ldc + col]);
  dst[r * ldd + col] = __fdiv_rn(v, 3.0f);
}

// ---- column sums: block b sums rows [b*rows_per_block, ...) -> ws[b, F]; then a
// second kernel adds the block rows in order.
constexpr int kColsumThreads = 256;

__global__ void __launch_bounds__(kColsumThreads) colsum_partial_kernel(const float* __restrict__ x, int64_t ldx,
                                                                        int64_t n, int F, int rows_per_block,
                                                                        float* __restrict__ ws) {
  // thread -> (column c, row lane j); row lanes of one column are reduced through smem in fixed order
  extern __shared__ float sm[];
  const int cols_per_pass = F < kColsumThreads ? F : kColsumThreads;
  const int row_lanes = kColsumThreads / cols_per_pass;
  const int c_in = threadIdx.x % cols_per_pass, j = threadIdx.x / cols_per_pass;
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(n, r0 + rows_per_block);
  for (int c0 = 0; c0 < F; c0 += cols_per_pass) {
    const int c = c0 + c_in;
    float s = 0.f;
    if (c < F && j < row_lanes)
      for (int64_t r = r0 + j; r < r1; r += row_lanes) s += x[r * ldx + c];
    sm[threadIdx.x] = s;
    __syncthreads();
    if (j == 0 && c < F) {
      float t = 0.f;
      for (int q = 0; q < row_lanes; ++q) t += sm[q * cols_per_pass + c_in];
      ws[int64_t(blockIdx.x) * F + c] = t;
    }
    __syncthreads();
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ ws, int n_blocks, int F, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= F) return;
  float s = 0.f;
  for (int b = 0; b < n_blocks; ++b) s += ws[int64_t(b) * F + c];
  out[c] = s;
}

inline int colsum_blocks(int64_t n) {
  int64_t b = ceil_div(n > 0 ? n : 1, 256);
  return int(b < 592 ? b : 592);  // 4 blocks per SM on 148 SMs
}

// ---- fused losses -----------------------------------------------------------
constexpr int kLossThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < kLossThreads / 32; ++w) t += sm[w];
  return t;  // valid on thread 0
}

// ws[b] = sum over the block's strided slice of  -log(pos+eps)/n_pos  and  -log(1-neg+eps)/n_neg
__global__ void __launch_bounds__(kLossThreads) lp_loss_partial_kernel(const float* __restrict__ pos, int64_t n_pos,
                                                                       const float* __restrict__ neg, int64_t n_neg,
                                                                       float eps, float* __restrict__ ws) {
  __shared__ float sm[kLossThreads / 32];
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  float sp = 0.f, sn = 0.f;
  for (int64_t i = tid; i < n_pos; i += stride) sp += logf(pos[i] + eps);
  for (int64_t i = tid; i < n_neg; i += stride) sn += logf(1.0f - neg[i] + eps);
  const float a = block_sum(sp, sm);
  __syncthreads();
  const float b = block_sum(sn, sm);
  if (threadIdx.x == 0) {
    ws[2 * blockIdx.x] = a;
    ws[2 * blockIdx.x + 1] = b;
  }
}

// one warp: lane l adds partials l, l+32, ... in order, then a fixed shuffle tree (deterministic)
__global__ void lp_loss_final_kernel(const float* __restrict__ ws, int n_blocks, int64_t n_pos, int64_t n_neg,
                                     float* __restrict__ loss) {
  const int lane = threadIdx.x;
  float a = 0.f, b = 0.f;
  for (int i = lane; i < n_blocks; i += 32) {
    a += ws[2 * i];
    b += ws[2 * i + 1];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(kFull, a, off);
    b += __shfl_xor_sync(kFull, b, off);
  }
  if (lane != 0) return;
  const float lp = n_pos > 0 ? -(a / float(n_pos)) : 0.f;
  const float ln = n_neg > 0 ? -(b / float(n_neg)) : 0.f;
  loss[0] = lp + ln;
}

__global__ void lp_loss_bwd_kernel(const float* __restrict__ pos, int64_t n_pos, const float* __restrict__ neg,
                                   int64_t n_neg, float eps, const float* __restrict__ grad_loss,
                                   float* __restrict__ grad_pos, float* __restrict__ grad_neg) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const float g = grad_loss[0];
  if (i < n_pos) grad_pos[i] = -g / (float(n_pos) * (pos[i] + eps));
  if (i < n_neg) grad_neg[i] = g / (float(n_neg) * (1.0f - neg[i] + eps));
}

__global__ void __launch_bounds__(kLossThreads) nc_loss_partial_kernel(const float* __restrict__ score, int64_t n, int C,
                                                                       const int64_t* __restrict__ label, float eps,
                                                                       float* __restrict__ ws) {
  __shared__ float sm[kLossThreads / 32];
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  float s = 0.f;
  for (int64_t i = tid; i < n; i += stride) s += logf(score[i * C + label[i]] + eps);
  const float a = block_sum(s, sm);
  if (threadIdx.x == 0) ws[blockIdx.x] = a;
}

__global__ void nc_loss_final_kernel(const float* __restrict__ ws, int n_blocks, int64_t n, float* __restrict__ loss) {
  const int lane = threadIdx.x;
  float a = 0.f;
  for (int i = lane; i < n_blocks; i += 32) a += ws[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(kFull, a, off);
  if (lane == 0) loss[0] = n > 0 ? -(a / float(n)) : 0.f;
}

__global__ void nc_loss_bwd_kernel(const float* __restrict__ score, int64_t n, int C,
                                   const int64_t* __restrict__ label, float eps, const float* __restrict__ grad_loss,
                                   float* __restrict__ grad_score) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * C) return;
  const int64_t i = idx / C;
  const int c = int(idx - i * C);
  grad_score[idx] = (c == label[i]) ? -grad_loss[0] / (float(n) * (score[idx] + eps)) : 0.f;
}

inline int loss_blocks(int64_t n) {
  int64_t b = ceil_div(n > 0 ? n : 1, kLossThreads * 4);
  return int(b < 296 ? (b > 0 ? b : 1) : 296);
}

}  // namespace gn

using namespace gn;

extern "C" {

int gn_map2d(int op, const float* src, int64_t lds, float* dst, int64_t ldd, int64_t n, int32_t F, void* stream) {
  if (n < 0 || F <= 0) return GN_ERR_ARG;
  if (n == 0) return GN_OK;
  if (!src || !dst) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const bool v4 = (F % 4 == 0) && (lds % 4 == 0) && (ldd % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
  const int64_t work = v4 ? n * (F / 4) : n * int64_t(F);
  const unsigned grid = (unsigned)ceil_div(work, 256);
#define GN_MAP_CASE(OP)                                                                        \
  case OP:                                                                                     \
    if (v4) { GN_LAUNCH((map2d_vec_kernel<OP>), grid, 256, 0, st, src, lds, dst, ldd, n, F / 4); } \
    else { GN_LAUNCH((map2d_kernel<OP>), grid, 256, 0, st, src, lds, dst, ldd, n, F); }        \
    break;
  switch (op) {
    GN_MAP_CASE(GN_EW_COPY)
    GN_MAP_CASE(GN_EW_ABS)
    GN_MAP_CASE(GN_EW_RELU)
    GN_MAP_CASE(GN_EW_ADD)
    default: return GN_ERR_ARG;
  }
#undef GN_MAP_CASE
  return GN_OK;
}

int gn_relu_bwd(const float* g, int64_t ldg, const float* y, int64_t ldy, float* dst, int64_t ldd, int64_t n,
                int32_t F, void* stream) {
  if (n < 0 || F <= 0) return GN_ERR_ARG;
  if (n == 0) return GN_OK;
  if (!g || !y || !dst) return GN_ERR_ARG;
  const bool v4 = F % 4 == 0 && ldg % 4 == 0 && ldy % 4 == 0 && ldd % 4 == 0 &&
                  ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  if (v4) {
    GN_LAUNCH(relu_bwd_vec_kernel, (unsigned)ceil_div(n * (F / 4), 256), 256, 0, as_stream(stream), g, ldg, y, ldy, dst,
              ldd, n, F / 4);
    return GN_OK;
  }
  GN_LAUNCH(relu_bwd_kernel, (unsigned)ceil_div(n * F, 256), 256, 0, as_stream(stream), g, ldg, y, ldy, dst, ldd, n, F);
  return GN_OK;
}

int gn_abs_bwd(const float* g, int64_t ldg, const float* t, int64_t ldt, float* dst, int64_t ldd, int64_t n,
               int32_t F, float scale, void* stream) {
  if (n < 0 || F <= 0) return GN_ERR_ARG;
  if (n == 0) return GN_OK;
  if (!g || !t || !dst) return GN_ERR_ARG;
  const bool v4 = F % 4 == 0 && ldg % 4 == 0 && ldt % 4 == 0 && ldd % 4 == 0 &&
                  ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  if (v4) {
    GN_LAUNCH(abs_bwd_vec_kernel, (unsigned)ceil_div(n * (F / 4), 256), 256, 0, as_stream(stream), g, ldg, t, ldt, dst,
              ldd, n, F / 4, scale);
    return GN_OK;
  }
  GN_LAUNCH(abs_bwd_kernel, (unsigned)ceil_div(n * F, 256), 256, 0, as_stream(stream), g, ldg, t, ldt, dst, ldd, n, F,
            scale);
  return GN_OK;
}

int gn_axpby(const float* a, int64_t lda, float alpha, const float* b, int64_t ldb, float beta, float* dst,
             int64_t ldd, int64_t n, int32_t F, void* stream) {
  if (n < 0 || F <= 0) return GN_ERR_ARG;
  if (n == 0) return GN_OK;
  if (!a || !dst) return GN_ERR_ARG;
  GN_LAUNCH(axpby_kernel, (unsigned)ceil_div(n * F, 256), 256, 0, as_stream(stream), a, lda, alpha, b, ldb, beta, dst,
            ldd, n, F);
  return GN_OK;
}

int gn_zero(void* p, size_t bytes, void* stream) {
  if (bytes == 0) return GN_OK;
  if (!p) return GN_ERR_ARG;
  if (cudaMemsetAsync(p, 0, bytes, as_stream(stream)) != cudaSuccess) return GN_ERR_CUDA;
  return GN_OK;
}

int gn_mean3(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c, int64_t ldc, float* dst,
             int64_t ldd, int64_t n, int32_t F, void* stream) {
  if (n < 0 || F <= 0) return GN_ERR_ARG;
  if (n == 0) return GN_OK;
  if (!a || !b || !c || !dst) return GN_ERR_ARG;
  GN_LAUNCH(mean3_kernel, (unsigned)ceil_div(n * F, 256), 256, 0, as_stream(stream), a, lda, b, ldb, c, ldc, dst, ldd,
            n, F);
  return GN_OK;
}

size_t gn_colsum_workspace_bytes(int64_t n, int32_t F) { return size_t(colsum_blocks(n)) * size_t(F) * 4 + 256; }

int gn_colsum(const float* x, int64_t ldx, int64_t n, int32_t F, float* out, void* ws, size_t ws_bytes, void* stream) {
  if (n < 0 || F <= 0 || !out) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    if (cudaMemsetAsync(out, 0, size_t(F) * 4, st) != cudaSuccess) return GN_ERR_CUDA;
    return GN_OK;
  }
  if (!x) return GN_ERR_ARG;
  const int blocks = colsum_blocks(n);
  if (!ws || ws_bytes < size_t(blocks) * F * 4) return GN_ERR_WORKSPACE;
  const int rows_per_block = int(ceil_div(n, blocks));
  GN_LAUNCH(colsum_partial_kernel, (unsigned)blocks, kColsumThreads, kColsumThreads * sizeof(float), st, x, ldx, n, F,
            rows_per_block, static_cast<float*>(ws));
  GN_LAUNCH(colsum_final_kernel, (unsigned)ceil_div(F, 128), 128, 0, st, static_cast<const float*>(ws), blocks, F, out);
  return GN_OK;
}

size_t gn_loss_workspace_bytes(int64_t n) { return size_t(loss_blocks(n)) * 2 * 4 + 256; }

int gn_lp_loss_fwd(const float* pos, int64_t n_pos, const float* neg, int64_t n_neg, float eps, float* loss, void* ws,
                   size_t ws_bytes, void* stream) {
  if (n_pos < 0 || n_neg < 0 || !loss) return GN_ERR_ARG;
  if ((n_pos > 0 && !pos) || (n_neg > 0 && !neg)) return GN_ERR_ARG;
  const int64_t n = n_pos > n_neg ? n_pos : n_neg;
  const int blocks = loss_blocks(n);
  if (!ws || ws_bytes < size_t(blocks) * 8) return GN_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  GN_LAUNCH(lp_loss_partial_kernel, (unsigned)blocks, kLossThreads, 0, st, pos, n_pos, neg, n_neg, eps,
            static_cast<float*>(ws));
  GN_LAUNCH(lp_loss_final_kernel, 1, 32, 0, st, static_cast<const float*>(ws), blocks, n_pos, n_neg, loss);
  return GN_OK;
}

int gn_lp_loss_bwd(const float* pos, int64_t n_pos, const float* neg, int64_t n_neg, float eps,
                   const float* grad_loss, float* grad_pos, float* grad_neg, void* stream) {
  if (n_pos < 0 || n_neg < 0 || !grad_loss) return GN_ERR_ARG;
  const int64_t n = n_pos > n_neg ? n_pos : n_neg;
  if (n == 0) return GN_OK;
  if ((n_pos > 0 && (!pos || !grad_pos)) || (n_neg > 0 && (!neg || !grad_neg))) return GN_ERR_ARG;
  GN_LAUNCH(lp_loss_bwd_kernel, (unsigned)ceil_div(n, 256), 256, 0, as_stream(stream), pos, n_pos, neg, n_neg, eps,
            grad_loss, grad_pos, grad_neg);
  return GN_OK;
}

int gn_nc_loss_fwd(const float* score, int64_t n, int32_t C, const int64_t* label, float eps, float* loss, void* ws,
                   size_t ws_bytes, void* stream) {
  if (n < 0 || C <= 0 || !loss) return GN_ERR_ARG;
  if (n > 0 && (!score || !label)) return GN_ERR_ARG;
  const int blocks = loss_blocks(n);
  if (!ws || ws_bytes < size_t(blocks) * 4) return GN_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  GN_LAUNCH(nc_loss_partial_kernel, (unsigned)blocks, kLossThreads, 0, st, score, n, C, label, eps,
            static_cast<float*>(ws));
  GN_LAUNCH(nc_loss_final_kernel, 1, 32, 0, st, static_cast<const float*>(ws), blocks, n, loss);
  return GN_OK;
}

int gn_nc_loss_bwd(const float* score, int64_t n, int32_t C, const int64_t* label, float eps, const float* grad_loss,
                   float* grad_score, void* stream) {
  if (n < 0 || C <= 0 || !grad_loss) return GN_ERR_ARG;
  if (n == 0) return GN_OK;
  if (!score || !label || !grad_score) return GN_ERR_ARG;
  GN_LAUNCH(nc_loss_bwd_kernel, (unsigned)ceil_div(n * C, 256), 256, 0, as_stream(stream), score, n, C, label, eps,
            grad_loss, grad_score);
  return GN_OK;
}

}  // extern "C"
