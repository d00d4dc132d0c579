// K3/K4/K5/K7/K8: CSR SpMM  out = act(row_scale * (A x) + bias + addend).
//
// HBM-bound gather kernel (E'(4+4F) + N(8+4F) algorithmic bytes):
//  * warp per chunk of a row (row-split for hub rows, see rowsplit.cuh);
//  * a warp reads 32 (col,val) pairs with one coalesced load each, then issues the
//    feature-row gathers as independent 128-bit loads (LPE lanes cover one row,
//    32/LPE rows per step, LPE steps unrolled -> LPE loads in flight per lane);
//  * bias / root addend / ReLU / column-slice write fused in the epilogue;
//  * no data atomics; summation order fixed -> deterministic fwd and bwd.
#include <cstdlib>

#include "rowsplit.cuh"

namespace gn {

template <int LPE, int VEC>
__global__ void __launch_bounds__(256, (LPE == 16 ? 5 : 6)) spmm_kernel(
    const gn_csr csr, const float* __restrict__ x, int64_t ldx, int F, const float* __restrict__ row_scale,
    const float* __restrict__ bias, const float* addend, int64_t ld_addend, int relu, float* out, int64_t ldo,
    float* __restrict__ partial) {
  ChunkInfo ci;
  if (!chunk_info(csr, ci)) return;
  constexpr int EPI = 32 / LPE;
  const int lane = threadIdx.x & 31;
  const int slot = lane / LPE, fl = lane % LPE;
  const int f = fl * VEC;
  const bool f_ok = f < F;

  Vec<VEC> acc[1];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0].v[i] = 0.f;

  for (int base = ci.beg; base < ci.end; base += 32) {
    int c = -1;
    float w = 0.f;
    const int mine = base + lane;
    if (mine < ci.end) {
      c = __ldg(csr.col + mine);
      w = csr.val ? __ldg(csr.val + mine) : 1.0f;
    }
#pragma unroll
    for (int t = 0; t < LPE; ++t) {
      const int from = t * EPI + slot;
      const int cc = __shfl_sync(kFull, c, from);
      const float ww = __shfl_sync(kFull, w, from);
      if (cc >= 0 && f_ok) {
        const Vec<VEC> xv = load_vec<VEC>(x + int64_t(cc) * ldx + f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[0].v[i] = fmaf(ww, xv.v[i], acc[0].v[i]);
      }
    }
  }
  reduce_slots<LPE, VEC>(acc[0]);

  const int row = ci.row;
  const float scale = row_scale ? __ldg(row_scale + row) : 1.0f;
  auto emit = [&](int, int ff, const Vec<VEC>& s) {
    Vec<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = s.v[i] * scale;
    if (bias) {
      const Vec<VEC> b = load_vec<VEC>(bias + ff);
#pragma unroll
      for (int i = 0; i < VEC; ++i) r.v[i] += b.v[i];
    }
    if (addend) {
      const Vec<VEC> a = load_vec_plain<VEC>(addend + int64_t(row) * ld_addend + ff);
#pragma unroll
      for (int i = 0; i < VEC; ++i) r.v[i] += a.v[i];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) r.v[i] = fmaxf(r.v[i], 0.f);
    }
    store_vec<VEC>(out + int64_t(row) * ldo + ff, r);
  };
  finish_row<LPE, VEC, 1>(csr, ci, acc, F, partial, emit);
}

template <int VEC>
static int launch_spmm(int lpe, const gn_csr& csr, const float* x, int64_t ldx, int F, const float* row_scale,
                       const float* bias, const float* addend, int64_t ld_addend, int relu, float* out, int64_t ldo,
                       float* partial, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(csr.n_chunks, 8);
#define GN_SPMM_CASE(L)                                                                                  \
  case L:                                                                                                \
    GN_LAUNCH((spmm_kernel<L, VEC>), grid, 256, 0, st, csr, x, ldx, F, row_scale, bias, addend, ld_addend, \
              relu, out, ldo, partial);                                                                  \
    break;
  switch (lpe) {
    GN_SPMM_CASE(1)
    GN_SPMM_CASE(2)
    GN_SPMM_CASE(4)
    GN_SPMM_CASE(8)
    GN_SPMM_CASE(16)
    GN_SPMM_CASE(32)
    default: return GN_ERR_ARG;
  }
#undef GN_SPMM_CASE
  return GN_OK;
}

}  // namespace gn

using namespace gn;

extern "C" int gn_spmm(const gn_csr* csr, const float* x, int64_t ldx, int32_t F, const float* row_scale,
                       const float* bias, const float* addend, int64_t ld_addend, int relu, float* out, int64_t ldo,
                       float* partial, void* stream) {
  if (!csr || !x || !out || F <= 0 || !csr->rowptr || !csr->chunk_ptr || !csr->chunk_row || !csr->chunk_beg ||
      !csr->row_counter || csr->chunk_len <= 0)
    return GN_ERR_ARG;
  if (csr->nnz > 0 && !csr->col) return GN_ERR_ARG;
  if (csr->n_rows == 0 || csr->n_chunks == 0) return GN_OK;
  if (csr->n_chunks > csr->n_rows && partial == nullptr) return GN_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  const bool vec4 = (F % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned16(x) && aligned16(out) &&
                    (!partial || aligned16(partial)) && (!bias || aligned16(bias)) &&
                    (!addend || (aligned16(addend) && ld_addend % 4 == 0));
  if (vec4) {
    // one launch covers up to 128 columns; wider rows are processed in 128-column panels
    for (int f0 = 0; f0 < F; f0 += 128) {
      const int w = F - f0 < 128 ? F - f0 : 128;
      const int lpe = pow2_ceil(w / 4);
      GN_CHECK(launch_spmm<4>(lpe, *csr, x + f0, ldx, w, row_scale, bias ? bias + f0 : nullptr,
                              addend ? addend + f0 : nullptr, ld_addend, relu, out + f0, ldo, partial, st));
    }
  } else {
    for (int f0 = 0; f0 < F; f0 += 32) {
      const int w = F - f0 < 32 ? F - f0 : 32;
      const int lpe = pow2_ceil(w);
      GN_CHECK(launch_spmm<1>(lpe, *csr, x + f0, ldx, w, row_scale, bias ? bias + f0 : nullptr,
                              addend ? addend + f0 : nullptr, ld_addend, relu, out + f0, ldo, partial, st));
    }
  }
  return GN_OK;
}
