"""GPU: tcgen05 3xTF32 dense transform (gn_tc_gemm) against a float64 product.

Tolerance: the parity bar of north_star (1e-5 relative, max|a-b|/max|b|) with a margin — the
error-compensated split must land at fp32-GEMM accuracy (~1e-6), far from plain TF32 (~5e-4)."""
import numpy as np
import pytest
import torch

from golden_util import rel_err

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _tc(tb, A, B, C, addend=None, mask=None):
    from gripnet_b200 import _lib
    from gripnet_b200.graph import _ptr, _stream
    lib = _lib.load()
    m, k = A.shape
    n = C.shape[1]
    nbytes = int(lib.gn_tc_gemm_workspace_bytes(m, n, k))
    assert nbytes > 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    rc = lib.gn_tc_gemm(int(tb), m, n, k, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(),
                        C.stride(0), _ptr(addend), addend.stride(0) if addend is not None else 0, _ptr(mask),
                        mask.stride(0) if mask is not None else 0, _ptr(ws), nbytes, _stream())
    _lib.check(rc, "gn_tc_gemm")
    torch.cuda.synchronize()


@pytest.mark.parametrize("tb", [0, 1])
@pytest.mark.parametrize("m,n,k", [(128, 16, 8), (300, 16, 32), (1000, 64, 128), (4173, 128, 256), (513, 48, 48),
                                   (2000, 512, 128), (777, 272, 80), (130, 20, 4), (5000, 32, 512), (64, 256, 36), (3000, 128, 1024), (1500, 256, 640),
                                   # one CTA per SM walking several M tiles: the deep-prefetch loader (tc_gemm_kernel<3>)
                                   (40000, 128, 256), (25003, 256, 64), (60000, 64, 288)])
def test_tc_gemm_matches_float64(tb, m, n, k):
    rs = np.random.RandomState(m + n + k)
    A = rs.randn(m, k).astype(np.float32)
    B = rs.randn(k, n).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64)
    d = _dev()
    At = torch.from_numpy(A).to(d)
    Bt = torch.from_numpy(B.T.copy() if tb else B).to(d)
    C = torch.full((m, n), float("nan"), device=d)
    _tc(tb, At, Bt, C)
    err = rel_err(C, want)
    # one TMEM accumulator covers <= 128 k (<= 320 k when N = 256 leaves room for only two accumulators)
    assert err < (2e-6 if k <= 256 else 4e-6), err


def test_tc_gemm_not_plain_tf32():
    """Operands with 24 significant bits: a plain-TF32 product would be off by ~1e-3."""
    d = _dev()
    rs = np.random.RandomState(0)
    A = (1.0 + rs.rand(512, 64) * 2 ** -12).astype(np.float32)
    B = (1.0 + rs.rand(64, 32) * 2 ** -12).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64)
    C = torch.empty(512, 32, device=d)
    _tc(0, torch.from_numpy(A).to(d), torch.from_numpy(B).to(d), C)
    diff = (C.double().cpu().numpy() - want)
    # all-positive sums are the worst case of the tensor core's TRUNCATING fp32 accumulate (a bias of about
    # half an ulp per hi*hi MMA, 8 of them here); plain TF32 would be ~2e-4
    assert np.abs(diff).max() / np.abs(want).max() < 1e-6
    # and the fine structure is resolved: compare the deviation from the mean product
    assert np.abs(diff).max() < 0.05 * np.abs(want - want.mean()).max()


def test_tc_gemm_epilogue_and_strided_slices():
    d = _dev()
    rs = np.random.RandomState(5)
    m, k, n = 900, 64, 32
    big_a = torch.from_numpy(rs.randn(m, 96).astype(np.float32)).to(d)       # A = columns 16..80 of a concat buffer
    A = big_a[:, 16:80]
    W = torch.from_numpy(rs.randn(k, n).astype(np.float32)).to(d)
    big_c = torch.zeros(m, 64, device=d)
    C = big_c[:, 32:64]
    addend = torch.from_numpy(rs.randn(m, n).astype(np.float32)).to(d)
    mask = torch.from_numpy(rs.randn(m, n).astype(np.float32)).to(d)
    _tc(0, A, W, C, addend=addend, mask=mask)
    want = (A.double() @ W.double() + addend.double()) * (mask > 0).double()
    assert rel_err(C, want) < 2e-6
    assert float(big_c[:, :32].abs().sum()) == 0.0                           # nothing written outside the slice


def test_dispatch_routes_tall_products_to_tensor_cores():
    """ops.sgemm sends tall plain products to gn_tc_gemm: ONE launch when B is a small weight matrix (split
    inside the kernel by the B-producer warp), two (B image + UMMA kernel) when B is wide."""
    import gripnet_b200 as gb
    from gripnet_b200 import ops
    d = _dev()
    m, k, n = 8192, 64, 16
    A = torch.randn(m, k, device=d)
    W = torch.randn(k, n, device=d)
    C = torch.empty(m, n, device=d)
    before = gb.launch_count()
    ops.sgemm(False, False, m, n, k, A.data_ptr(), k, W.data_ptr(), n, C.data_ptr(), n, d)
    assert gb.launch_count() - before == 1
    assert rel_err(C, A.double() @ W.double()) < 2e-6
    n2 = 80                                              # nt > 64: the image path
    W2 = torch.randn(k, n2, device=d)
    C2 = torch.empty(m, n2, device=d)
    before = gb.launch_count()
    ops.sgemm(False, False, m, n2, k, A.data_ptr(), k, W2.data_ptr(), n2, C2.data_ptr(), n2, d)
    assert gb.launch_count() - before == 2
    assert rel_err(C2, A.double() @ W2.double()) < 2e-6


@pytest.mark.parametrize("n,mo,no,inner", [(645, 48, 512, 32), (645, 48, 35104, 32), (5000, 32, 16, 0), (70000, 64, 64, 0),
                                           (200000, 192, 64, 0), (33, 4, 8, 0), (4097, 128, 128, 0),
                                           (1000, 256, 32, 0)])
def test_tc_tn_weight_gradient_products(n, mo, no, inner):
    """C = A^T B on tcgen05 (gn_tc_tn: transposing 3xTF32 loaders, TMEM runs drained into fp32 registers, split
    partials added in order) against float64 — GCN dW shapes, long reductions over several splits, and the
    relation-batched dW of config 4 written in the [R][k][f] layout.  1e-5 relative like every fp32 result."""
    from gripnet_b200 import ops
    d = _dev()
    gen = torch.Generator().manual_seed(n + mo + no)
    a = torch.randn(n, mo, generator=gen)
    b = torch.randn(n, no, generator=gen)
    big = torch.zeros(n, mo + 8)
    big[:, 4:4 + mo] = a
    ad = big.to(d)[:, 4:4 + mo]                                   # a column slice: leading dimension != width
    bd = b.to(d)
    ref = a.double().t() @ b.double()
    outs = []
    for _ in range(2):
        if inner:
            out = torch.empty(no // inner, mo, inner, device=d)
            ok = ops.weight_grad(ops.M(ad), ops.M(bd), out, d, c_inner=inner, c_stride=mo * inner, force_tc=True)
        else:
            out = torch.empty(mo, no, device=d)
            ok = ops.weight_grad(ops.M(ad), ops.M(bd), out, d, force_tc=True)
        assert ok, "the tensor path must accept these aligned shapes"
        outs.append(out)
    torch.cuda.synchronize()
    got = outs[0].permute(1, 0, 2).reshape(mo, no) if inner else outs[0]
    err = float((got.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err
    assert torch.equal(outs[0], outs[1])                          # deterministic


@pytest.mark.parametrize("n,k,r,f", [(645, 48, 16, 32), (645, 48, 1097, 32), (130, 16, 3, 8), (5000, 64, 40, 16)])
def test_tc_relation_batched_transform(n, k, r, f):
    """Y[:, r, :] = X W[r] for every relation at once on tcgen05 (gn_tc_gemm_rel, W kept in its [R][k][f] layout)."""
    from gripnet_b200 import ops
    d = _dev()
    gen = torch.Generator().manual_seed(n + r)
    x = torch.randn(n, k, generator=gen)
    w = torch.randn(r, k, f, generator=gen) / np.sqrt(k)
    xd, wd = x.to(d), w.to(d)
    y = torch.empty(n, r * f, device=d)
    ops.rel_transform(ops.M(xd), wd, ops.M(y), r, k, f, d)
    ref = torch.einsum("nk,rkf->nrf", x.double(), w.double()).reshape(n, r * f)
    assert rel_err(y, ref) < 1e-5
