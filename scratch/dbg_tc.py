import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch, numpy as np
from golden_util import rel_err
from gripnet_b200 import _lib, ops
from gripnet_b200.graph import _ptr, _stream
dev = torch.device("cuda:0")
lib = _lib.load()

def tc(tb, A, B, C, addend=None, mask=None):
    m, k = A.shape; n = C.shape[1]
    nbytes = int(lib.gn_tc_gemm_workspace_bytes(m, n, k))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    rc = lib.gn_tc_gemm(int(tb), m, n, k, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(),
                        C.stride(0), _ptr(addend), addend.stride(0) if addend is not None else 0, _ptr(mask),
                        mask.stride(0) if mask is not None else 0, _ptr(ws), nbytes, _stream())
    _lib.check(rc, "gn_tc_gemm")

torch.manual_seed(0)
for (M, N, K, tb, ad, mk, scale) in [(200000, 128, 64, 1, 1, 0, 1.0), (200000, 256, 64, 1, 0, 0, 1.0), (150000, 128, 128, 1, 1, 0, 1.0),
                              (200000, 64, 128, 0, 0, 0, 1.0), (200000, 64, 64, 1, 1, 1, 1.0), (200000, 128, 64, 1, 1, 0, 1e-6),
                              (200000, 256, 64, 1, 0, 0, 1e-7)]:
    A = torch.randn(M, K, device=dev) * scale
    # sparse-ish A like a gradient: most rows tiny, few rows large
    A[::1000] *= 1e3
    B = torch.randn((N, K) if tb else (K, N), device=dev)
    addend = torch.randn(M, N, device=dev) * scale if ad else None
    mask = torch.randn(M, N, device=dev) if mk else None
    want = A.double() @ (B.double().t() if tb else B.double())
    if ad: want = want + addend.double()
    if mk: want = want * (mask > 0).double()
    outs = []
    for rep in range(3):
        C = torch.full((M, N), float("nan"), device=dev)
        tc(tb, A, B, C, addend, mask)
        torch.cuda.synchronize()
        outs.append(C)
    d = (outs[0].double() - want).abs()
    rowmax = want.abs().max(dim=1).values.clamp_min(1e-30)
    rowrel = (d.max(dim=1).values / rowmax)
    print(f"M={M} N={N} K={K} tb={tb} ad={ad} mk={mk} scale={scale}: rel_err={rel_err(outs[0], want):.3e} worst row-rel={float(rowrel.max()):.3e} "
          f"rows>1e-4: {int((rowrel > 1e-4).sum())} deterministic={torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])} nan={int(torch.isnan(outs[0]).sum())}")
    bad = (rowrel > 1e-4).nonzero().view(-1)[:10].tolist()
    if bad:
        print("   bad rows", bad, "A row absmax", [float(A[r].abs().max()) for r in bad[:5]])
