"""Decoder micro-benchmark (pose-0 size: 645 x 80 table, 16 relations, 400 k relation-major edges): warm,
back-to-back launches timed with CUDA events, and the results checked against a float64 torch evaluation."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from gripnet_b200 import ops

d = torch.device("cuda:0")
rs = np.random.RandomState(0)
n, D, r, e = 645, 80, 16, 400_000
z = (torch.randn(n, D) * 0.3).to(d)
w = torch.randn(r, D).to(d)
ei = torch.from_numpy(rs.randint(0, n, (2, e))).to(d)
et = torch.from_numpy(np.sort(rs.randint(0, r, e))).to(d)
g = torch.randn(e, device=d)
out = ops._distmult_fwd(z, w, ei, et, True)
coef = ops._distmult_coef(g, out, True)
dw = torch.empty_like(w)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


print("env", {k: v for k, v in os.environ.items() if k.startswith("GRIPNET_B200")}, "path", ops.DECODER_PATH)
print("fwd us %.1f  dz us %.1f  dw us %.1f" % (timeit(lambda: ops._distmult_fwd(z, w, ei, et, True)),
                                                timeit(lambda: ops._distmult_dz((ei, et), coef, z, w)),
                                                timeit(lambda: ops._distmult_dw(dw, et, ei, coef, z))))
z64, w64, c64 = z.double(), w.double(), coef.double()
s64 = torch.sigmoid((z64[ei[0]] * z64[ei[1]] * w64[et]).sum(1))
dz64 = torch.zeros_like(z64).index_add_(0, ei[0], c64[:, None] * z64[ei[1]] * w64[et]) \
    .index_add_(0, ei[1], c64[:, None] * z64[ei[0]] * w64[et])
dw64 = torch.zeros_like(w64).index_add_(0, et, c64[:, None] * z64[ei[0]] * z64[ei[1]])
rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
dz = ops._distmult_dz((ei, et), coef, z, w)
ops._distmult_dw(dw, et, ei, coef, z)
print("rel err vs float64: fwd %.2e dz %.2e dw %.2e" % (rel(out, s64), rel(dz, dz64), rel(dw, dw64)))
