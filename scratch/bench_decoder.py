import sys, os
sys.path.insert(0, ".")
import torch, numpy as np
from gripnet_b200 import ops, graph as G
d = torch.device("cuda:0")
rs = np.random.RandomState(0)
n, D, r, e = 645, 80, 16, 400_000
z = (torch.randn(n, D) * 0.3).to(d); w = torch.randn(r, D).to(d)
ei = torch.from_numpy(rs.randint(0, n, (2, e))).to(d)
et = torch.from_numpy(np.sort(rs.randint(0, r, e))).to(d)
g = torch.randn(e, device=d)
out = ops._distmult_fwd(z, w, ei, et, True)
coef = ops._distmult_coef(g, out, True)
dw = torch.empty_like(w)
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
print("decoder path", ops.DECODER_PATH)
print("fwd us", timeit(lambda: ops._distmult_fwd(z, w, ei, et, True)))
print("dz  us", timeit(lambda: ops._distmult_dz((ei, et), coef, z, w)))
print("dw  us", timeit(lambda: ops._distmult_dw(dw, et, ei, coef, z)))
if ops.DECODER_PATH == "auto":
    # the two families agree to fp32 round-off
    ops.DECODER_PATH = "global"
    o2 = ops._distmult_fwd(z, w, ei, et, True); dz2 = ops._distmult_dz((ei, et), coef, z, w)
    dw2 = torch.empty_like(w); ops._distmult_dw(dw2, et, ei, coef, z)
    ops.DECODER_PATH = "auto"
    dz1 = ops._distmult_dz((ei, et), coef, z, w); ops._distmult_dw(dw, et, ei, coef, z)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print("resident vs global rel diff: fwd %.2e dz %.2e dw %.2e" % (rel(out, o2), rel(dz1, dz2), rel(dw, dw2)))
