"""Numerical check (float64, CPU) of the dense-relation decoder identities noted in DESIGN.md §9a:
    C_r[n,m] = sum of coef_e over edges e of relation r with (src,dst) = (n,m) or (m,n)   (both directions)
    T_r = C_r z
    dz  = sum_r T_r * w_r            dw_r = 1/2 sum_n z_n * T_r[n]          score_e = S_r[src,dst], S_r = (z*w_r) z^T
against autograd of the gather-multiply-reduce form (gripnet/decoder.py:19-23)."""
import numpy as np
import torch

rs = np.random.RandomState(0)
n, D, R, E = 60, 12, 4, 3000
z = torch.randn(n, D, dtype=torch.float64, requires_grad=True)
w = torch.randn(R, D, dtype=torch.float64, requires_grad=True)
ei = torch.from_numpy(rs.randint(0, n, (2, E)))
et = torch.from_numpy(np.sort(rs.randint(0, R, E)))
g = torch.randn(E, dtype=torch.float64)
s = (z[ei[0]] * z[ei[1]] * w[et]).sum(1)
(s * g).sum().backward()                      # coef_e = g_e (no sigmoid)

zd, wd = z.detach(), w.detach()
C = torch.zeros(R, n, n, dtype=torch.float64)
C.index_put_((et, ei[0], ei[1]), g, accumulate=True)
C = C + C.transpose(1, 2)
T = torch.einsum("rnm,md->rnd", C, zd)
dz = (T * wd[:, None, :]).sum(0)
dw = 0.5 * (zd[None] * T).sum(1)
S = torch.einsum("nd,rd,md->rnm", zd, wd, zd)
print("score", float((S[et, ei[0], ei[1]] - s.detach()).abs().max()))
print("dz   ", float((dz - z.grad).abs().max() / z.grad.abs().max()))
print("dw   ", float((dw - w.grad).abs().max() / w.grad.abs().max()))
