import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
from golden_util import rel_err
from oracle import port, synth
from gripnet_b200 import ops, graph as G
from gripnet_b200.pipelines import AminerModel, load_flat_params, to_device
dev = torch.device("cuda:0")
g = synth.aminer_full(); p = synth.aminer_params(g)
pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
ref = port.aminer_forward(pl, g); ref[0].backward()
for path in ("ffma", "auto", "ffma", "auto"):
    ops.GEMM_PATH = path
    G.clear_cache()
    m = load_flat_params(AminerModel(g["n_p"], g["n_a"], g["n_class"]), p).to(dev)
    out = m(to_device(g, dev)); out[0].backward(); torch.cuda.synchronize()
    print(path, "loss", rel_err(out[0], ref[0]), "z", rel_err(out[1], ref[1]))
    for k, v in m.named_parameters():
        if v.grad is not None:
            e = rel_err(v.grad, pl[k].grad)
            print("   ", k, f"{e:.3e}")
            if e > 1e-4:
                d = (v.grad.cpu().double() - pl[k].grad.double()).abs()
                rows = (d.max(dim=1).values > 1e-4 * pl[k].grad.abs().max()).nonzero().view(-1) if d.dim() == 2 else None
                if rows is not None:
                    print("      bad rows:", rows.numel(), rows[:20].tolist(), "cols:", (d.max(dim=0).values > 1e-4 * pl[k].grad.abs().max()).nonzero().view(-1)[:40].tolist())
