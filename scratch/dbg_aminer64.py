import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
from golden_util import rel_err
from oracle import port, synth
from gripnet_b200 import ops, graph as G
from gripnet_b200.pipelines import AminerModel, load_flat_params, to_device
dev = torch.device("cuda:0")
g = synth.aminer_full(); p = synth.aminer_params(g)
t0 = time.time()
pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
ref = port.aminer_forward(pl, g); ref[0].backward()
t1 = time.time()
pd = {k: v.double().requires_grad_(True) for k, v in p.items()}
ref64 = port.aminer_forward(pd, g); ref64[0].backward()
print("port32 s", t1 - t0, "port64 s", time.time() - t1)
print("port32 vs port64: loss", rel_err(ref[0], ref64[0]), "z", rel_err(ref[1], ref64[1]))
for k in pl:
    if pl[k].grad is not None:
        print("   ", k, f"{rel_err(pl[k].grad, pd[k].grad):.3e}")
for path in ("ffma", "auto"):
    ops.GEMM_PATH = path
    G.clear_cache()
    m = load_flat_params(AminerModel(g["n_p"], g["n_a"], g["n_class"]), p).to(dev)
    out = m(to_device(g, dev)); out[0].backward(); torch.cuda.synchronize()
    print(path, "vs port64: loss", rel_err(out[0], ref64[0]), "z", rel_err(out[1], ref64[1]))
    for k, v in m.named_parameters():
        if v.grad is not None:
            print("   ", k, f"{rel_err(v.grad, pd[k].grad):.3e}")
