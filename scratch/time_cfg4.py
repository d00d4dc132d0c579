import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle import port, synth
t0 = time.time()
g = synth.pose2_graph()
p = synth.pose_params(g)
print("synth", time.time() - t0, g["dd_edge_index"].shape, g["n_rel"]); t0 = time.time()
pl = {k: v.clone().requires_grad_(True) for k, v in p.items()}
ref = port.pose_forward(pl, g)
print("fwd", time.time() - t0); t0 = time.time()
ref[0].backward()
print("bwd", time.time() - t0)
