#!/usr/bin/env python
"""SpMM roofline across operand sizes (one B200): the GCN SpMM (gn_spmm over the dst-CSR built by gn_gcn_prep)
timed alone with CUDA events, L2 flushed before every launch, on graphs whose gathered operand ranges from
L2-resident (pose-0) to far beyond the 126 MB L2 (freebase-d-shaped, config-5-shaped).  One JSON line per
shape: algorithmic bytes (E'(8+4F) + N(8+4F)), achieved GB/s, fraction of the measured HBM peak.

    python profiles/spmm_sweep.py > gpurun_out/spmm_sweep.jsonl
    ncu --set full -k regex:spmm_kernel ... python profiles/spmm_sweep.py --once    # dram__bytes per shape
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

SHAPES = [  # name, nodes, directed edges, F, generator
    ("pose-0 gg (uniform)", 19081, 1431224, 16, "uniform"),
    ("aminer pp (uniform)", 200000, 2000000, 64, "uniform"),
    ("freebase-d pp (uniform)", 300000, 3000000, 128, "uniform"),
    ("1M nodes deg 32 (uniform)", 1 << 20, 32 << 20, 64, "uniform"),
    ("config-5 aa slice (R-MAT, 4M nodes, 100M edges)", 1 << 22, 100_000_000, 64, "rmat"),
]


def main():
    once = "--once" in sys.argv
    from gripnet_b200 import graph as G, ops
    from gripnet_b200.synthetic import rmat_edges
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1111)
    shapes = SHAPES[:4] if "--small" in sys.argv else SHAPES
    for name, n, e, F, kind in shapes:
        if kind == "uniform":
            ei = torch.randint(0, n, (2, e), device=dev, generator=gen)
        else:
            ei = rmat_edges(22, e, dev, gen)
        g = G.gcn_graph(ei, n, n)
        csr = g.fwd
        x = torch.randn(n, F, device=dev)
        out = torch.empty(n, F, device=dev)
        bias = torch.zeros(F, device=dev)

        def run():
            ops.spmm(csr, ops.M(x), ops.M(out), F, bias=bias, relu=True)

        reps = 1 if once else 20
        for _ in range(0 if once else 3):
            run()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        nnz = csr.nnz
        alg = nnz * (8 + 4 * F) + n * (8 + 4 * F)
        # compulsory DRAM bytes when every operand row is fetched once (perfect reuse): the lower bound
        compulsory = nnz * 8 + n * 4 * F + n * (8 + 4 * F)
        gbs = alg / (ms * 1e-3) / 1e9
        print(json.dumps({"shape": name, "nodes": n, "entries": nnz, "F": F, "operand_MB": n * F * 4 / 1e6,
                          "chunk_len": csr.chunk_len, "us": ms * 1e3, "algorithmic_bytes": alg,
                          "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"],
                          "compulsory_bytes": compulsory,
                          "compulsory_GBps": compulsory / (ms * 1e-3) / 1e9, "hbm_peak_GBps": peaks["hbm_gbs"]}),
              flush=True)
        del g, csr, x, out, ei
        G.clear_cache()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
