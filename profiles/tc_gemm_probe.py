#!/usr/bin/env python
"""tcgen05 dense transforms alone: C[M,N] = A[M,K] W[K,N] at the shapes of the configs, CUDA events, cold L2.
    python profiles/tc_gemm_probe.py            # one JSON line per shape (us, GB/s of 4(MK+MN) bytes)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from gripnet_b200 import ops
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = [(4_000_000, 64, 64), (4_000_000, 192, 64), (4_000_000, 256, 64), (4_000_000, 64, 256),
              (200_000, 128, 64), (19_081, 32, 16), (19_081, 64, 16)]
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        shapes = shapes[:1]
    for m, k, n in shapes:
        a = torch.randn(m, k, device=dev)
        w = torch.randn(k, n, device=dev)
        c = torch.empty(m, n, device=dev)
        fn = lambda: ops.sgemm(False, False, m, n, k, a.data_ptr(), k, w.data_ptr(), n, c.data_ptr(), n, dev)
        for _ in range(2):
            fn()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts)[len(ts) // 2]
        nbytes = 4 * (m * k + m * n)
        print(json.dumps({"M": m, "K": k, "N": n, "us": round(us, 1), "GBps": round(nbytes / us / 1e3, 1)}), flush=True)
        del a, w, c


if __name__ == "__main__":
    main()
