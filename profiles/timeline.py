#!/usr/bin/env python
"""Kernel timeline of ONE captured training step (what `nsys` would show; nsys is not in this image).

    python profiles/timeline.py --workload pose --tag r02_v3_pose_n1
    python -m torch.distributed.run --nproc-per-node 8 ... profiles/timeline.py --workload pose --tag r02_v3_pose_n8

Builds the bench workload exactly as bench.py does (same builders, same CUDA-graph capture), replays the step
under torch.profiler (CUPTI activity records: kernel start / end timestamps per stream, also for kernels launched
from a CUDA graph) and writes, per rank:

    gpurun_out/<tag>_rank<r>_kernels.csv   every kernel of one replay: name, stream, start_us (from the step's first
                                           kernel), dur_us
    gpurun_out/<tag>_rank<r>_summary.txt   step span, union of busy time, idle time, time per kernel name, and the
                                           serial chain = kernels on the longest-running stream

The numbers are taken under a profiler: they attribute the step, they are not bench values.
"""
import argparse
import csv
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="pose")
    ap.add_argument("--tag", default="timeline")
    ap.add_argument("--replays", type=int, default=4)
    ap.add_argument("--no-flush", action="store_true", help="replay with a warm L2 (a 1-byte fill marks the replays)")
    args = ap.parse_args()
    import bench
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dctx = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from gripnet_b200.parallel import DistContext
        dctx = DistContext(defer_grad_reduce=True)
    bargs = argparse.Namespace(workload=args.workload, warmup=3, eager=False, steps=10)
    w = bench.build_workload(bargs, args.workload, world, rank, dev, dctx)
    step, execution = bench.capture_step(w, dctx, bargs, world, rank)
    flush = torch.empty(2 * bench.L2_BYTES, dtype=torch.uint8, device=dev)
    for _ in range(5):
        step.replay()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.replays):
            (flush[:1] if args.no_flush else flush).zero_()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            step.replay()
            torch.cuda.synchronize()
    import json
    import tempfile
    trace = os.path.join(tempfile.gettempdir(), f"gripnet_trace_{os.getpid()}.json")
    prof.export_chrome_trace(trace)
    with open(trace) as f:
        tr = json.load(f)
    os.remove(trace)
    ks = []
    for e in tr.get("traceEvents", []):
        cat = str(e.get("cat", "")).lower()
        if e.get("ph") != "X" or cat not in ("kernel", "gpu_memset", "gpu_memcpy"):
            continue
        s0 = float(e["ts"])
        ks.append((s0, s0 + float(e.get("dur", 0.0)), e.get("name", "?"), e.get("args", {}).get("stream", -1),
                   "kernel" if cat == "kernel" else "mem"))
    ks.sort()
    # split into replays at the L2 flush (the 252 MiB fill that precedes every replay)
    groups, cur = [], []
    for k in ks:
        if "FillFunctor<unsigned char>" in k[2] and (args.no_flush or k[1] - k[0] > 20):
            if cur:
                groups.append(cur)
            cur = []
            continue
        cur.append(k)
    if cur:
        groups.append(cur)
    groups = [g for g in groups if len(g) >= max(10, step.launches_per_replay // 2)]
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    if not groups:
        open(os.path.join(out_dir, f"{args.tag}_rank{rank}_summary.txt"), "w").write("no kernel activity records\n")
        return
    g = groups[-1]
    t0 = min(k[0] for k in g)
    with open(os.path.join(out_dir, f"{args.tag}_rank{rank}_kernels.csv"), "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["start_us", "dur_us", "stream", "name"])
        for s, e, name, stream, _ in g:
            wr.writerow([f"{s - t0:.2f}", f"{e - s:.2f}", stream, name[:160]])
    span = max(k[1] for k in g) - t0
    # union of busy intervals
    busy, last_end = 0.0, None
    for s, e, *_ in sorted(g):
        if last_end is None or s > last_end:
            busy += e - s
            last_end = e
        elif e > last_end:
            busy += e - last_end
            last_end = e
    per = defaultdict(lambda: [0, 0.0])
    for s, e, name, *_ in g:
        short = name.split("(")[0].split("<")[0].replace("void ", "").replace("gn::", "")
        per[short][0] += 1
        per[short][1] += e - s
    streams = defaultdict(float)
    for s, e, _, stream, _ in g:
        streams[stream] += e - s
    lines = [f"workload {args.workload}  world {world}  rank {rank}  {execution}",
             f"kernels in the step: {len(g)} (library launch count {step.launches_per_replay})",
             f"step span {span:.1f} us   busy (union over streams) {busy:.1f} us   idle {span - busy:.1f} us   "
             f"sum of kernel time {sum(e - s for s, e, *_ in g):.1f} us",
             "time per stream: " + ", ".join(f"{k}: {v:.1f}" for k, v in sorted(streams.items(), key=lambda kv: -kv[1])),
             "", f"{'kernel':60s} {'n':>4s} {'total_us':>10s} {'share_of_span':>14s}"]
    for name, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{name[:60]:60s} {n:4d} {t:10.1f} {t / span:14.3f}")
    open(os.path.join(out_dir, f"{args.tag}_rank{rank}_summary.txt"), "w").write("\n".join(lines) + "\n")
    # the step in launch order: start, duration, stream, kernel (the critical path is read off this list)
    lines.append("")
    lines.append("timeline (us from the first kernel of the step): start  dur  end  stream  kernel")
    for s0, e0, name, stream, _ in g:
        short = name.replace("void ", "").replace("gn::", "").split("(")[0][:70]
        lines.append(f"{s0 - t0:9.1f} {e0 - s0:8.1f} {e0 - t0:9.1f}  st{stream}  {short}")
    open(os.path.join(out_dir, f"{args.tag}_rank{rank}_summary.txt"), "w").write("\n".join(lines) + "\n")
    if rank == 0:
        print("\n".join(lines[:36]))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
