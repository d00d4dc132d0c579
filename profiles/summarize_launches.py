"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step
(the launches between two L2-flush fills of bench.py) aggregated per kernel.

    python profiles/summarize_launches.py profiles/<launches>.csv [step_index]
"""
import collections
import csv
import re
import sys


def main(path, which=0):
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    name = lambda r: re.sub(r"\(.*", "", r["Kernel Name"]).replace("gn::", "").replace("void ", "")[:72]
    seq = [(name(r), float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Kernel Name"]) for r in rows]
    flush = [i for i, s in enumerate(seq) if "FillFunctor<unsigned char>" in s[3]]
    a, b = flush[which], flush[which + 1]
    step = seq[a + 1:b]
    agg = collections.OrderedDict()
    for n, t, g, _ in step:
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: step {which}: {len(step)} launches, {tot:.1f} us summed (cold-cache, serialised: compare SHARES)")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:9.1f} us {100 * t / tot:5.1f}%  x{c:<3d} {n}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
