"""profiles/traffic.json from the committed `ncu --set full` summaries: measured DRAM bytes (read + written) per launch
of each kernel, the `roofline.traffic` of bench.py.

    python profiles/extract_traffic.py pose=profiles/r02_v6_ncu_full_pose.csv aminer=profiles/r02_v6_ncu_full_aminer.csv ...

Kernel key = the function name without namespace / template arguments (what bench.py's probes are named after).
Only launches of the step's own shapes count: for every kernel the launches are grouped by grid size and the group
with the largest total duration is taken (the probe kernel of bench.py is the dominant shape)."""
import csv
import json
import os
import sys
from collections import defaultdict


def load(path):
    rows = list(csv.reader(open(path)))
    head, body = rows[0], rows[2:]
    ix = {c: head.index(c) for c in ("Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                     "dram__bytes_write.sum") if c in head}
    units = dict(zip(head, rows[1]))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    groups = defaultdict(lambda: defaultdict(list))
    for r in body:
        name = r[ix["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").split("::")[-1]
        rd = float(r[ix["dram__bytes_read.sum"]].replace(",", "")) * scale.get(units["dram__bytes_read.sum"], 1)
        wr = float(r[ix["dram__bytes_write.sum"]].replace(",", "")) * scale.get(units["dram__bytes_write.sum"], 1)
        dur = float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
        groups[name][r[ix["Grid Size"]]].append((dur, rd + wr))
    out = {}
    for name, by_grid in groups.items():
        best = max(by_grid.values(), key=lambda v: sum(d for d, _ in v))
        out[name] = int(sum(b for _, b in best) / len(best))
    return out


def main():
    res = {}
    for arg in sys.argv[1:]:
        wl, path = arg.split("=", 1)
        res[wl] = load(path)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    json.dump(res, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
