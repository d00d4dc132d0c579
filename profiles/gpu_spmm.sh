#!/bin/bash
# SpMM software-pipelining experiment (one B200): parity tests with the switch on, sweep and bench on/off.
tag=${1:-spmm}
out=gpurun_out
mkdir -p $out
GRIPNET_B200_SPMM_PIPE=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_full_size.py -m gpu -q -x > $out/${tag}_pytest_pipe.log 2>&1
echo "pytest(pipe) rc=$?"; tail -3 $out/${tag}_pytest_pipe.log
for v in 0 1; do
  GRIPNET_B200_SPMM_PIPE=$v timeout 120 python profiles/spmm_sweep.py --small > $out/${tag}_sweep_pipe$v.jsonl 2> $out/${tag}_sweep_pipe$v.err
  echo "sweep pipe=$v rc=$?"; python - <<PY
import json
for l in open("$out/${tag}_sweep_pipe$v.jsonl"):
    d = json.loads(l); print("pipe=$v", d["shape"], "us", round(d["us"], 1), "frac", round(d["frac_of_hbm_peak"], 3))
PY
done
for v in 0 1; do
  GRIPNET_B200_SPMM_PIPE=$v timeout 200 python bench.py --no-cpu-baseline --no-train-epoch > $out/${tag}_bench_pipe$v.json 2> $out/${tag}_bench_pipe$v.err
  echo "bench pipe=$v rc=$?"; python - <<PY
import json
d = json.load(open("$out/${tag}_bench_pipe$v.json"))
print("pipe=$v", "ms/step", d["ms_per_step"], "G edges/s", d["value"] / 1e9, "roofline frac", d["roofline"]["frac"], "us", d["roofline"]["us_per_launch"])
PY
done
