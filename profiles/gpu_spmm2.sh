#!/bin/bash
# SpMM experiment 2 (one B200): row-is-chunk fast path and software pipelining at 6 CTAs/SM.
tag=${1:-spmm2}
out=gpurun_out
mkdir -p $out
GRIPNET_B200_SPMM_PIPE=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x > $out/${tag}_pytest_pipe.log 2>&1
echo "pytest(rowchunk+pipe) rc=$?"; tail -2 $out/${tag}_pytest_pipe.log
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  GRIPNET_B200_ROW_IS_CHUNK=$1 GRIPNET_B200_SPMM_PIPE=$2 timeout 120 python profiles/spmm_sweep.py --small > $out/${tag}_sweep_$1$2.jsonl 2> $out/${tag}_sweep_$1$2.err
  python - <<PY
import json
print("rowchunk=$1 pipe=$2", [(round(json.loads(l)["us"], 1), round(json.loads(l)["frac_of_hbm_peak"], 3)) for l in open("$out/${tag}_sweep_$1$2.jsonl")])
PY
  GRIPNET_B200_ROW_IS_CHUNK=$1 GRIPNET_B200_SPMM_PIPE=$2 timeout 200 python bench.py --no-cpu-baseline --no-train-epoch > $out/${tag}_bench_$1$2.json 2> $out/${tag}_bench_$1$2.err
  python - <<PY
import json
d = json.load(open("$out/${tag}_bench_$1$2.json"))
print("rowchunk=$1 pipe=$2", "ms/step", round(d["ms_per_step"], 4), "G edges/s", round(d["value"] / 1e9, 3), "roofline frac", round(d["roofline"]["frac"], 3), "us", round(d["roofline"]["us_per_launch"], 2))
PY
done
