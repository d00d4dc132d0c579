#!/bin/bash
# A/B pass (one B200): GPU tests, decoder micro-benchmark and whole-step bench with the shared-memory-resident
# decoder (auto) against the global-memory decoder, launch list + full ncu capture of the training-epoch kernels.
#   gpurun --timeout 900 -- 'bash profiles/gpu_ab.sh r01_v7'
tag=${1:-ab}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
for path in auto global; do
  GRIPNET_B200_DECODER=$path timeout 120 python scratch/bench_decoder.py > $out/${tag}_decoder_$path.txt 2>&1
  echo "decoder[$path] rc=$?"; cat $out/${tag}_decoder_$path.txt
done
for path in auto global; do
  GRIPNET_B200_DECODER=$path timeout 300 python bench.py --no-cpu-baseline > $out/${tag}_bench_$path.json 2> $out/${tag}_bench_$path.err
  echo "bench[$path] rc=$?"; python - <<PY
import json
d = json.load(open("$out/${tag}_bench_$path.json"))
print("$path", "ms/step", d["ms_per_step"], "G edges/s", d["value"] / 1e9, "e2e", d["e2e"]["value"] / 1e9, "train", d.get("train_epoch", {}).get("ms_per_epoch"))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager \
    > $out/${tag}_launches_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'res_kernel|adam_kernel|lp_metrics_kernel|negsample|rs_scatter|rs_histogram|lp_rank' -s 60 -c 36 -f -o $out/${tag}_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager > $out/${tag}_full_bench.log 2>&1
echo "ncu full rc=$?"
ls -la $out | tail -20
