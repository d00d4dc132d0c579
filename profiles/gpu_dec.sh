#!/bin/bash
# Decoder experiment pass (one B200): micro-benchmark under every switch, then the whole step for the best ones.
tag=${1:-dec}
out=gpurun_out
mkdir -p $out
i=0
for envs in "X=1" "GRIPNET_B200_DW_IDENTITY=0" "GRIPNET_B200_DECODER_LPE8=z" "GRIPNET_B200_DECODER_LPE8=w" "GRIPNET_B200_DECODER_LPE8=zw"; do
  env $envs timeout 120 python scratch/bench_decoder.py >> $out/${tag}_decoder.txt 2>&1
  echo "rc=$?" >> $out/${tag}_decoder.txt
done
cat $out/${tag}_decoder.txt
for envs in "X=1" "GRIPNET_B200_DECODER_LPE8=zw" "GRIPNET_B200_DECODER_LPE8=z"; do
  i=$((i+1))
  env $envs timeout 200 python bench.py --no-cpu-baseline --no-train-epoch > $out/${tag}_bench_$i.json 2> $out/${tag}_bench_$i.err
  echo "bench[$envs] rc=$?"; python - <<PY
import json
d = json.load(open("$out/${tag}_bench_$i.json"))
print("$envs", "ms/step", d["ms_per_step"], "G edges/s", d["value"] / 1e9, "e2e ms", d["e2e"]["ms_per_step"], "launches", d["launches_per_step"])
PY
done
