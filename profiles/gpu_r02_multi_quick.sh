#!/bin/bash
# bench (pose only) + timeline at N ranks:  bash profiles/gpu_r02_multi_quick.sh <tag> <n>
tag=${1:-r02_mq}; n=${2:-8}; out=gpurun_out; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --gpus $n --steps 100 --warmup 3 --no-config5 $3 > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err
echo "bench rc=$?"; python - <<PY
import json
d=json.loads([l for l in open("$out/${tag}_bench_n${n}.json") if l.startswith("{")][0])
print("ms/step", d["ms_per_step"], "value", d["value"], "launches", d["launches_per_step"], d.get("exchange"), d.get("loss_check",{}).get("rel_err"))
PY
bash profiles/gpu_r02_timeline.sh $tag $n pose | grep -E "peer_|slot_sum|grads|walk|step span" 
