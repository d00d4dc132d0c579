#!/bin/bash
# kernel timelines (torch.profiler / CUPTI) of the captured step: bash profiles/gpu_r02_timeline.sh <tag> <n_gpus> <workload...>
tag=${1:-r02_tl}
n=${2:-1}
shift 2
out=gpurun_out
mkdir -p $out
for wl in "$@"; do
  if [ "$n" == "1" ]; then
    timeout 600 python profiles/timeline.py --workload $wl --tag ${tag}_${wl}_n1 $TIMELINE_ARGS > $out/${tag}_${wl}_n1.log 2>&1
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29547 \
      profiles/timeline.py --workload $wl --tag ${tag}_${wl}_n${n} $TIMELINE_ARGS > $out/${tag}_${wl}_n${n}.log 2>&1
  fi
  echo "timeline $wl n=$n rc=$?"; grep -v Warning $out/${tag}_${wl}_n${n}.log | tail -45
done
