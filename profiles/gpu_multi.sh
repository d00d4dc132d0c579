#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): partitioned-path parity tests, then the bench at N ranks with the peer-memory
# all-gather (default) and with NCCL only, on the weak-scaled pose workload and on BASELINE config 5.
#   gpurun --gpus 2 --timeout 900 -- 'bash profiles/gpu_multi.sh r01_v7 2'
tag=${1:-multi}; n=${2:-2}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > $out/${tag}_pytest_dist_n$n.log 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_dist_n$n.log
run() {  # name, env, extra args
  env $2 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port 29611 bench.py --gpus $n $3 > $out/${tag}_bench_$1_n$n.json 2> $out/${tag}_bench_$1_n$n.err
  echo "bench[$1] rc=$?"; tail -c 1500 $out/${tag}_bench_$1_n$n.json | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'G edges/s', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3), d['config']['parallelism'][:70])
" 2>/dev/null || tail -5 $out/${tag}_bench_$1_n$n.err
}
run pose_peer "GRIPNET_B200_PEER=auto" "--steps 100"
run pose_nccl "GRIPNET_B200_PEER=off" "--steps 100"
run config5_peer "GRIPNET_B200_PEER=auto" "--workload scaled --steps 10"
ls -la $out | tail
