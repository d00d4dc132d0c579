#!/bin/bash
# Round-2 multi-GPU check: multi-rank parity worker + the driver's bench command at N ranks.
#   gpurun --gpus 2 --timeout 1200 -- 'bash profiles/gpu_r02_multi.sh r02_v2 2'
tag=${1:-r02_multi}
n=${2:-2}
steps=${3:-100}
out=gpurun_out
mkdir -p $out
export MASTER_ADDR=127.0.0.1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
  tests/dist_worker.py > $out/${tag}_pytest_dist_n${n}.txt 2>&1
echo "dist_worker rc=$?"; tail -15 $out/${tag}_pytest_dist_n${n}.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --gpus $n --steps $steps --warmup 3 $4 > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err
echo "bench rc=$?"; cat $out/${tag}_bench_n${n}.json; tail -5 $out/${tag}_bench_n${n}.err
