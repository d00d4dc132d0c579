#!/bin/bash
# Quick GPU check (one B200): GPU parity tests + the default bench line, and the same bench without side streams.
#   gpurun --timeout 900 -- 'bash profiles/gpu_quick.sh r01_v5'
tag=${1:-quick}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $out/${tag}_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; cat $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
GRIPNET_B200_STREAMS=0 timeout 300 python bench.py --no-cpu-baseline > $out/${tag}_bench_nostreams.json 2>> $out/${tag}_bench.err
echo "bench(no streams) rc=$?"; cat $out/${tag}_bench_nostreams.json
