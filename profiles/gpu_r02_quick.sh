#!/bin/bash
# Round-2 quick GPU check (one B200): GPU parity tests (all failures listed) + the default bench line without the
# config-5 leg + config 5 alone at N=1.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_r02_quick.sh r02_v1'
tag=${1:-r02_quick}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 $out/${tag}_pytest.log
timeout 300 python bench.py --steps 100 --no-config5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; cat $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
if [ "$2" == "c5" ]; then
  timeout 400 python bench.py --workload scaled --steps 5 --no-cpu-baseline > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
  echo "bench c5 rc=$?"; cat $out/${tag}_bench_c5.json; tail -3 $out/${tag}_bench_c5.err
fi
