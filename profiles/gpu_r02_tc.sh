#!/bin/bash
# tc_gemm loader A/B on ONE B200: tensor-core GEMM tests, then configs 2 / 3 / 5 with the deep-prefetch loader
# (default) and with the shallow one everywhere (GRIPNET_B200_TC_PREFETCH=1).
tag=${1:-r02_tc}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py -q > $out/${tag}_pytest_tc.log 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_tc.log
for wl in aminer freebase-d scaled; do
  for v in deep shallow; do
    [ "$v" == "shallow" ] && export GRIPNET_B200_TC_PREFETCH=1 || unset GRIPNET_B200_TC_PREFETCH
    steps=50; [ "$wl" == "scaled" ] && steps=5
    timeout 500 python bench.py --workload $wl --steps $steps --no-cpu-baseline --no-config5 --no-train-epoch > $out/${tag}_bench_${wl}_${v}.json 2> $out/${tag}_bench_${wl}_${v}.err
    echo "bench $wl $v rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_bench_${wl}_${v}.json')); print(round(d['ms_per_step'],4), '%.4g' % d['value'], d['launches_per_step'], [(r['kernel'][:16], round(r['us_per_launch'],1), round(r['frac'],2)) for r in d.get('roofline_kernels',[])][:5])"
  done
done
