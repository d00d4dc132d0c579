#!/bin/bash
# one development iteration on ONE B200: parity tests, then pose bench lines under a list of environment variants
# (A/B switches), one pose2 line, and the kernel timeline of the captured pose step.
#   gpurun --timeout 900 -- 'bash profiles/gpu_r02_iter.sh r02_v19 "GRIPNET_B200_PRIORITIES=0" "GRIPNET_B200_PROLOGUE=0"'
tag=${1:-r02_iter}; shift
out=gpurun_out
mkdir -p $out
timeout 800 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $out/${tag}_pytest.log
line() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], round(d['ms_per_step'],4), '%.4g' % d['value'], d['launches_per_step'], 'e2e', round(d['e2e']['ms_per_step'],4), [(r['kernel'][:18], round(r['us_per_launch'],1)) for r in d.get('roofline_kernels',[])][:4])" $1 "$2"; }
i=0
for variant in "" "$@"; do
  name=v$i; i=$((i+1))
  env $variant timeout 300 python bench.py --workload pose --steps 200 --no-config5 --no-train-epoch --no-cpu-baseline > $out/${tag}_bench_pose_${name}.json 2> $out/${tag}_bench_pose_${name}.err
  echo "bench pose [$variant] rc=$?"; line $out/${tag}_bench_pose_${name}.json "pose[$variant]"
done
timeout 300 python bench.py --workload pose2 --steps 50 --no-config5 --no-train-epoch --no-cpu-baseline > $out/${tag}_bench_pose2.json 2> $out/${tag}_bench_pose2.err
echo "bench pose2 rc=$?"; line $out/${tag}_bench_pose2.json pose2
bash profiles/gpu_r02_timeline.sh $tag 1 pose > $out/${tag}_timeline.log 2>&1
grep -E "step span" $out/${tag}_timeline.log | head -3
