#!/bin/bash
# Lean evidence pass (one B200): GPU tests, cold- and warm-cache launch lists of the bench command, and a
# `--set full` capture small enough for the 64 MiB return limit.
#   gpurun --timeout 700 -- 'bash profiles/gpu_final.sh r01_v9'
tag=${1:-final}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager \
    > $out/${tag}_launches_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1600 --csv \
    --log-file $out/${tag}_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager \
    > $out/${tag}_launches_warm_bench.log 2>&1
echo "ncu warm launches rc=$?"
timeout 400 ncu --set full --clock-control none \
    -k regex:'spmm_kernel|distmult|tc_gemm|adam_kernel|lp_metrics_kernel|neg_draw' -s 330 -c 30 -f -o $out/${tag}_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager > $out/${tag}_full_bench.log 2>&1
echo "ncu full rc=$?"
du -sh $out; ls -la $out
