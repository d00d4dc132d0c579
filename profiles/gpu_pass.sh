#!/bin/bash
# One GPU pass (run under gpurun on one B200): GPU parity tests, both bench arms, the ncu launch list of
# the bench command, and one `--set full` capture of the hot kernels.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_pass.sh r01_v4'
tag=${1:-pass}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $out/${tag}_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $out/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"
cat $out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
cat $out/${tag}_bench_ref.json
# launch list of the same command (graph replay: ncu reports the kernel nodes one by one)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager \
    > $out/${tag}_launches_bench.log 2>&1
echo "ncu launches rc=$?"
# full capture of the hot kernels (skip the warm-up launches; 1 step is ~64 launches)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'spmm_kernel|distmult|sgemm|tc_gemm|rs_scatter' -s 200 -c 40 -f -o $out/${tag}_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager > $out/${tag}_full_bench.log 2>&1
echo "ncu full rc=$?"
ls -la $out
