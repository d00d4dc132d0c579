#!/bin/bash
# One GPU pass (run under gpurun on one B200): GPU parity tests, both bench arms, the ncu launch list of
# the bench command, one `--set full` capture of the hot kernels, and the SpMM size sweep (events + ncu DRAM
# bytes).  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_pass.sh r01_v6'
tag=${1:-pass}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $out/${tag}_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $out/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"
cat $out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
cat $out/${tag}_bench_ref.json
if [ -z "$SKIP_SWEEP" ]; then
timeout 300 python profiles/spmm_sweep.py > $out/${tag}_spmm_sweep.jsonl 2> $out/${tag}_spmm_sweep.err
echo "sweep rc=$?"; cat $out/${tag}_spmm_sweep.jsonl
fi
# launch list of the same command (graph replay: ncu reports the kernel nodes one by one)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager \
    > $out/${tag}_launches_bench.log 2>&1
echo "ncu launches rc=$?"
# the same list with warm caches (no flush between kernels): closer to the in-graph durations
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1600 --csv \
    --log-file $out/${tag}_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager \
    > $out/${tag}_launches_warm_bench.log 2>&1
echo "ncu warm launches rc=$?"
# DRAM bytes of the SpMM at every sweep size (one launch each)
[ -z "$SKIP_SWEEP" ] && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:spmm_kernel --csv --log-file $out/${tag}_spmm_sweep_dram.csv \
    python profiles/spmm_sweep.py --once > $out/${tag}_spmm_sweep_ncu.log 2>&1
echo "ncu sweep rc=$?"
# full capture of the hot kernels (skip the warm-up launches; 1 step is ~70 launches)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'spmm_kernel|distmult|sgemm|tc_gemm|rs_scatter|adam_kernel|lp_metrics|neg_draw' -s 260 -c 60 -f -o $out/${tag}_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager > $out/${tag}_full_bench.log 2>&1
echo "ncu full rc=$?"
ls -la $out
