#!/bin/bash
# Final single-GPU pass of round 2: the whole GPU suite, the default bench line (headline + e2e + roofline +
# cpu_baseline + train_epoch + config 5 at N = 1), then the per-workload evidence pass (bench lines, ncu launch
# lists, ncu --set full summaries) and the kernel timelines of configs 1 and 5.
#   gpurun --timeout 2700 -- 'bash profiles/gpu_r02_final.sh'
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/r02_final_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -3 $out/r02_final_pytest_gpu.txt
timeout 900 python bench.py > $out/r02_final_bench_n1.json 2> $out/r02_final_bench_n1.err
echo "default bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/r02_final_bench_reference_arm.json 2> $out/r02_final_bench_reference_arm.err
echo "reference arm rc=$?"
bash profiles/gpu_r02_evidence.sh r02_final
bash profiles/gpu_r02_timeline.sh r02_final 1 pose scaled > $out/r02_final_timelines.log 2>&1
grep "step span" $out/r02_final_timelines.log
