#!/bin/bash
# Round-2 evidence pass on ONE B200: per-workload bench lines (configs 1-4), ncu launch lists of the same commands,
# and one `ncu --set full` capture of the hot kernels per workload (incl. the tcgen05 X.W_r transform of config 4 with
# its tensor-pipe counters).  Numbers printed under ncu are never bench values.
#   gpurun --timeout 2400 -- 'bash profiles/gpu_r02_evidence.sh r02_v6'
tag=${1:-r02_ev}
out=gpurun_out
mkdir -p $out
for wl in pose aminer freebase-d pose2; do
  extra="--no-config5 --no-train-epoch"
  [ "$wl" == "pose" ] || extra="$extra --no-cpu-baseline"
  timeout 600 python bench.py --workload $wl --steps 50 $extra > $out/${tag}_bench_${wl}.json 2> $out/${tag}_bench_${wl}.err
  echo "bench $wl rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_bench_${wl}.json')); print(d['ms_per_step'], d['value'], d['launches_per_step'], d['roofline']['kernel'][:40], d['roofline']['frac'])"
  # launch list of the same command, warm caches (closer to the in-graph durations)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 2500 --csv \
      --log-file $out/${tag}_launches_${wl}.csv python bench.py --workload $wl --steps 2 --warmup 1 --eager \
      --no-cpu-baseline --no-config5 --no-train-epoch > $out/${tag}_launches_${wl}.log 2>&1
  echo "ncu launches $wl rc=$?"
done
# config 4 with the relation transform on the CUDA cores instead (A/B of north_star item 3)
GRIPNET_B200_GEMM=ffma timeout 600 python bench.py --workload pose2 --steps 50 --no-cpu-baseline --no-config5 --no-train-epoch \
  > $out/${tag}_bench_pose2_ffma.json 2> $out/${tag}_bench_pose2_ffma.err
echo "bench pose2 ffma rc=$?"
# full captures (one launch of each hot kernel per workload; -s skips the graph-prep and warm-up launches).
# The .ncu-rep files stay on the box (gpurun returns at most 64 MiB): only their per-launch CSV summaries come back.
timeout 900 ncu --set full --clock-control none \
    -k regex:'spmm_kernel|pair_walk|distmult_fwd|distmult_grads|tc_gemm_kernel|tc_tn_kernel|sgemm_kernel' -s 300 -c 30 -f -o /tmp/${tag}_full_pose \
    python bench.py --steps 2 --warmup 1 --eager --no-cpu-baseline --no-config5 --no-train-epoch > $out/${tag}_full_pose.log 2>&1
echo "ncu full pose rc=$?"
timeout 900 ncu --set full --clock-control none \
    -k regex:'tc_gemm_kernel|tc_tn_kernel|b_image_kernel|spmm_kernel|pair_walk|distmult_fwd|sgemm_kernel' -s 150 -c 24 -f -o /tmp/${tag}_full_pose2 \
    python bench.py --workload pose2 --steps 1 --warmup 1 --eager --no-cpu-baseline --no-config5 --no-train-epoch > $out/${tag}_full_pose2.log 2>&1
echo "ncu full pose2 rc=$?"
for wl in aminer freebase-d; do
  timeout 900 ncu --set full --clock-control none \
      -k regex:'spmm_kernel|tc_gemm_kernel|tc_tn_kernel|sgemm_kernel' -s 150 -c 16 -f -o /tmp/${tag}_full_${wl} \
      python bench.py --workload $wl --steps 1 --warmup 1 --eager --no-cpu-baseline --no-config5 --no-train-epoch > $out/${tag}_full_${wl}.log 2>&1
  echo "ncu full $wl rc=$?"
done
for wl in pose pose2 aminer freebase-d; do
  python profiles/summarize_full.py /tmp/${tag}_full_${wl}.ncu-rep > $out/${tag}_ncu_full_${wl}.csv 2>/dev/null
done
du -sh $out; ls $out | grep ${tag} | head -40
