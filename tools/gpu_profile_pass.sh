#!/bin/bash
# One profiling pass on the GPU box (single GPU): launch lists of one step of configs 2 / 3 / 4,
# `ncu --set full` of one step of each (summarised to CSV on the box: the .ncu-rep files are
# too large to travel back), the K4 metric kernel alone, the wait-cycle profile of the TF32
# engine and the tcgen05 rate micro-benchmark.  Outputs -> gpurun_out/<tag>_*.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
for c in 2 3 4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_c$c.csv python tools/step_only.py --config $c --steps 3 \
    > gpurun_out/${TAG}_launches_c$c.log 2>&1
  python tools/launch_table.py gpurun_out/${TAG}_launches_c$c.csv > gpurun_out/${TAG}_launches_c$c.txt 2>&1
done
for c in 2 3 4; do
  timeout 600 ncu --set full --clock-control none -f -o /tmp/${TAG}_step_c$c \
    python tools/step_only.py --config $c --steps 1 > gpurun_out/${TAG}_full_c$c.log 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_step_c$c.ncu-rep gpurun_out/${TAG}_step_c${c}_ncu_full_summary.csv \
    gpurun_out/${TAG}_traffic_c$c.json >> gpurun_out/${TAG}_full_c$c.log 2>&1
done
timeout 300 ncu --set full --clock-control none -f -k regex:rank_metrics -c 2 -o /tmp/${TAG}_k4 \
  python tools/metric_only.py > gpurun_out/${TAG}_full_k4.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_k4.ncu-rep gpurun_out/${TAG}_k4_ncu_full_summary.csv >> gpurun_out/${TAG}_full_k4.log 2>&1
for g in fwd1 dz1 dw1; do
  timeout 300 ncu --set full --import-source on --clock-control none -f -k regex:tc_gemm -c 1 \
    -o gpurun_out/${TAG}_hot_$g python tools/gemm_only.py $g > gpurun_out/${TAG}_hot_$g.log 2>&1
done
timeout 300 python tools/tc_wait_profile.py > gpurun_out/${TAG}_tc_gemm_wait_cycles.txt 2>&1
[ -x build/mma_rate ] && timeout 120 ./build/mma_rate > gpurun_out/${TAG}_mma_rate.txt 2>&1
ls -la gpurun_out | tail -30
