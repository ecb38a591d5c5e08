#!/bin/bash
# One profiling pass on the GPU box: launch lists for configs 2/3/4, ncu --set full of one
# step of configs 2 and 3 (summarised to CSV on the box; the .ncu-rep files are too large to
# travel back, only the top-kernel capture is kept), the wait-cycle profile of the TF32 engine.
# Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
for c in 2 3 4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_c$c.csv python tools/step_only.py --config $c --steps 3 \
    > gpurun_out/${TAG}_launches_c$c.log 2>&1
done
for c in 2 3; do
  timeout 600 ncu --set full --clock-control none -f \
    -o /tmp/${TAG}_step_c$c python tools/step_only.py --config $c --steps 2 \
    > gpurun_out/${TAG}_full_c$c.log 2>&1
  python tools/ncu_summary.py /tmp/${TAG}_step_c$c.ncu-rep gpurun_out/${TAG}_step_c${c}_ncu_full_summary.csv \
    gpurun_out/${TAG}_traffic_c$c.json >> gpurun_out/${TAG}_full_c$c.log 2>&1
done
timeout 300 python tools/tc_wait_profile.py > gpurun_out/${TAG}_tc_gemm_wait_cycles.txt 2>&1
ls -la gpurun_out | tail -20
