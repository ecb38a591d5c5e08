#!/bin/bash
# Final single-GPU pass: tests, the bench lines of every config (with the CPU baseline), the
# reference arm, then the profile pass.
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r02_pytest_gpu.log
tail -2 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
for c in 2 3 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r02_bench_c${c}_1gpu.json 2> gpurun_out/r02_bench_c${c}_1gpu.err
  python - <<PY
import json
try:
  d = json.load(open('gpurun_out/r02_bench_c${c}_1gpu.json'))
  print('c$c', round(d['value']), d['ms_per_step'], (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
except Exception as e:
  print('c$c ERR', e)
PY
done
for c in 2 3; do
  timeout 900 python bench.py --impl reference --config $c --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm_c$c.json 2> gpurun_out/r02_bench_reference_arm_c$c.err
  head -c 400 gpurun_out/r02_bench_reference_arm_c$c.json; echo
done
bash tools/gpu_profile_pass.sh r02 > gpurun_out/profile_pass.log 2>&1
du -sh gpurun_out
