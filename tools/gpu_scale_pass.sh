#!/bin/bash
# Multi-GPU pass (run with `gpurun --gpus N -- bash tools/gpu_scale_pass.sh N`): 2-rank
# equivalence of the data-parallel step (fused K7 kernel and NCCL), then the bench lines of
# configs 2 / 3 / 5 at N GPUs.  Outputs -> gpurun_out/r02_*.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "$N" -ge 2 ]; then
  for coll in fused nccl; do
    TFR_COLLECTIVE=$coll timeout 300 $TR --nproc-per-node 2 --master-port 29611 tools/dp_equivalence.py \
      > gpurun_out/r02_dp_equivalence_$coll.log 2>&1
    grep -h "DP_EQUIVALENCE_OK\|max rel\|Error\|error" gpurun_out/r02_dp_equivalence_$coll.log | head -5
  done
fi
port=29620
for c in 2 3 5; do
  port=$((port + 1))
  timeout 400 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --config $c --steps 20 --warmup 5 \
    > gpurun_out/r02_bench_c${c}_${N}gpu.json 2> gpurun_out/r02_bench_c${c}_${N}gpu.err
  python - <<PY
import json
try:
  d = json.load(open('gpurun_out/r02_bench_c${c}_${N}gpu.json'))
  print('c$c N=$N', round(d['value']), d['ms_per_step'], d.get('collective'), (d.get('e2e') or {}).get('ms_per_step'))
except Exception as e:
  print('c$c N=$N ERR', e)
PY
done
port=$((port + 1))
timeout 400 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --config 2 --collective nccl --steps 20 --warmup 5 \
  > gpurun_out/r02_bench_c2_${N}gpu_nccl.json 2> gpurun_out/r02_bench_c2_${N}gpu_nccl.err
python - <<PY
import json
try:
  d = json.load(open('gpurun_out/r02_bench_c2_${N}gpu_nccl.json'))
  print('c2 nccl N=$N', round(d['value']), d['ms_per_step'], d.get('collective'))
except Exception as e:
  print('c2 nccl N=$N ERR', e)
PY
