"""A few fused training steps of a bench workload, nothing else (for ncu captures:
`ncu --set full -k regex:'gemm|loss|out_' -s <first step's launches> -c <one step> ...
python tools/step_only.py --config 3`)."""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ranking_b200 as tfr
import bench

ap = argparse.ArgumentParser()
ap.add_argument('--config', type=int, default=2, choices=[2, 3, 4])
ap.add_argument('--precision', default=None)
ap.add_argument('--steps', type=int, default=3)
args = ap.parse_args()
w = bench.WORKLOADS[args.config]
dev = torch.device('cuda')
gs = w['group_size']
tower = tfr.keras.layers.create_tower(bench.HIDDEN, gs, activation='relu', use_batch_norm=False,
                                      dropout=0, input_dim=w['D'] * gs,
                                      precision=args.precision or w['precision'], seed=1)
lam = tfr.keras.losses.NDCGLambdaWeight() if w['lam'] == 'ndcg' else None
loss = tfr.keras.losses.get(w['loss'], lambda_weight=lam)
if gs > 1:
  trainer = tfr.train.GroupwiseRankingTrainer(tower, loss, gs, optimizer='adagrad',
                                              learning_rate=0.05)
else:
  trainer = tfr.train.RankingTrainer(tower, loss, optimizer='adagrad', learning_rate=0.05)
batches = [tuple(t.to(dev) for t in bench.make_batch(s, w['B'], w['N'], w['D'],
                                                     tower.input_dtype))
           for s in range(2)]
from ranking_b200 import _C
for i in range(args.steps):
  l0 = _C.lib.tfr_launch_count()
  trainer.train_step(*batches[i % 2])
  torch.cuda.synchronize()
  print('step %d: %d launches' % (i, _C.lib.tfr_launch_count() - l0), flush=True)
