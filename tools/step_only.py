"""Three fused training steps on the bench workload (for ncu captures of one step:
`ncu --set full -k regex:'tc_gemm|approx_loss|out_layer' -s 22 -c 11 ... python tools/step_only.py`)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ranking_b200 as tfr
import bench

dev = torch.device('cuda')
tower = tfr.keras.layers.create_tower(bench.HIDDEN, 1, activation='relu', use_batch_norm=False,
                                      dropout=0, input_dim=bench.D, precision='tf32x3', seed=1)
trainer = tfr.train.RankingTrainer(tower, tfr.keras.losses.get('approx_ndcg_loss'),
                                   optimizer='adagrad', learning_rate=0.05)
batches = [tuple(t.to(dev) for t in bench.make_batch(s)) for s in range(3)]
for x, y in batches:
  trainer.train_step(x, y)
torch.cuda.synchronize()
