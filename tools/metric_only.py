"""Runs the rank-metrics kernel (K4: NDCG@{1,3,5,10,all} + MRR + the extended metrics in one
launch) alone on the bench shape, for ncu captures."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ranking_b200 as tfr
import bench

B = int(os.environ.get('B', 1024))
N = int(os.environ.get('N', 200))
_, y = bench.make_batch(0, B, N, 8)
y = y.cuda()
scores = torch.randn(B, N, device='cuda')
group = tfr.keras.metrics.MetricGroup.default()
for _ in range(3):
  group.update_state(y, scores)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
  group.update_state(y, scores)
e1.record()
torch.cuda.synchronize()
print('us per update_state', e0.elapsed_time(e1) * 1e3 / 20)
