"""Runs the ApproxNDCG fwd+bwd kernel alone on the bench shape (for ncu captures)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ranking_b200 as tfr
import bench

B, N = 1024, 200
_, y = bench.make_batch(0, B, N, 8)
y = y.cuda()
scores = torch.randn(B, N, device='cuda')
loss = tfr.keras.losses.get(sys.argv[1] if len(sys.argv) > 1 else 'approx_ndcg_loss')
grad = torch.empty_like(scores)
per_list = torch.empty(2, B, device='cuda')
total2 = torch.zeros(2, device='cuda')
for _ in range(3):
  loss.fused_fwd_bwd(y, scores, None, grad, per_list, total2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
  loss.fused_fwd_bwd(y, scores, None, grad, per_list, total2)
e1.record()
torch.cuda.synchronize()
print('us per call', e0.elapsed_time(e1) * 1e3 / 20)
