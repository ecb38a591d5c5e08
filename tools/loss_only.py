"""Runs the ApproxNDCG fwd+bwd kernel alone on the bench shape (for ncu captures)."""
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
key = sys.argv[1] if len(sys.argv) > 1 else 'approx_ndcg_loss'
lam = tfr.keras.losses.NDCGLambdaWeight() if len(sys.argv) > 2 and sys.argv[2] == 'ndcg' else None
loss = tfr.keras.losses.get(key, lambda_weight=lam) if lam is not None else tfr.keras.losses.get(key)
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
