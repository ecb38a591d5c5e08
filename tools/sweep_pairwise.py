"""BASELINE config 5: list-size sweep of the fused pairwise-logistic loss kernel (K1).

Reports per N: kernel time (CUDA events), lists/s, algorithmic HBM GB/s
((12 N + 16) bytes per list: scores + labels in, gradient out, 4 per-list scalars)
against the measured HBM peak, and pair evaluations per second (N^2 per list)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ranking_b200 as tfr
import bench


def run(n, lam=None, reps=20, nbuf=8):
  b = max(64, (1 << 21) // n)
  bufs = []
  for i in range(nbuf):
    x, y = bench.make_batch(100 + i, b=b, n=n, d=1)
    bufs.append((torch.randn(b, n, generator=torch.Generator().manual_seed(i)).mul(2).cuda(),
                 y.cuda()))
  loss = tfr.keras.losses.PairwiseLogisticLoss(lambda_weight=lam)
  grad = torch.empty(b, n, device='cuda')
  per_list = torch.empty(2, b, device='cuda')
  total2 = torch.zeros(2, device='cuda')
  for i in range(3):
    loss.fused_fwd_bwd(bufs[i % nbuf][1], bufs[i % nbuf][0], None, grad, per_list, total2)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for i in range(reps):
    loss.fused_fwd_bwd(bufs[i % nbuf][1], bufs[i % nbuf][0], None, grad, per_list, total2)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / reps
  bytes_ = (12 * n + 16) * b
  return {'N': n, 'B': b, 'ms': ms, 'lists_per_s': b / (ms * 1e-3),
          'hbm_gbs_algorithmic': bytes_ / (ms * 1e-3) / 1e9,
          'pair_evals_per_s': float(n) * n * b / (ms * 1e-3)}


if __name__ == '__main__':
  peaks = bench.load_peaks()
  out = {'peak_hbm_gbs': peaks['hbm_gbs'], 'peak_source': peaks['source'], 'rows': []}
  for name, lam in (('none', None), ('ndcg', tfr.keras.losses.NDCGLambdaWeight())):
    for n in (32, 64, 128, 256, 512, 1024):
      r = run(n, lam)
      r['lambda_weight'] = name
      r['hbm_frac'] = r['hbm_gbs_algorithmic'] / peaks['hbm_gbs']
      out['rows'].append(r)
      print(json.dumps(r), flush=True)
  json.dump(out, open(os.path.join('gpurun_out', 'sweep_pairwise.json'), 'w'), indent=1)
