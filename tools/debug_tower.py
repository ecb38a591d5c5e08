"""Debug helper: per-parameter-block gradient error of the CUDA tower vs the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ranking_b200 as tfr
from oracle import scorer as S


def run(m, d, hidden, out, act, precision):
  tower = tfr.keras.layers.create_tower(hidden, out, activation=act, use_batch_norm=False,
                                        dropout=0, input_dim=d, seed=m, precision=precision)
  with torch.no_grad():
    for i in range(len(tower.dims) - 1):
      tower.bias(i).uniform_(-0.2, 0.2)
  nl = len(tower.dims) - 1
  params = {'dense_w': [tower.kernel(i).detach().cpu().double().clone().requires_grad_()
                        for i in range(nl)],
            'dense_b': [tower.bias(i).detach().cpu().double().clone().requires_grad_()
                        for i in range(nl)]}
  g = torch.Generator().manual_seed(m)
  x = torch.randn(m, d, generator=g)
  up = torch.randn(m, out, generator=g)
  y = tower(x.cuda())
  (y * up.cuda()).sum().backward()
  ref = S.tower_forward(x.double(), params, activation=act)
  (ref * up.double()).sum().backward()
  got = tower.flat.grad.cpu().double()
  print('shape', m, d, hidden, act, precision, 'fwd err %.2e' % float(
      (y.detach().cpu().double() - ref).abs().max() / ref.abs().max()))
  for i, (a, b, c) in enumerate(tower.offsets):
    gw, gb = got[a:b], got[b:c]
    rw, rb = params['dense_w'][i].grad.reshape(-1), params['dense_b'][i].grad
    print('  layer %d  W err %.2e (max %.2e)   b err %.2e (max %.2e)' % (
        i, float((gw - rw).abs().max()), float(rw.abs().max()),
        float((gb - rb).abs().max()), float(rb.abs().max())))


if __name__ == '__main__':
  for m in (4100, 4096, 2048, 2304, 300):
    for act in ('relu', None):
      run(m, 136, [256, 128, 64], 1, act, 'tf32x3')
  run(4100, 136, [256, 128, 64], 1, 'relu', 'fp32')
