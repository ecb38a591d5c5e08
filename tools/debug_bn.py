"""Per-parameter-segment error of the BN tower against the fp64 oracle (GPU box)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import ranking_b200 as tfr
import oracle
from oracle import scorer
import test_parity_gpu as T


def run(m, d, hidden, out, act, input_bn, use_bn, precision):
  tower, params = T._bn_tower_and_params(tfr, d, hidden, out, m, act, precision, input_bn, use_bn)
  g = torch.Generator().manual_seed(m)
  x = torch.randn(m, d, generator=g) * 1.5 + 0.3
  up = torch.randn(m, out, generator=g)
  tower.train()
  y = tower(x.cuda())
  (y * up.cuda()).sum().backward()
  ref = scorer.tower_forward(x.double(), params, activation=act, use_batch_norm=use_bn and bool(hidden),
                             input_batch_norm=input_bn, training=True)
  (ref * up.double()).sum().backward()
  print(m, d, hidden, act, input_bn, use_bn, precision, 'fwd', T._rel_err(y, ref))
  got = tower.flat.grad.detach().cpu().double()
  nl = len(tower.dims) - 1
  for i in range(nl):
    a, b, c = tower.offsets[i]
    w, bb = params['dense_w'][i].grad.reshape(-1), params['dense_b'][i].grad
    print('  W%d l2 %.3e max %.3e | b abs %.3e (ref max %.3e)' % (
        i, float((got[a:b] - w).norm() / w.norm()), float((got[a:b] - w).abs().max() / w.abs().max()),
        float((got[b:c] - bb).abs().max()), float(bb.abs().max())))
  for key in tower.bn_offsets:
    a, b, w_ = tower.bn_offsets[key]
    rg = (params['in_bn_gamma'] if key == 'input' else params['bn_gamma'][key]).grad
    rb = (params['in_bn_beta'] if key == 'input' else params['bn_beta'][key]).grad
    print('  bn %s gamma l2 %.3e beta l2 %.3e' % (
        key, float((got[a:a + w_] - rg).norm() / rg.norm()), float((got[b:b + w_] - rb).norm() / rb.norm())))


for prec in ['fp32', 'tf32x3']:
  run(700, 136, [256, 128, 64], 1, 'relu', True, True, prec)
  run(700, 136, [256, 128, 64], 1, 'relu', False, True, prec)
  run(700, 136, [256, 128, 64], 1, None, False, True, prec)
  run(1300, 24, [48, 16], 1, None, True, False, prec)
