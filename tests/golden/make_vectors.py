"""Generates tests/golden/reference_vectors.json.

The reference (tensorflow_ranking) cannot be executed offline (it needs
TensorFlow), so these vectors come from the fp64 ORACLE, which is itself pinned by
the reference's closed-form test expectations (tests/test_golden_*.py).  They
freeze seeded inputs and outputs (loss, d loss / d scores, NDCG@k / MRR) so that
(a) the oracle cannot drift silently and (b) the CUDA path can be checked on the
GPU box against committed numbers.  Run from the repo root:
    python tests/golden/make_vectors.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import keras_losses as K  # noqa: E402
from oracle import metrics_impl as M  # noqa: E402


def batch(b, n, seed):
  g = torch.Generator().manual_seed(seed)
  scores = (torch.randn(b, n, generator=g) * 2.0).float()
  labels = torch.multinomial(torch.tensor([.55, .25, .12, .06, .02]), b * n,
                             replacement=True, generator=g).reshape(b, n).float()
  lens = torch.randint((n + 1) // 2, n + 1, (b,), generator=g)
  labels = torch.where(torch.arange(n).unsqueeze(0) < lens.unsqueeze(1), labels,
                       torch.full_like(labels, -1.))
  labels[1] = torch.where(labels[1] >= 0, torch.zeros_like(labels[1]), labels[1])
  weights = (torch.rand(b, n, generator=g) + 0.5).float()
  return scores, labels, weights


LOSSES = {
    'pairwise_logistic_loss': lambda: K.PairwiseLogisticLoss(),
    'pairwise_hinge_loss': lambda: K.PairwiseHingeLoss(),
    'pairwise_soft_zero_one_loss': lambda: K.PairwiseSoftZeroOneLoss(),
    'pairwise_mse_loss': lambda: K.PairwiseMSELoss(),
    'pairwise_logistic_loss+ndcg_lambda': lambda: K.PairwiseLogisticLoss(
        lambda_weight=K.NDCGLambdaWeight()),
    'pairwise_logistic_loss+ndcg_lambda_top5_smooth': lambda: K.PairwiseLogisticLoss(
        lambda_weight=K.NDCGLambdaWeight(topn=5, smooth_fraction=0.25)),
    'pairwise_logistic_loss+ndcg_lambda_v2_top5': lambda: K.PairwiseLogisticLoss(
        lambda_weight=K.NDCGLambdaWeightV2(topn=5)),
    'softmax_loss': lambda: K.SoftmaxLoss(),
    'approx_ndcg_loss': lambda: K.ApproxNDCGLoss(),
    'approx_mrr_loss': lambda: K.ApproxMRRLoss(),
}


def main():
  out = {'generator': 'tests/golden/make_vectors.py', 'oracle_dtype': 'float64', 'cases': []}
  for (b, n, seed) in [(4, 9, 1), (6, 37, 2), (3, 200, 3)]:
    scores, labels, weights = batch(b, n, seed)
    case = {'B': b, 'N': n, 'scores': scores.tolist(), 'labels': labels.tolist(),
            'weights': weights.tolist(), 'losses': {}, 'metrics': {}}
    for name, make in LOSSES.items():
      for wname, w in (('none', None), ('item', weights)):
        s = scores.double().requires_grad_()
        val = make()(labels.double(), s, None if w is None else w.double())
        val.backward()
        case['losses'][name + '|' + wname] = {'value': float(val.detach()),
                                              'grad': s.grad.tolist()}
    for topn in (1, 5, 10, None):
      nd, ndw = M.NDCGMetric(topn=topn).compute(labels.double(), scores.double(),
                                                weights.double())
      mr, mrw = M.MRRMetric(topn=topn).compute(labels.double(), scores.double(),
                                               weights.double())
      case['metrics'][str(topn)] = {'ndcg': nd[:, 0].tolist(), 'ndcg_w': ndw[:, 0].tolist(),
                                    'mrr': mr[:, 0].tolist(), 'mrr_w': mrw[:, 0].tolist()}
    out['cases'].append(case)
  path = os.path.join(ROOT, 'tests', 'golden', 'reference_vectors.json')
  json.dump(out, open(path, 'w'))
  print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
