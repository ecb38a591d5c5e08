"""`ranking_b200.metrics`: RankingMetricKey / compute_mean / make_ranking_metric_fn
(reference metrics.py:37-301).  The dispatch and weighting logic is exercised on the CPU with
the oracle's metric classes swapped in (same `compute` contract), against the reference's own
expectations (metrics_test.py-style closed forms); the CUDA classes run under `-m gpu`."""
import math

import pytest
import torch


def log2p1(x):
  return math.log2(1. + x)


@pytest.fixture
def surface(monkeypatch, request):
  from ranking_b200 import metrics as S
  if request.param == 'oracle':
    from oracle import metrics_impl as OM
    monkeypatch.setattr(S, 'metrics_impl', OM)
    return S, (lambda x: torch.tensor(x, dtype=torch.float64))
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  return S, (lambda x: torch.tensor(x, dtype=torch.float32, device='cuda'))


def _cases():
  scores = [[1., 3., 2.], [1., 2., 3.]]
  labels = [[0., 0., 1.], [0., 1., 2.]]
  return scores, labels


@pytest.mark.parametrize('surface', ['oracle', pytest.param('cuda', marks=pytest.mark.gpu)],
                         indirect=True)
def test_make_ranking_metric_fn(surface):
  S, t = surface
  K = S.RankingMetricKey
  scores, labels = _cases()
  y, s = t(labels), t(scores)
  w = t([[1., 2., 3.], [4., 5., 6.]])

  def val(key, features=None, **kw):
    return float(S.make_ranking_metric_fn(key, **kw)(y, s, features or {}))

  # metrics_impl_test.py-style closed forms
  assert val(K.MRR) == pytest.approx((1. / 2. + 1.) / 2.)
  assert val(K.ARP) == pytest.approx((2. + 1. * 2. + 2. * 1.) / (1. + 1. + 2.))
  ndcg0 = (1. / log2p1(2.)) / (1. / log2p1(1.))
  assert val(K.NDCG) == pytest.approx((ndcg0 + 1.) / 2., rel=1e-5)
  assert val(K.NDCG, topn=1) == pytest.approx((0. + 1.) / 2., rel=1e-5)
  assert val(K.DCG) == pytest.approx(
      ((1. / log2p1(2.)) + (3. / log2p1(1.) + 1. / log2p1(2.))) / 2., rel=1e-5)
  assert val(K.PRECISION, topn=1) == pytest.approx((0. + 1.) / 2.)
  assert val(K.RECALL, topn=1) == pytest.approx((0. + 0.5) / 2.)
  assert val(K.MAP) == pytest.approx((0.5 + 1.) / 2.)
  assert val(K.HITS, topn=1) == pytest.approx((0. + 1.) / 2.)
  assert val(K.ORDERED_PAIR_ACCURACY) == pytest.approx((1. + 3.) / (2. + 3.))
  assert val(K.BPREF) == pytest.approx((0. + 1.) / 2.)
  assert val(K.BPREF, use_trec_version=False) == pytest.approx((0. + 1.) / 2.)
  # weights feature: per-item weights -> per-list weight = mean weight of the relevant items
  got = val(K.MRR, features={'w': w}, weights_feature_name='w')
  assert got == pytest.approx((0.5 * 3. + 1. * 5.5) / (3. + 5.5))
  # custom gain / discount reach NDCG
  lin = val(K.NDCG, gain_fn=lambda l: l, rank_discount_fn=lambda r: 1. / r)
  assert lin == pytest.approx(((1. / 2.) / 1. + 1.) / 2., rel=1e-5)
  # compute_mean (metrics.py:78-121)
  assert float(S.compute_mean(K.MRR, y, s)) == pytest.approx((0.5 + 1.) / 2.)
  assert float(S.compute_mean(K.NDCG, y, s, topn=1)) == pytest.approx(0.5, rel=1e-5)


def test_metric_surface_errors():
  from ranking_b200 import metrics as S
  with pytest.raises(ValueError):
    S.make_ranking_metric_fn('no_such_metric')
  with pytest.raises(ValueError):
    S.make_ranking_metric_fn(S.RankingMetricKey.PWA)
