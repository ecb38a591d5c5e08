"""`tfr.metrics` surface: `RankingMetricKey`, `compute_mean`, `make_ranking_metric_fn`
(reference metrics.py:37-301).

The reference returns `tf.compat.v1.metrics.mean(per_list_metric, per_list_weights)` pairs; here
a metric function returns the weighted mean of the batch as a 0-d tensor,
`sum(metric * weight) / sum(weight)` (what `compute_mean`, metrics.py:78-121, returns and what
the streaming mean accumulates per batch).  Every key maps to the CUDA metric classes of
`ranking_b200.metrics_impl` (one K4 launch: 64-bit-key bitonic sort per list).
"""
import torch

from ranking_b200 import metrics_impl


class RankingMetricKey(object):
  """Ranking metric key strings (metrics.py:37-76)."""
  MRR = 'mrr'
  ARP = 'arp'
  NDCG = 'ndcg'
  DCG = 'dcg'
  PRECISION = 'precision'
  RECALL = 'recall'
  MAP = 'map'
  PRECISION_IA = 'precision_ia'
  ORDERED_PAIR_ACCURACY = 'ordered_pair_accuracy'
  ALPHA_DCG = 'alpha_dcg'
  BPREF = 'bpref'
  HITS = 'hits'
  PWA = 'pwa'


def _build(metric_key, topn, name, gain_fn, rank_discount_fn, kwargs):
  K, M = RankingMetricKey, metrics_impl
  if metric_key == K.ARP:
    return M.ARPMetric(name)
  if metric_key == K.MRR:
    return M.MRRMetric(name, topn)
  if metric_key in (K.NDCG, K.DCG):
    cls = M.NDCGMetric if metric_key == K.NDCG else M.DCGMetric
    extra = {}
    if gain_fn is not None:
      extra['gain_fn'] = gain_fn
    if rank_discount_fn is not None:
      extra['rank_discount_fn'] = rank_discount_fn
    return cls(name, topn, **extra)
  if metric_key == K.PRECISION:
    return M.PrecisionMetric(name, topn)
  if metric_key == K.RECALL:
    return M.RecallMetric(name, topn)
  if metric_key == K.MAP:
    return M.MeanAveragePrecisionMetric(name, topn)
  if metric_key == K.ORDERED_PAIR_ACCURACY:
    return M.OPAMetric(name)
  if metric_key == K.BPREF:
    return M.BPrefMetric(name, topn, use_trec_version=kwargs.get('use_trec_version', True))
  if metric_key == K.HITS:
    return M.HitsMetric(name, topn)
  if metric_key == K.PRECISION_IA:
    return M.PrecisionIAMetric(name, topn)
  if metric_key == K.ALPHA_DCG:
    extra = {'alpha': kwargs.get('alpha', 0.5), 'seed': kwargs.get('seed')}
    if rank_discount_fn is not None:
      extra['rank_discount_fn'] = rank_discount_fn
    return M.AlphaDCGMetric(name, topn, **extra)
  if metric_key == K.PWA:
    raise ValueError('metric_key pwa is not available in ranking_b200.')
  raise ValueError('metric_key %s not supported.' % metric_key)


def _weighted_mean(metric, weight):
  num = (metric * weight).sum()
  den = weight.sum()
  return torch.where(den != 0, num / torch.where(den != 0, den, torch.ones_like(den)),
                     torch.zeros_like(num))


def compute_mean(metric_key, labels, predictions, weights=None, topn=None, name=None):
  """metrics.py:78-121: the weighted mean of the metric over the batch (a scalar)."""
  metric = _build(metric_key, topn, name, None, None, {})
  value, weight = metric.compute(labels, predictions, weights)
  return _weighted_mean(value, weight)


def make_ranking_metric_fn(metric_key, weights_feature_name=None, topn=None, name=None,
                           gain_fn=None, rank_discount_fn=None, **kwargs):
  """metrics.py:124-301.  Returns `metric_fn(labels, predictions, features)`; the weights come
  from `features[weights_feature_name]` ([B, list_size] per item or [B, 1] per list).
  `gain_fn` / `rank_discount_fn` default to `2^label - 1` and `ln 2 / ln(1 + rank)`
  (metrics.py:32-34) and apply to NDCG / DCG (the discount also to alpha-DCG); `alpha`, `seed`
  (alpha-DCG) and `use_trec_version` (BPref) travel in `kwargs`."""
  metric = _build(metric_key, topn, name, gain_fn, rank_discount_fn, kwargs)

  def _metric_fn(labels, predictions, features):
    weights = None
    if weights_feature_name:
      weights = features[weights_feature_name]
      if weights.dim() == 1:
        weights = weights.reshape(-1, 1)
      elif weights.dim() > 2:
        weights = weights.reshape(weights.shape[0], -1)
    value, weight = metric.compute(labels, predictions, weights)
    return _weighted_mean(value, weight)

  return _metric_fn
