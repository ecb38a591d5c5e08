"""Oracle restatement of tensorflow_ranking/python/metrics_impl.py (NDCG, MRR)
and of the Keras `Mean`-based wrappers in keras/metrics.py:156-193.

Test infrastructure only (see oracle/__init__.py).
"""
import torch

from oracle import losses_impl as L
from oracle import utils

_DEFAULT_GAIN_FN = L.pow_minus_1                  # metrics_impl.py:31
_DEFAULT_RANK_DISCOUNT_FN = L.log2_inverse        # metrics_impl.py:33 (ln2/log1p)


def _per_example_weights_to_per_list_weights(weights, relevance):
  """metrics_impl.py:63-119."""
  nonzero_weights = weights.sum(1, keepdim=True) > 0.0
  per_list_relevance = relevance.sum(1, keepdim=True)
  nonzero_relevance = torch.where(
      nonzero_weights, (per_list_relevance > 0.0).to(weights.dtype),
      torch.zeros_like(per_list_relevance))
  nonzero_relevance_count = nonzero_relevance.sum(0, keepdim=True)
  per_list_weights = L._divide_no_nan(
      (weights * relevance).sum(1, keepdim=True), per_list_relevance)
  sum_weights = per_list_weights.sum(0, keepdim=True)
  avg_weight = torch.where(
      nonzero_relevance_count > 0.0,
      L._divide_no_nan(sum_weights, nonzero_relevance_count),
      torch.ones_like(nonzero_relevance_count))
  return torch.where(
      nonzero_weights,
      torch.where(per_list_relevance > 0.0, per_list_weights,
                  torch.ones_like(per_list_weights) * avg_weight),
      torch.zeros_like(per_list_weights))


def _discounted_cumulative_gain(labels, weights, gain_fn, rank_discount_fn):
  """metrics_impl.py:122-151."""
  list_size = labels.shape[1]
  position = torch.arange(1, list_size + 1).to(labels.dtype)
  gain = gain_fn(labels)
  discount = rank_discount_fn(position)
  return (weights * gain * discount).sum(1, keepdim=True)


class _RankingMetric(object):

  def _prepare_and_validate_params(self, labels, predictions, weights, mask):
    """metrics_impl.py:228-266."""
    predictions = torch.as_tensor(predictions)
    labels = torch.as_tensor(labels, dtype=predictions.dtype)
    weights = 1.0 if weights is None else torch.as_tensor(
        weights, dtype=predictions.dtype)
    example_weights = torch.ones_like(labels) * weights
    assert predictions.dim() == 2
    if mask is None:
      mask = utils.is_label_valid(labels)
    mask = torch.logical_and(torch.as_tensor(mask), example_weights > 0.0)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    predictions = torch.where(
        mask, predictions, -1e-6 * torch.ones_like(predictions) +
        predictions.min(dim=1, keepdim=True).values)
    return labels, predictions, example_weights, mask

  def compute(self, labels, predictions, weights=None, mask=None):
    """metrics_impl.py:268-291."""
    labels, predictions, weights, mask = self._prepare_and_validate_params(
        labels, predictions, weights, mask)
    return self._compute_impl(labels, predictions, weights, mask)


class MRRMetric(_RankingMetric):
  """metrics_impl.py:429-459."""

  def __init__(self, name=None, topn=None):
    self._topn = topn

  def _compute_impl(self, labels, predictions, weights, mask):
    topn = predictions.shape[1] if self._topn is None else self._topn
    sorted_labels, = utils.sort_by_scores(predictions, [labels], topn=topn,
                                          mask=mask)
    n = sorted_labels.shape[1]
    relevance = (sorted_labels >= 1.0).to(predictions.dtype)
    reciprocal_rank = 1.0 / torch.arange(1, n + 1).to(predictions.dtype)
    mrr = (relevance * reciprocal_rank).max(dim=1, keepdim=True).values
    per_list_weights = _per_example_weights_to_per_list_weights(
        weights=weights, relevance=(labels >= 1.0).to(predictions.dtype))
    return mrr, per_list_weights


class NDCGMetric(_RankingMetric):
  """metrics_impl.py:631-670."""

  def __init__(self, name=None, topn=None, gain_fn=_DEFAULT_GAIN_FN,
               rank_discount_fn=_DEFAULT_RANK_DISCOUNT_FN):
    self._topn = topn
    self._gain_fn = gain_fn
    self._rank_discount_fn = rank_discount_fn

  def _compute_impl(self, labels, predictions, weights, mask):
    topn = predictions.shape[1] if self._topn is None else self._topn
    sorted_labels, sorted_weights = utils.sort_by_scores(
        predictions, [labels, weights], topn=topn, mask=mask)
    dcg = _discounted_cumulative_gain(sorted_labels, sorted_weights,
                                      self._gain_fn, self._rank_discount_fn)
    weighted_gains = weights * self._gain_fn(labels)
    ideal_sorted_labels, ideal_sorted_weights = utils.sort_by_scores(
        weighted_gains, [labels, weights], topn=topn, mask=mask)
    ideal_dcg = _discounted_cumulative_gain(
        ideal_sorted_labels, ideal_sorted_weights, self._gain_fn,
        self._rank_discount_fn)
    per_list_ndcg = L._divide_no_nan(dcg, ideal_dcg)
    per_list_weights = _per_example_weights_to_per_list_weights(
        weights=weights, relevance=self._gain_fn(labels))
    return per_list_ndcg, per_list_weights


def _per_list_recall(labels, predictions, topn, mask):
  """metrics_impl.py:154-177."""
  sorted_labels = utils.sort_by_scores(predictions, [labels], topn=topn,
                                       mask=mask)[0]
  topn_positives = (sorted_labels >= 1.0).to(predictions.dtype)
  labels = (labels >= 1.0).to(predictions.dtype)
  return L._divide_no_nan(topn_positives.sum(1, keepdim=True),
                          labels.sum(1, keepdim=True))


def _per_list_precision(labels, predictions, topn, mask):
  """metrics_impl.py:180-207."""
  sorted_labels = utils.sort_by_scores(predictions, [labels], topn=topn,
                                       mask=mask)[0]
  relevance = (sorted_labels >= 1.0).to(predictions.dtype)
  if topn is None:
    topn = relevance.shape[1]
  valid_topn = torch.clamp(mask.to(torch.int64).sum(1, keepdim=True), max=topn)
  return L._divide_no_nan(relevance.sum(1, keepdim=True),
                          valid_topn.to(predictions.dtype))


def _binary(labels):
  return (labels >= 1.0).to(labels.dtype)


class HitsMetric(_RankingMetric):
  """metrics_impl.py:462-506."""

  def __init__(self, name=None, topn=None):
    self._topn = topn

  def _compute_impl(self, labels, predictions, weights, mask):
    topn = predictions.shape[1] if self._topn is None else self._topn
    sorted_labels, = utils.sort_by_scores(predictions, [labels], topn=topn,
                                          mask=mask)
    relevance = (sorted_labels >= 1.0).to(predictions.dtype)
    hits = relevance.max(dim=1, keepdim=True).values
    return hits, _per_example_weights_to_per_list_weights(weights, _binary(labels))


class ARPMetric(_RankingMetric):
  """metrics_impl.py:509-536."""

  def __init__(self, name=None):
    pass

  def _compute_impl(self, labels, predictions, weights, mask):
    topn = predictions.shape[1]
    sorted_labels, sorted_weights = utils.sort_by_scores(
        predictions, [labels, weights], topn=topn, mask=mask)
    weighted_labels = sorted_labels * sorted_weights
    position = torch.arange(1, topn + 1).to(predictions.dtype) * torch.ones_like(
        weighted_labels)
    per_list_weights = weighted_labels.sum(1, keepdim=True)
    per_list_arp = L._divide_no_nan(
        (position * weighted_labels).sum(1, keepdim=True), per_list_weights)
    return per_list_arp, per_list_weights


class RecallMetric(_RankingMetric):
  """metrics_impl.py:539-561."""

  def __init__(self, name=None, topn=None):
    self._topn = topn

  def _compute_impl(self, labels, predictions, weights, mask):
    topn = predictions.shape[1] if self._topn is None else self._topn
    return (_per_list_recall(labels, predictions, topn, mask),
            _per_example_weights_to_per_list_weights(weights, _binary(labels)))


class PrecisionMetric(_RankingMetric):
  """metrics_impl.py:564-586."""

  def __init__(self, name=None, topn=None):
    self._topn = topn

  def _compute_impl(self, labels, predictions, weights, mask):
    topn = predictions.shape[1] if self._topn is None else self._topn
    return (_per_list_precision(labels, predictions, topn, mask),
            _per_example_weights_to_per_list_weights(weights, _binary(labels)))


class MeanAveragePrecisionMetric(_RankingMetric):
  """metrics_impl.py:589-628."""

  def __init__(self, name=None, topn=None):
    self._topn = topn

  def _compute_impl(self, labels, predictions, weights, mask):
    topn = predictions.shape[1] if self._topn is None else self._topn
    relevance = _binary(labels)
    sorted_relevance, sorted_weights = utils.sort_by_scores(
        predictions, [relevance, weights], topn=topn, mask=mask)
    per_list_relevant_counts = torch.cumsum(sorted_relevance, 1)
    per_list_cutoffs = torch.cumsum(torch.ones_like(sorted_relevance), 1)
    per_list_precisions = L._divide_no_nan(per_list_relevant_counts,
                                           per_list_cutoffs)
    total_precision = (per_list_precisions * sorted_weights *
                       sorted_relevance).sum(1, keepdim=True)
    total_relevance = (weights * relevance).sum(1, keepdim=True)
    per_list_map = L._divide_no_nan(total_precision, total_relevance)
    return per_list_map, _per_example_weights_to_per_list_weights(weights,
                                                                  relevance)


class BPrefMetric(_RankingMetric):
  """metrics_impl.py:825-898."""

  def __init__(self, name=None, topn=None, use_trec_version=True):
    self._topn = topn
    self._use_trec_version = use_trec_version

  def _compute_impl(self, labels, predictions, weights, mask):
    topn = predictions.shape[1] if self._topn is None else self._topn
    relevance = _binary(labels)
    irrelevance = mask.to(relevance.dtype) - relevance
    total_relevance = relevance.sum(1, keepdim=True)
    total_irrelevance = irrelevance.sum(1, keepdim=True)
    sorted_relevance, sorted_irrelevance = utils.sort_by_scores(
        predictions, [relevance, irrelevance], mask=mask, topn=topn)
    numerator = torch.minimum(torch.cumsum(sorted_irrelevance, 1), total_relevance)
    denominator = (torch.minimum(total_irrelevance, total_relevance)
                   if self._use_trec_version else total_relevance)
    bpref = L._divide_no_nan(
        ((1. - L._divide_no_nan(numerator, denominator.expand_as(numerator))) *
         sorted_relevance).sum(1, keepdim=True), total_relevance)
    return bpref, _per_example_weights_to_per_list_weights(weights, relevance)


class DCGMetric(_RankingMetric):
  """metrics_impl.py:673-705."""

  def __init__(self, name=None, topn=None, gain_fn=_DEFAULT_GAIN_FN,
               rank_discount_fn=_DEFAULT_RANK_DISCOUNT_FN):
    self._topn = topn
    self._gain_fn = gain_fn
    self._rank_discount_fn = rank_discount_fn

  def _compute_impl(self, labels, predictions, weights, mask):
    topn = predictions.shape[1] if self._topn is None else self._topn
    sorted_labels, sorted_weights = utils.sort_by_scores(
        predictions, [labels, weights], topn=topn, mask=mask)
    dcg = _discounted_cumulative_gain(sorted_labels, sorted_weights,
                                      self._gain_fn, self._rank_discount_fn)
    per_list_weights = _per_example_weights_to_per_list_weights(
        weights=weights, relevance=self._gain_fn(labels))
    return L._divide_no_nan(dcg, per_list_weights), per_list_weights


class OPAMetric(_RankingMetric):
  """metrics_impl.py:708-743."""

  def __init__(self, name=None):
    pass

  def _compute_impl(self, labels, predictions, weights, mask):
    dt = predictions.dtype
    valid_pair = torch.logical_and(mask.unsqueeze(2), mask.unsqueeze(1))
    pair_label_diff = labels.unsqueeze(2) - labels.unsqueeze(1)
    pair_pred_diff = predictions.unsqueeze(2) - predictions.unsqueeze(1)
    correct_pairs = (pair_label_diff > 0).to(dt) * (pair_pred_diff > 0).to(dt)
    pair_weights = (pair_label_diff > 0).to(dt) * weights.unsqueeze(2) * valid_pair.to(dt)
    per_list_weights = pair_weights.sum((1, 2)).unsqueeze(1)
    per_list_opa = L._divide_no_nan(
        (correct_pairs * pair_weights).sum((1, 2)).unsqueeze(1), per_list_weights)
    return per_list_opa, per_list_weights


def _alpha_dcg_gain_fn(labels, alpha):
  """metrics_impl.py:36-60."""
  cum_subtopics = torch.cumsum(labels, 1) - labels        # exclusive
  return (labels * torch.pow(torch.as_tensor(1. - alpha, dtype=labels.dtype),
                             cum_subtopics)).sum(-1)


class _DivRankingMetric(_RankingMetric):
  """metrics_impl.py:313-427."""

  def __init__(self, name=None, topn=None):
    self._topn = topn

  def _prepare_and_validate_params(self, labels, predictions, weights, mask):
    predictions = torch.as_tensor(predictions)
    labels = torch.as_tensor(labels, dtype=predictions.dtype)
    assert labels.dim() == 3
    if mask is None:
      mask = utils.is_label_valid(labels)
    mask = torch.as_tensor(mask)
    if mask.dim() == 3:
      mask = mask.any(dim=2)
    predictions = torch.where(
        mask, predictions, -1e-6 * torch.ones_like(predictions) +
        predictions.min(dim=1, keepdim=True).values)
    labels = torch.where(mask.unsqueeze(2), labels, torch.zeros_like(labels))
    weights = 1.0 if weights is None else torch.as_tensor(weights, dtype=predictions.dtype)
    return labels, predictions, torch.ones_like(predictions) * weights, mask

  def _compute_per_list_weights(self, weights, labels):
    return _per_example_weights_to_per_list_weights(
        weights, (labels >= 1.0).any(dim=-1).to(weights.dtype))

  def _compute_impl(self, labels, predictions, weights, mask):
    topn = predictions.shape[1] if self._topn is None else self._topn
    return (self._compute_per_list_metric(labels, predictions, weights, topn, mask),
            self._compute_per_list_weights(weights, labels))


class PrecisionIAMetric(_DivRankingMetric):
  """metrics_impl.py:746-782."""

  def _compute_per_list_metric(self, labels, predictions, weights, topn, mask):
    sorted_labels = utils.sort_by_scores(predictions, [labels], topn=topn, mask=mask)[0]
    relevance = (sorted_labels >= 1.0).to(predictions.dtype).sum(-1)
    num_subtopics = (labels >= 1.0).any(dim=1, keepdim=True).to(predictions.dtype).sum(-1)
    valid_topn = torch.clamp(mask.to(torch.int64).sum(1, keepdim=True), max=topn)
    return L._divide_no_nan(
        relevance.sum(1, keepdim=True),
        (valid_topn.to(predictions.dtype) * num_subtopics).sum(1, keepdim=True))


class AlphaDCGMetric(_DivRankingMetric):
  """metrics_impl.py:785-822."""

  def __init__(self, name=None, topn=None, alpha=0.5,
               rank_discount_fn=_DEFAULT_RANK_DISCOUNT_FN, seed=None):
    super().__init__(name, topn)
    self._alpha = alpha
    self._rank_discount_fn = rank_discount_fn

  def _compute_per_list_metric(self, labels, predictions, weights, topn, mask):
    sorted_labels, sorted_weights = utils.sort_by_scores(
        predictions, [labels, weights], topn=topn, mask=mask)
    alpha_dcg = _discounted_cumulative_gain(
        sorted_labels, sorted_weights, lambda l: _alpha_dcg_gain_fn(l, self._alpha),
        self._rank_discount_fn)
    return L._divide_no_nan(alpha_dcg, self._compute_per_list_weights(weights, labels))


class KerasMean(object):
  """tf.keras.metrics.Mean over (values, sample_weight): keras/metrics.py:171-193."""

  def __init__(self, metric):
    self._metric = metric
    self.reset_state()

  def reset_state(self):
    self.total = 0.0
    self.count = 0.0

  def update_state(self, y_true, y_pred, sample_weight=None):
    v, w = self._metric.compute(y_true, y_pred, sample_weight)
    self.total += float((v * w).sum())
    self.count += float(w.sum())

  def result(self):
    return self.total / self.count if self.count != 0 else 0.0
