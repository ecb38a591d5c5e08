"""Oracle restatement of tensorflow_ranking/python/losses_impl.py (hot-path subset).

Test infrastructure only (see oracle/__init__.py).  Same [B, N, N] algorithm as
the reference; gradients come from torch autograd.  All computation happens in
the dtype of `logits` (fp32 to mimic TF, fp64 for a high-precision check).
"""
import math

import torch

from oracle import utils

_EPSILON = 1e-10  # losses_impl.py:30


# ----------------------------------------------------------------------------
# gain / discount functions (keras/utils.py:50-135 and the lambdas in
# losses_impl.py:110-111,224-225, losses.py:454-455)
# ----------------------------------------------------------------------------
def identity(label):
  return label


def inverse(rank):
  """keras/utils.py:65-76: divide_no_nan(1, rank)."""
  rank = torch.as_tensor(rank)
  return torch.where(rank == 0, torch.zeros_like(rank, dtype=_f(rank)),
                     1. / rank.to(_f(rank)))


def pow_minus_1(label):
  """keras/utils.py:79-91."""
  return torch.pow(2., label) - 1.


def log2_inverse(rank):
  """keras/utils.py:94-107: divide_no_nan(log 2, log1p(rank))."""
  rank = torch.as_tensor(rank)
  rank = rank.to(_f(rank))
  d = torch.log1p(rank)
  return torch.where(d == 0, torch.zeros_like(d), math.log(2.) / d)


def log1p_inverse(rank):
  """losses_impl.py:111 / losses.py:455: 1 / log1p(rank)."""
  rank = torch.as_tensor(rank)
  return 1. / torch.log1p(rank.to(_f(rank)))


def is_greater_equal_1(label):
  return label >= 1.0


def _f(t):
  return t.dtype if t.is_floating_point() else torch.float32


def _divide_no_nan(x, y):
  y = torch.as_tensor(y, dtype=x.dtype) if not torch.is_tensor(y) else y
  return torch.where(y == 0, torch.zeros_like(x * y), x / torch.where(
      y == 0, torch.ones_like(y), y))


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------
def _safe_default_gain_fn(labels):
  """losses_impl.py:33-49."""
  max_labels = labels.max(dim=-1, keepdim=True).values
  return torch.pow(2., labels - max_labels) - torch.pow(2., -max_labels)


def _apply_pairwise_op(op, tensor):
  """losses_impl.py:61-64: entry [b, i, j] = op(t[b, i], t[b, j])."""
  assert tensor.dim() == 2
  return op(tensor.unsqueeze(2), tensor.unsqueeze(1))


def _get_valid_pairs_and_clean_labels(labels):
  """losses_impl.py:67-74."""
  is_valid = utils.is_label_valid(labels)
  valid_pairs = _apply_pairwise_op(torch.logical_and, is_valid)
  labels = torch.where(is_valid, labels, torch.zeros_like(labels))
  return valid_pairs, labels


def approx_ranks(logits):
  """losses_impl.py:77-106: r_i = sum_j sigmoid(s_j - s_i) + 0.5."""
  x = logits.unsqueeze(2)
  y = logits.unsqueeze(1)
  return torch.sigmoid(y - x).sum(-1) + .5


def inverse_max_dcg(labels, gain_fn=pow_minus_1, rank_discount_fn=log1p_inverse,
                    topn=None):
  """losses_impl.py:109-134."""
  ideal_sorted_labels, = utils.sort_by_scores(labels, [labels], topn=topn)
  rank = torch.arange(ideal_sorted_labels.shape[1]) + 1
  discounted_gain = gain_fn(ideal_sorted_labels) * rank_discount_fn(
      rank.to(labels.dtype))
  discounted_gain = discounted_gain.sum(1, keepdim=True)
  return torch.where(discounted_gain > 0., 1. / discounted_gain,
                     torch.zeros_like(discounted_gain))


def ndcg(labels, ranks=None):
  """losses_impl.py:137-167 (perm_mat branch is not on the path)."""
  if ranks is None:
    ranks = torch.arange(labels.shape[1]) + 1
  discounts = 1. / torch.log1p(ranks.to(labels.dtype))
  gains = _safe_default_gain_fn(labels)
  dcg = (gains * discounts).sum(-1, keepdim=True)
  return dcg * inverse_max_dcg(labels, gain_fn=_safe_default_gain_fn)


# ----------------------------------------------------------------------------
# LambdaWeight family (losses_impl.py:170-454)
# ----------------------------------------------------------------------------
class _LambdaWeight(object):

  def pair_weights(self, labels, ranks):
    raise NotImplementedError

  def individual_weights(self, labels, ranks):
    """losses_impl.py:195-207."""
    return labels


class LabelDiffLambdaWeight(_LambdaWeight):
  """losses_impl.py:210-216."""

  def pair_weights(self, labels, ranks):
    return torch.abs(_apply_pairwise_op(torch.sub, labels))


class AbstractDCGLambdaWeight(_LambdaWeight):
  """losses_impl.py:219-296."""

  def __init__(self, topn=None, gain_fn=identity, rank_discount_fn=inverse,
               normalized=False):
    self._topn = topn
    self._gain_fn = gain_fn
    self._rank_discount_fn = rank_discount_fn
    self._normalized = normalized

  def _pair_rank_discount(self, ranks, topn):
    raise NotImplementedError

  def pair_weights(self, labels, ranks):
    """losses_impl.py:255-279."""
    valid_pair, labels = _get_valid_pairs_and_clean_labels(labels)
    gain = self._gain_fn(labels)
    if self._normalized:
      gain = gain * inverse_max_dcg(
          labels, gain_fn=self._gain_fn,
          rank_discount_fn=self._rank_discount_fn, topn=self._topn)
    pair_gain = _apply_pairwise_op(torch.sub, gain)
    pair_gain = pair_gain * valid_pair.to(labels.dtype)
    list_size = labels.shape[1]
    topn = self._topn or list_size
    pair_weight = torch.abs(pair_gain) * self._pair_rank_discount(
        ranks, topn).to(labels.dtype)
    # Scale by the (padded) list size: losses_impl.py:274-278.
    return pair_weight * float(list_size)

  def individual_weights(self, labels, ranks):
    """losses_impl.py:281-296."""
    labels = torch.where(utils.is_label_valid(labels), labels,
                         torch.zeros_like(labels))
    gain = self._gain_fn(labels)
    if self._normalized:
      gain = gain * inverse_max_dcg(
          labels, gain_fn=self._gain_fn,
          rank_discount_fn=self._rank_discount_fn, topn=self._topn)
    rank_discount = self._rank_discount_fn(ranks.to(labels.dtype))
    return gain * rank_discount


class DCGLambdaWeight(AbstractDCGLambdaWeight):
  """losses_impl.py:299-369."""

  def __init__(self, topn=None, gain_fn=identity, rank_discount_fn=inverse,
               normalized=False, smooth_fraction=0.):
    super().__init__(topn, gain_fn, rank_discount_fn, normalized)
    if not 0. <= smooth_fraction <= 1.:
      raise ValueError('smooth_fraction %s should be in range [0, 1].' %
                       smooth_fraction)
    self._smooth_fraction = smooth_fraction

  def _pair_rank_discount(self, ranks, topn):
    fdt = torch.float64
    # u: relative rank difference (losses_impl.py:337-352).
    pair_valid_rank = _apply_pairwise_op(torch.logical_or, ranks <= topn)
    rank_diff = torch.abs(_apply_pairwise_op(torch.sub, ranks)).to(fdt)
    safe_diff = torch.where(rank_diff > 0, rank_diff, torch.ones_like(rank_diff))
    u = torch.where(
        torch.logical_and(rank_diff > 0, pair_valid_rank),
        torch.abs(self._rank_discount_fn(safe_diff) -
                  self._rank_discount_fn(safe_diff + 1)),
        torch.zeros_like(rank_diff))
    # v: absolute rank (losses_impl.py:354-363).
    rank_discount = torch.where(ranks > topn, torch.zeros_like(ranks, dtype=fdt),
                                self._rank_discount_fn(ranks.to(fdt)))
    v = torch.abs(_apply_pairwise_op(torch.sub, rank_discount))
    pair_discount = (1. - self._smooth_fraction) * u + self._smooth_fraction * v
    pair_mask = _apply_pairwise_op(torch.logical_or, ranks <= topn)
    return pair_discount * pair_mask.to(fdt)


class DCGLambdaWeightV2(AbstractDCGLambdaWeight):
  """losses_impl.py:372-394."""

  def _pair_rank_discount(self, ranks, topn):
    fdt = torch.float64
    rank_diff = torch.abs(_apply_pairwise_op(torch.sub, ranks)).to(fdt)
    max_rank = _apply_pairwise_op(torch.maximum, ranks).to(fdt)
    multiplier = torch.where(max_rank > float(topn),
                             1. / (1. - self._rank_discount_fn(max_rank)),
                             torch.ones_like(max_rank))
    safe_diff = torch.where(rank_diff > 0, rank_diff, torch.ones_like(rank_diff))
    return torch.where(
        rank_diff > 0.,
        torch.abs(self._rank_discount_fn(safe_diff) -
                  self._rank_discount_fn(safe_diff + 1)) * multiplier,
        torch.zeros_like(rank_diff))


class YetiDCGLambdaWeight(DCGLambdaWeightV2):
  """losses_impl.py:397-407."""

  def pair_weights(self, labels, ranks):
    pair_weight = super().pair_weights(labels, ranks)
    neighbor_pair = torch.abs(_apply_pairwise_op(torch.sub, ranks)) == 1
    return pair_weight * neighbor_pair.to(pair_weight.dtype)


class PrecisionLambdaWeight(_LambdaWeight):
  """losses_impl.py:410-454."""

  def __init__(self, topn, positive_fn=is_greater_equal_1):
    self._topn = topn
    self._positive_fn = positive_fn

  def pair_weights(self, labels, ranks):
    valid_pair, labels = _get_valid_pairs_and_clean_labels(labels)
    binary_labels = self._positive_fn(labels).to(labels.dtype)
    label_diff = torch.abs(_apply_pairwise_op(torch.sub, binary_labels))
    label_diff = label_diff * valid_pair.to(labels.dtype)
    rank_mask = _apply_pairwise_op(torch.logical_xor, ranks <= self._topn)
    return label_diff * rank_mask.to(labels.dtype)


# ----------------------------------------------------------------------------
# ranks / pairwise comparison (losses_impl.py:483-537)
# ----------------------------------------------------------------------------
def _compute_ranks(logits, is_valid):
  """losses_impl.py:483-500."""
  logits = logits.detach()
  scores = torch.where(
      is_valid, logits,
      -1e-6 * torch.ones_like(logits) + logits.min(dim=1, keepdim=True).values)
  return utils.sorted_ranks(scores)


def _pairwise_comparison(labels, logits, mask):
  """losses_impl.py:503-537."""
  pairwise_label_diff = _apply_pairwise_op(torch.sub, labels)
  pairwise_logits = _apply_pairwise_op(torch.sub, logits)
  pairwise_labels = (pairwise_label_diff > 0).to(logits.dtype)
  valid_pair = _apply_pairwise_op(torch.logical_and, mask)
  pairwise_labels = pairwise_labels * valid_pair.to(logits.dtype)
  return pairwise_labels, pairwise_logits


# ----------------------------------------------------------------------------
# tf.compat.v1.losses.compute_weighted_loss semantics
# ----------------------------------------------------------------------------
class Reduction(object):
  NONE = 'none'
  SUM = 'weighted_sum'
  MEAN = 'weighted_mean'
  SUM_OVER_BATCH_SIZE = 'weighted_sum_over_batch_size'
  SUM_BY_NONZERO_WEIGHTS = 'weighted_sum_by_nonzero_weights'


def compute_weighted_loss(losses, weights, reduction):
  weights = torch.as_tensor(weights, dtype=losses.dtype)
  weighted = losses * weights
  if reduction == Reduction.NONE:
    return weighted
  total = weighted.sum()
  if reduction == Reduction.SUM:
    return total
  if reduction == Reduction.MEAN:
    denom = (torch.ones_like(losses) * weights).sum()
  elif reduction == Reduction.SUM_BY_NONZERO_WEIGHTS:
    denom = ((torch.ones_like(losses) * weights) != 0).to(losses.dtype).sum()
  elif reduction == Reduction.SUM_OVER_BATCH_SIZE:
    denom = torch.tensor(float(losses.numel()), dtype=losses.dtype)
  else:
    raise ValueError('bad reduction %r' % (reduction,))
  return torch.where(denom > 0, total / torch.where(
      denom > 0, denom, torch.ones_like(denom)), torch.zeros_like(total))


# ----------------------------------------------------------------------------
# _RankingLoss (losses_impl.py:652-860)
# ----------------------------------------------------------------------------
class _RankingLoss(object):

  def __init__(self, name=None, lambda_weight=None, temperature=1.0):
    self._name = name
    self._lambda_weight = lambda_weight
    self._temperature = temperature

  def _prepare_and_validate_params(self, labels, logits, weights, mask):
    """losses_impl.py:674-707."""
    logits = torch.as_tensor(logits)
    labels = torch.as_tensor(labels, dtype=logits.dtype)
    if mask is None:
      mask = utils.is_label_valid(labels)
    if weights is None:
      weights = 1.0
    weights = torch.as_tensor(weights, dtype=logits.dtype)
    mask = torch.as_tensor(mask)
    return labels, logits, weights, mask

  def compute_unreduced_loss(self, labels, logits, mask=None):
    """losses_impl.py:709-726."""
    labels, logits, _, mask = self._prepare_and_validate_params(
        labels, logits, None, mask)
    return self._compute_unreduced_loss_impl(labels, logits, mask)

  def normalize_weights(self, labels, weights):
    """losses_impl.py:745-766."""
    return self._normalize_weights_impl(torch.as_tensor(labels), weights)

  def _normalize_weights_impl(self, labels, weights):
    return 1.0 if weights is None else weights

  def get_logits(self, logits):
    """losses_impl.py:773-785."""
    return torch.as_tensor(logits) / self._temperature

  def compute(self, labels, logits, weights, reduction, mask=None):
    """losses_impl.py:787-814 (estimator-style reduced loss)."""
    labels, logits, _, mask = self._prepare_and_validate_params(
        labels, logits, None, mask)
    logits = self.get_logits(logits)
    losses, loss_weights = self._compute_unreduced_loss_impl(
        labels, logits, mask)
    nw = self._normalize_weights_impl(labels, weights)
    weights = torch.as_tensor(nw, dtype=logits.dtype) * loss_weights
    return compute_weighted_loss(losses, weights, reduction)


class _PairwiseLoss(_RankingLoss):
  """losses_impl.py:863-930."""

  def _pairwise_loss(self, pairwise_logits):
    raise NotImplementedError

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    ranks = _compute_ranks(logits, mask)
    pairwise_labels, pairwise_logits = _pairwise_comparison(
        labels, logits, mask)
    pairwise_weights = pairwise_labels
    if self._lambda_weight is not None:
      pairwise_weights = pairwise_weights * self._lambda_weight.pair_weights(
          labels, ranks).to(logits.dtype)
    pairwise_weights = pairwise_weights.detach()
    return self._pairwise_loss(pairwise_logits), pairwise_weights

  def compute_per_list(self, labels, logits, weights, mask=None):
    """losses_impl.py:886-915.  (No temperature here, as in the reference.)"""
    labels, logits, weights, mask = self._prepare_and_validate_params(
        labels, logits, weights, mask)
    losses, loss_weights = self._compute_unreduced_loss_impl(
        labels, logits, mask)
    weights = self._normalize_weights_impl(labels, weights) * loss_weights
    per_list_weights = weights.sum(dim=[1, 2])
    per_list_losses = (losses * weights).sum(dim=[1, 2])
    return _divide_no_nan(per_list_losses, per_list_weights), per_list_weights

  def _normalize_weights_impl(self, labels, weights):
    """losses_impl.py:917-930: row-item weights, [B, N, 1]."""
    if weights is None:
      weights = 1.
    labels = torch.as_tensor(labels)
    weights = torch.as_tensor(weights, dtype=labels.dtype)
    weights = torch.where(utils.is_label_valid(labels),
                          torch.ones_like(labels) * weights,
                          torch.zeros_like(labels))
    return weights.unsqueeze(2)


class PairwiseLogisticLoss(_PairwiseLoss):
  """losses_impl.py:933-940."""

  def _pairwise_loss(self, pairwise_logits):
    return torch.relu(-pairwise_logits) + torch.log1p(
        torch.exp(-torch.abs(pairwise_logits)))


class PairwiseHingeLoss(_PairwiseLoss):
  """losses_impl.py:943-948."""

  def _pairwise_loss(self, pairwise_logits):
    return torch.relu(1 - pairwise_logits)


class PairwiseSoftZeroOneLoss(_PairwiseLoss):
  """losses_impl.py:951-958."""

  def _pairwise_loss(self, pairwise_logits):
    return torch.where(pairwise_logits > 0, 1. - torch.sigmoid(pairwise_logits),
                       torch.sigmoid(-pairwise_logits))


class PairwiseMSELoss(_PairwiseLoss):
  """losses_impl.py:961-998."""

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    pairwise_label_diff = _apply_pairwise_op(torch.sub, labels)
    pairwise_logit_diff = _apply_pairwise_op(torch.sub, logits)
    pairwise_mse_loss = torch.square(pairwise_logit_diff - pairwise_label_diff)
    valid_pair = _apply_pairwise_op(torch.logical_and, mask)
    pairwise_weights = torch.ones_like(pairwise_mse_loss)
    pairwise_weights = pairwise_weights - torch.eye(
        labels.shape[1], dtype=logits.dtype).unsqueeze(0)
    pairwise_weights = pairwise_weights * valid_pair.to(logits.dtype)
    if self._lambda_weight is not None:
      ranks = _compute_ranks(logits, mask)
      pairwise_weights = pairwise_weights * self._lambda_weight.pair_weights(
          labels, ranks).to(logits.dtype)
    return pairwise_mse_loss, pairwise_weights.detach()


class _ListwiseLoss(_RankingLoss):
  """losses_impl.py:1001-1033."""

  def _normalize_weights_impl(self, labels, weights):
    if weights is None:
      return 1.0
    labels = torch.as_tensor(labels)
    weights = torch.as_tensor(weights, dtype=labels.dtype)
    is_valid = utils.is_label_valid(labels)
    labels = torch.where(is_valid, labels, torch.zeros_like(labels))
    return _divide_no_nan((weights * labels).sum(1, keepdim=True),
                          labels.sum(1, keepdim=True))

  def compute_per_list(self, labels, logits, weights, mask=None):
    labels, logits, weights, mask = self._prepare_and_validate_params(
        labels, logits, weights, mask)
    losses, loss_weights = self._compute_unreduced_loss_impl(
        labels, logits, mask)
    weights = torch.as_tensor(self._normalize_weights_impl(labels, weights),
                              dtype=logits.dtype) * loss_weights
    return losses.squeeze(1), weights.squeeze(1)


class SoftmaxLoss(_ListwiseLoss):
  """losses_impl.py:1119-1197."""

  def precompute(self, labels, logits, weights, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    ranks = _compute_ranks(logits, mask)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    logits = torch.where(mask, logits,
                         math.log(_EPSILON) * torch.ones_like(logits))
    if self._lambda_weight is not None and isinstance(self._lambda_weight,
                                                      DCGLambdaWeight):
      labels = self._lambda_weight.individual_weights(labels, ranks)
    if weights is not None:
      labels = labels * weights
    return labels, logits

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    label_sum = labels.sum(1, keepdim=True)
    nonzero_mask = label_sum.reshape(-1) > 0.0
    padded_labels = torch.where(nonzero_mask.unsqueeze(1), labels,
                                _EPSILON * torch.ones_like(labels))
    padded_labels = torch.where(mask, padded_labels,
                                torch.zeros_like(padded_labels))
    padded_label_sum = padded_labels.sum(1, keepdim=True)
    labels_for_softmax = _divide_no_nan(padded_labels, padded_label_sum)
    losses = -(labels_for_softmax * torch.log_softmax(logits, dim=1)).sum(1)
    return losses, label_sum.reshape(-1)

  def compute(self, labels, logits, weights, reduction, mask=None):
    labels, logits, weights, mask = self._prepare_and_validate_params(
        labels, logits, weights, mask)
    logits = self.get_logits(logits)
    labels, logits = self.precompute(labels, logits, weights, mask)
    losses, weights = self._compute_unreduced_loss_impl(labels, logits, mask)
    return compute_weighted_loss(losses, weights, reduction)

  def compute_per_list(self, labels, logits, weights, mask=None):
    labels, logits, weights, mask = self._prepare_and_validate_params(
        labels, logits, weights, mask)
    logits = self.get_logits(logits)
    labels, logits = self.precompute(labels, logits, weights, mask)
    return self._compute_unreduced_loss_impl(labels, logits, mask)

  def compute_unreduced_loss(self, labels, logits, mask=None):
    labels, logits, _, mask = self._prepare_and_validate_params(
        labels, logits, None, mask)
    logits = self.get_logits(logits)
    labels, logits = self.precompute(labels, logits, weights=None, mask=mask)
    return self._compute_unreduced_loss_impl(labels, logits, mask)


class ApproxNDCGLoss(_ListwiseLoss):
  """losses_impl.py:1579-1603."""

  def __init__(self, name=None, lambda_weight=None, temperature=0.1):
    super().__init__(name, lambda_weight, temperature)

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    logits = torch.where(
        mask, logits,
        -1e3 * torch.ones_like(logits) + logits.min(dim=-1, keepdim=True).values)
    label_sum = labels.sum(1, keepdim=True)
    nonzero_mask = label_sum.reshape(-1) > 0.0
    labels = torch.where(nonzero_mask.unsqueeze(1), labels,
                         _EPSILON * torch.ones_like(labels))
    ranks = approx_ranks(logits)
    return -ndcg(labels, ranks), nonzero_mask.to(logits.dtype).reshape(-1, 1)


class ApproxMRRLoss(_ListwiseLoss):
  """losses_impl.py:1606-1632."""

  def __init__(self, name=None, lambda_weight=None, temperature=0.1):
    super().__init__(name, lambda_weight, temperature)

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    logits = torch.where(
        mask, logits,
        -1e3 * torch.ones_like(logits) + logits.min(dim=-1, keepdim=True).values)
    label_sum = labels.sum(1, keepdim=True)
    nonzero_mask = label_sum.reshape(-1) > 0.0
    labels = torch.where(nonzero_mask.unsqueeze(1), labels,
                         _EPSILON * torch.ones_like(labels))
    rr = 1. / approx_ranks(logits)
    rr = (rr * labels).sum(-1, keepdim=True)
    mrr = rr / labels.sum(-1, keepdim=True)
    return -mrr, nonzero_mask.to(logits.dtype).reshape(-1, 1)


# ----------------------------------------------------------------------------
# Pointwise losses (losses_impl.py:1284-1469)
# ----------------------------------------------------------------------------
class _PointwiseLoss(_RankingLoss):
  """losses_impl.py:1284-1321."""

  def _normalize_weights_impl(self, labels, weights):
    labels = torch.as_tensor(labels)
    if weights is None:
      weights = 1.
    weights = torch.as_tensor(weights, dtype=labels.dtype)
    return torch.where(utils.is_label_valid(labels),
                       torch.ones_like(labels) * weights,
                       torch.zeros_like(labels))

  def compute_per_list(self, labels, logits, weights, mask=None):
    labels, logits, weights, mask = self._prepare_and_validate_params(
        labels, logits, weights, mask)
    losses, loss_weights = self._compute_unreduced_loss_impl(
        labels, logits, mask)
    weights = self._normalize_weights_impl(labels, weights) * loss_weights
    per_list_weights = weights.sum(1)
    per_list_losses = (losses * weights).sum(1)
    return _divide_no_nan(per_list_losses, per_list_weights), per_list_weights


class SigmoidCrossEntropyLoss(_PointwiseLoss):
  """losses_impl.py:1425-1446."""

  def __init__(self, name=None, temperature=1.0):
    super().__init__(name, None, temperature)

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    logits = torch.where(mask, logits, torch.zeros_like(logits))
    # tf.nn.sigmoid_cross_entropy_with_logits: max(x, 0) - x z + log(1 + exp(-|x|))
    losses = torch.clamp(logits, min=0) - logits * labels + torch.log1p(
        torch.exp(-logits.abs()))
    return losses, mask.to(logits.dtype)


class MeanSquaredLoss(_PointwiseLoss):
  """losses_impl.py:1449-1469 (temperature is not used)."""

  def __init__(self, name=None):
    super().__init__(name, None, 1.0)

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    logits = torch.where(mask, logits, torch.zeros_like(logits))
    return (labels - logits) ** 2, mask.to(logits.dtype)


# ----------------------------------------------------------------------------
# UniqueSoftmaxLoss (losses_impl.py:1250-1281), ListMLELoss (:1541-1576)
# ----------------------------------------------------------------------------
class UniqueSoftmaxLoss(_ListwiseLoss):

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    logits = torch.where(mask, logits,
                         math.log(_EPSILON) * torch.ones_like(logits))
    pairwise_labels, _ = _pairwise_comparison(labels, logits, mask)
    denominator_logits = logits.unsqueeze(1) * pairwise_labels
    denominator_logits = torch.cat([denominator_logits, logits.unsqueeze(2)], 2)
    denominator_mask = torch.cat(
        [pairwise_labels, torch.ones_like(logits).unsqueeze(2)], 2)
    denominator_logits = torch.where(
        denominator_mask > 0.0, denominator_logits,
        -1e-3 + denominator_logits.min() * torch.ones_like(denominator_logits))
    logits_max = denominator_logits.max(dim=-1, keepdim=True).values
    denominator_logits = denominator_logits - logits_max
    logits = logits - logits_max.squeeze(-1)
    gains = torch.pow(2.0, labels) - 1
    per_doc_softmax = -logits + torch.log(
        (torch.exp(denominator_logits) * denominator_mask).sum(-1))
    losses = (per_doc_softmax * gains).sum(1, keepdim=True)
    return losses, torch.ones_like(losses)


class ListMLELambdaWeight(_LambdaWeight):
  """losses_impl.py:457-480."""

  def __init__(self, rank_discount_fn):
    self._rank_discount_fn = rank_discount_fn

  def individual_weights(self, labels, ranks):
    labels = torch.as_tensor(labels)
    return torch.ones_like(labels) * self._rank_discount_fn(
        torch.as_tensor(ranks).to(labels.dtype))


class ListMLELoss(_ListwiseLoss):
  """losses_impl.py:1541-1576.  The reference breaks label ties randomly
  (`shuffle_ties=True, seed=37`); this restatement breaks them by index, which is
  one of the orders the reference can draw."""

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    logits = torch.where(mask, logits,
                         math.log(_EPSILON) * torch.ones_like(logits))
    scores = torch.where(
        mask, labels,
        labels.min(dim=1, keepdim=True).values - 1e-6 * torch.ones_like(labels))
    sorted_labels, sorted_logits = utils.sort_by_scores(scores, [labels, logits])
    raw_max = sorted_logits.max(dim=1, keepdim=True).values
    sorted_logits = sorted_logits - raw_max
    sums = torch.flip(torch.cumsum(torch.flip(torch.exp(sorted_logits), [1]), 1),
                      [1])
    sums = torch.log(sums) - sorted_logits
    if isinstance(self._lambda_weight, ListMLELambdaWeight):
      b, n = sorted_labels.shape
      ranks = (torch.arange(n) + 1).unsqueeze(0).expand(b, n)
      sums = sums * self._lambda_weight.individual_weights(sorted_labels, ranks)
    nll = sums.sum(1, keepdim=True)
    return nll, torch.ones_like(nll)


# ----------------------------------------------------------------------------
# GumbelSampler (losses_impl.py:540-649) with the uniforms passed in explicitly
# ----------------------------------------------------------------------------
class GumbelSampler(object):

  def __init__(self, name=None, sample_size=8, temperature=1.0, seed=None):
    self._sample_size = sample_size
    self._temperature = temperature

  def sample(self, labels, logits, weights=None, uniforms=None):
    """`uniforms` [B, S, N] in [0, 1): what `tf.random.uniform` returns in the
    reference (:647-649); the caller supplies them so that the restatement is
    deterministic."""
    logits = torch.as_tensor(logits)
    labels = torch.as_tensor(labels, dtype=logits.dtype)
    b, n = labels.shape
    s = self._sample_size
    expanded_labels = labels.unsqueeze(1).repeat(1, s, 1).reshape(b * s, n)
    u = torch.as_tensor(uniforms, dtype=logits.dtype).reshape(b, s, n)
    eps = 1e-20
    gumbel = -torch.log(-torch.log(u + eps) + eps)
    sampled = (logits.unsqueeze(1).repeat(1, s, 1) + gumbel).reshape(b * s, n)
    is_label_valid = utils.is_label_valid(expanded_labels)
    sampled = torch.where(is_label_valid, sampled / self._temperature,
                          math.log(1e-20) * torch.ones_like(sampled))
    sampled = torch.log(torch.softmax(sampled, dim=1) + 1e-20)
    expanded_weights = weights
    if expanded_weights is not None:
      w = torch.as_tensor(expanded_weights, dtype=logits.dtype)
      w = w.reshape(b, 1, 1) if w.dim() == 1 else w.unsqueeze(1)
      expanded_weights = w.repeat(1, s, 1).reshape(b * s, -1)
    return expanded_labels, sampled, expanded_weights


class OrdinalLoss(_PointwiseLoss):
  """losses_impl.py:1850-1918."""

  def __init__(self, name=None, ordinal_size=1, temperature=1.0,
               use_fraction_label=False):
    super().__init__(name, None, temperature)
    self._ordinal_size = ordinal_size
    self._use_fraction_label = use_fraction_label

  def _labels_to_ordinals(self, labels, mask):
    one_to_n = torch.arange(1, self._ordinal_size + 1).to(labels.dtype)
    unsqueezed = labels.unsqueeze(2).repeat(1, 1, self._ordinal_size)
    ordinals = torch.where(unsqueezed >= one_to_n, torch.ones_like(unsqueezed),
                           torch.zeros_like(unsqueezed))
    if self._use_fraction_label:
      fractions = unsqueezed - one_to_n + 1.0
      fractions = torch.where((fractions > 0.0) & (fractions < 1.0), fractions,
                              torch.zeros_like(fractions))
      ordinals = ordinals + fractions
    return torch.where(mask.unsqueeze(-1), ordinals, torch.zeros_like(ordinals))

  def _prepare_and_validate_params(self, labels, logits, weights, mask):
    logits = torch.as_tensor(logits)
    labels = torch.as_tensor(labels, dtype=logits.dtype)
    if mask is None:
      mask = utils.is_label_valid(labels)
    if weights is None:
      weights = 1.0
    return labels, logits, torch.as_tensor(weights, dtype=logits.dtype), torch.as_tensor(mask)

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    if logits.dim() != 3:
      raise ValueError('Predictions for ordinal loss must have rank 3.')
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    logits = torch.where(mask.unsqueeze(-1), logits, torch.zeros_like(logits))
    ordinals = self._labels_to_ordinals(labels, mask)
    ce = torch.clamp(logits, min=0) - logits * ordinals + torch.log1p(
        torch.exp(-logits.abs()))
    losses = torch.where(mask.unsqueeze(-1), ce, torch.zeros_like(ce))
    return losses.sum(-1), mask.to(logits.dtype)


class CoupledRankDistilLoss(_ListwiseLoss):
  """losses_impl.py:1984-2116 with the uniforms of the teacher sampler passed in
  (`uniforms` [B, S, N]); ties of the sampled teacher scores are broken by index."""

  def __init__(self, name=None, sample_size=8, topk=None, temperature=1.):
    super().__init__(name, None, temperature)
    self._sample_size = sample_size
    self._topk = topk
    self.uniforms = None            # set by the caller before compute()
    self.sampled_teacher = None     # or: the sampled teacher scores directly

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    label_sum = labels.sum(1, keepdim=True)
    nonzero_mask = label_sum.reshape(-1) > 0.0
    log_eps = math.log(_EPSILON)
    teacher_scores = torch.where(mask, labels, log_eps * torch.ones_like(labels))
    student_scores = torch.where(mask, logits, log_eps * torch.ones_like(logits))
    b, n = labels.shape
    s = self._sample_size
    if self.sampled_teacher is not None:
      sampled = torch.as_tensor(self.sampled_teacher, dtype=logits.dtype).reshape(b, s, n)
    else:
      u = torch.as_tensor(self.uniforms, dtype=logits.dtype).reshape(b, s, n)
      gumbel = -torch.log(-torch.log(u + 1e-20) + 1e-20)
      sampled = teacher_scores.unsqueeze(1).repeat(1, s, 1) + gumbel
      sampled = torch.log(torch.softmax(sampled, dim=-1) + _EPSILON)
    expanded_student = student_scores.unsqueeze(1).repeat(1, s, 1)
    sorted_student = utils.sort_by_scores(
        sampled.reshape(b * s, n), [expanded_student.reshape(b * s, n)])[0].reshape(b, s, n)
    topk = self._topk or n
    topk_student = sorted_student[:, :, :topk]
    ones_upper = torch.triu(torch.ones(topk, n, dtype=logits.dtype))
    denom_mask = ones_upper.bool().unsqueeze(0).unsqueeze(0).expand(b, s, topk, n)
    tiled = sorted_student.unsqueeze(2).expand(b, s, topk, n)
    denom = torch.where(denom_mask, tiled, log_eps * torch.ones_like(tiled))
    logprob = topk_student - torch.logsumexp(denom, dim=3)
    nll = (-logprob.sum(2)).mean(1, keepdim=True)
    return nll, nonzero_mask.to(logits.dtype).reshape(-1, 1)


# ----------------------------------------------------------------------------
# CircleLoss (losses_impl.py:1036-1116)
# ----------------------------------------------------------------------------
class CircleLoss(_ListwiseLoss):
  """losses_impl.py:1036-1116.  Scores are clipped to [0, 1] by `get_logits`
  (:1079-1082; no temperature); pairs (i, j) with label_i > label_j, both valid."""

  def __init__(self, name=None, lambda_weight=None, gamma=64, margin=0.25):
    super().__init__(name, lambda_weight, 1.0)
    self._margin = margin
    self._gamma = gamma

  def get_logits(self, logits):
    return torch.clamp(torch.as_tensor(logits), 0., 1.)

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    si = logits.unsqueeze(2)          # score_i of entry [b, i, j]
    sj = logits.unsqueeze(1)
    alpha_i = torch.relu(1 - si + self._margin).detach()       # :1090-1091
    alpha_j = torch.relu(sj + self._margin).detach()           # :1092-1093
    pairwise_logits = alpha_i * (1 - si - self._margin) + alpha_j * (sj - self._margin)
    pairwise_labels, _ = _pairwise_comparison(labels, logits, mask)
    pairwise_weights = pairwise_labels.detach()
    losses = torch.exp(self._gamma * pairwise_logits)
    per_list_losses = torch.log1p((losses * pairwise_weights).sum((1, 2)))
    # :1108-1110  0 / 0 = NaN for a list without a valid pair, as in the reference
    per_list_weights = pairwise_weights.sum((1, 2)) / (pairwise_weights > 0).to(
        logits.dtype).sum((1, 2))
    return per_list_losses.unsqueeze(1), per_list_weights.unsqueeze(1)


# ----------------------------------------------------------------------------
# NeuralSort (losses_impl.py:1635-1801)
# ----------------------------------------------------------------------------
def neural_sort(logits, mask=None):
  """losses_impl.py:1711-1801: rows = ranks (valid ranks first), columns = items."""
  logits = torch.as_tensor(logits)
  if mask is None:
    mask = torch.ones_like(logits, dtype=torch.bool)
  mask = torch.as_tensor(mask)
  logits = torch.where(mask, logits, torch.zeros_like(logits))
  num_valid = mask.to(torch.int64).sum(1, keepdim=True)
  logit_diff = (logits.unsqueeze(2) - logits.unsqueeze(1)).abs()
  valid_pair = _apply_pairwise_op(torch.logical_and, mask)
  logit_diff = torch.where(valid_pair, logit_diff, torch.zeros_like(logit_diff))
  logit_diff_sum = logit_diff.sum(1, keepdim=True)                    # [B, 1, N]
  masked_range = mask.to(torch.int64).cumsum(1)
  scaling = (num_valid + 1 - 2 * masked_range).to(logits.dtype).unsqueeze(2)   # [B, N, 1]
  p_logits = scaling * logits.unsqueeze(1) - logit_diff_sum
  p_logits = torch.where(valid_pair, p_logits,
                         torch.full_like(p_logits, -math.inf))
  p_logits = torch.where(_apply_pairwise_op(torch.logical_or, mask), p_logits,
                         torch.zeros_like(p_logits))
  order = torch.argsort(mask.to(torch.int64), dim=1, descending=True, stable=True)
  p_logits = torch.gather(p_logits, 1, order.unsqueeze(2).expand_as(p_logits))
  return torch.softmax(p_logits, -1)


class NeuralSortCrossEntropyLoss(_ListwiseLoss):
  """losses_impl.py:1635-1675."""

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    logits = torch.where(mask, logits, torch.zeros_like(logits))
    label_sum = labels.sum(1, keepdim=True)
    nonzero_mask = label_sum.reshape(-1) > 0.0
    true_perm = neural_sort(labels, mask=mask)
    smooth_perm = neural_sort(logits, mask=mask)
    # softmax_cross_entropy_with_logits_v2(labels=true_perm, logits=log(1e-20 + P))
    losses = -(true_perm * torch.log_softmax(torch.log(1e-20 + smooth_perm), 2)).sum(2)
    sorted_mask = torch.sort(mask.to(logits.dtype), dim=1, descending=True).values > 0
    losses = torch.where(sorted_mask, losses, torch.zeros_like(losses))
    losses = _divide_no_nan(losses.sum(-1, keepdim=True),
                            mask.to(logits.dtype).sum(-1, keepdim=True))
    return losses, nonzero_mask.to(logits.dtype).reshape(-1, 1)


def ndcg_perm(labels, perm_mat):
  """losses_impl.py:137-167, perm_mat branch."""
  ranks = torch.arange(labels.shape[1]) + 1
  discounts = 1. / torch.log1p(ranks.to(labels.dtype))
  gains = _safe_default_gain_fn(labels)
  gains = (perm_mat * gains.unsqueeze(1)).sum(-1)
  dcg = (gains * discounts).sum(-1, keepdim=True)
  return dcg * inverse_max_dcg(labels, gain_fn=_safe_default_gain_fn)


class NeuralSortNDCGLoss(_ListwiseLoss):
  """losses_impl.py:1678-1708."""

  def _compute_unreduced_loss_impl(self, labels, logits, mask=None):
    if mask is None:
      mask = utils.is_label_valid(labels)
    labels = torch.where(mask, labels, torch.zeros_like(labels))
    logits = torch.where(mask, logits, torch.zeros_like(logits))
    label_sum = labels.sum(1, keepdim=True)
    nonzero_mask = label_sum.reshape(-1) > 0.0
    labels = torch.where(nonzero_mask.unsqueeze(1), labels,
                         _EPSILON * torch.ones_like(labels))
    smooth_perm = neural_sort(logits, mask=mask)
    return -ndcg_perm(labels, smooth_perm), nonzero_mask.to(logits.dtype).reshape(-1, 1)
