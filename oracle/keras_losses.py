"""Oracle restatement of tensorflow_ranking/python/keras/losses.py (hot-path subset).

Test infrastructure only (see oracle/__init__.py).  Reproduces the Keras
`Loss.__call__` reduction semantics the reference inherits from tf.keras:
`losses = call(y_true, y_pred)`; `losses *= sample_weight` (rank-aligned);
AUTO / SUM_OVER_BATCH_SIZE -> sum / numel(losses); SUM -> sum; NONE -> as is.
"""
import torch

from oracle import losses_impl


class Reduction(object):
  AUTO = 'auto'
  NONE = 'none'
  SUM = 'sum'
  SUM_OVER_BATCH_SIZE = 'sum_over_batch_size'


def _keras_compute_weighted_loss(losses, sample_weight, reduction):
  if sample_weight is not None:
    sample_weight = torch.as_tensor(sample_weight, dtype=losses.dtype)
    # squeeze_or_expand_dimensions: align the weight rank to the loss rank.
    if sample_weight.dim() == losses.dim() + 1 and sample_weight.shape[-1] == 1:
      sample_weight = sample_weight.squeeze(-1)
    elif sample_weight.dim() == losses.dim() - 1:
      sample_weight = sample_weight.unsqueeze(-1)
    losses = losses * sample_weight
  if reduction == Reduction.NONE:
    return losses
  total = losses.sum()
  if reduction == Reduction.SUM:
    return total
  return total / float(losses.numel())   # AUTO == SUM_OVER_BATCH_SIZE


# LambdaWeights with the Keras defaults (keras/losses.py:114-231).
class LabelDiffLambdaWeight(losses_impl.LabelDiffLambdaWeight):
  pass


class DCGLambdaWeight(losses_impl.DCGLambdaWeight):

  def __init__(self, topn=None, gain_fn=None, rank_discount_fn=None,
               normalized=False, smooth_fraction=0., **kwargs):
    super().__init__(topn, gain_fn or losses_impl.identity,
                     rank_discount_fn or losses_impl.inverse, normalized,
                     smooth_fraction)


class NDCGLambdaWeight(DCGLambdaWeight):

  def __init__(self, topn=None, gain_fn=None, rank_discount_fn=None,
               smooth_fraction=0., **kwargs):
    super().__init__(topn, gain_fn or losses_impl.pow_minus_1,
                     rank_discount_fn or losses_impl.log2_inverse,
                     normalized=True, smooth_fraction=smooth_fraction)


class NDCGLambdaWeightV2(losses_impl.DCGLambdaWeightV2):

  def __init__(self, topn=None, gain_fn=None, rank_discount_fn=None, **kwargs):
    super().__init__(topn, gain_fn or losses_impl.pow_minus_1,
                     rank_discount_fn or losses_impl.log2_inverse,
                     normalized=True)


class YetiDCGLambdaWeight(losses_impl.YetiDCGLambdaWeight):

  def __init__(self, topn=None, gain_fn=None, rank_discount_fn=None,
               normalized=False, **kwargs):
    super().__init__(topn, gain_fn or losses_impl.pow_minus_1,
                     rank_discount_fn or losses_impl.log2_inverse,
                     normalized=normalized)


class PrecisionLambdaWeight(losses_impl.PrecisionLambdaWeight):

  def __init__(self, topn=None, positive_fn=None, **kwargs):
    super().__init__(topn, positive_fn or losses_impl.is_greater_equal_1)


class _RankingLoss(object):
  """keras/losses.py:247-285."""

  def __init__(self, reduction=Reduction.AUTO, name=None):
    self.reduction = reduction
    self.name = name
    self._loss = None

  def __call__(self, y_true, y_pred, sample_weight=None):
    y_pred = torch.as_tensor(y_pred)
    y_true = torch.as_tensor(y_true, dtype=y_pred.dtype)
    sample_weight = self._loss.normalize_weights(y_true, sample_weight)
    losses = self.call(y_true, y_pred)
    return _keras_compute_weighted_loss(losses, sample_weight, self.reduction)

  def call(self, y_true, y_pred):
    y_pred = self._loss.get_logits(y_pred)
    losses, weights = self._loss.compute_unreduced_loss(
        labels=y_true, logits=y_pred)
    return losses * weights


class _PairwiseLoss(_RankingLoss):
  """keras/losses.py:288-335."""

  _impl = None

  def __init__(self, reduction=Reduction.AUTO, name=None, lambda_weight=None,
               temperature=1.0):
    super().__init__(reduction, name)
    self._loss = self._impl(name=name, lambda_weight=lambda_weight,
                            temperature=temperature)

  def call(self, y_true, y_pred):
    y_pred = self._loss.get_logits(y_pred)
    losses, weights = self._loss.compute_unreduced_loss(
        labels=y_true, logits=y_pred)
    return (losses * weights).sum(dim=2)


class PairwiseHingeLoss(_PairwiseLoss):
  _impl = losses_impl.PairwiseHingeLoss


class PairwiseLogisticLoss(_PairwiseLoss):
  _impl = losses_impl.PairwiseLogisticLoss


class PairwiseSoftZeroOneLoss(_PairwiseLoss):
  _impl = losses_impl.PairwiseSoftZeroOneLoss


class PairwiseMSELoss(_PairwiseLoss):
  _impl = losses_impl.PairwiseMSELoss


class _ListwiseLoss(_RankingLoss):
  _impl = None
  _default_temperature = 1.0

  def __init__(self, reduction=Reduction.AUTO, name=None, lambda_weight=None,
               temperature=None):
    super().__init__(reduction, name)
    if temperature is None:
      temperature = self._default_temperature
    self._loss = self._impl(name=name, lambda_weight=lambda_weight,
                            temperature=temperature)


class SoftmaxLoss(_ListwiseLoss):
  """keras/losses.py:758-832."""
  _impl = losses_impl.SoftmaxLoss

  def __call__(self, y_true, y_pred, sample_weight=None):
    y_pred = torch.as_tensor(y_pred)
    y_true = torch.as_tensor(y_true, dtype=y_pred.dtype)
    if sample_weight is not None:
      sample_weight = torch.as_tensor(sample_weight, dtype=y_pred.dtype)
    losses, sw = self._loss.compute_per_list(y_true, y_pred, sample_weight)
    return _keras_compute_weighted_loss(losses, sw, self.reduction)


class CalibratedSoftmaxLoss(SoftmaxLoss):
  """keras/losses.py:836-943."""

  def __init__(self, reduction=Reduction.AUTO, name=None, lambda_weight=None,
               temperature=1.0, virtual_label=0.0):
    super().__init__(reduction, name, lambda_weight, temperature)
    self._virtual_label = virtual_label

  def __call__(self, y_true, y_pred, sample_weight=None):
    y_pred = torch.as_tensor(y_pred)
    y_true = torch.as_tensor(y_true, dtype=y_pred.dtype)
    b = y_true.shape[0]
    y_true = torch.cat([y_true, torch.ones(b, 1, dtype=y_true.dtype) *
                        self._virtual_label], 1)
    y_pred = torch.cat([y_pred, torch.zeros(b, 1, dtype=y_pred.dtype)], 1)
    if sample_weight is not None:
      sample_weight = torch.as_tensor(sample_weight, dtype=y_pred.dtype)
      if sample_weight.dim() == 2 and sample_weight.shape[1] > 1:
        sample_weight = torch.cat(
            [sample_weight, torch.ones(b, 1, dtype=sample_weight.dtype)], 1)
    return super().__call__(y_true, y_pred, sample_weight)


class ApproxNDCGLoss(_ListwiseLoss):
  """keras/losses.py:1164-1237."""
  _impl = losses_impl.ApproxNDCGLoss
  _default_temperature = 0.1


class ApproxMRRLoss(_ListwiseLoss):
  _impl = losses_impl.ApproxMRRLoss
  _default_temperature = 0.1


class UniqueSoftmaxLoss(_ListwiseLoss):
  """keras/losses.py:946-1005."""
  _impl = losses_impl.UniqueSoftmaxLoss


class ListMLELoss(_ListwiseLoss):
  """keras/losses.py:1008-1090."""
  _impl = losses_impl.ListMLELoss


class ListMLELambdaWeight(losses_impl.ListMLELambdaWeight):
  """keras/losses.py:233-244."""

  def __init__(self, rank_discount_fn=None, **kwargs):
    super().__init__(rank_discount_fn)


class SigmoidCrossEntropyLoss(_RankingLoss):
  """keras/losses.py:1499-1546."""

  def __init__(self, reduction=Reduction.AUTO, name=None):
    super().__init__(reduction, name)
    self._loss = losses_impl.SigmoidCrossEntropyLoss(name=name)


class MeanSquaredLoss(_RankingLoss):
  """keras/losses.py:1549-1600."""

  def __init__(self, reduction=Reduction.AUTO, name=None):
    super().__init__(reduction, name)
    self._loss = losses_impl.MeanSquaredLoss(name=name)


class _GumbelMixin(object):
  """keras/losses.py:609-718, 1241-1341; `uniforms` must be supplied per call."""

  def _init_gumbel(self, sample_size, gumbel_temperature):
    self._gumbel_sampler = losses_impl.GumbelSampler(
        sample_size=sample_size, temperature=gumbel_temperature)

  def __call__(self, y_true, y_pred, sample_weight=None, uniforms=None):
    gl, gs, gw = self._gumbel_sampler.sample(y_true, y_pred, sample_weight,
                                             uniforms=uniforms)
    return super().__call__(gl, gs, gw)


class YetiLogisticLoss(_GumbelMixin, PairwiseLogisticLoss):

  def __init__(self, reduction=Reduction.AUTO, name=None, lambda_weight=None,
               temperature=1.0, sample_size=8, gumbel_temperature=1.0, seed=None):
    PairwiseLogisticLoss.__init__(
        self, reduction, name, lambda_weight or YetiDCGLambdaWeight(), temperature)
    self._init_gumbel(sample_size, gumbel_temperature)


class GumbelApproxNDCGLoss(_GumbelMixin, ApproxNDCGLoss):

  def __init__(self, reduction=Reduction.AUTO, name=None, lambda_weight=None,
               temperature=0.1, sample_size=8, gumbel_temperature=1.0, seed=None):
    ApproxNDCGLoss.__init__(self, reduction, name, lambda_weight, temperature)
    self._init_gumbel(sample_size, gumbel_temperature)


class CoupledRankDistilLoss(_ListwiseLoss):
  """keras/losses.py:1659-1750 (set `loss._loss.uniforms` before calling)."""

  def __init__(self, reduction=Reduction.AUTO, name=None, sample_size=8, topk=None,
               temperature=1.):
    _RankingLoss.__init__(self, reduction, name)
    self._loss = losses_impl.CoupledRankDistilLoss(
        name=name, sample_size=sample_size, topk=topk, temperature=temperature)


class OrdinalLoss(_RankingLoss):
  """keras/losses.py:1603-1656."""

  def __init__(self, reduction=Reduction.AUTO, name=None, ordinal_size=1,
               use_fraction_label=False):
    super().__init__(reduction, name)
    self._loss = losses_impl.OrdinalLoss(name=name, ordinal_size=ordinal_size,
                                         use_fraction_label=use_fraction_label)


_KEY_TO_CLS = {
    'ordinal_loss': OrdinalLoss,
    'coupled_rankdistil_loss': CoupledRankDistilLoss,
    'yeti_logistic_loss': YetiLogisticLoss,
    'gumbel_approx_ndcg_loss': GumbelApproxNDCGLoss,
    'unique_softmax_loss': UniqueSoftmaxLoss,
    'list_mle_loss': ListMLELoss,
    'sigmoid_cross_entropy_loss': SigmoidCrossEntropyLoss,
    'mean_squared_loss': MeanSquaredLoss,
    'pairwise_hinge_loss': PairwiseHingeLoss,
    'pairwise_logistic_loss': PairwiseLogisticLoss,
    'pairwise_soft_zero_one_loss': PairwiseSoftZeroOneLoss,
    'pairwise_mse_loss': PairwiseMSELoss,
    'softmax_loss': SoftmaxLoss,
    'calibrated_softmax_loss': CalibratedSoftmaxLoss,
    'approx_ndcg_loss': ApproxNDCGLoss,
    'approx_mrr_loss': ApproxMRRLoss,
}


def get(loss, reduction=Reduction.AUTO, lambda_weight=None, name=None, **kwargs):
  """keras/losses.py:51-111 (hot-path keys)."""
  if loss not in _KEY_TO_CLS:
    raise ValueError('unsupported loss: {}'.format(loss))
  kw = dict(reduction=reduction, name=name, **kwargs)
  if loss not in ('approx_ndcg_loss', 'approx_mrr_loss', 'sigmoid_cross_entropy_loss',
                  'mean_squared_loss', 'gumbel_approx_ndcg_loss', 'ordinal_loss',
                  'coupled_rankdistil_loss'):
    kw['lambda_weight'] = lambda_weight
  return _KEY_TO_CLS[loss](**kw)
