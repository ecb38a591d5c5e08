"""`tfr.losses` surface: `RankingLossKey` and `make_loss_fn` (reference losses.py:29-300).

The estimator-style entry point of the reference: one loss key, a list of keys, or a string
such as 'mean_squared_loss:0.1,softmax_loss:0.9' -> a function
`loss_fn(labels, logits, features)` that returns the weighted sum of the reduced losses.  Every
key maps to the CUDA loss objects of `ranking_b200.losses_impl` (one fused forward + backward
kernel per loss); the Gumbel keys draw their perturbed copies with the CUDA Gumbel sampler
(losses_impl.py:540-649) exactly like the reference's `_LossFunctionMaker`.
"""
from ranking_b200 import losses_impl
from ranking_b200.losses_impl import Reduction


class RankingLossKey(object):
  """Ranking loss key strings (losses.py:29-55)."""
  PAIRWISE_HINGE_LOSS = 'pairwise_hinge_loss'
  PAIRWISE_LOGISTIC_LOSS = 'pairwise_logistic_loss'
  PAIRWISE_SOFT_ZERO_ONE_LOSS = 'pairwise_soft_zero_one_loss'
  PAIRWISE_MSE_LOSS = 'pairwise_mse_loss'
  YETI_LOGISTIC_LOSS = 'yeti_logistic_loss'
  CIRCLE_LOSS = 'circle_loss'
  SOFTMAX_LOSS = 'softmax_loss'
  POLY_ONE_SOFTMAX_LOSS = 'poly_one_softmax_loss'
  UNIQUE_SOFTMAX_LOSS = 'unique_softmax_loss'
  SIGMOID_CROSS_ENTROPY_LOSS = 'sigmoid_cross_entropy_loss'
  MEAN_SQUARED_LOSS = 'mean_squared_loss'
  LIST_MLE_LOSS = 'list_mle_loss'
  APPROX_NDCG_LOSS = 'approx_ndcg_loss'
  APPROX_MRR_LOSS = 'approx_mrr_loss'
  GUMBEL_APPROX_NDCG_LOSS = 'gumbel_approx_ndcg_loss'
  NEURAL_SORT_CROSS_ENTROPY_LOSS = 'neural_sort_cross_entropy_loss'
  GUMBEL_NEURAL_SORT_CROSS_ENTROPY_LOSS = 'gumbel_neural_sort_cross_entropy_loss'
  NEURAL_SORT_NDCG_LOSS = 'neural_sort_ndcg_loss'
  GUMBEL_NEURAL_SORT_NDCG_LOSS = 'gumbel_neural_sort_ndcg_loss'

  @classmethod
  def all_keys(cls):
    return [v for k, v in vars(cls).items() if k.isupper()]


def parse_keys_and_weights(key):
  """utils.py:446-480: 'a:0.9,b:0.1' -> {'a': 0.9, 'b': 0.1}; a missing weight is 1."""
  out = {}
  for part in key.replace(' ', '').split(','):
    if not part:
      continue
    if ':' in part:
      k, w = part.split(':')
      out[k] = float(w)
    else:
      out[part] = 1.0
  return out


_REDUCTIONS = (Reduction.SUM, Reduction.MEAN, Reduction.SUM_BY_NONZERO_WEIGHTS,
               Reduction.SUM_OVER_BATCH_SIZE)

# key -> (loss class, takes lambda_weight, drawn from the Gumbel sampler, default temperature)
_TABLE = {
    RankingLossKey.PAIRWISE_HINGE_LOSS: ('PairwiseHingeLoss', True, False, None),
    RankingLossKey.PAIRWISE_LOGISTIC_LOSS: ('PairwiseLogisticLoss', True, False, None),
    RankingLossKey.PAIRWISE_SOFT_ZERO_ONE_LOSS: ('PairwiseSoftZeroOneLoss', True, False, None),
    RankingLossKey.PAIRWISE_MSE_LOSS: ('PairwiseMSELoss', True, False, None),
    RankingLossKey.YETI_LOGISTIC_LOSS: ('PairwiseLogisticLoss', True, True, None),
    RankingLossKey.CIRCLE_LOSS: ('CircleLoss', True, False, None),
    RankingLossKey.SOFTMAX_LOSS: ('SoftmaxLoss', True, False, None),
    RankingLossKey.UNIQUE_SOFTMAX_LOSS: ('UniqueSoftmaxLoss', True, False, None),
    RankingLossKey.SIGMOID_CROSS_ENTROPY_LOSS: ('SigmoidCrossEntropyLoss', False, False, None),
    RankingLossKey.MEAN_SQUARED_LOSS: ('MeanSquaredLoss', False, False, None),
    RankingLossKey.LIST_MLE_LOSS: ('ListMLELoss', True, False, None),
    RankingLossKey.APPROX_NDCG_LOSS: ('ApproxNDCGLoss', False, False, 0.1),
    RankingLossKey.APPROX_MRR_LOSS: ('ApproxMRRLoss', False, False, 0.1),
    RankingLossKey.GUMBEL_APPROX_NDCG_LOSS: ('ApproxNDCGLoss', False, True, 0.1),
    RankingLossKey.NEURAL_SORT_CROSS_ENTROPY_LOSS: ('NeuralSortCrossEntropyLoss', False, False, 1.0),
    RankingLossKey.GUMBEL_NEURAL_SORT_CROSS_ENTROPY_LOSS:
        ('NeuralSortCrossEntropyLoss', False, True, 1.0),
    RankingLossKey.NEURAL_SORT_NDCG_LOSS: ('NeuralSortNDCGLoss', False, False, 1.0),
    RankingLossKey.GUMBEL_NEURAL_SORT_NDCG_LOSS: ('NeuralSortNDCGLoss', False, True, 1.0),
}


def make_loss_fn(loss_keys, loss_weights=None, weights_feature_name=None, lambda_weight=None,
                 reduction=Reduction.SUM_BY_NONZERO_WEIGHTS, name=None, params=None,
                 gumbel_params=None):
  """losses.py:263-300 (and `_LossFunctionMaker`, :58-260).  Same arguments and errors.
  `params` holds loss-specific arguments (`temperature`, `gamma`, `margin`); they reach the
  losses that take them.  'poly_one_softmax_loss' is not built."""
  if isinstance(loss_keys, str) and (':' in loss_keys or ',' in loss_keys):
    if loss_weights is not None:
      raise ValueError('`loss_weights` has to be None when weights are encoded in `loss_keys`.')
    kw = parse_keys_and_weights(loss_keys)
    loss_keys, loss_weights = list(kw.keys()), list(kw.values())
  if reduction not in _REDUCTIONS:
    raise ValueError('Invalid reduction: {}'.format(reduction))
  if not loss_keys:
    raise ValueError('loss_keys cannot be None or empty.')
  if not isinstance(loss_keys, list):
    loss_keys = [loss_keys]
  if loss_weights:
    if len(loss_keys) != len(loss_weights):
      raise ValueError('loss_keys and loss_weights must have the same size.')
  for k in loss_keys:
    if k == RankingLossKey.POLY_ONE_SOFTMAX_LOSS:
      raise ValueError('loss_key poly_one_softmax_loss is not available in ranking_b200.')
    if k not in _TABLE:
      raise ValueError('Invalid loss_key: {}.'.format(k))
  params = dict(params or {})
  sampler = losses_impl.GumbelSampler(**(gumbel_params or {}))

  def _build(key):
    cls_name, takes_lambda, _, default_t = _TABLE[key]
    cls = getattr(losses_impl, cls_name)
    kwargs = {}
    if takes_lambda:
      kwargs['lambda_weight'] = lambda_weight
    if cls_name == 'CircleLoss':
      for p in ('gamma', 'margin'):
        if p in params:
          kwargs[p] = params[p]
    elif default_t is not None or 'temperature' in params:
      if cls_name not in ('MeanSquaredLoss',):
        kwargs['temperature'] = params.get('temperature', default_t if default_t else 1.0)
    return cls(name, **kwargs)

  objs = [(k, _build(k)) for k in loss_keys]

  def _loss_fn(labels, logits, features):
    """labels / logits [batch_size, list_size] CUDA tensors; `features` only supplies the
    weights feature.  Returns the (weighted sum of the) reduced loss(es)."""
    weights = None
    if weights_feature_name:
      weights = features[weights_feature_name]
      if weights.dim() == 1:
        weights = weights.reshape(-1, 1)
      elif weights.dim() > 2:
        weights = weights.reshape(weights.shape[0], -1)
    gbl = None
    total = None
    for i, (key, obj) in enumerate(objs):
      if _TABLE[key][2]:
        if gbl is None:
          gbl = sampler.sample(labels, logits, weights=weights)
        value = obj.compute(gbl[0], gbl[1], gbl[2], reduction)
      else:
        value = obj.compute(labels, logits, weights, reduction)
      if loss_weights:
        value = value * float(loss_weights[i])
      total = value if total is None else total + value
    return total

  return _loss_fn
