"""Known-answer tests for the tfr.keras.losses surface (Keras reductions).

Expected values are the closed-form expressions of the reference's
keras/losses_test.py (cited per test) and the docstring examples of
keras/losses.py.  Run against the oracle and (under -m gpu) the CUDA path.
"""
import math

import pytest
import torch

ln = math.log


def _close(actual, expected, tol=1e-5):
  a = float(torch.as_tensor(actual).detach().double().cpu())
  assert abs(a - expected) <= tol * max(1.0, abs(expected)), (a, expected)


def _softmax(values):
  total = sum(math.exp(v) for v in values)
  return [math.exp(v) / total for v in values]


def _ref_pairwise_loss(labels, scores, weights, loss_form,
                       rank_discount_form=None):
  """Pure-python pairwise loss over one list, after sorting by score
  (keras/losses_test.py:42-100).  Returns (sum of loss, list_size)."""
  scores, labels, weights = zip(
      *sorted(zip(scores, labels, weights), reverse=True))

  def disc(rank):
    return {'LINEAR': 1. / rank, 'LOG': 1. / ln(1. + rank)}[rank_discount_form]

  def lam(label_diff, i, j):
    delta = math.fabs(label_diff)
    if rank_discount_form is not None:
      return delta * math.fabs(disc(i + 1) - disc(j + 1))
    return 1. if delta > 0 else 0

  phi = {
      'hinge': lambda d: max(0, 1 - d),
      'logistic': lambda d: ln(1. + math.exp(-d)),
      'soft_zero_one': lambda d: 1. / (1. + math.exp(d)),
  }[loss_form]
  loss = 0.
  for i in range(len(labels)):
    for j in range(len(labels)):
      if labels[i] > labels[j] and weights[i] > 0:
        loss += phi(scores[i] - scores[j]) * lam(labels[i] - labels[j], i,
                                                 j) * weights[i]
  return loss, float(len(labels))


def _agg(parts):
  return sum(p[0] for p in parts) / sum(p[1] for p in parts)


SCORES = [[1., 3., 2.], [1., 2., 3.]]
LABELS = [[0., 0., 1.], [0., 0., 2.]]


@pytest.mark.parametrize('form,cls', [
    ('hinge', 'PairwiseHingeLoss'),
    ('logistic', 'PairwiseLogisticLoss'),
    ('soft_zero_one', 'PairwiseSoftZeroOneLoss'),
])
def test_pairwise_losses(api, form, cls):
  """keras/losses_test.py:146-243."""
  K = api.keras_losses
  listwise_weights = [[2.], [1.]]
  lw_exp = [[2.] * 3, [1.] * 3]
  itemwise = [[2., 3., 4.], [1., 1., 1.]]
  ones = [1.] * 3
  loss_fn = getattr(K, cls)(name=form)
  for b in range(2):
    _close(loss_fn(api.t([LABELS[b]]), api.t([SCORES[b]])),
           _agg([_ref_pairwise_loss(LABELS[b], SCORES[b], ones, form)]))
    _close(loss_fn(api.t([LABELS[b]]), api.t([SCORES[b]]),
                   sample_weight=api.t([itemwise[b]])),
           _agg([_ref_pairwise_loss(LABELS[b], SCORES[b], itemwise[b], form)]))
  _close(loss_fn(api.t(LABELS), api.t(SCORES),
                 sample_weight=api.t(listwise_weights)),
         _agg([_ref_pairwise_loss(LABELS[b], SCORES[b], lw_exp[b], form)
               for b in range(2)]))
  # LambdaWeight: smooth_fraction=1 with 1/log1p discount; scaled by list size.
  lam = K.DCGLambdaWeight(rank_discount_fn=api.fns.log1p_inverse,
                          smooth_fraction=1.)
  loss_fn = getattr(K, cls)(name=form, lambda_weight=lam)
  _close(loss_fn(api.t(LABELS), api.t(SCORES),
                 sample_weight=api.t(listwise_weights)),
         _agg([_ref_pairwise_loss(LABELS[b], SCORES[b], lw_exp[b], form,
                                  rank_discount_form='LOG')
               for b in range(2)]) * 3.)


def test_pairwise_mse_loss(api):
  """keras/losses_test.py:245-264."""
  K = api.keras_losses
  loss = K.PairwiseMSELoss(reduction=api.KerasReduction.SUM_OVER_BATCH_SIZE)
  expected = (((2. - 3.) - (1. - 0.))**2 + ((2. - 1.) - (1. - 0.))**2 +
              ((3. - 1.) - (0. - 0.))**2 + ((3. - 2.) - (2. - 0.))**2 +
              ((3. - 1.) - (2. - 0.))**2 + ((2. - 1.) - (0. - 0.))**2) * 2. / 6.
  _close(loss(api.t(LABELS), api.t(SCORES)), expected)
  expected = (((2. - 3.) - (1. - 0.))**2 + ((2. - 1.) - (1. - 0.))**2 +
              ((3. - 1.) - (0. - 0.))**2 + 2 * ((3. - 2.) - (2. - 0.))**2 + 2 *
              ((3. - 1.) - (2. - 0.))**2 + 2 * ((2. - 1.) -
                                                (0. - 0.))**2) * 2. / 6.
  _close(loss(api.t(LABELS), api.t(SCORES), api.t([[1.], [2.]])), expected)


def test_softmax_loss(api):
  """keras/losses_test.py:284-308."""
  K = api.keras_losses
  scores = [[1., 3., 2.], [1., 2., 3.], [1., 2., 3.]]
  labels = [[0., 0., 1.], [0., 0., 2.], [0., 0., 0.]]
  weights = [[2.], [1.], [1.]]
  loss = K.SoftmaxLoss()
  _close(loss(api.t(labels), api.t(scores)),
         -(ln(_softmax(scores[0])[2]) + ln(_softmax(scores[1])[2]) * 2.) / 3.)
  _close(loss(api.t(labels), api.t(scores), api.t(weights)),
         -(ln(_softmax(scores[0])[2]) * 2. +
           ln(_softmax(scores[1])[2]) * 2. * 1.) / 3.)
  lam = K.DCGLambdaWeight(rank_discount_fn=api.fns.log1p_inverse)
  loss = K.SoftmaxLoss(lambda_weight=lam)
  _close(loss(api.t(labels), api.t(scores)),
         -(ln(_softmax(scores[0])[2]) / ln(1. + 2.) +
           ln(_softmax(scores[1])[2]) * 2. / ln(1. + 1.)) / 3.)


A_SCORES = [[1.4, -2.8, -0.4], [0., 1.8, 10.2], [1., 1.2, -3.2]]
A_LABELS = [[0., 2., 1.], [1., 0., 3.], [0., 0., 0.]]
A_EW = [[1., 2., 3.], [4., 5., 6.], [7., 8., 9.]]


def _norm_w(weights, labels):
  s = sum(max(0, l) for l in labels)
  return sum(w * max(0, l) for w, l in zip(weights, labels)) / s if s else 0


@pytest.mark.parametrize('reduction,div', [
    ('AUTO', 3.), ('SUM', 1.), ('SUM_OVER_BATCH_SIZE', 3.)])
def test_approx_ndcg_loss(api, reduction, div):
  """keras/losses_test.py:576-602, 650-690."""
  K = api.keras_losses
  nw = [_norm_w(w, l) for w, l in zip(A_EW, A_LABELS)]
  l0 = (1 / (3 / ln(2) + 1 / ln(3))) * (3 / ln(4) + 1 / ln(3))
  l1 = (1 / (7 / ln(2) + 1 / ln(3))) * (7 / ln(2) + 1 / ln(4))
  loss = K.ApproxNDCGLoss(reduction=getattr(api.KerasReduction, reduction))
  _close(loss(api.t(A_LABELS), api.t(A_SCORES)), -(l0 + l1) / div)
  _close(loss(api.t(A_LABELS), api.t(A_SCORES), api.t([[2.], [1.], [1.]])),
         -(2 * l0 + 1 * l1) / div)
  _close(loss(api.t(A_LABELS), api.t(A_SCORES), api.t(A_EW)),
         -(nw[0] * l0 + nw[1] * l1) / div)


def test_approx_mrr_loss(api):
  """keras/losses_test.py:692-706."""
  K = api.keras_losses
  labels = [[0., 0., 1.], [1., 0., 1.], [0., 0., 0.]]
  loss = K.ApproxMRRLoss()
  _close(loss(api.t(labels), api.t(A_SCORES)),
         -((1 / 2.) + 1 / 2. * (1 / 3. + 1 / 1.)) / 3.)
  _close(loss(api.t(labels), api.t(A_SCORES), api.t([[2.], [1.], [1.]])),
         -(2 * 1 / 2. + 1 * 1 / 2. * (1 / 3. + 1 / 1.)) / 3.)


def test_invalid_labels_and_temperature(api):
  """keras/losses_test.py:708-740."""
  K = api.keras_losses
  scores, labels = api.t([[1., 3., 2.]]), api.t([[0., -1., 1.]])
  _close(K.PairwiseLogisticLoss()(labels, scores), ln(1 + math.exp(-1.)) / 3.)
  _close(K.PairwiseLogisticLoss(reduction=api.KerasReduction.SUM)(
      labels, scores), ln(1 + math.exp(-1.)))
  _close(K.PairwiseLogisticLoss(reduction=api.KerasReduction.SUM,
                                temperature=0.1)(labels, scores),
         ln(1 + math.exp(-10.)))
  _close(K.SoftmaxLoss()(labels, scores), -(ln(_softmax([1, 2])[1])))


@pytest.mark.parametrize('cls,expected', [
    ('PairwiseHingeLoss', 4.),
    ('PairwiseLogisticLoss', 2.9397852),
    ('PairwiseSoftZeroOneLoss', 1.7310586),
    ('SoftmaxLoss', 4.034129),
    ('ApproxNDCGLoss', -1.2618682),
    ('ApproxMRRLoss', -1.0000114),
])
def test_loss_sum_on_padded_lists(api, cls, expected):
  """keras/losses_test.py:769-790 (ragged inputs densified with the padding of
  utils.py:21-23: label -1, prediction -1e6)."""
  scores = api.t([[1., 3., 2.], [3., 2., -1e6]])
  labels = api.t([[0., 0., 1.], [0., 2., -1.]])
  loss = getattr(api.keras_losses, cls)(reduction=api.KerasReduction.SUM)
  _close(loss(labels, scores), expected)


@pytest.mark.parametrize('cls,y_true,y_pred,expected', [
    # docstring examples: keras/losses.py:350-361, 417-428, 484-495, 770-781,
    # 1183-1194
    ('PairwiseHingeLoss', [[1., 0.]], [[0.6, 0.8]], 0.6),
    ('PairwiseLogisticLoss', [[1., 0.]], [[0.6, 0.8]], 0.39906943),
    ('PairwiseSoftZeroOneLoss', [[1., 0.]], [[0.6, 0.8]], 0.274917),
    ('SoftmaxLoss', [[1., 0.]], [[0.6, 0.8]], 0.7981389),
    ('ApproxNDCGLoss', [[1., 0.]], [[0.6, 0.8]], -0.655107),
])
def test_docstring_examples(api, cls, y_true, y_pred, expected):
  loss = getattr(api.keras_losses, cls)()
  _close(loss(api.t(y_true), api.t(y_pred)), expected)


def test_get_factory(api):
  """keras/losses.py:51-111."""
  K = api.keras_losses
  assert isinstance(K.get('approx_ndcg_loss'), K.ApproxNDCGLoss)
  assert isinstance(K.get('pairwise_logistic_loss'), K.PairwiseLogisticLoss)
  assert isinstance(K.get('softmax_loss'), K.SoftmaxLoss)
  with pytest.raises(ValueError):
    K.get('no_such_loss')


# ---------------------------------------------------------------------------
# Gumbel-sampled losses (keras/losses_test.py:604-645; losses_impl.py:540-649)
# ---------------------------------------------------------------------------
def test_gumbel_sampler_transform(oracle_api):
  """log(softmax((s + G) / T) + 1e-20) with G = -log(-log(u + eps) + eps); labels and
  weights tiled sample-major per list."""
  L = oracle_api.losses_impl
  scores = [[1.0, 2.0, 0.5]]
  labels = [[1.0, 0.0, -1.0]]
  u = [[[0.5, 0.25, 0.9], [0.1, 0.7, 0.3]]]
  smp = L.GumbelSampler(sample_size=2, temperature=2.0)
  el, sl, ew = smp.sample(oracle_api.t(labels), oracle_api.t(scores),
                          oracle_api.t([[3.0]]), uniforms=oracle_api.t(u))
  rows = []
  for srow in u[0]:
    z = [(s + -math.log(-math.log(x + 1e-20) + 1e-20)) / 2.0 for s, x in zip(scores[0], srow)]
    z[2] = math.log(1e-20)
    m = max(z)
    den = sum(math.exp(v - m) for v in z)
    rows.append([math.log(math.exp(v - m) / den + 1e-20) for v in z])
  torch.testing.assert_close(sl.double(), torch.tensor(rows, dtype=torch.float64),
                             rtol=1e-5, atol=1e-5)
  assert el.tolist() == [labels[0], labels[0]]
  assert ew.tolist() == [[3.0], [3.0]]


def test_gumbel_approx_ndcg_reference_case(api):
  """keras/losses_test.py:604-645: the reference lists the scores its sampler drew
  (seed 1); ApproxNDCG on those expanded lists must give its expected value."""
  sampled = [[-.291, -1.643, -2.826], [-.0866, -2.924, -3.530],
             [-12.42, -9.492, -7.939e-5], [-8.859, -6.830, -1.223e-3],
             [-.8930, -.5266, -45.80183], [-.6650, -.7220, -45.94149]]
  ex_labels = [[0., 2., 1.], [0., 2., 1.], [1., 0., 3.], [1., 0., 3.],
               [0., 0., 0.], [0., 0., 0.]]
  loss = api.keras_losses.ApproxNDCGLoss()
  got = float(loss(api.t(ex_labels), api.t(sampled)))
  want = -(2 * (1 / (3 / ln(2) + 1 / ln(3))) * (3 / ln(3) + 1 / ln(4)) + 2 *
           (1 / (7 / ln(2) + 1 / ln(3))) * (7 / ln(2) + 1 / ln(4))) / 6
  assert abs(got - want) < 5e-4
  ex_w = [[2.], [2.], [1.], [1.], [1.], [1.]]
  got = float(loss(api.t(ex_labels), api.t(sampled), api.t(ex_w)))
  want = -(2 * 2 * (1 / (3 / ln(2) + 1 / ln(3))) * (3 / ln(3) + 1 / ln(4)) + 1 * 2 *
           (1 / (7 / ln(2) + 1 / ln(3))) * (7 / ln(2) + 1 / ln(4))) / 6
  assert abs(got - want) < 5e-4


def test_calibrated_softmax_loss(api):
  """keras/losses_test.py:936-972."""
  scores = [[1.0, 3.0, 2.0], [1.0, 2.0, 3.0], [1.0, 2.0, 3.0]]
  labels = [[0.0, 0.0, 1.0], [0.0, 0.0, 2.0], [0.0, 0.0, 0.0]]
  weights = [[2.0], [1.0], [1.0]]
  vl = 0.5
  den = lambda values: sum(math.exp(v) for v in values)
  sm = lambda values: [math.exp(v) / (den(values) + 1.0) for v in values]
  loss = api.keras_losses.get('calibrated_softmax_loss', virtual_label=vl)
  _close(loss(api.t(labels), api.t(scores)),
         -(ln(sm(scores[0])[2]) + ln(sm(scores[1])[2]) * 2.0 -
           vl * ln(1.0 + den(scores[0])) - vl * ln(1.0 + den(scores[1])) -
           vl * ln(1.0 + den(scores[2]))) / 3.0)
  _close(loss(api.t(labels), api.t(scores), api.t(weights)),
         -(ln(sm(scores[0])[2]) * 2.0 + ln(sm(scores[1])[2]) * 2.0 * 1.0 -
           vl * ln(1.0 + den(scores[0])) * 2.0 - vl * ln(1.0 + den(scores[1])) -
           vl * ln(1.0 + den(scores[2]))) / 3.0)
