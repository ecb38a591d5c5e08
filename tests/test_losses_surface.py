"""`ranking_b200.losses`: RankingLossKey / make_loss_fn (reference losses.py:29-300)."""
import pytest
import torch


def test_ranking_loss_keys_match_reference():
  """losses.py:29-55."""
  from ranking_b200 import losses as L
  want = {'pairwise_hinge_loss', 'pairwise_logistic_loss', 'pairwise_soft_zero_one_loss',
          'pairwise_mse_loss', 'yeti_logistic_loss', 'circle_loss', 'softmax_loss',
          'poly_one_softmax_loss', 'unique_softmax_loss', 'sigmoid_cross_entropy_loss',
          'mean_squared_loss', 'list_mle_loss', 'approx_ndcg_loss', 'approx_mrr_loss',
          'gumbel_approx_ndcg_loss', 'neural_sort_cross_entropy_loss',
          'gumbel_neural_sort_cross_entropy_loss', 'neural_sort_ndcg_loss',
          'gumbel_neural_sort_ndcg_loss'}
  assert set(L.RankingLossKey.all_keys()) == want


def test_make_loss_fn_argument_errors():
  """losses.py:86-93, 176-190, 243-246 (same ValueErrors) and utils.parse_keys_and_weights."""
  from ranking_b200 import losses as L
  assert L.parse_keys_and_weights('softmax_loss:0.9, sigmoid_cross_entropy_loss:0.1') == {
      'softmax_loss': 0.9, 'sigmoid_cross_entropy_loss': 0.1}
  assert L.parse_keys_and_weights('softmax_loss') == {'softmax_loss': 1.0}
  for kw in (dict(loss_keys='softmax_loss:0.9,mean_squared_loss:0.1', loss_weights=[1., 2.]),
             dict(loss_keys=['softmax_loss', 'mean_squared_loss'], loss_weights=[1.]),
             dict(loss_keys=None), dict(loss_keys=[]), dict(loss_keys='no_such_loss'),
             dict(loss_keys='softmax_loss', reduction='none'),
             dict(loss_keys='poly_one_softmax_loss')):
    with pytest.raises(ValueError):
      L.make_loss_fn(**kw)
  assert callable(L.make_loss_fn('softmax_loss:0.9,mean_squared_loss:0.1'))
  assert callable(L.make_loss_fn(L.RankingLossKey.all_keys()[:5]))


@pytest.mark.gpu
def test_make_loss_fn_values():
  """Single keys equal the loss objects' `compute`; weighted combinations add up; the
  weights feature is taken from `features`; Gumbel keys run and are finite."""
  from ranking_b200 import losses as L
  from ranking_b200 import losses_impl as I
  g = torch.Generator().manual_seed(3)
  b, n = 6, 17
  scores = torch.randn(b, n, generator=g).cuda()
  labels = torch.randint(0, 4, (b, n), generator=g).float()
  labels[:, -3:] = -1.
  labels = labels.cuda()
  w = (torch.rand(b, 1, generator=g) + 0.5).cuda()
  red = I.Reduction.SUM_BY_NONZERO_WEIGHTS
  soft = float(I.SoftmaxLoss(None).compute(labels, scores, w, red))
  mse = float(I.MeanSquaredLoss(None).compute(labels, scores, w, red))
  circ = float(I.CircleLoss(None, gamma=4., margin=0.1).compute(labels, scores.sigmoid(), w, red))
  f = L.make_loss_fn('softmax_loss', weights_feature_name='w')
  assert float(f(labels, scores, {'w': w})) == pytest.approx(soft, rel=1e-6)
  f = L.make_loss_fn('softmax_loss:0.25,mean_squared_loss:2.0', weights_feature_name='w')
  assert float(f(labels, scores, {'w': w})) == pytest.approx(0.25 * soft + 2.0 * mse, rel=1e-6)
  f = L.make_loss_fn(['circle_loss'], weights_feature_name='w', params={'gamma': 4., 'margin': 0.1})
  assert float(f(labels, scores.sigmoid(), {'w': w})) == pytest.approx(circ, rel=1e-6)
  nd = float(I.NeuralSortNDCGLoss(None, temperature=0.5).compute(labels, scores, None, red))
  f = L.make_loss_fn('neural_sort_ndcg_loss', params={'temperature': 0.5})
  assert float(f(labels, scores, {})) == pytest.approx(nd, rel=1e-6)
  # differentiable, and the Gumbel keys draw sample_size perturbed copies per list
  s = scores.clone().requires_grad_()
  f = L.make_loss_fn(['gumbel_approx_ndcg_loss', 'gumbel_neural_sort_cross_entropy_loss',
                      'yeti_logistic_loss'], loss_weights=[1., 0.5, 0.25],
                     gumbel_params={'sample_size': 4, 'seed': 7})
  v = f(labels, s, {})
  v.backward()
  assert torch.isfinite(v) and torch.isfinite(s.grad).all() and float(s.grad.abs().max()) > 0
