"""Known-answer tests of CircleLoss and the NeuralSort losses, transcribed from the
reference's own tests (losses_impl_test.py:32-86 `_circle_loss`, :126-150 `_neural_sort`,
:278-299, :1002-1086, :1758-1847).  Run against the CPU oracle and, under `-m gpu`, the
CUDA kernels (csrc/loss_extra.cu) through the C ABI."""
import math

import pytest
import torch


def ln(x):
  return math.log(x)


def _close(actual, expected, tol=1e-5):
  a = torch.as_tensor(actual).detach().double().cpu()
  e = torch.as_tensor(expected).double()
  torch.testing.assert_close(a.reshape(e.shape), e, rtol=tol, atol=tol)


def _circle_loss(labels, scores, gamma=64., margin=0.25):
  """losses_impl_test.py:32-86 without a lambda weight: sum over pairs l_i > l_j."""
  scores, labels = zip(*sorted(zip(scores, labels), reverse=True))
  loss = 0.
  for i in range(len(labels)):
    for j in range(len(labels)):
      if labels[i] > labels[j]:
        si, sj = scores[i], scores[j]
        loss += math.exp(gamma * max(0., (1 + margin) - si) * ((1 - margin) - si) +
                         gamma * max(0., sj + margin) * (sj - margin))
  return loss


def _softmax20(values):
  total = sum(math.exp(v) for v in values)
  return [math.exp(v) / (1e-20 + total) for v in values]


def _neural_sort(logits, temperature=1.0):
  """losses_impl_test.py:126-147."""
  result = []
  for row in logits:
    n = len(row)
    diff_sum = [sum(abs(m - l) for m in row) for l in row]
    perm = []
    for i in range(n):
      scaling = n + 1 - 2 * (i + 1)
      p = [(scaling * l - s) / temperature for l, s in zip(row, diff_sum)]
      p = [l - max(p) for l in p]
      perm.append(_softmax20(p))
    result.append(perm)
  return result


def _softmax_cross_entropy(p_trues, p_preds):
  return sum(sum(-y_t * math.log(1e-20 + y_p) for y_t, y_p in zip(t, p))
             for t, p in zip(p_trues, p_preds))


def test_circle_loss(api):
  """losses_impl_test.py:1004-1086."""
  L, R = api.losses_impl, api.Reduction
  scores = [[0.1, 0.3, 0.2], [0.1, 0.2, 0.3]]
  labels = [[0., 0., 1.], [0., 1., 2.]]
  l0 = math.log1p(_circle_loss(labels[0], scores[0]))
  l1 = math.log1p(_circle_loss(labels[1], scores[1]))
  fn = L.CircleLoss(name=None)
  _close(fn.compute(api.t(labels), api.t(scores), None, R.MEAN), (l0 + l1) / 2)
  _close(fn.compute(api.t(labels), api.t(scores), api.t([[1.], [2.]]), R.MEAN),
         (l0 * 1. + l1 * 2.) / 3.)
  _close(fn.compute(api.t(labels), api.t(scores), api.t([[1., 1., 2.], [1., 1., 1.]]), R.MEAN),
         (l0 * 2. + l1 * 1.) / 3.)
  labels2 = [[0., 0., 1.], [0., 0., 2.]]
  fn2 = L.CircleLoss(name=None, gamma=4., margin=0.1)
  _close(fn2.compute(api.t(labels2), api.t(scores), None, R.MEAN),
         (math.log1p(_circle_loss(labels2[0], scores[0], 4., 0.1)) +
          math.log1p(_circle_loss(labels2[1], scores[1], 4., 0.1))) / 2)
  _close(fn.compute(api.t([[0., -1., 1.]]), api.t([[.1, .3, .2]]), None, R.MEAN),
         math.log1p(_circle_loss([0., 1.], [.1, .2])))
  mask = torch.tensor([[True, False, True], [True, True, True]], device=api.device)
  _close(fn.compute(api.t([[1., 0., 0.], [0., 0., 2.]]), api.t([[.1, .3, .2], [.1, .2, .3]]),
                    None, R.MEAN, mask),
         (math.log1p(_circle_loss([1., 0.], [.1, .2])) +
          math.log1p(_circle_loss([0., 0., 2.], [.1, .2, .3]))) / 2)


def test_neural_sort_cross_entropy_loss(api):
  """losses_impl_test.py:1760-1807."""
  L, R = api.losses_impl, api.Reduction
  scores = [[1.4, -2.8, -0.4], [0., 1.8, 10.2], [1., 1.2, -3.2]]
  labels = [[0., 2., 1.], [1., 0., -3.], [0., 0., 0.]]
  p_scores = _neural_sort([[1.4, -2.8, -0.4], [0., 1.8], [1., 1.2, -3.2]])
  p_labels = _neural_sort([[0., 2., 1.], [1., 0.], [0., 0., 0.]])
  fn = L.NeuralSortCrossEntropyLoss(name=None)
  _close(fn.compute(api.t(labels), api.t(scores), None, R.SUM),
         _softmax_cross_entropy(p_labels[0], p_scores[0]) / 3. +
         _softmax_cross_entropy(p_labels[1], p_scores[1]) / 2., tol=1e-4)
  _close(fn.compute(api.t(labels), api.t(scores), api.t([[2.], [1.], [1.]]), R.SUM),
         _softmax_cross_entropy(p_labels[0], p_scores[0]) * 2.0 / 3. +
         _softmax_cross_entropy(p_labels[1], p_scores[1]) / 2., tol=1e-4)
  ps, pl = _neural_sort([[1., 2.]]), _neural_sort([[0., 1.]])
  _close(fn.compute(api.t([[0., -1., 1.]]), api.t([[1., 3., 2.]]), None,
                    R.SUM_BY_NONZERO_WEIGHTS), _softmax_cross_entropy(pl[0], ps[0]) / 2.)
  ps, pl = _neural_sort([[2., 3.]]), _neural_sort([[0., 1.]])
  mask = torch.tensor([[True, False, True, False, False]], device=api.device)
  _close(fn.compute(api.t([[0., 0., 1., 0., 1.]]), api.t([[2., 4., 3., 3., -1e10]]), None,
                    R.SUM_BY_NONZERO_WEIGHTS, mask), _softmax_cross_entropy(pl[0], ps[0]) / 2.)


def test_neural_sort_ndcg_loss(api):
  """losses_impl_test.py:1812-1847."""
  L, R = api.losses_impl, api.Reduction
  scores = [[1.4, -2.8, -0.4], [0., 1.8, 10.2], [1., 1.2, -3.2]]
  labels = [[0., 2., 1.], [1., 0., -3.], [0., 0., 0.]]
  fn = L.NeuralSortNDCGLoss(name=None, temperature=0.1)
  a = (1 / (3 / ln(2) + 1 / ln(3))) * (3 / ln(4) + 1 / ln(3))
  b = (1 / (1 / ln(2))) * (1 / ln(3))
  _close(fn.compute(api.t(labels), api.t(scores), None, R.SUM), -(a + b), tol=1e-4)
  _close(fn.compute(api.t(labels), api.t(scores), api.t([[2.], [1.], [1.]]), R.SUM),
         -(2 * a + b), tol=1e-4)
  _close(fn.compute(api.t([[0., -1., 1.]]), api.t([[1., 3., 2.]]), None,
                    R.SUM_BY_NONZERO_WEIGHTS), -1., tol=1e-4)
  mask = torch.tensor([[True, False, True, False, False]], device=api.device)
  _close(fn.compute(api.t([[0., 0., 1., 0., 1.]]), api.t([[2., 4., 3., -5., 1000.0]]), None,
                    R.SUM_BY_NONZERO_WEIGHTS, mask), -1., tol=1e-4)


def test_neural_sort_matrix_oracle():
  """losses_impl_test.py:278-299 (the matrix itself only exists in the oracle)."""
  from oracle import losses_impl as L
  p = L.neural_sort(torch.tensor([[140., -280., -40.], [0., 180., 1020.], [100., 120., -320.]]))
  want = [[[1, 0, 0], [0, 0, 1], [0, 1, 0]], [[0, 0, 1], [0, 1, 0], [1, 0, 0]],
          [[0, 1, 0], [1, 0, 0], [0, 0, 1]]]
  torch.testing.assert_close(p, torch.tensor(want, dtype=torch.float32), rtol=1e-3, atol=1e-6)
  p = L.neural_sort(torch.tensor([[3.0, 1.0, -1.0, 1000.0, 5.0, 2.0]]),
                    mask=torch.tensor([[True, True, True, False, False, True]]))
  want = [[[0.72140, 0.01321, 0.00000, 0., 0., 0.26539],
           [0.21183, 0.21183, 0.00053, 0., 0., 0.57581],
           [0.01204, 0.65723, 0.08895, 0., 0., 0.24178],
           [0.00004, 0.11849, 0.87557, 0., 0., 0.0059],
           [0., 0., 0., 0.5, 0.5, 0.], [0., 0., 0., 0.5, 0.5, 0.]]]
  torch.testing.assert_close(p, torch.tensor(want), rtol=0, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('cls,kw', [('CircleLoss', dict(gamma=8., margin=0.25)),
                                    ('CircleLoss', dict()),
                                    ('NeuralSortCrossEntropyLoss', dict(temperature=1.0)),
                                    ('NeuralSortCrossEntropyLoss', dict(temperature=0.5)),
                                    ('NeuralSortNDCGLoss', dict(temperature=1.0)),
                                    ('NeuralSortNDCGLoss', dict(temperature=0.3))])
@pytest.mark.parametrize('wkind', ['none', 'list', 'item'])
@pytest.mark.parametrize('n', [1, 5, 64, 200])
def test_extra_losses_and_grad(cuda_api, oracle_api, cls, kw, wkind, n):
  """Loss values and d loss / d scores of K3c against the fp64 oracle (autograd through the
  [B, N, N] formulation), MEAN reduction over lists, ragged lists with padding."""
  g = torch.Generator().manual_seed(31 + n)
  b = 6
  circle = cls == 'CircleLoss'
  if circle:
    # gamma = 64 (default): exp(gamma (a_i + c_j)) overflows fp32 (as it does in the fp32
    # reference) once a positive sits near 0 and a negative near 1; keep those cases finite
    wide = kw.get('gamma', 64.) <= 8.
    scores = (torch.rand(b, n, generator=g) * 1.4 - 0.2 if wide
              else torch.rand(b, n, generator=g) * 0.4 + 0.3)
  else:
    scores = torch.randn(b, n, generator=g)
  labels = torch.randint(0, 4, (b, n), generator=g).float()
  if n > 2:
    labels[:, -max(1, n // 5):] = -1.
    labels[0] = torch.where(labels[0] >= 0, torch.zeros_like(labels[0]), labels[0])  # no positives
  if circle:
    labels[:, 0] = 3.
    labels[:, min(1, n - 1)] = 0. if n > 1 else 3.
  item_w = torch.rand(b, n, generator=g) + 0.5
  list_w = torch.rand(b, 1, generator=g) + 0.5
  weights = {'none': None, 'list': list_w, 'item': item_w}[wkind]
  fc = getattr(cuda_api.losses_impl, cls)(name=None, **kw)
  fo = getattr(oracle_api.losses_impl, cls)(name=None, **kw)
  red_c, red_o = cuda_api.Reduction.MEAN, oracle_api.Reduction.MEAN
  w_gpu = None if weights is None else weights.cuda()
  w_ref = None if weights is None else weights.double()
  if circle and n == 1:
    # a list without a pair: weight = 0 / 0 = NaN on both sides (losses_impl.py:1108-1110)
    _, wc = fc.compute_per_list(labels.cuda(), scores.cuda(), w_gpu)
    _, wo = fo.compute_per_list(labels.double(), scores.double(), w_ref)
    assert bool(torch.isnan(wc).all()) and bool(torch.isnan(wo).all())
    return
  s_gpu = scores.cuda().requires_grad_()
  got = fc.compute(labels.cuda(), s_gpu, w_gpu, red_c)
  s_ref = scores.double().requires_grad_()
  ref = fo.compute(labels.double(), s_ref, w_ref, red_o)
  got.backward()
  ref.backward()
  # NeuralSort rows are softmaxes of u = c s_k - D_k with |u| ~ N |s| / T: one fp32 ulp of u
  # is 6e-8 |u|, which bounds how well ANY fp32 evaluation can match the fp64 oracle
  tol = 2e-5
  if not circle:
    tol = max(tol, 4e-7 * n * float(scores.abs().max()) / kw.get('temperature', 1.0))
  assert abs(float(got) - float(ref)) <= tol * max(1., abs(float(ref))), (float(got), float(ref))
  gr = s_ref.grad
  err = float((s_gpu.grad.double().cpu() - gr).abs().max() / (gr.abs().max() + 1e-30))
  if float(gr.abs().max()) > 0:
    assert err <= 2.5 * tol, (err, tol)
  else:
    assert float(s_gpu.grad.abs().max()) <= 1e-6
