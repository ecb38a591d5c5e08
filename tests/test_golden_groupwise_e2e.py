"""The reference's only gradient-bearing golden (model_test.py:122-150, 326-468):
groupwise model, group_size 2, group score function = Dense(2) over
[context, age(member 0), age(member 1)], pairwise hinge loss (estimator reduction
SUM_BY_NONZERO_WEIGHTS, list weights 1), kernel [[1,1],[2,2],[3,3]], bias [1,1].

  eval:   loss 6.75, logits_mean 300.833343                        (:401-419)
  predict values for list sizes 2 and 4                             (:421-446)
  one AdagradOptimizer(0.1) step (accumulators 1e-12, so every step is
  lr * sign(grad)):  kernel -> [[1,1],[1.9,2.1],[3.1,2.9]], bias unchanged   (:392-399)
  predict values after that step                                    (:448-468)

Runs against the oracle (CPU) and against the product path (CUDA loss kernel through the
C ABI + device-side group formation / scatter-average).  TRAIN mode shuffles the valid
items with TF's RNG in the reference; the hinge gradients' signs — all this golden pins —
do not depend on the shuffle, so the identity permutation reproduces it.
"""
import pytest
import torch


def _score_fn(kernel, bias):
  def fn(group_features):      # [B*G, 2, 2]: channels (age, context)
    ctx = group_features[:, 0, 1:2]
    inp = torch.cat([ctx, group_features[:, 0, 0:1], group_features[:, 1, 0:1]], 1)
    return inp @ kernel + bias
  return fn


def _logits(api, kernel, bias, context, age, labels=None):
  x = torch.stack([age, context.expand_as(age)], -1)      # [B, N, 2]
  valid = torch.ones(age.shape, dtype=torch.bool, device=age.device) if labels is None \
      else labels >= 0
  if api.name == 'oracle':
    return api.scorer.groupwise_logits(x, valid, 2, _score_fn(kernel, bias))
  import ranking_b200 as tfr
  return tfr.model.GroupwiseRankingModel(_score_fn(kernel, bias), 2).compute_logits(x, valid)


def _case(api):
  dt = torch.float64 if api.name == 'oracle' else torch.float32
  t = lambda v: torch.tensor(v, dtype=dt, device=api.device)   # noqa: E731
  kernel = t([[1., 1.], [2., 2.], [3., 3.]]).requires_grad_()
  bias = t([1., 1.]).requires_grad_()
  context = t([[178.], [155.]])
  age = t([[10., 20., 20.], [50., 30., 30.]])
  labels = t([[1., 0., 0.], [1., 0., 0.]])
  weights = t([[1.], [1.]])
  return t, kernel, bias, context, age, labels, weights


def test_reference_groupwise_eval_golden(api):
  t, kernel, bias, context, age, labels, weights = _case(api)
  logits = _logits(api, kernel, bias, context, age, labels)
  assert float(logits.detach().mean()) == pytest.approx(300.833343, rel=1e-6)
  loss = api.losses_impl.PairwiseHingeLoss(name=None).compute(
      labels, logits, weights, api.Reduction.SUM_BY_NONZERO_WEIGHTS)
  assert float(loss.detach()) == pytest.approx(6.75, rel=1e-6)


def test_reference_groupwise_predict_golden(api):
  t, kernel, bias, _, _, _, _ = _case(api)
  out = _logits(api, kernel, bias, t([[178.], [155.]]), t([[10., 20.], [50., 30.]]))
  assert out.tolist() == [[254., 254.], [356., 356.]]
  out = _logits(api, kernel, bias, t([[178.]]), t([[20., 10., 10., 10.]]))
  assert out.tolist() == [[254., 239., 229., 244.]]


def test_reference_groupwise_adagrad_step_golden(api):
  t, kernel, bias, context, age, labels, weights = _case(api)
  logits = _logits(api, kernel, bias, context, age, labels)
  loss = api.losses_impl.PairwiseHingeLoss(name=None).compute(
      labels, logits, weights, api.Reduction.SUM_BY_NONZERO_WEIGHTS)
  loss.backward()
  # tf.compat.v1.train.AdagradOptimizer(0.1), accumulators initialised to 1e-12:
  # accum += g^2; var -= lr * g / sqrt(accum)
  with torch.no_grad():
    if api.name == 'cuda':
      from ranking_b200 import _C
      flat = torch.cat([kernel.reshape(-1), bias]).contiguous()
      grad = torch.cat([kernel.grad.reshape(-1), bias.grad]).contiguous()
      accum = torch.full_like(flat, 1e-12)
      _C.check(_C.lib.tfr_optimizer_step(_C.ptr(flat), _C.ptr(grad), _C.ptr(accum),
                                         flat.numel(), 1, 0.1, 0.0, 1.0, _C.stream()))
      new_kernel, new_bias = flat[:6].reshape(3, 2), flat[6:]
    else:
      upd = lambda p: p - 0.1 * p.grad / (1e-12 + p.grad * p.grad).sqrt()   # noqa: E731
      new_kernel, new_bias = upd(kernel), upd(bias)
  exp = torch.tensor([[1., 1.], [1.9, 2.1], [3.1, 2.9]], dtype=new_kernel.dtype)
  assert torch.allclose(new_kernel.cpu(), exp, rtol=1e-6, atol=1e-6), new_kernel
  assert torch.allclose(new_bias.cpu(), torch.ones(2, dtype=new_bias.dtype), atol=1e-6)
  # predictions after the step (model_test.py:448-468)
  nk, nb = new_kernel.detach(), new_bias.detach()
  out = _logits(api, nk, nb, t([[178.], [155.]]), t([[10., 20.], [50., 30.]]))
  assert torch.allclose(out.cpu(), torch.tensor([[255., 253.], [354., 358.]], dtype=out.dtype),
                        rtol=1e-6)
  out = _logits(api, nk, nb, t([[178.]]), t([[20., 10., 10., 10.]]))
  assert torch.allclose(out.cpu(), torch.tensor([[253., 239.5, 229., 244.5]], dtype=out.dtype),
                        rtol=1e-6)
