"""GPU parity tests added in round 2 (CUDA path through the C ABI vs the CPU oracle).

Tolerances, stated where they are used:
  * losses: 1e-5 relative (north_star);
  * score gradients: PER ELEMENT |got - ref| <= 1e-5 |ref| + 1e-5 mean_list|ref|
    (the absolute floor is the list's mean gradient magnitude: single entries are sums
    of up to N signed fp32 terms and may cancel to ~0);
  * integer outputs (ranks): exact.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _batch(b, n, seed, pad=True, holes=False):
  g = torch.Generator().manual_seed(seed)
  scores = torch.randn(b, n, generator=g) * 2.0
  probs = torch.tensor([.55, .25, .12, .06, .02])
  labels = torch.multinomial(probs, b * n, replacement=True,
                             generator=g).reshape(b, n).float()
  if pad:
    lens = torch.randint((n + 1) // 2, n + 1, (b,), generator=g)
    labels = torch.where(torch.arange(n).unsqueeze(0) < lens.unsqueeze(1),
                         labels, torch.full_like(labels, -1.))
  if holes:   # padding in the middle of the list, not only at the tail
    drop = torch.rand(b, n, generator=g) < 0.2
    labels = torch.where(drop, torch.full_like(labels, -1.), labels)
  item_w = torch.rand(b, n, generator=g) + 0.5
  return scores, labels, item_w


def assert_grad_close(got, ref, rtol=RTOL):
  got = got.detach().double().cpu()
  ref = ref.detach().double().cpu()
  floor = rtol * ref.abs().mean(dim=-1, keepdim=True)
  bad = (got - ref).abs() > rtol * ref.abs() + floor + 1e-30
  assert not bool(bad.any()), (
      'per-element gradient check failed on %d entries; worst |d|=%.3e at ref=%.3e' %
      (int(bad.sum()), float((got - ref).abs().max()),
       float(ref.flatten()[(got - ref).abs().flatten().argmax()])))


def _check(cuda_loss, oracle_loss, scores, labels, weights):
  s_gpu = scores.cuda().requires_grad_()
  w_gpu = None if weights is None else weights.cuda()
  got = cuda_loss(labels.cuda(), s_gpu, w_gpu)
  got.backward()
  s_ref = scores.double().requires_grad_()
  w_ref = None if weights is None else weights.double()
  ref = oracle_loss(labels.double(), s_ref, w_ref)
  ref.backward()
  assert abs(float(got) - float(ref)) <= RTOL * max(1.0, abs(float(ref))), (
      float(got), float(ref))
  assert_grad_close(s_gpu.grad, s_ref.grad)


LAMBDAS = {
    'none': lambda K: None,
    'label_diff': lambda K: K.LabelDiffLambdaWeight(),
    'dcg_smooth': lambda K: K.DCGLambdaWeight(topn=20, smooth_fraction=0.4),
    'ndcg': lambda K: K.NDCGLambdaWeight(),
    'ndcg_top7': lambda K: K.NDCGLambdaWeight(topn=7),
    'ndcg_v2_top5': lambda K: K.NDCGLambdaWeightV2(topn=5),
    'yeti': lambda K: K.YetiDCGLambdaWeight(topn=4, normalized=True),
    'precision_top3': lambda K: K.PrecisionLambdaWeight(topn=3),
}


# N covers every tiling of the triangular K1 kernel: 1 / 2 / 4 / 8-tile chunks, several
# chunks (N > 256), partial tiles, the 8 / 4 / 2 lists-per-CTA packings and N = 1024.
@pytest.mark.parametrize('n', [1, 2, 5, 31, 32, 33, 64, 100, 128, 200, 257, 300, 513])
@pytest.mark.parametrize('lam', ['none', 'ndcg', 'dcg_smooth'])
def test_k1_triangular_sizes(cuda_api, oracle_api, n, lam):
  scores, labels, item_w = _batch(11, n, seed=100 + n)
  KC, KO = cuda_api.keras_losses, oracle_api.keras_losses
  _check(KC.PairwiseLogisticLoss(lambda_weight=LAMBDAS[lam](KC)),
         KO.PairwiseLogisticLoss(lambda_weight=LAMBDAS[lam](KO)), scores, labels, item_w)


@pytest.mark.parametrize('cls', ['PairwiseLogisticLoss', 'PairwiseHingeLoss',
                                 'PairwiseSoftZeroOneLoss'])
@pytest.mark.parametrize('lam', sorted(LAMBDAS))
def test_k1_triangular_lambdas_with_holes(cuda_api, oracle_api, cls, lam):
  """Every lambda weight x phi at N = 150 with padding scattered inside the lists."""
  scores, labels, item_w = _batch(7, 150, seed=5, holes=True)
  KC, KO = cuda_api.keras_losses, oracle_api.keras_losses
  _check(getattr(KC, cls)(lambda_weight=LAMBDAS[lam](KC), temperature=0.8),
         getattr(KO, cls)(lambda_weight=LAMBDAS[lam](KO), temperature=0.8),
         scores, labels, item_w)


def test_k1_triangular_n1024(cuda_api, oracle_api):
  scores, labels, _ = _batch(3, 1024, seed=77)
  KC, KO = cuda_api.keras_losses, oracle_api.keras_losses
  _check(KC.PairwiseLogisticLoss(lambda_weight=KC.NDCGLambdaWeight()),
         KO.PairwiseLogisticLoss(lambda_weight=KO.NDCGLambdaWeight()), scores, labels, None)


@pytest.mark.parametrize('n', [19, 150, 400])
def test_k1_triangular_row_losses(cuda_api, oracle_api, n):
  """reduction=NONE needs the [B, N] row sums of the pair-loss matrix (hi item's row)."""
  scores, labels, item_w = _batch(5, n, seed=3 + n, holes=True)
  KC, KO = cuda_api.keras_losses, oracle_api.keras_losses
  up = torch.rand(5, n)
  for lam in ['none', 'ndcg']:
    s_gpu = scores.cuda().requires_grad_()
    out = KC.PairwiseLogisticLoss(reduction=KC.Reduction.NONE,
                                  lambda_weight=LAMBDAS[lam](KC))(
                                      labels.cuda(), s_gpu, item_w.cuda())
    (out * up.cuda()).sum().backward()
    s_ref = scores.double().requires_grad_()
    ref = KO.PairwiseLogisticLoss(reduction=KO.Reduction.NONE,
                                  lambda_weight=LAMBDAS[lam](KO))(
                                      labels.double(), s_ref, item_w.double())
    (ref * up.double()).sum().backward()
    assert_grad_close(out, ref)
    assert_grad_close(s_gpu.grad, s_ref.grad)


def test_k1_estimator_reductions_large(cuda_api, oracle_api):
  """SUM_BY_NONZERO_WEIGHTS / MEAN need the per-list sum W and #(W != 0)."""
  scores, labels, item_w = _batch(6, 170, seed=4)
  LC, LO = cuda_api.losses_impl, oracle_api.losses_impl
  for red in ['SUM', 'MEAN', 'SUM_BY_NONZERO_WEIGHTS', 'SUM_OVER_BATCH_SIZE']:
    for cls in ['PairwiseLogisticLoss', 'PairwiseHingeLoss']:
      got = getattr(LC, cls)(name=None, lambda_weight=LC.DCGLambdaWeight(topn=9)).compute(
          labels.cuda(), scores.cuda(), item_w.cuda(), getattr(LC.Reduction, red))
      ref = getattr(LO, cls)(name=None, lambda_weight=LO.DCGLambdaWeight(topn=9)).compute(
          labels.double(), scores.double(), item_w.double(), getattr(LO.Reduction, red))
      assert abs(float(got) - float(ref)) <= RTOL * max(1., abs(float(ref))), (
          red, cls, float(got), float(ref))


def test_k1_sort_order_equals_counting_ranks(cuda_api, oracle_api):
  """The walk order of the triangular kernel (bitonic sort; `ranks_out` of the C entry)
  must be the counting rank, exactly, with ties broken by index and padding last."""
  from ranking_b200 import _C
  b, n = 16, 300
  scores, labels, _ = _batch(b, n, seed=8, holes=True)
  scores[:, ::7] = scores[:, 1:2]          # many exact ties -> broken by index
  s, l = scores.cuda().contiguous(), labels.cuda().contiguous()
  grad = torch.empty_like(s)
  loss = torch.empty(b, device='cuda')
  ranks = torch.zeros(b, n, dtype=torch.int32, device='cuda')
  _C.check(_C.lib.tfr_pairwise_loss_fwd_bwd(
      _C.ptr(s), _C.ptr(l), None, 0, None, b, n, 1.0, _C.PHI_LOGISTIC, None, 1.0,
      _C.ptr(grad), None, _C.ptr(loss), None, None, _C.ptr(ranks), _C.stream()))
  ref = oracle_api.losses_impl._compute_ranks(scores.double(), labels >= 0)
  assert torch.equal(ranks.cpu().long(), ref)
