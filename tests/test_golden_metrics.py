"""Known-answer tests for NDCG / MRR metrics (metrics_impl_test.py:27-137,
643-839; keras/metrics.py docstrings).  Oracle + (under -m gpu) CUDA path."""
import math

import pytest
import torch


def log2p1(x):
  return math.log2(1. + x)


def _close(actual, expected, tol=1e-5):
  a = torch.as_tensor(actual).detach().double().cpu()
  e = torch.as_tensor(expected).double()
  torch.testing.assert_close(a.reshape(e.shape), e, rtol=tol, atol=tol)


# ------------------------------- MRR ---------------------------------------
def test_mrr_values(api):
  M = api.metrics_impl
  out, _ = M.MRRMetric(name=None, topn=None).compute(
      api.t([[0., 0., 1.]]), api.t([[1., 3., 2.]]), None)
  _close(out, [[1. / 2.]])
  out, _ = M.MRRMetric(name=None, topn=None).compute(
      api.t([[0., 0., 0.]]), api.t([[1., 3., 2.]]), None)
  _close(out, [[0.]])
  out, _ = M.MRRMetric(name=None, topn=1).compute(
      api.t([[0., 0., 1.]]), api.t([[1., 3., 2.]]), None)
  _close(out, [[0.]])


def test_mrr_topn(api):
  M = api.metrics_impl
  scores = api.t([[3., 2., 1.], [3., 2., 1.], [3., 2., 1.]])
  labels = api.t([[1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
  for topn, exp in [(1, [[1.], [0.], [0.]]), (2, [[1.], [1. / 2.], [0.]]),
                    (6, [[1.], [1. / 2.], [1. / 3.]])]:
    out, _ = M.MRRMetric(name=None, topn=topn).compute(labels, scores, None)
    _close(out, exp)


def test_mrr_padding_mask_batch(api):
  M = api.metrics_impl
  out, _ = M.MRRMetric(name=None, topn=None).compute(
      api.t([[0., 1., -1.]]), api.t([[1., 2., 3.]]), None)
  _close(out, [[1.]])
  out, _ = M.MRRMetric(name=None, topn=None).compute(
      api.t([[0., 1., 0.]]), api.t([[1., 2., 3.]]), None,
      mask=api.t([[True, True, False]], dtype=torch.bool))
  _close(out, [[1.]])
  out, _ = M.MRRMetric(name=None, topn=None).compute(
      api.t([[0., 0., 1.], [0., 1., 1.]]), api.t([[1., 3., 2.], [1., 2., 3.]]),
      None)
  _close(out, [[1. / 2.], [1.]])


def test_mrr_weights(api):
  M = api.metrics_impl
  _, w = M.MRRMetric(name=None, topn=None).compute(
      api.t([[1., 0., 0.], [0., 1., 1.]]), api.t([[1., 3., 2.], [1., 2., 3.]]),
      api.t([[2., 5., 1.], [1., 2., 3.]]))
  _close(w, [[2.], [(2. + 3.) / 2.]])
  _, w = M.MRRMetric(name=None, topn=None).compute(
      api.t([[0., 0., 0.], [0., 0., 0.]]), api.t([[1., 3., 2.], [1., 3., 2.]]),
      api.t([[2., 5., 1.], [1., 1., 0.]]))
  _close(w, [[1.], [1.]])
  _, w = M.MRRMetric(name=None, topn=2).compute(
      api.t([[1., 0., 1.], [0., 1., 1.]]), api.t([[3., 2., 1.], [1., 3., 2.]]),
      api.t([[2., 0., 5.], [1., 4., 2.]]))
  _close(w, [[(5. + 2.) / 2.], [(2. + 4.) / 2.]])


# ------------------------------- NDCG --------------------------------------
def test_ndcg_values(api):
  M = api.metrics_impl
  out, _ = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[0., 1., 0.]]), api.t([[3., 2., 1.]]), None)
  _close(out, [[(1. / log2p1(2.)) / (1. / log2p1(1.))]])
  out, _ = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[0., 0., 0.]]), api.t([[3., 2., 1.]]), None)
  _close(out, [[0.]])
  out, _ = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[0., 3., 1., 0.]]), api.t([[4., 3., 2., 1.]]), None)
  dcg = (2.**3. - 1.) / log2p1(2.) + 1. / log2p1(3.)
  max_dcg = (2.**3. - 1.) / log2p1(1.) + 1. / log2p1(2.)
  _close(out, [[dcg / max_dcg]])


def test_ndcg_custom_gain_and_discount(api):
  M = api.metrics_impl
  scores, labels = api.t([[4., 3., 2., 1.]]), api.t([[0., 3., 1., 0.]])
  out, _ = M.NDCGMetric(name=None, topn=None,
                        gain_fn=lambda label: label / 2.).compute(
                            labels, scores, None)
  dcg = (3. / 2.) / log2p1(2.) + (1. / 2.) / log2p1(3.)
  max_dcg = (3. / 2.) / log2p1(1.) + (1. / 2.) / log2p1(2.)
  _close(out, [[dcg / max_dcg]])
  out, _ = M.NDCGMetric(
      name=None, topn=None,
      rank_discount_fn=lambda rank: 1.0 / (rank + 10.0)).compute(
          labels, scores, None)
  dcg = (2.**3. - 1.) / (2. + 10.) + 1. / (3. + 10.)
  max_dcg = (2.**3. - 1.) / (1. + 10.) + 1. / (2. + 10.)
  _close(out, [[dcg / max_dcg]])


def test_ndcg_padding_and_mask(api):
  M = api.metrics_impl
  dcg = (2.**2. - 1.) / log2p1(3.) + 1. / log2p1(1.)
  max_dcg = (2.**2. - 1.) / log2p1(1.) + 1. / log2p1(2.)
  out, _ = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[2., -1., 1., 0.]]), api.t([[1., 4., 3., 2.]]), None)
  _close(out, [[dcg / max_dcg]])
  out, _ = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[2., 2., 1., 0.]]), api.t([[1., 4., 3., 2.]]), None,
      mask=api.t([[True, False, True, True]], dtype=torch.bool))
  _close(out, [[dcg / max_dcg]])


def test_ndcg_per_list_and_topn(api):
  M = api.metrics_impl
  out, _ = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[0., 1., 0.], [1., 1., 0.]]), api.t([[3., 2., 1.], [3., 1., 2.]]),
      None)
  dcg = [1. / log2p1(2.), 1. / log2p1(1.) + 1. / log2p1(3.)]
  max_dcg = [1. / log2p1(1.), 1. / log2p1(1.) + 1. / log2p1(2.)]
  _close(out, [[dcg[0] / max_dcg[0]], [dcg[1] / max_dcg[1]]])

  scores = api.t([[3., 2., 1.], [3., 2., 1.], [3., 2., 1.]])
  labels = api.t([[1., 0., 2.], [0., 1., 0.], [0., 0., 1.]])
  max_dcg_top1 = [(2.**2. - 1.) / log2p1(1.), 1. / log2p1(1.), 1. / log2p1(1.)]
  max_dcg = [(2.**2. - 1.) / log2p1(1.) + 1. / log2p1(2.), 1. / log2p1(1.),
             1. / log2p1(1.)]
  out, _ = M.NDCGMetric(name=None, topn=1).compute(labels, scores, None)
  _close(out, [[(1. / log2p1(1.)) / max_dcg_top1[0]], [0.], [0.]])
  out, _ = M.NDCGMetric(name=None, topn=2).compute(labels, scores, None)
  _close(out, [[(1. / log2p1(1.)) / max_dcg[0]],
               [(1. / log2p1(2.)) / max_dcg[1]], [0.]])
  out, _ = M.NDCGMetric(name=None, topn=6).compute(labels, scores, None)
  _close(out, [[(1. / log2p1(1.) + (2.**2. - 1.) / log2p1(3.)) / max_dcg[0]],
               [(1. / log2p1(2.)) / max_dcg[1]],
               [(1. / log2p1(3.)) / max_dcg[2]]])


def test_ndcg_weights(api):
  M = api.metrics_impl
  _, w = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[1., 0., 2.]]), api.t([[1., 3., 2.]]), api.t([[3., 7., 9.]]))
  _close(w, [[(1. * 3. + (2.**2. - 1.) * 9.) / (1. + (2.**2. - 1.))]])
  _, w = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[0., 0., 0.]]), api.t([[1., 3., 2.]]), api.t([[2., 4., 4.]]))
  _close(w, [[1.]])
  _, w = M.NDCGMetric(name=None, topn=None,
                      gain_fn=lambda label: label + 5.).compute(
                          api.t([[1., 0., 2.]]), api.t([[1., 3., 2.]]),
                          api.t([[3., 7., 9.]]))
  _close(w, [[((1. + 5.) * 3. + (0. + 5.) * 7. + (2. + 5.) * 9.) /
              ((1. + 5.) + (0. + 5.) + (2. + 5.))]])


def test_ndcg_with_weights(api):
  M = api.metrics_impl
  out, _ = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[1., 2., 3.]]), api.t([[1., 2., 3.]]), api.t([[4., 1., 1.]]))
  _close(out, [[((2**3. - 1.) / log2p1(1) + (2**2. - 1.) / log2p1(2) +
                 (2**1. - 1.) / log2p1(3) * 4.) /
                ((2**3. - 1.) / log2p1(1) + (2**2. - 1.) / log2p1(3) +
                 (2**1. - 1.) / log2p1(2) * 4.)]])
  out, w = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[1., 2., 3.]]), api.t([[1., 2., 3.]]), api.t([[0., 0., 0.]]))
  _close(out, [[0.0]])
  _close(w, [[0.0]])


def test_ndcg_mean_of_lists_without_relevance(api):
  """keras/metrics_test.py:901-907 behaviour: a list without relevant items
  counts as NDCG 0 with the batch-average weight."""
  M = api.metrics_impl
  out, w = M.NDCGMetric(name=None, topn=None).compute(
      api.t([[0., 0., 1.], [0., 0., 0.]]), api.t([[1., 3., 2.], [1., 2., 3.]]),
      None)
  _close(out, [[(1. / log2p1(2.)) / 1.], [0.]])
  _close(w, [[1.], [1.]])
