"""Known-answer tests for the metrics served by K4's extended outputs (Hits, ARP,
Recall, Precision, MAP, DCG, OPA), transcribed from the reference's own tests
(metrics_impl_test.py:139-1128).  Each case: (metric class, ctor kwargs, labels,
scores, weights, mask, expected value or None, expected weight or None).
Runs against the CPU oracle and, under `-m gpu`, the CUDA kernel through the C ABI.
"""
import math

import pytest
import torch


def log2p1(x):
  return math.log2(1. + x)


T, F = True, False
CASES = [
    # ---- Hits :139-249
    ('HitsMetric', dict(topn=None), [[0., 0., 1.]], [[1., 3., 2.]], None, None, [[1.]], None),
    ('HitsMetric', dict(topn=None), [[0., 0., 0.]], [[1., 3., 2.]], None, None, [[0.]], None),
    ('HitsMetric', dict(topn=1), [[0., 0., 1.]], [[1., 3., 2.]], None, None, [[0.]], None),
    ('HitsMetric', dict(topn=1), [[1., 0., 0.], [0., 1., 0.], [0., 0., 1.]],
     [[3., 2., 1.]] * 3, None, None, [[1.], [0.], [0.]], None),
    ('HitsMetric', dict(topn=2), [[1., 0., 0.], [0., 1., 0.], [0., 0., 1.]],
     [[3., 2., 1.]] * 3, None, None, [[1.], [1.], [0.]], None),
    ('HitsMetric', dict(topn=6), [[1., 0., 0.], [0., 1., 0.], [0., 0., 1.]],
     [[3., 2., 1.]] * 3, None, None, [[1.], [1.], [1.]], None),
    ('HitsMetric', dict(topn=1), [[0., 1., -1.]], [[1., 2., 3.]], None, None, [[1.]], None),
    ('HitsMetric', dict(topn=1), [[0., 1., 0.]], [[1., 2., 3.]], None, [[T, T, F]], [[1.]], None),
    ('HitsMetric', dict(topn=1), [[0., 0., 1.], [0., 1., 1.]], [[1., 3., 2.], [1., 2., 3.]],
     None, None, [[0.], [1.]], None),
    ('HitsMetric', dict(topn=None), [[1., 0., 0.], [0., 1., 1.]], [[1., 3., 2.], [1., 2., 3.]],
     [[2., 5., 1.], [1., 2., 3.]], None, None, [[2.], [(2. + 3.) / 2.]]),
    ('HitsMetric', dict(topn=None), [[0., 0., 0.], [0., 0., 0.]], [[1., 3., 2.], [1., 3., 2.]],
     [[2., 5., 1.], [1., 1., 0.]], None, None, [[1.], [1.]]),
    ('HitsMetric', dict(topn=1), [[1., 0., 1.], [0., 1., 1.]], [[3., 2., 1.], [1., 3., 2.]],
     [[2., 0., 5.], [1., 3., 2.]], None, None, [[(5. + 2.) / 2.], [(2. + 3.) / 2.]]),
    # ---- ARP :251-333
    ('ARPMetric', {}, [[0., 0., 1.]], [[1., 3., 2.]], None, None, [[2.]], None),
    ('ARPMetric', {}, [[0., 0., 1.], [0., 1., 2.]], [[1., 3., 2.], [1., 2., 3.]], None, None,
     [[2.], [((1. * 2.) + (2. * 1.)) / (2. + 1.)]], None),
    ('ARPMetric', {}, [[0., 0., 0.]], [[1., 3., 2.]], None, None, [[0.]], None),
    ('ARPMetric', {}, [[1., -1., 1., -1., 0.]], [[1., 5., 4., 3., 2.]], None, None, [[2.]], None),
    ('ARPMetric', {}, [[1., 0., 1., 1., 0.]], [[1., 5., 4., 3., 2.]], None, [[T, F, T, F, T]],
     [[2.]], None),
    ('ARPMetric', {}, [[0., 0., 1.], [0., 1., 2.]], [[1., 3., 2.], [1., 2., 3.]],
     [[1., 2., 3.], [4., 5., 6.]], None,
     [[2.], [2. * (5. / (5. + 6. * 2.)) + 1. * (6. * 2. / (5. + 6. * 2.))]],
     [[3.], [5. + 6. * 2.]]),
    # ---- Recall :335-436
    ('RecallMetric', dict(topn=1), [[0., 0., 1.]], [[1., 3., 2.]], None, None, [[0.]], None),
    ('RecallMetric', dict(topn=2), [[0., 0., 1.]], [[1., 3., 2.]], None, None, [[1.]], None),
    ('RecallMetric', dict(topn=6), [[0., 0., 1.]], [[1., 3., 2.]], None, None, [[1.]], None),
    ('RecallMetric', dict(topn=2), [[1., 0., 1.], [0., 1., 1.]], [[1., 3., 2.], [1., 3., 4.]],
     None, None, [[1. / 2.], [1.]], None),
    ('RecallMetric', dict(topn=None), [[1., 1., 0.]], [[1., 3., 2.]], [[3., 9., 2.]], None,
     None, [[(3. + 9.) / 2.]]),
    ('RecallMetric', dict(topn=None), [[4., 0., 2.]], [[1., 3., 2.]], [[3., 9., 2.]], None,
     None, [[(3. + 2.) / 2.]]),
    ('RecallMetric', dict(topn=1), [[1., 1., 0.]], [[1., 3., 2.]], [[3., 9., 2.]], None,
     None, [[(3. + 9.) / 2.]]),
    ('RecallMetric', dict(topn=None), [[0., 0., 0.]], [[1., 3., 2.]], None, None, None, [[1.]]),
    # ---- Precision :438-539
    ('PrecisionMetric', dict(topn=None), [[0., 0., 1.]], [[1., 3., 2.]], None, None,
     [[1. / 3.]], None),
    ('PrecisionMetric', dict(topn=None), [[0., 0., 0.]], [[1., 3., 2.]], None, None, [[0.]], None),
    ('PrecisionMetric', dict(topn=None), [[0., 0., 1., 1.], [0., 0., 1., 0.]],
     [[1., 3., 2., 4.], [4., 1., 3., 2.]], None, None, [[2. / 4.], [1. / 4.]], None),
    ('PrecisionMetric', dict(topn=1), [[1., 0., 1.], [0., 1., 0.], [0., 0., 1.]],
     [[3., 2., 1.]] * 3, None, None, [[1.], [0.], [0.]], None),
    ('PrecisionMetric', dict(topn=2), [[1., 0., 1.], [0., 1., 0.], [0., 0., 1.]],
     [[3., 2., 1.]] * 3, None, None, [[1. / 2.], [1. / 2.], [0.]], None),
    ('PrecisionMetric', dict(topn=6), [[1., 0., 1.], [0., 1., 0.], [0., 0., 1.]],
     [[3., 2., 1.]] * 3, None, None, [[2. / 3.], [1. / 3.], [1. / 3.]], None),
    ('PrecisionMetric', dict(topn=None), [[0., 0., 1., -1.], [0., -1., 1., -1.]],
     [[1., 3., 2., 4.], [4., 1., 3., 2.]], None, None, [[1. / 3.], [1. / 2.]], None),
    ('PrecisionMetric', dict(topn=None), [[0., 0., 1., 0.], [0., 1., 1., 0.]],
     [[1., 3., 2., 4.], [4., 1., 3., 2.]], None, [[T, T, T, F], [T, F, T, F]],
     [[1. / 3.], [1. / 2.]], None),
    ('PrecisionMetric', dict(topn=None), [[1., 0., 2.]], [[1., 3., 2.]], [[13., 7., 29.]], None,
     None, [[(13. + 29.) / 2.]]),
    ('PrecisionMetric', dict(topn=1), [[1., 1., 0.]], [[1., 3., 2.]], [[3., 7., 15.]], None,
     None, [[(3. + 7.) / 2.]]),
    ('PrecisionMetric', dict(topn=1), [[0., 0., 0.]], [[1., 3., 2.]], [[3., 7., 15.]], None,
     None, [[1.]]),
    # ---- MAP :541-650
    ('MeanAveragePrecisionMetric', dict(topn=None), [[0., 1., 0.]], [[3., 2., 1.]], None, None,
     [[(1. / 2.) / 1.]], None),
    ('MeanAveragePrecisionMetric', dict(topn=None), [[0., 2., 1., 3.]], [[3., 4., 1., 2.]],
     None, None, [[(1. + 2. / 3. + 3. / 4.) / 3.]], None),
    ('MeanAveragePrecisionMetric', dict(topn=None), [[0., 0., 0.]], [[3., 2., 1.]], None, None,
     [[0.]], None),
    ('MeanAveragePrecisionMetric', dict(topn=None), [[0., 0., 1.], [0., 1., 1.]],
     [[1., 3., 2.], [1., 3., 2.]], None, None, [[(1. / 2.) / 1.], [(1. / 1. + 2. / 2.) / 2.]],
     None),
    ('MeanAveragePrecisionMetric', dict(topn=1), [[1., 0., 2.], [0., 1., 0.], [0., 0., 1.]],
     [[3., 2., 1.]] * 3, None, None, [[1. / 2.], [0.], [0.]], None),
    ('MeanAveragePrecisionMetric', dict(topn=2), [[1., 0., 2.], [0., 1., 0.], [0., 0., 1.]],
     [[3., 2., 1.]] * 3, None, None, [[1. / 2.], [(1. / 2.) / 1.], [0.]], None),
    ('MeanAveragePrecisionMetric', dict(topn=6), [[1., 0., 2.], [0., 1., 0.], [0., 0., 1.]],
     [[3., 2., 1.]] * 3, None, None,
     [[(1. + 2. / 3.) / 2.], [(1. / 2.) / 1.], [(1. / 3.) / 1.]], None),
    ('MeanAveragePrecisionMetric', dict(topn=None), [[1., 0., 2.]], [[1., 3., 2.]],
     [[13., 7., 29.]], None, None, [[(13. + 29.) / 2.]]),
    ('MeanAveragePrecisionMetric', dict(topn=None), [[0., 0., 0.]], [[1., 3., 2.]],
     [[3., 7., 15.]], None, None, [[1.]]),
    # ---- DCG :842-1001
    ('DCGMetric', dict(topn=None), [[0., 1., 0.]], [[3., 2., 1.]], None, None,
     [[1. / log2p1(2.)]], None),
    ('DCGMetric', dict(topn=None), [[0., 0., 0.]], [[3., 2., 1.]], None, None, [[0.]], None),
    ('DCGMetric', dict(topn=None), [[0., 3., 1., 0.]], [[4., 3., 2., 1.]], None, None,
     [[(2. ** 3. - 1.) / log2p1(2.) + 1. / log2p1(3.)]], None),
    ('DCGMetric', dict(topn=None), [[2., -1., 1., 0.]], [[1., 4., 3., 2.]], None, None,
     [[(2. ** 2. - 1.) / log2p1(3.) + 1. / log2p1(1.)]], None),
    ('DCGMetric', dict(topn=None), [[2., 2., 1., 0.]], [[1., 4., 3., 2.]], None, [[T, F, T, T]],
     [[(2. ** 2. - 1.) / log2p1(3.) + 1. / log2p1(1.)]], None),
    ('DCGMetric', dict(topn=None), [[0., 1., 0.], [1., 1., 0.]], [[3., 2., 1.], [3., 1., 2.]],
     None, None, [[1. / log2p1(2.)], [1. / log2p1(1.) + 1. / log2p1(3.)]], None),
    ('DCGMetric', dict(topn=1), [[1., 0., 2.], [0., 1., 0.], [0., 0., 1.]], [[3., 2., 1.]] * 3,
     None, None, [[1. / log2p1(1.)], [0.], [0.]], None),
    ('DCGMetric', dict(topn=2), [[1., 0., 2.], [0., 1., 0.], [0., 0., 1.]], [[3., 2., 1.]] * 3,
     None, None, [[1. / log2p1(1.)], [1. / log2p1(2.)], [0.]], None),
    ('DCGMetric', dict(topn=6), [[1., 0., 2.], [0., 1., 0.], [0., 0., 1.]], [[3., 2., 1.]] * 3,
     None, None, [[1. / log2p1(1.) + (2. ** 2. - 1.) / log2p1(3.)], [1. / log2p1(2.)],
                  [1. / log2p1(3.)]], None),
    ('DCGMetric', dict(topn=None), [[1., 0., 2.]], [[1., 3., 2.]], [[3., 7., 9.]], None, None,
     [[(1. * 3. + (2. ** 2. - 1.) * 9.) / (1. + (2. ** 2. - 1.))]]),
    ('DCGMetric', dict(topn=None), [[1., 0., 2.]], [[1., 3., 2.]], [[0., 0., 0.]], None,
     [[0.]], [[0.]]),
    ('DCGMetric', dict(topn=None), [[0., 0., 0.]], [[1., 3., 2.]], [[2., 4., 4.]], None, None,
     [[1.]]),
    # ---- OPA :1003-1127
    ('OPAMetric', {}, [[0., 1., 0.]], [[3., 2., 1.]], None, None, [[1. / 2.]], [[2.]]),
    ('OPAMetric', {}, [[0., 0., 0.]], [[3., 2., 1.]], None, None, [[0.]], [[0.]]),
    ('OPAMetric', {}, [[1., 3., 0., 1.]], [[4., 3., 2., 1.]], None, None, [[3. / 5.]], [[5.]]),
    ('OPAMetric', {}, [[2., -1., 1., 0.]], [[4., 1., 2., 3.]], None, None, [[2. / 3.]], [[3.]]),
    ('OPAMetric', {}, [[2., 1., 1., 0.]], [[4., 1., 2., 3.]], None, [[T, F, T, T]],
     [[2. / 3.]], [[3.]]),
    ('OPAMetric', {}, [[0., 1., 0.], [1., 0., 1.]], [[3., 2., 1.], [3., 1., 2.]], None, None,
     [[1. / 2.], [2. / 2.]], [[2.], [2.]]),
    ('OPAMetric', {}, [[1., 0., 2.]], [[1., 3., 2.]], [[3., 7., 9.]], None,
     [[9. / (9. + 9. + 3.)]], [[9. + 9. + 3.]]),
    ('OPAMetric', {}, [[0., 0., 0.]], [[1., 3., 2.]], [[2., 4., 4.]], None, None, [[0.]]),
]


def _close(actual, expected):
  a = torch.as_tensor(actual).detach().double().cpu()
  e = torch.as_tensor(expected).double()
  torch.testing.assert_close(a.reshape(e.shape), e, rtol=1e-5, atol=1e-6)


# ---- BPref metrics_impl_test.py:1448-1610
_BP = 1. / 2. * ((1. - 1. / 2.) + (1. - 2. / 2.))
CASES += [
    ('BPrefMetric', dict(topn=None), [[0., 1., 0., 1.]], [[4., 3., 2., 1.]], None, None,
     [[_BP]], None),
    ('BPrefMetric', dict(topn=None), [[0., 1., 0., 2.]], [[4., 3., 2., 1.]], None, None,
     [[_BP]], None),
    ('BPrefMetric', dict(topn=None), [[0., 0., 0.]], [[3., 2., 1.]], None, None, [[0.]], None),
    ('BPrefMetric', dict(topn=None), [[1., 1., 1.]], [[3., 2., 1.]], None, None, [[1.]], None),
    ('BPrefMetric', dict(topn=None, use_trec_version=False), [[1., 1., 1.]], [[3., 2., 1.]],
     None, None, [[1.]], None),
    ('BPrefMetric', dict(topn=None), [[0., 1., 1.]], [[3., 2., 1.]], None, None, [[0.]], None),
    ('BPrefMetric', dict(topn=None, use_trec_version=False), [[0., 1., 1.]], [[3., 2., 1.]],
     None, None, [[0.5]], None),
    ('BPrefMetric', dict(topn=3), [[0., 0., 0., 1.]], [[3., 2., 1., 0.]], None, None, [[0.]],
     None),
    ('BPrefMetric', dict(topn=5, use_trec_version=False),
     [[0., 1., 1., 1., 1., 0., 0., 0., 1., 1.]], [[5., 4., 3., 2., 1., 0., 0., 0., 0., 0.]],
     None, None, [[(4. * (1. - 1. / 6.)) / 6.]], None),
    ('BPrefMetric', dict(topn=5), [[0., 1., 1., 1., 1., 0., 0., 0., 1., 1.]],
     [[5., 4., 3., 2., 1., 0., 0., 0., 0., 0.]], None, None, [[(4. * (1. - 1. / 4.)) / 6.]],
     None),
    ('BPrefMetric', dict(topn=None), [[-1., 0., -1., 1., 0., 1.]], [[6., 5., 4., 3., 2., 1.]],
     None, None, [[_BP]], None),
    ('BPrefMetric', dict(topn=None), [[1., 0., 2.]], [[1., 3., 2.]], [[13., 7., 29.]], None,
     None, [[(13. + 29.) / 2.]]),
    ('BPrefMetric', dict(topn=1), [[1., 1., 0.]], [[1., 3., 2.]], [[3., 7., 15.]], None, None,
     [[(3. + 7.) / 2.]]),
    ('BPrefMetric', dict(topn=None), [[0., 0., 0.]], [[1., 3., 2.]], [[3., 7., 15.]], None,
     None, [[1.]]),
    # ragged case :1566-1573 as padded lists
    ('BPrefMetric', dict(topn=None), [[0., 0., 1., 0.], [0., 1., 1., -1.]],
     [[1., 3., 2., 4.], [1., 2., 3., 0.]], None, None, [[0.], [1.]], None),
]


@pytest.mark.parametrize('case', range(len(CASES)))
def test_reference_cases(api, case):
  cls, kw, labels, scores, weights, mask, want_v, want_w = CASES[case]
  metric = getattr(api.metrics_impl, cls)(name=None, **kw)
  m = None if mask is None else torch.tensor(mask, device=api.device)
  w = None if weights is None else api.t(weights)
  got_v, got_w = metric.compute(api.t(labels), api.t(scores), w, mask=m)
  if want_v is not None:
    _close(got_v, want_v)
  if want_w is not None:
    _close(got_w, want_w)


def test_dcg_custom_gain_and_discount(api):
  """metrics_impl_test.py:866-885, 991-1001."""
  M = api.metrics_impl
  labels, scores = api.t([[0., 3., 1., 0.]]), api.t([[4., 3., 2., 1.]])
  v, _ = M.DCGMetric(name=None, topn=None, gain_fn=lambda l: l / 2.).compute(labels, scores)
  _close(v, [[(3. / 2.) / log2p1(2.) + (1. / 2.) / log2p1(3.)]])
  v, _ = M.DCGMetric(name=None, topn=None,
                     rank_discount_fn=lambda r: 1.0 / (r + 10.0)).compute(labels, scores)
  _close(v, [[(2. ** 3. - 1.) / (2. + 10.) + 1. / (3. + 10.)]])
  _, w = M.DCGMetric(name=None, topn=None, gain_fn=lambda l: l + 3.).compute(
      api.t([[1., 0., 2.]]), api.t([[1., 3., 2.]]), api.t([[4., 1., 9.]]))
  _close(w, [[((1. + 3.) * 4. + (0. + 3.) * 1. + (2. + 3.) * 9.) /
              ((1. + 3.) + (0. + 3.) + (2. + 3.))]])


# ---------------------------------------------------------------------------
# Diversity metrics: metrics_impl_test.py:1129-1420
# ---------------------------------------------------------------------------
def _a(c, r, alpha=0.5, disc=None):
  """One alphaDCG term: (1 - alpha)^c at rank r."""
  return (1. - alpha) ** c * (disc(r) if disc else 1. / log2p1(r))


DIV_CASES = [
    ('PrecisionIAMetric', dict(topn=None), [[[0., 0.], [1., 0.], [0., 1.]]], [[1., 3., 2.]],
     None, None, [[2. / (2. * 3.)]], None),
    ('PrecisionIAMetric', dict(topn=None), [[[0.], [1.], [0.]]], [[1., 3., 2.]], None, None,
     [[1. / (1. * 3.)]], None),
    ('PrecisionIAMetric', dict(topn=None),
     [[[0., 0., 1., 0.], [1., 1., 1., 1.], [0., 1., 1., 0.]]], [[1., 3., 2.]], None, None,
     [[7. / (4. * 3.)]], None),
    ('PrecisionIAMetric', dict(topn=None), [[[0., 0.], [1., 0.], [1., 1.], [0., 1.]]],
     [[1., 3., 4., 2.]], None, [[T, T, F, T]], [[2. / (2. * 3.)]], None),
    ('PrecisionIAMetric', dict(topn=None), [[[0., 0.], [0., 1.], [0., 1.], [0., 0.]]],
     [[1., 3., 2., 4.]], None, None, [[2. / (1. * 4.)]], None),
    ('PrecisionIAMetric', dict(topn=None), [[[0., 0.], [0., 0.], [0., 0.]]], [[1., 3., 2.]],
     None, None, [[0.]], None),
    ('PrecisionIAMetric', dict(topn=None),
     [[[0., 0.], [0., 0.], [1., 1.], [1., 0.]], [[1., 0.], [1., 1.], [1., 0.], [0., 1.]]],
     [[1., 3., 2., 4.], [4., 1., 3., 2.]], None, None, [[3. / (2. * 4.)], [5. / (2. * 4.)]],
     None),
] + [
    ('PrecisionIAMetric', dict(topn=k),
     [[[1., 1.], [0., 0.], [1., 0.]], [[0., 0.], [0., 1.], [1., 0.]],
      [[0., 1.], [0., 0.], [1., 0.]], [[1., 1.], [1., 1.], [1., 1.]]],
     [[3., 2., 1.]] * 4, None, None, want, None)
    for k, want in ((1, [[2. / 2.], [0.], [1. / 2.], [2. / 2.]]),
                    (2, [[2. / 4.], [1. / 4.], [1. / 4.], [4. / 4.]]),
                    (6, [[3. / 6.], [2. / 6.], [2. / 6.], [6. / 6.]]))
] + [
    ('PrecisionIAMetric', dict(topn=None), [[[0., 1.], [0., 0.], [1., 1.]]], [[1., 3., 2.]],
     [[3., 7., 9.]], None, None, [[(3. + 9.) / 2.]]),
    ('PrecisionIAMetric', dict(topn=1), [[[1., 1.], [1., 0.], [0., 0.]]], [[1., 3., 2.]],
     [[3., 4., 5.]], None, None, [[(3. + 4.) / 2.]]),
    ('PrecisionIAMetric', dict(topn=None), [[[0., 0.], [0., 0.], [0., 0.]]], [[1., 3., 2.]],
     [[3., 7., 2.]], None, None, [[1.]]),
    # ---- alphaDCG :1267-1420
    ('AlphaDCGMetric', dict(topn=None), [[[0., 0.], [1., 0.], [0., 1.]]], [[1., 3., 2.]],
     None, None, [[_a(0, 1) + _a(0, 2)]], None),
    ('AlphaDCGMetric', dict(topn=None), [[[0.], [0.], [1.]]], [[1., 3., 2.]], None, None,
     [[_a(0, 2)]], None),
    ('AlphaDCGMetric', dict(topn=None),
     [[[0., 1., 0., 0.], [1., 1., 0., 1.], [0., 1., 1., 0.]]], [[1., 3., 2.]], None, None,
     [[3 * _a(0, 1) + _a(1, 2) + _a(0, 2) + _a(2, 3)]], None),
    ('AlphaDCGMetric', dict(topn=None), [[[0., 0.], [0., 0.], [0., 0.]]], [[1., 3., 2.]],
     None, None, [[0.]], None),
    ('AlphaDCGMetric', dict(topn=None), [[[0., 0.], [1., 1.], [1., 0.], [0., 1.], [1., 0.]]],
     [[1., 4., 3., 2., 2.]], None, [[T, F, T, T, F]], [[_a(0, 1) + _a(0, 2)]], None),
    ('AlphaDCGMetric', dict(topn=None),
     [[[0., 0.], [0., 0.], [1., 1.], [1., 0.]], [[1., 0.], [1., 1.], [1., 0.], [0., 1.]]],
     [[1., 3., 2., 4.], [4., 1., 3., 2.]], None, None,
     [[_a(0, 1) + _a(0, 3) + _a(1, 3)],
      [_a(0, 1) + _a(1, 2) + _a(0, 3) + _a(1, 4) + _a(2, 4)]], None),
    ('AlphaDCGMetric', dict(topn=None, alpha=0.2), [[[1., 1.], [0., 1.], [0., 1.], [1., 0.]]],
     [[1., 3., 2., 4.]], None, None,
     [[_a(0, 1, .2) + _a(0, 2, .2) + _a(1, 3, .2) + _a(1, 4, .2) + _a(2, 4, .2)]], None),
    ('AlphaDCGMetric', dict(topn=None, alpha=0.95),
     [[[1., 1.], [0., 1.], [0., 1.], [1., 0.]]], [[1., 3., 2., 4.]], None, None,
     [[_a(0, 1, .95) + _a(0, 2, .95) + _a(1, 3, .95) + _a(1, 4, .95) + _a(2, 4, .95)]], None),
] + [
    ('AlphaDCGMetric', dict(topn=k), [[[1., 0.], [0., 0.], [1., 0.], [1., 1.], [0., 1.]]],
     [[3., 2., 1., 4., 5.]], None, None, want, None)
    for k, want in ((1, [[_a(0, 1)]]), (2, [[_a(0, 1) + _a(0, 2) + _a(1, 2)]]),
                    (6, [[_a(0, 1) + _a(0, 2) + _a(1, 2) + _a(1, 3) + _a(2, 5)]]))
] + [
    ('AlphaDCGMetric', dict(topn=None), [[[0., 1.], [0., 0.], [1., 1.]]], [[1., 3., 2.]],
     [[3., 7., 9.]], None, None, [[(3. + 9.) / 2.]]),
    ('AlphaDCGMetric', dict(topn=1), [[[1., 1.], [1., 0.], [0., 0.]]], [[1., 3., 2.]],
     [[3., 4., 5.]], None, None, [[(3. + 4.) / 2.]]),
    ('AlphaDCGMetric', dict(topn=None), [[[0., 0.], [0., 0.], [0., 0.]]], [[1., 3., 2.]],
     [[3., 7., 2.]], None, None, [[1.]]),
]


@pytest.mark.parametrize('case', range(len(DIV_CASES)))
def test_diversity_reference_cases(api, case):
  cls, kw, labels, scores, weights, mask, want_v, want_w = DIV_CASES[case]
  metric = getattr(api.metrics_impl, cls)(name=None, **kw)
  m = None if mask is None else torch.tensor(mask, device=api.device)
  w = None if weights is None else api.t(weights)
  got_v, got_w = metric.compute(api.t(labels), api.t(scores), w, mask=m)
  if want_v is not None:
    _close(got_v, want_v)
  if want_w is not None:
    _close(got_w, want_w)


def test_alpha_dcg_custom_rank_discount(api):
  """metrics_impl_test.py:1366-1378."""
  disc = lambda rank: 1. / (10. + rank)
  metric = api.metrics_impl.AlphaDCGMetric(name=None, topn=None, rank_discount_fn=disc)
  v, _ = metric.compute(api.t([[[1., 0.], [1., 1.], [0., 1.], [1., 0.]]]),
                        api.t([[1., 3., 2., 4.]]))
  _close(v, [[_a(0, 1, disc=disc) + _a(0, 2, disc=disc) + _a(1, 2, disc=disc) +
              _a(1, 3, disc=disc) + _a(2, 4, disc=disc)]])
