"""Known-answer tests for the index/sort helpers and the scorer plumbing of the
oracle: utils_test.py:64-126,203-257; keras/layers_test.py:34-135;
model_test.py:52-112,223-277 (deterministic, shuffle_ties=False cases)."""
import math

import torch

from oracle import scorer as S
from oracle import utils as U

LOG_EPS = math.log(1e-10)


def test_sort_by_scores():
  scores = torch.tensor([[1., 3., 2.], [1., 2., 3.]])
  positions = torch.tensor([[1, 2, 3], [4, 5, 6]])
  assert U.sort_by_scores(scores, [positions])[0].tolist() == [[2, 3, 1],
                                                               [6, 5, 4]]
  assert U.sort_by_scores(scores, [positions], topn=2)[0].tolist() == [[2, 3],
                                                                       [6, 5]]
  feat = torch.tensor([[[1., 2., 3.], [4., 5., 6.], [7., 8., 9.]],
                       [[10., 20., 30.], [40., 50., 60.], [70., 80., 90.]]])
  out = U.sort_by_scores(scores, [feat], topn=2)[0]
  assert out.tolist() == [[[4., 5., 6.], [7., 8., 9.]],
                          [[70., 80., 90.], [40., 50., 60.]]]


def test_sort_by_scores_ties_and_mask():
  names = torch.tensor([[0, 1, 2]])
  assert U.sort_by_scores(torch.tensor([[2., 1., 1.]]),
                          [names])[0].tolist() == [[0, 1, 2]]
  scores = torch.tensor([[0., math.inf, 2., -math.inf, 1.]])
  names = torch.tensor([[0, 1, 2, 3, 4]])   # a b c d e
  m1 = torch.tensor([[True, False, True, True, False]])
  m2 = torch.tensor([[False, True, False, True, True]])
  assert U.sort_by_scores(scores, [names], mask=m1)[0].tolist() == [[2, 0, 3, 1, 4]]
  assert U.sort_by_scores(scores, [names], mask=m2)[0].tolist() == [[1, 4, 3, 0, 2]]
  assert U.sort_by_scores(scores, [names])[0].tolist() == [[1, 2, 4, 0, 3]]


def test_sorted_ranks():
  assert U.sorted_ranks(torch.tensor([[1., 3., 2.]])).tolist() == [[3, 1, 2]]
  assert U.sorted_ranks(torch.tensor([[1., 2., 1.]])).tolist() == [[2, 1, 3]]


def test_circular_and_padded_indices():
  idx, mask = U._circular_indices(3, [3])
  assert idx.tolist() == [[0, 1, 2]] and mask.tolist() == [[True, True, True]]
  idx, mask = U._circular_indices(3, [2])
  assert idx.tolist() == [[0, 1, 0]] and mask.tolist() == [[True, True, False]]
  idx, mask = U._circular_indices(3, [0])
  assert idx.tolist() == [[0, 0, 0]] and mask.tolist() == [[False] * 3]
  idx, mask = U._circular_indices(3, [3, 2])
  assert idx.tolist() == [[0, 1, 2], [0, 1, 0]]
  idx, mask = U.padded_nd_indices(
      torch.tensor([[True, True, True], [True, True, False]]))
  assert idx.tolist() == [[0, 1, 2], [0, 1, 0]]
  assert mask.tolist() == [[True, True, True], [True, True, False]]


def test_flatten_list():
  ctx = {'context_feature_1': torch.tensor([[1.], [0.]])}
  ex = {'example_feature_1': torch.tensor([[[1.], [0.], [-1.]],
                                           [[0.], [1.], [0.]]])}
  mask = torch.tensor([[True, True, False], [True, False, False]])
  fc, fe = S.flatten_list(ctx, ex, mask)
  assert fc['context_feature_1'].tolist() == [[1.], [1.], [1.], [0.], [0.], [0.]]
  assert fe['example_feature_1'].tolist() == [[1.], [0.], [1.], [0.], [0.], [0.]]
  fc, fe = S.flatten_list(ctx, ex, mask, circular_padding=False)
  assert fe['example_feature_1'].tolist() == [[1.], [0.], [-1.], [0.], [1.], [0.]]


def test_restore_list():
  flat = torch.tensor([1, 0.5, 2, 0, -1, 0])
  mask = torch.tensor([[True, True, False], [True, False, False]])
  out = S.restore_list(flat, mask)
  torch.testing.assert_close(
      out, torch.tensor([[1, 0.5, LOG_EPS], [0, LOG_EPS, LOG_EPS]]))
  out = S.restore_list(flat.reshape(-1, 1), mask, by_scatter=True)
  torch.testing.assert_close(
      out, torch.tensor([[1.5, 0.5, LOG_EPS], [-1. / 3., LOG_EPS, LOG_EPS]]))


def test_rolling_window_indices():
  idx, mask = S.rolling_window_indices(3, 2, [3])
  assert idx.tolist() == [[[0, 1], [1, 2], [2, 0]]]
  assert mask.tolist() == [[True, True, True]]
  idx, mask = S.rolling_window_indices(3, 2, [2])
  assert idx.tolist() == [[[0, 1], [1, 0], [0, 1]]]
  assert mask.tolist() == [[True, True, False]]
  idx, mask = S.rolling_window_indices(3, 2, [0])
  assert idx.tolist() == [[[0, 0], [0, 0], [0, 0]]]
  assert mask.tolist() == [[False, False, False]]
  idx, mask = S.rolling_window_indices(2, 3, [2])
  assert idx.tolist() == [[[0, 1, 0], [1, 0, 1]]]


def test_form_group_indices():
  idx, mask = S.form_group_indices(
      torch.tensor([[True, True, True], [True, True, False]]), 2)
  assert idx.tolist() == [[[0, 1], [1, 2], [2, 0]], [[0, 1], [1, 0], [0, 1]]]
  assert mask.tolist() == [[True, True, True], [True, True, False]]


def test_groupwise_compute_logits():
  """model_test.py:223-277: dummy score fn = context + feature + #rows."""
  def score_fn(group_features):   # [B*G, 2, 1]
    logits = (1. + group_features).reshape(-1, 2)
    return logits + float(logits.shape[0])
  x = torch.tensor([[[1.], [2.], [3.]]])
  valid = torch.tensor([[True, True, False]])
  out = S.groupwise_logits(x, valid, 2, score_fn)
  assert out.tolist() == [[5., 6., 0.]]
  out = S.groupwise_logits(x, valid, 2, score_fn, num_shuffles=2)
  assert out.tolist() == [[8., 9., 0.]]
  x = torch.tensor([[[1.], [2.], [0.]]])
  out = S.groupwise_logits(x, torch.tensor([[True, True, True]]), 2, score_fn,
                           num_shuffles=2)
  assert out.tolist() == [[8., 9., 7.]]


def test_tower_shapes_and_order():
  """keras/layers_test.py:25-31 + layer order of keras/layers.py:65-77."""
  p = S.init_tower_params(1, [3, 2, 1], 1, use_batch_norm=True)
  x = torch.tensor([[1.], [0.], [-1.], [0.], [1.], [0.]])
  out = S.tower_forward(x, p, activation=None, use_batch_norm=True)
  assert list(out.shape) == [6, 1]
  # Without BN and with identity activation the tower is a product of affines.
  p = S.init_tower_params(4, [3, 2], 1)
  for b in p['dense_b']:
    b.uniform_(-1, 1)
  x = torch.randn(5, 4)
  ref = x
  for w, b in zip(p['dense_w'], p['dense_b']):
    ref = ref @ w + b
  torch.testing.assert_close(S.tower_forward(x, p), ref)
