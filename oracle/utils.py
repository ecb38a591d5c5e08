"""Oracle restatement of tensorflow_ranking/python/utils.py (sort / rank / index helpers).

Test infrastructure only (see oracle/__init__.py).  Ties are broken by original
index (the reference's `shuffle_ties=False` behaviour, utils.py:99-112).
"""
import math

import torch

_PADDING_LABEL = -1.0          # utils.py:21
_PADDING_PREDICTION = -1e6     # utils.py:22
_PADDING_WEIGHT = 0.0          # utils.py:23


def is_label_valid(labels):
  """utils.py:78-81."""
  return torch.as_tensor(labels) >= 0.


def _stable_argsort(values, descending=False):
  return torch.sort(values, dim=-1, descending=descending, stable=True).indices


def _get_shuffle_indices(shape, mask=None):
  """utils.py:84-112 with shuffle_ties=False: zeros (+2 where masked out)."""
  shuffle_values = torch.zeros(shape, dtype=torch.float32)
  if mask is not None:
    shuffle_values = torch.where(mask, shuffle_values, shuffle_values + 2.0)
  return _stable_argsort(shuffle_values)


def sort_by_scores(scores, features_list, topn=None, mask=None):
  """utils.py:115-164 (shuffle_ties=False).

  top_k breaks ties by lower index first, i.e. a stable descending sort.
  """
  scores = torch.as_tensor(scores).to(torch.float32) if not torch.is_tensor(
      scores) else scores
  assert scores.dim() == 2
  list_size = scores.shape[1]
  topn = list_size if topn is None else min(topn, list_size)
  shuffle_ind = None
  if mask is not None:
    mask = torch.as_tensor(mask)
    # NB: the reference uses the global min over the whole batch (utils.py:150).
    scores = torch.where(mask, scores, scores.min())
    shuffle_ind = _get_shuffle_indices(scores.shape, mask)
    scores = torch.gather(scores, 1, shuffle_ind)
  indices = _stable_argsort(scores, descending=True)[:, :topn]
  if shuffle_ind is not None:
    indices = torch.gather(shuffle_ind, 1, indices)
  out = []
  for f in features_list:
    f = torch.as_tensor(f)
    if f.dim() == 2:
      out.append(torch.gather(f, 1, indices))
    else:
      idx = indices.unsqueeze(-1).expand(-1, -1, f.shape[2])
      out.append(torch.gather(f, 1, idx))
  return out


def sorted_ranks(scores):
  """utils.py:167-195 (shuffle_ties=False): 1-based ranks, ties by index."""
  scores = torch.as_tensor(scores)
  batch_size, list_size = scores.shape
  positions = torch.arange(list_size).unsqueeze(0).expand(batch_size, -1)
  sorted_positions = sort_by_scores(scores, [positions])[0]
  return _stable_argsort(sorted_positions) + 1


def organize_valid_indices(is_valid):
  """utils.py:203-235 with shuffle=False.  Returns [B, N] column indices
  (the batch index of the reference's nd-indices is implicit)."""
  is_valid = torch.as_tensor(is_valid)
  n = is_valid.shape[1]
  values = torch.arange(n - 1, -1, -1, dtype=torch.float32).expand_as(is_valid)
  rand = torch.where(is_valid, values, torch.full_like(values, -1e-6))
  return _stable_argsort(rand, descending=True)


def _circular_indices(size, num_valid_entries):
  """utils.py:272-305."""
  num_valid_entries = torch.as_tensor(num_valid_entries).reshape(-1, 1)
  batch_indices = torch.arange(size).unsqueeze(0).expand(
      num_valid_entries.shape[0], -1)
  mask = batch_indices < num_valid_entries
  nv = torch.where(num_valid_entries < 1, torch.ones_like(num_valid_entries),
                   num_valid_entries)
  return torch.remainder(batch_indices, nv), mask


def padded_nd_indices(is_valid):
  """utils.py:308-356 with shuffle=False.  Returns ([B, N] column indices, mask)."""
  is_valid = torch.as_tensor(is_valid)
  list_size = is_valid.shape[1]
  num_valid = is_valid.to(torch.int64).sum(1)
  indices, mask = _circular_indices(list_size, num_valid)
  shuffled = organize_valid_indices(is_valid)
  return torch.gather(shuffled, 1, indices), mask


LOG_EPSILON = math.log(1e-10)   # losses_impl.py:30, keras/layers.py:265
