"""Host-side mirror of the pieces of tensorflow_ranking/python/utils.py that sit
on the hot path (utils.py:78-81, 167-195, 203-235, 272-356)."""
import torch

from ranking_b200 import _C

_PADDING_LABEL = -1.
_PADDING_PREDICTION = -1e6
_PADDING_WEIGHT = 0.


def is_label_valid(labels):
  """utils.py:78-81."""
  return torch.as_tensor(labels) >= 0.


def sorted_ranks(scores, labels=None, mask=None):
  """1-based ranks by descending score, ties by index (utils.py:167-195 with
  shuffle_ties=False).  Entries that are invalid (label < 0, or mask == 0) rank
  after every valid entry, as in losses_impl.py:483-500."""
  scores = torch.as_tensor(scores, dtype=torch.float32, device='cuda') \
      if not torch.is_tensor(scores) else scores.float().contiguous()
  _C.require_cuda(scores, 'scores')
  b, n = scores.shape
  if labels is None:
    labels = torch.zeros_like(scores)
  labels = torch.as_tensor(labels, dtype=torch.float32,
                           device=scores.device).contiguous()
  m = None if mask is None else torch.as_tensor(
      mask, device=scores.device).to(torch.uint8).contiguous()
  ranks = torch.empty(b, n, dtype=torch.int32, device=scores.device)
  _C.check(_C.lib.tfr_sorted_ranks(_C.ptr(scores), _C.ptr(labels), _C.ptr(m), b,
                                   n, _C.ptr(ranks), _C.stream()))
  return ranks


def organize_valid_indices(is_valid):
  """utils.py:203-235 with shuffle=False: column indices, valid entries first in
  their original order."""
  is_valid = torch.as_tensor(is_valid)
  n = is_valid.shape[1]
  values = torch.arange(n - 1, -1, -1, dtype=torch.float32,
                        device=is_valid.device).expand_as(is_valid)
  rand = torch.where(is_valid, values, torch.full_like(values, -1e-6))
  return torch.sort(rand, dim=1, descending=True, stable=True).indices


def padded_nd_indices(is_valid):
  """utils.py:308-356 with shuffle=False: ([B, N] column indices, mask)."""
  is_valid = torch.as_tensor(is_valid)
  n = is_valid.shape[1]
  num_valid = is_valid.to(torch.int64).sum(1, keepdim=True)
  idx = torch.arange(n, device=is_valid.device).unsqueeze(0).expand_as(is_valid)
  mask = idx < num_valid
  nv = torch.clamp(num_valid, min=1)
  circular = torch.remainder(idx, nv)
  return torch.gather(organize_valid_indices(is_valid), 1, circular), mask


_PADDING_LABEL = -1.          # utils.py:21-23
_PADDING_PREDICTION = -1e6
_PADDING_WEIGHT = 0.


def _rows(x):
  """A ragged batch: a list / tuple of 1-D sequences, or a torch nested tensor."""
  if torch.is_tensor(x) and getattr(x, 'is_nested', False):
    return list(x.unbind())
  if torch.is_tensor(x):
    raise ValueError('ragged=True expects a list of per-list sequences (or a nested '
                     'tensor), got a dense tensor')
  return [torch.as_tensor(r) for r in x]


def _pad_rows(rows, pad, width, device):
  trailing = rows[0].shape[1:] if rows and rows[0].dim() > 1 else ()
  out = torch.full((len(rows), width) + tuple(trailing), pad, dtype=torch.float32,
                   device=device)
  for i, r in enumerate(rows):
    if r.numel():
      out[i, :r.shape[0]] = r.to(device=device, dtype=torch.float32)
  return out


def ragged_to_dense(labels, predictions, weights, device='cuda'):
  """utils.py:421-443 for ragged batches given as lists of per-list sequences:
  labels pad -1, predictions pad -1e6, ragged weights pad 0; returns
  (labels, predictions, weights, mask)."""
  lab_rows = _rows(labels)
  width = max([r.shape[0] for r in lab_rows] + [1])
  mask = torch.zeros(len(lab_rows), width, dtype=torch.bool, device=device)
  for i, r in enumerate(lab_rows):
    mask[i, :r.shape[0]] = True
  dense_labels = _pad_rows(lab_rows, _PADDING_LABEL, width, device)
  dense_pred = None
  if predictions is not None:
    dense_pred = _pad_rows(_rows(predictions), _PADDING_PREDICTION, width, device)
  dense_w = weights
  if isinstance(weights, (list, tuple)) or (torch.is_tensor(weights) and getattr(
      weights, 'is_nested', False)):
    w_rows = _rows(weights)
    if len(w_rows) == len(lab_rows) and all(
        r.dim() >= 1 and r.shape[0] == l_.shape[0] for r, l_ in zip(w_rows, lab_rows)):
      dense_w = _pad_rows(w_rows, _PADDING_WEIGHT, width, device)   # per-item, ragged
    else:
      dense_w = torch.as_tensor(weights, dtype=torch.float32, device=device)
  return dense_labels, dense_pred, dense_w, mask
