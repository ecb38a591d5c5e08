"""Groupwise scoring (`tfr.model._GroupwiseRankingModel`, model.py:164-421).

Every list forms `list_size` rolling-window groups of `group_size` valid items
(model.py:164-244); the group score function sees the concatenated features of a
group and emits one score per member; an item's logit is the MEAN of the scores
it received (scatter-add + count, model.py:388-412); invalid slots get 0.

The heavy part — the group score function over [B * G, group_size * D] — is the
CUDA tower (`keras.layers.Tower`); index formation and the scatter-average are
device-side torch ops (gather / scatter_add), which are plumbing here.  Training
mode in the reference shuffles the valid items with TF's RNG (model.py:316-322);
that stream is not reproducible, so the permutation is an explicit, optional
input (identity = the reference's PREDICT-mode behaviour, model_test.py:171-185).
"""
import ctypes

import torch

from ranking_b200 import _C
from ranking_b200 import utils as tfr_utils


def _rolling_window_indices(size, rw_size, num_valid_entries):
  """model.py:164-202."""
  nv = num_valid_entries.reshape(-1)
  dev = nv.device
  rw = torch.arange(rw_size, device=dev).unsqueeze(0) + torch.arange(
      size, device=dev).unsqueeze(1)
  batch_rw = rw.unsqueeze(0).expand(nv.shape[0], -1, -1)
  mask = batch_rw.min(dim=2).values < nv.reshape(-1, 1)
  nv1 = torch.clamp(nv, min=1)
  return torch.remainder(batch_rw, nv1.reshape(-1, 1, 1)), mask


def _form_group_indices(is_valid, group_size, permutation=None):
  """model.py:205-244: ([B, G, group_size] column indices, [B, G] group mask)."""
  b, n = is_valid.shape
  rw, mask = _rolling_window_indices(n, group_size, is_valid.sum(1))
  organized = tfr_utils.organize_valid_indices(is_valid)   # valid first, in order
  if permutation is not None:   # caller-supplied shuffle of the valid-first order
    organized = torch.gather(organized, 1, permutation)
  idx = torch.gather(organized, 1, rw.reshape(b, -1)).reshape(b, n, group_size)
  return idx, mask


class _GroupTowerFn(torch.autograd.Function):
  """logits [B, N] = folded groupwise tower (tfr_group_mlp_fwd / bwd, csrc/mlp_group.cu)."""

  @staticmethod
  def forward(ctx, x, flat, idx, gmask, tower):
    b, n, _ = x.shape
    g, gs = idx.shape[1], idx.shape[2]
    cfg = tower._run_cfg()
    nbytes = _C.lib.tfr_group_mlp_workspace_bytes(ctypes.byref(cfg), b, n, g, gs)
    if nbytes == 0:
      raise ValueError(_C.last_error())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    logits = torch.empty(b, n, dtype=torch.float32, device=x.device)
    _C.check(_C.lib.tfr_group_mlp_fwd(_C.ptr(x), b, n, g, gs, _C.ptr(idx), _C.ptr(gmask),
                                      ctypes.byref(cfg), _C.ptr(flat), _C.ptr(ws),
                                      _C.ptr(logits), tower._precision, _C.stream()))
    ctx.tower, ctx.cfg, ctx.ws = tower, cfg, ws
    ctx.save_for_backward(x, flat, idx, gmask)
    return logits

  @staticmethod
  def backward(ctx, g_out):
    x, flat, idx, gmask = ctx.saved_tensors
    b, n, _ = x.shape
    g, gs = idx.shape[1], idx.shape[2]
    grads = torch.empty_like(flat)
    _C.check(_C.lib.tfr_group_mlp_bwd(_C.ptr(x), b, n, g, gs, _C.ptr(idx), _C.ptr(gmask),
                                      ctypes.byref(ctx.cfg), _C.ptr(flat),
                                      _C.ptr(g_out.contiguous()), _C.ptr(ctx.ws),
                                      _C.ptr(grads), ctx.tower._precision, _C.stream()))
    return None, grads, None, None, None


def fold_supported(score_fn, group_size, d):
  """The first-layer fold serves tower group score functions on the
  tensor-core path with >= 1 hidden layer and no BN / Dropout."""
  t = getattr(score_fn, 'tower', None)
  return (t is not None and t.precision in ('tf32x3', 'tf32') and
          len(t.hidden_layer_dims) >= 1 and not t.use_batch_norm and
          not t.input_batch_norm and t.dropout == 0.0 and
          t.input_dim == group_size * d and t.output_units == group_size and
          d % 4 == 0 and t.hidden_layer_dims[0] % 4 == 0 and
          all(h % 4 == 0 for h in t.hidden_layer_dims))


class GroupwiseRankingModel(torch.nn.Module):
  """`group_score_fn`: module/callable [B*G, group_size, D] -> [B*G, group_size]."""

  def __init__(self, group_score_fn, group_size):
    super().__init__()
    if group_size <= 0:
      raise ValueError('Invalid group_size %d' % group_size)
    self._group_size = group_size
    self._score_fn = group_score_fn
    self.fold = True   # False: always materialise the gathered features (reference plan)

  def compute_logits(self, example_features, is_valid, num_shuffles=1,
                     permutations=None):
    """example_features [B, N, D]; is_valid [B, N] bool -> logits [B, N]."""
    x = example_features
    b, n, d = x.shape
    gs = self._group_size
    idx_list, mask_list = [], []
    for s in range(num_shuffles):
      perm = None if permutations is None else permutations[s]
      idx, gmask = _form_group_indices(is_valid, gs, perm)
      idx_list.append(idx)
      mask_list.append(gmask)
    idx = torch.cat(idx_list, 1)
    gmask = torch.cat(mask_list, 1)
    g = idx.shape[1]
    if self.fold and fold_supported(self._score_fn, gs, d):
      # K8: the gathered [B, G, gs, D] tensor is never formed (csrc/mlp_group.cu)
      return _GroupTowerFn.apply(x.float().contiguous(), self._score_fn.tower.flat,
                                 idx.to(torch.int32).contiguous(),
                                 gmask.to(torch.uint8).contiguous(), self._score_fn.tower)
    gathered = torch.gather(x, 1, idx.reshape(b, g * gs, 1).expand(-1, -1, d))
    scores = self._score_fn(gathered.reshape(b * g, gs, d)).reshape(b, g, gs)
    smask = gmask.unsqueeze(2).expand(-1, -1, gs)
    scores = torch.where(smask, scores, torch.zeros_like(scores))
    flat_idx = idx.reshape(b, g * gs)
    counts = torch.zeros(b, n, dtype=scores.dtype, device=x.device).scatter_add_(
        1, flat_idx, smask.to(scores.dtype).reshape(b, -1))
    logits = torch.zeros(b, n, dtype=scores.dtype, device=x.device).scatter_add(
        1, flat_idx, scores.reshape(b, -1))
    return torch.where(counts > 0, logits / counts.clamp(min=1.),
                       torch.zeros_like(logits))

  forward = compute_logits


class TowerGroupScoreFn(torch.nn.Module):
  """Group score function backed by the CUDA tower: the group's member features
  are concatenated ([B*G, group_size * D]) and scored by one tower with
  `group_size` outputs (examples/tf_ranking_libsvm.py:313-349)."""

  def __init__(self, tower):
    super().__init__()
    self.tower = tower

  def forward(self, group_features):
    bg, gs, d = group_features.shape
    return self.tower(group_features.reshape(bg, gs * d).contiguous())
