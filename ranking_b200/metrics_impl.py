"""Host-side mirror of tensorflow_ranking/python/metrics_impl.py (NDCG, MRR).

`compute(labels, predictions, weights=None, mask=None)` returns
`(per_list_metric [B, 1], per_list_weight [B, 1])` like the reference
(metrics_impl.py:268-291); the sort, DCG, ideal DCG, MRR and the per-list weight
rule all run in one CUDA launch (+ a one-block finalise), for any number of
cut-offs at once (`compute_topns`).
"""
import ctypes

import torch

from ranking_b200 import _C
from ranking_b200.keras import utils as keras_utils
from ranking_b200.losses_impl import (_as_f32, _prep_2d, _prep_mask,
                                      _prep_weights, _GAIN_ENUM, _DISC_ENUM)

_DEFAULT_GAIN_FN = keras_utils.pow_minus_1            # metrics_impl.py:31
_DEFAULT_RANK_DISCOUNT_FN = keras_utils.log2_inverse  # metrics_impl.py:33


_EXT_KEYS = ('dcg', 'precision', 'recall', 'map', 'hits', 'arp', 'opa', 'bpref', 'bpref_alt')


def rank_metrics(labels, predictions, weights=None, mask=None, topns=(None,),
                 gain_fn=None, rank_discount_fn=None, want_ndcg=True,
                 want_mrr=True, ext=(), ragged=False):
  """One launch of K4.  Returns dict(ndcg [B, T], ndcg_w [B], mrr [B, T],
  mrr_w [B], raw [B, 5]) plus, for every name in `ext` (subset of dcg, precision,
  recall, map, hits [B, T] and arp, opa [B, 2] = value, weight), that tensor."""
  if ragged:   # lists of per-list sequences, padded like utils.ragged_to_dense
    from ranking_b200 import utils as tfr_utils
    labels, predictions, weights, mask = tfr_utils.ragged_to_dense(
        labels, predictions, weights)
  labels, predictions = _prep_2d(labels, predictions)
  w, wpi = _prep_weights(weights, predictions)
  m = _prep_mask(mask, predictions)
  b, n = predictions.shape
  dev = predictions.device
  gain_fn = gain_fn or _DEFAULT_GAIN_FN
  rank_discount_fn = rank_discount_fn or _DEFAULT_RANK_DISCOUNT_FN
  keep = []
  gain_enum, gain_table = _C.GAIN_TABLE, None
  if gain_fn in _GAIN_ENUM:
    gain_enum = _GAIN_ENUM[gain_fn]
  else:
    # gain of the cleaned labels (metrics_impl.py:256-262: labels := 0 where the
    # item is masked out or has weight <= 0)
    ok = (m != 0) if m is not None else (labels >= 0)
    if w is not None:
      ok = ok & ((w if wpi else w.reshape(-1, 1)) > 0)
    cleaned = torch.where(ok, labels, torch.zeros_like(labels))
    gain_table = torch.as_tensor(gain_fn(cleaned)).to(torch.float32).contiguous()
    keep.append(gain_table)
  disc_enum, disc_table = _C.DISC_TABLE, None
  if rank_discount_fn in _DISC_ENUM:
    disc_enum = _DISC_ENUM[rank_discount_fn]
  else:
    r = torch.arange(0, n + 2, dtype=torch.float32, device=dev)
    r[0] = 1.0
    disc_table = torch.as_tensor(rank_discount_fn(r)).to(
        torch.float32).contiguous()
    keep.append(disc_table)
  t = len(topns)
  topn_arr = (ctypes.c_int32 * t)(*[int(x) if x else 0 for x in topns])
  out = {
      'ndcg': torch.empty(b, t, dtype=torch.float32, device=dev)
              if want_ndcg else None,
      'mrr': torch.empty(b, t, dtype=torch.float32, device=dev)
             if want_mrr else None,
      'ndcg_w': torch.empty(b, dtype=torch.float32, device=dev)
                if want_ndcg else None,
      'mrr_w': torch.empty(b, dtype=torch.float32, device=dev)
               if want_mrr else None,
      'raw': torch.empty(b, 5, dtype=torch.float32, device=dev),
  }
  ext_struct = _C.MetricExt()
  for key in ext:
    if key not in _EXT_KEYS:
      raise ValueError('unknown extended metric %r' % (key,))
    out[key] = torch.empty(b, 2 if key in ('arp', 'opa') else t,
                           dtype=torch.float32, device=dev)
    setattr(ext_struct, key, out[key].data_ptr())
  _C.check(_C.lib.tfr_rank_metrics_ext(
      _C.ptr(predictions), _C.ptr(labels), _C.ptr(w), wpi, _C.ptr(m), b, n,
      topn_arr, t, gain_enum, disc_enum, _C.ptr(gain_table),
      _C.ptr(disc_table), _C.ptr(out['ndcg']), _C.ptr(out['ndcg_w']),
      _C.ptr(out['mrr']), _C.ptr(out['mrr_w']), _C.ptr(out['raw']),
      ctypes.byref(ext_struct) if ext else None, _C.stream()))
  del keep
  return out


class _RankingMetric(object):
  """metrics_impl.py:199-291."""

  def __init__(self, ragged=False):
    self._ragged = ragged

  def compute(self, labels, predictions, weights=None, mask=None):
    raise NotImplementedError


class MRRMetric(_RankingMetric):
  """metrics_impl.py:429-459."""

  def __init__(self, name=None, topn=None, ragged=False):
    super().__init__(ragged)
    self._name = name
    self._topn = topn

  @property
  def name(self):
    return self._name

  def compute(self, labels, predictions, weights=None, mask=None):
    o = rank_metrics(labels, predictions, weights, mask, (self._topn,),
                     want_ndcg=False, ragged=self._ragged)
    return o['mrr'], o['mrr_w'].unsqueeze(1)


class NDCGMetric(_RankingMetric):
  """metrics_impl.py:631-670."""

  def __init__(self, name=None, topn=None, gain_fn=_DEFAULT_GAIN_FN,
               rank_discount_fn=_DEFAULT_RANK_DISCOUNT_FN, ragged=False):
    super().__init__(ragged)
    self._name = name
    self._topn = topn
    self._gain_fn = gain_fn
    self._rank_discount_fn = rank_discount_fn

  @property
  def name(self):
    return self._name

  def compute(self, labels, predictions, weights=None, mask=None):
    o = rank_metrics(labels, predictions, weights, mask, (self._topn,),
                     self._gain_fn, self._rank_discount_fn, want_mrr=False, ragged=self._ragged)
    return o['ndcg'], o['ndcg_w'].unsqueeze(1)


def _safe_div(num, den):
  return torch.where(den != 0, num / torch.where(den != 0, den, torch.ones_like(den)),
                     torch.zeros_like(num))


class _ExtMetric(_RankingMetric):
  """Metrics served by the extended outputs of K4 (one sort for all of them)."""
  _key = None
  _weight = 'mrr_w'     # per-list weight rule with relevance = [label >= 1]

  def __init__(self, name=None, topn=None, ragged=False):
    super().__init__(ragged)
    self._name = name
    self._topn = topn

  @property
  def name(self):
    return self._name

  def compute(self, labels, predictions, weights=None, mask=None):
    o = rank_metrics(labels, predictions, weights, mask, (self._topn,),
                     want_ndcg=False, want_mrr=True, ext=(self._key,), ragged=self._ragged)
    return o[self._key], o[self._weight].unsqueeze(1)


class HitsMetric(_ExtMetric):
  """metrics_impl.py:462-506."""
  _key = 'hits'


class RecallMetric(_ExtMetric):
  """metrics_impl.py:539-561."""
  _key = 'recall'


class PrecisionMetric(_ExtMetric):
  """metrics_impl.py:564-586."""
  _key = 'precision'


class MeanAveragePrecisionMetric(_ExtMetric):
  """metrics_impl.py:589-628."""
  _key = 'map'


class BPrefMetric(_ExtMetric):
  """metrics_impl.py:825-898: binary preference; `use_trec_version=False` divides the
  count of irrelevant items ranked above a relevant one by R instead of min(R, N)."""

  def __init__(self, name=None, topn=None, use_trec_version=True, ragged=False):
    super().__init__(name, topn, ragged)
    self._use_trec_version = use_trec_version
    self._key = 'bpref' if use_trec_version else 'bpref_alt'


class ARPMetric(_RankingMetric):
  """metrics_impl.py:509-536."""

  def __init__(self, name=None, ragged=False):
    super().__init__(ragged)
    self._name = name

  @property
  def name(self):
    return self._name

  def compute(self, labels, predictions, weights=None, mask=None):
    o = rank_metrics(labels, predictions, weights, mask, (None,), want_ndcg=False,
                     want_mrr=False, ext=('arp',), ragged=self._ragged)
    return o['arp'][:, 0:1], o['arp'][:, 1:2]


class OPAMetric(_RankingMetric):
  """metrics_impl.py:708-743."""

  def __init__(self, name=None, ragged=False):
    super().__init__(ragged)
    self._name = name

  @property
  def name(self):
    return self._name

  def compute(self, labels, predictions, weights=None, mask=None):
    o = rank_metrics(labels, predictions, weights, mask, (None,), want_ndcg=False,
                     want_mrr=False, ext=('opa',), ragged=self._ragged)
    return o['opa'][:, 0:1], o['opa'][:, 1:2]


class DCGMetric(_RankingMetric):
  """metrics_impl.py:673-705."""

  def __init__(self, name=None, topn=None, gain_fn=_DEFAULT_GAIN_FN,
               rank_discount_fn=_DEFAULT_RANK_DISCOUNT_FN, ragged=False):
    super().__init__(ragged)
    self._name = name
    self._topn = topn
    self._gain_fn = gain_fn
    self._rank_discount_fn = rank_discount_fn

  @property
  def name(self):
    return self._name

  def compute(self, labels, predictions, weights=None, mask=None):
    o = rank_metrics(labels, predictions, weights, mask, (self._topn,),
                     self._gain_fn, self._rank_discount_fn, want_mrr=False,
                     ext=('dcg',), ragged=self._ragged)
    w = o['ndcg_w'].unsqueeze(1)
    return _safe_div(o['dcg'], w), w


# ----------------------------------------------------------------------------
# Diversity metrics (metrics_impl.py:313-427, 746-823)
# ----------------------------------------------------------------------------
def div_metrics(labels, predictions, weights=None, mask=None, topns=(None,), alpha=0.5,
                rank_discount_fn=None, ragged=False):
  """One launch of the diversity kernel: dict(precision_ia [B, T], alpha_dcg [B, T]
  (unnormalised), list_w [B], raw [B, 5]).  labels [B, N, S]."""
  if ragged:
    from ranking_b200 import utils as tfr_utils
    labels, predictions, weights, mask = tfr_utils.ragged_to_dense(
        labels, predictions, weights)
  predictions = _as_f32(predictions, what='predictions')
  labels = _as_f32(labels, predictions.device, 'labels')
  if labels.dim() != 3 or predictions.dim() != 2 or labels.shape[:2] != predictions.shape:
    raise ValueError('labels must be [batch_size, list_size, subtopic_size] and '
                     'predictions [batch_size, list_size]')
  b, n, s_ = labels.shape
  dev = predictions.device
  w, wpi = _prep_weights(weights, predictions)
  m = None
  if mask is not None:
    mk = torch.as_tensor(mask, device=dev)
    if mk.dim() == 3:
      mk = mk.any(dim=2)
    m = mk.to(torch.uint8).contiguous()
  rank_discount_fn = rank_discount_fn or _DEFAULT_RANK_DISCOUNT_FN
  disc_enum, disc_table = _C.DISC_TABLE, None
  if rank_discount_fn in _DISC_ENUM:
    disc_enum = _DISC_ENUM[rank_discount_fn]
  else:
    r = torch.arange(0, n + 2, dtype=torch.float32, device=dev)
    r[0] = 1.0
    disc_table = torch.as_tensor(rank_discount_fn(r)).to(torch.float32).contiguous()
  t = len(topns)
  topn_arr = (ctypes.c_int32 * t)(*[int(x) if x else 0 for x in topns])
  out = {'precision_ia': torch.empty(b, t, dtype=torch.float32, device=dev),
         'alpha_dcg': torch.empty(b, t, dtype=torch.float32, device=dev),
         'list_w': torch.empty(b, dtype=torch.float32, device=dev),
         'raw': torch.empty(b, 5, dtype=torch.float32, device=dev)}
  _C.check(_C.lib.tfr_div_metrics(
      _C.ptr(predictions), _C.ptr(labels), _C.ptr(w), wpi, _C.ptr(m), b, n, s_, topn_arr, t,
      float(alpha), disc_enum, _C.ptr(disc_table), _C.ptr(out['precision_ia']),
      _C.ptr(out['alpha_dcg']), _C.ptr(out['list_w']), _C.ptr(out['raw']), _C.stream()))
  return out


class _DivRankingMetric(_RankingMetric):
  """metrics_impl.py:313-427."""

  def __init__(self, name=None, topn=None, ragged=False):
    super().__init__(ragged)
    self._name = name
    self._topn = topn

  @property
  def name(self):
    return self._name


class PrecisionIAMetric(_DivRankingMetric):
  """metrics_impl.py:746-782."""

  def compute(self, labels, predictions, weights=None, mask=None):
    o = div_metrics(labels, predictions, weights, mask, (self._topn,), ragged=self._ragged)
    return o['precision_ia'], o['list_w'].unsqueeze(1)


class AlphaDCGMetric(_DivRankingMetric):
  """metrics_impl.py:785-822."""

  def __init__(self, name=None, topn=None, alpha=0.5,
               rank_discount_fn=_DEFAULT_RANK_DISCOUNT_FN, seed=None, ragged=False):
    super().__init__(name, topn, ragged)
    self._alpha = alpha
    self._rank_discount_fn = rank_discount_fn
    self._seed = seed

  def compute(self, labels, predictions, weights=None, mask=None):
    o = div_metrics(labels, predictions, weights, mask, (self._topn,), self._alpha,
                    self._rank_discount_fn, ragged=self._ragged)
    w = o['list_w'].unsqueeze(1)
    return _safe_div(o['alpha_dcg'], w), w
