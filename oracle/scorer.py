"""Oracle restatement of the scorer side of the hot path.

  * create_tower            keras/layers.py:26-77
  * FlattenList/RestoreList keras/layers.py:126-175, 231-265
  * DNNScorer / UnivariateScorer.__call__   keras/model.py:755-817
  * _GroupwiseRankingModel (no-shuffle)     model.py:164-421

Test infrastructure only (see oracle/__init__.py).  Weights are explicit
tensors (Keras layout: Dense kernel is [in, out]) so that the CUDA path and the
oracle run on identical parameters; autograd provides the reference gradients.
"""
import math

import torch

from oracle import utils

BN_EPSILON = 1e-3   # tf.keras.layers.BatchNormalization default epsilon


def glorot_uniform(fan_in, fan_out, gen, dtype=torch.float32):
  """tf.keras Dense default kernel initializer (values come from `gen`)."""
  limit = math.sqrt(6.0 / (fan_in + fan_out))
  return (torch.rand(fan_in, fan_out, generator=gen, dtype=torch.float64) * 2 *
          limit - limit).to(dtype)


def init_tower_params(input_dim, hidden_layer_dims, output_units, seed=1238,
                      input_batch_norm=False, use_batch_norm=False,
                      dtype=torch.float32):
  """Parameter dict in create_tower order.  Dense bias zeros (Keras default),
  BN gamma ones / beta zeros."""
  gen = torch.Generator().manual_seed(seed)
  p = {'dense_w': [], 'dense_b': [], 'bn_gamma': [], 'bn_beta': [],
       'in_bn_gamma': None, 'in_bn_beta': None}
  if input_batch_norm:
    p['in_bn_gamma'] = torch.ones(input_dim, dtype=dtype)
    p['in_bn_beta'] = torch.zeros(input_dim, dtype=dtype)
  d = input_dim
  for h in list(hidden_layer_dims) + [output_units]:
    p['dense_w'].append(glorot_uniform(d, h, gen, dtype))
    p['dense_b'].append(torch.zeros(h, dtype=dtype))
    d = h
  if use_batch_norm:
    for h in hidden_layer_dims:
      p['bn_gamma'].append(torch.ones(h, dtype=dtype))
      p['bn_beta'].append(torch.zeros(h, dtype=dtype))
  return p


def _batch_norm_train(x, gamma, beta, moving=None, momentum=0.999):
  """tf.keras BatchNormalization, non-fused rank-2 path, training=True: population
  variance; `moving` = [moving_mean, moving_variance] is updated in place as
  moving * momentum + batch * (1 - momentum)."""
  mean = x.mean(0, keepdim=True)
  var = x.var(0, unbiased=False, keepdim=True)
  if moving is not None:
    with torch.no_grad():
      moving[0].mul_(momentum).add_(mean.reshape(-1) * (1 - momentum))
      moving[1].mul_(momentum).add_(var.reshape(-1) * (1 - momentum))
  return (x - mean) * torch.rsqrt(var + BN_EPSILON) * gamma + beta


def _batch_norm_infer(x, gamma, beta, moving):
  """training=False: the moving statistics normalise."""
  return (x - moving[0]) * torch.rsqrt(moving[1] + BN_EPSILON) * gamma + beta


def init_bn_moving(input_dim, hidden_layer_dims, input_batch_norm=False,
                   use_batch_norm=False, dtype=torch.float32):
  """Keras initial moving statistics: mean zeros, variance ones.
  {'input': [mean, var], 0: [...], 1: [...]}"""
  out = {}
  if input_batch_norm:
    out['input'] = [torch.zeros(input_dim, dtype=dtype),
                    torch.ones(input_dim, dtype=dtype)]
  if use_batch_norm:
    for i, h in enumerate(hidden_layer_dims):
      out[i] = [torch.zeros(h, dtype=dtype), torch.ones(h, dtype=dtype)]
  return out


def _act(x, activation):
  if activation is None:
    return x
  if activation == 'relu':
    return torch.relu(x)
  if activation == 'tanh':
    return torch.tanh(x)
  if activation == 'sigmoid':
    return torch.sigmoid(x)
  raise ValueError(activation)


def tower_forward(x, params, activation=None, use_batch_norm=False,
                  input_batch_norm=False, training=True, bn_moving=None,
                  momentum=0.999, keep_masks=None):
  """create_tower forward (keras/layers.py:65-77):
  [BN] -> (Dense -> [BN] -> act -> [Dropout]) x L -> Dense.
  Dropout is random in the reference; `keep_masks` (one {0,1} tensor per hidden
  layer, already divided by keep probability or not — see below) lets a test
  replay a given mask: h = h * keep_masks[i] (the caller includes the 1/(1-p)
  scale)."""
  def bn(h, g, b, key):
    mv = None if bn_moving is None else bn_moving[key]
    if training:
      return _batch_norm_train(h, g, b, mv, momentum)
    return _batch_norm_infer(h, g, b, mv)

  h = x
  if input_batch_norm:
    h = bn(h, params['in_bn_gamma'], params['in_bn_beta'], 'input')
  n_hidden = len(params['dense_w']) - 1
  for i in range(n_hidden):
    h = h @ params['dense_w'][i] + params['dense_b'][i]
    if use_batch_norm:
      h = bn(h, params['bn_gamma'][i], params['bn_beta'][i], i)
    h = _act(h, activation)
    if keep_masks is not None:
      h = h * keep_masks[i]
  return h @ params['dense_w'][-1] + params['dense_b'][-1]


def flatten_list(context_features, example_features, mask,
                 circular_padding=True):
  """keras/layers.py:126-175."""
  if not example_features:
    raise ValueError('Need a valid example feature.')
  mask = torch.as_tensor(mask)
  batch_size, list_size = mask.shape
  flat_ctx = {}
  for name, t in context_features.items():
    t = torch.as_tensor(t)
    flat_ctx[name] = t.unsqueeze(1).expand(-1, list_size, *t.shape[1:]).reshape(
        batch_size * list_size, *t.shape[1:])
  idx = None
  if circular_padding:
    idx, _ = utils.padded_nd_indices(mask)
  flat_ex = {}
  for name, t in example_features.items():
    t = torch.as_tensor(t)
    if idx is not None:
      t = torch.gather(t, 1, idx.reshape(batch_size, list_size, *(
          [1] * (t.dim() - 2))).expand(-1, -1, *t.shape[2:]))
    flat_ex[name] = t.reshape(batch_size * list_size, *t.shape[2:])
  return flat_ctx, flat_ex


def restore_list(flattened_logits, mask, by_scatter=False):
  """keras/layers.py:231-265."""
  mask = torch.as_tensor(mask)
  logits = torch.as_tensor(flattened_logits).reshape(mask.shape)
  if by_scatter:
    idx, _ = utils.padded_nd_indices(mask)
    counts = torch.zeros_like(logits).scatter_add_(1, idx,
                                                   torch.ones_like(logits))
    summed = torch.zeros_like(logits).scatter_add_(1, idx, logits)
    return torch.where(counts > 0., summed / torch.where(
        counts > 0, counts, torch.ones_like(counts)),
                       torch.full_like(logits, utils.LOG_EPSILON))
  return torch.where(mask, logits, torch.full_like(logits, utils.LOG_EPSILON))


def dnn_scorer(context_features, example_features, mask, params, **tower_kw):
  """UnivariateScorer.__call__ + DNNScorer._score_flattened
  (keras/model.py:755-817): features concatenated context-first, each group in
  sorted key order."""
  flat_ctx, flat_ex = flatten_list(context_features, example_features, mask)
  cols = [flat_ctx[k].reshape(flat_ctx[k].shape[0], -1) for k in sorted(flat_ctx)]
  cols += [flat_ex[k].reshape(flat_ex[k].shape[0], -1) for k in sorted(flat_ex)]
  x = torch.cat(cols, 1)
  return restore_list(tower_forward(x, params, **tower_kw), mask)


# ----------------------------------------------------------------------------
# Groupwise scoring (model.py:164-421), deterministic (shuffle=False) variant.
# ----------------------------------------------------------------------------
def rolling_window_indices(size, rw_size, num_valid_entries):
  """model.py:164-202."""
  nv = torch.as_tensor(num_valid_entries).reshape(-1)
  rw = torch.arange(rw_size).unsqueeze(0) + torch.arange(size).unsqueeze(1)
  batch_rw = rw.unsqueeze(0).expand(nv.shape[0], -1, -1)
  mask = batch_rw.min(dim=2).values < nv.reshape(-1, 1)
  nv1 = torch.where(nv < 1, torch.ones_like(nv), nv)
  return torch.remainder(batch_rw, nv1.reshape(-1, 1, 1)), mask


def form_group_indices(is_valid, group_size):
  """model.py:205-244 with shuffle=False: ([B, G, g] column indices, [B, G])."""
  is_valid = torch.as_tensor(is_valid)
  b, n = is_valid.shape
  rw, mask = rolling_window_indices(n, group_size, is_valid.sum(1))
  organized = utils.organize_valid_indices(is_valid)            # [B, N]
  idx = torch.gather(organized, 1, rw.reshape(b, -1)).reshape(b, n, group_size)
  return idx, mask


def groupwise_logits(x, is_valid, group_size, group_score_fn, num_shuffles=1):
  """model.py:341-421: gather groups, score, scatter-average.

  x: [B, N, D]; group_score_fn: [B*G, group_size, D] -> [B*G, group_size].
  With shuffle=False every "shuffle" repeats the same groups (as PREDICT mode
  with num_shuffles=None, model_test.py:171-185).
  """
  x = torch.as_tensor(x)
  b, n, d = x.shape
  idx, gmask = form_group_indices(is_valid, group_size)
  idx = torch.cat([idx] * num_shuffles, 1)
  gmask = torch.cat([gmask] * num_shuffles, 1)
  g = idx.shape[1]
  gathered = torch.gather(
      x, 1, idx.reshape(b, g * group_size, 1).expand(-1, -1, d)).reshape(
          b * g, group_size, d)
  scores = group_score_fn(gathered).reshape(b, g, group_size)
  smask = gmask.unsqueeze(2).expand(-1, -1, group_size)
  scores = torch.where(smask, scores, torch.zeros_like(scores))
  flat_idx = idx.reshape(b, g * group_size)
  counts = torch.zeros(b, n, dtype=scores.dtype).scatter_add_(
      1, flat_idx, smask.to(scores.dtype).reshape(b, -1))
  logits = torch.zeros(b, n, dtype=scores.dtype).scatter_add_(
      1, flat_idx, scores.reshape(b, -1))
  return torch.where(counts > 0, logits / torch.where(
      counts > 0, counts, torch.ones_like(counts)), torch.zeros_like(logits))
