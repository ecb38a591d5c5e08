"""Fused training step of the hot path (the reference's `model.fit` inner step,
keras/pipeline.py:605-632 and SURVEY.md §3.1):

    scores = scorer(x)                       tfr_mlp_fwd   (+ RestoreList fill)
    loss, d loss/d scores                    one fused loss kernel + 1-CTA reduce
    d loss/d params                          tfr_mlp_bwd   (flat gradient buffer)
    all-reduce(SUM) of the flat gradient     NCCL over NVLink (world_size > 1)
    params -= optimizer(grads / world_size)  tfr_optimizer_step

Data-parallel semantics follow the reference (extension/task.py:256-262):
each rank reduces its loss over its LOCAL batch, the loss is scaled by
1/num_replicas and gradients are summed — done here by one all-reduce of the
flat fp32 gradient and a 1/world_size factor inside the optimizer kernel.
"""
import ctypes

import torch

from ranking_b200 import _C
from ranking_b200 import dp

_OPT = {'sgd': 0, 'adagrad': 1}


class RankingTrainer(object):
  """Owns the static buffers of one training configuration [B, N, D]."""

  def __init__(self, tower, loss, optimizer='adagrad', learning_rate=0.001,
               epsilon=1e-7, initial_accumulator_value=0.1, process_group=None):
    if optimizer not in _OPT:
      raise ValueError('optimizer must be one of %s' % sorted(_OPT))
    if not hasattr(loss, 'fused_fwd_bwd'):
      raise ValueError('loss must be a ranking_b200.keras.losses object')
    self.tower = tower
    self.loss = loss
    self.opt_kind = _OPT[optimizer]
    self.lr = float(learning_rate)
    self.eps = float(epsilon)
    dev = tower.flat.device
    self.device = dev
    self.accum = torch.full_like(tower.flat.data, initial_accumulator_value)
    self.grads = torch.zeros_like(tower.flat.data)
    self.group = process_group
    self.world = dp.world_size(process_group)
    self._shape = None
    self.launches_per_step = None

  # -- buffers ---------------------------------------------------------------
  def _ensure(self, b, n):
    if self._shape == (b, n):
      return
    dev = self.device
    m = b * n
    self.scores = torch.empty(b, n, dtype=torch.float32, device=dev)
    self.dscores = torch.empty(b, n, dtype=torch.float32, device=dev)
    self.per_list = torch.empty(2, b, dtype=torch.float32, device=dev)
    self.total2 = torch.zeros(2, dtype=torch.float32, device=dev)
    self.ws = self.tower._new_workspace(m)
    self._shape = (b, n)

  # -- one step on device-resident inputs -------------------------------------
  def train_step(self, x, y_true, sample_weight=None, mask=None):
    """x [B, N, D] fp32 (device), y_true [B, N] (label < 0 = padding).
    Returns the scalar loss as a 0-d device tensor (no host sync)."""
    b, n, d = x.shape
    self._ensure(b, n)
    t = self.tower
    m = b * n
    st = _C.stream()
    run_cfg = t._run_cfg(training=True)
    cfg = ctypes.byref(run_cfg)
    m8 = None
    if mask is not None:
      m8 = mask.reshape(-1).to(torch.uint8).contiguous()
    _C.check(_C.lib.tfr_mlp_fwd(_C.ptr(x), m, cfg, _C.ptr(t.flat.data),
                                _C.ptr(m8), _C.ptr(self.ws), _C.ptr(self.scores),
                                t._precision, st))
    self.loss.fused_fwd_bwd(y_true, self.scores, sample_weight, self.dscores,
                            self.per_list, self.total2)
    _C.check(_C.lib.tfr_mlp_bwd(_C.ptr(x), m, cfg, _C.ptr(t.flat.data),
                                _C.ptr(self.dscores), _C.ptr(m8), _C.ptr(self.ws),
                                _C.ptr(self.grads), t._precision, st))
    dp.all_reduce_sum_(self.grads, self.group)   # the one collective of the step
    _C.check(_C.lib.tfr_optimizer_step(
        _C.ptr(t.flat.data), _C.ptr(self.grads), _C.ptr(self.accum),
        self.grads.numel(), self.opt_kind, self.lr, self.eps,
        dp.replica_grad_scale(self.group), st))
    return self.total2[0]

  # -- evaluation --------------------------------------------------------------
  @torch.no_grad()
  def predict(self, x, mask=None):
    b, n, d = x.shape
    self._ensure(b, n)
    t = self.tower
    m8 = None if mask is None else mask.reshape(-1).to(torch.uint8).contiguous()
    _C.check(_C.lib.tfr_mlp_fwd(_C.ptr(x), b * n,
                                ctypes.byref(t._run_cfg(training=False)),
                                _C.ptr(t.flat.data), _C.ptr(m8), _C.ptr(self.ws),
                                _C.ptr(self.scores), t._precision, _C.stream()))
    return self.scores


class HostBatchPipeline(object):
  """End-to-end step from HOST buffers: pinned-memory batches are copied to the
  device on a copy stream, double-buffered so the copy of batch k+1 overlaps the
  compute of batch k; the scalar loss of every step is read back to the host."""

  def __init__(self, trainer, b, n, d):
    self.trainer = trainer
    dev = trainer.device
    self.copy_stream = torch.cuda.Stream(device=dev)
    self.x = [torch.empty(b, n, d, dtype=torch.float32, device=dev)
              for _ in range(2)]
    self.y = [torch.empty(b, n, dtype=torch.float32, device=dev) for _ in range(2)]
    self.ready = [torch.cuda.Event() for _ in range(2)]
    self.free = [torch.cuda.Event() for _ in range(2)]
    self.loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    self.h2d_bytes = (b * n * d + b * n) * 4
    self.d2h_bytes = 4
    self._slot = 0
    self._primed = False

  def _upload(self, slot, x_host, y_host):
    with torch.cuda.stream(self.copy_stream):
      self.copy_stream.wait_event(self.free[slot])
      self.x[slot].copy_(x_host, non_blocking=True)
      self.y[slot].copy_(y_host, non_blocking=True)
      self.ready[slot].record(self.copy_stream)

  def run(self, host_batches):
    """host_batches: sequence of (x_pinned [B,N,D], y_pinned [B,N]).  Returns the
    list of per-step losses (python floats)."""
    tr = self.trainer
    cur = torch.cuda.current_stream()
    losses = []
    it = iter(host_batches)
    nxt = next(it, None)
    if nxt is None:
      return losses
    for s in range(2):
      self.free[s].record(cur)
    slot = 0
    self._upload(slot, *nxt)
    while nxt is not None:
      upcoming = next(it, None)
      if upcoming is not None:
        self._upload(slot ^ 1, *upcoming)
      cur.wait_event(self.ready[slot])
      loss = tr.train_step(self.x[slot], self.y[slot])
      self.free[slot].record(cur)
      self.loss_host[slot].copy_(loss, non_blocking=True)
      cur.synchronize()          # the step's result is read on the host
      losses.append(float(self.loss_host[slot]))
      nxt = upcoming
      slot ^= 1
    return losses
