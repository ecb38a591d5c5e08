"""A small fit / evaluate loop around `RankingTrainer` with checkpoint + resume: the
shape of the reference's `ModelFitPipeline.train_and_validate`
(keras/pipeline.py:561-650) for the fused CUDA step.
"""
import os

import torch
import torch.distributed as dist

from ranking_b200 import dp
from ranking_b200.keras import metrics as keras_metrics


def save_checkpoint(trainer, path, step):
  """Parameters, optimizer accumulator, BN moving statistics and the step counter.
  Data parallel: rank 0 alone writes (replicas hold identical parameters; BN statistics
  are per replica and rank 0's are the ones kept), every rank waits for the file."""
  tower = trainer.tower
  if dp.rank(trainer.group) == 0:
    tmp = '%s.tmp.%d' % (path, os.getpid())
    torch.save({'step': int(step), 'flat': tower.flat.detach().cpu(),
                'accum': trainer.accum.cpu(), 'bn_state': tower.bn_state.cpu(),
                'dims': list(tower.dims)}, tmp)
    os.replace(tmp, path)      # atomic: a crash never leaves a torn file
  if dp.world_size(trainer.group) > 1:
    dist.barrier(group=trainer.group)


def load_checkpoint(trainer, path):
  ckpt = torch.load(path, map_location='cpu', weights_only=True)
  tower = trainer.tower
  if list(ckpt['dims']) != list(tower.dims):
    raise ValueError('checkpoint was written for dims %s, tower has %s' %
                     (ckpt['dims'], tower.dims))
  with torch.no_grad():
    tower.flat.copy_(ckpt['flat'])
    trainer.accum.copy_(ckpt['accum'])
    if tower.bn_state.numel():
      tower.bn_state.copy_(ckpt['bn_state'])
  return int(ckpt['step'])


def evaluate(trainer, batches, metric_group=None):
  """Scores every batch with the tower and accumulates `MetricGroup.default()`
  (all of `default_keras_metrics()`, one launch per batch)."""
  group = metric_group or keras_metrics.MetricGroup.default()
  group.reset_state()
  dev = trainer.device
  for x, y in batches:
    x, y = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
    scores = trainer.predict(x, mask=(y >= 0))
    group.update_state(y, scores)
  group.all_reduce(trainer.group)
  return group.result()


def fit(trainer, train_batches, num_steps, checkpoint_dir=None,
        steps_per_checkpoint=1000, eval_batches_fn=None, log_fn=print):
  """Runs `num_steps` fused training steps over `train_batches` (an iterator of host
  or device (x, y) pairs).  With `checkpoint_dir`, resumes from `ckpt.pt` if it exists
  and rewrites it every `steps_per_checkpoint` steps and at the end."""
  step = 0
  ckpt = None
  if checkpoint_dir:
    os.makedirs(checkpoint_dir, exist_ok=True)
    ckpt = os.path.join(checkpoint_dir, 'ckpt.pt')
    if os.path.exists(ckpt):
      step = load_checkpoint(trainer, ckpt)
      log_fn('resumed from %s at step %d' % (ckpt, step))
  dev = trainer.device
  last = None
  for x, y in train_batches:
    if step >= num_steps:
      break
    x, y = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
    last = trainer.train_step(x, y)
    step += 1
    if ckpt and step % steps_per_checkpoint == 0:
      save_checkpoint(trainer, ckpt, step)
      msg = 'step %d loss %.6f' % (step, float(last))
      if eval_batches_fn is not None:
        msg += ' ' + str(evaluate(trainer, eval_batches_fn()))
      log_fn(msg)
  if ckpt:
    save_checkpoint(trainer, ckpt, step)
  return step, (None if last is None else float(last))


# ----------------------------------------------------------------------------------------
# keras/pipeline.py:262-334, 372-650: PipelineHparams + ModelFitPipeline.train_and_validate
# ----------------------------------------------------------------------------------------
import dataclasses
from typing import Dict, Optional, Union


@dataclasses.dataclass
class PipelineHparams:
  """The reference's `PipelineHparams` (keras/pipeline.py:262-334), same field names.  Fields
  that select TF machinery (`strategy`, `cluster_resolver`, `variable_partitioner`, `tpu`,
  `steps_per_execution`) are accepted and ignored: data parallelism here is one process per
  GPU under torchrun.  `optimizer` must be one the fused step implements
  ('adagrad', 'sgd'); the reference's default 'adam' raises."""
  model_dir: str
  num_epochs: int
  steps_per_epoch: int
  validation_steps: int
  learning_rate: float
  loss: Union[str, Dict[str, str]]
  loss_reduction: str = 'auto'
  optimizer: str = 'adagrad'
  loss_weights: Optional[Union[float, Dict[str, float]]] = None
  steps_per_execution: int = 10
  automatic_reduce_lr: bool = False
  early_stopping_patience: int = 0
  early_stopping_min_delta: float = 0.0
  use_weighted_metrics: bool = False
  export_best_model: bool = False
  best_exporter_metric_higher_better: bool = False
  best_exporter_metric: str = 'loss'
  strategy: Optional[str] = None
  cluster_resolver: Optional[object] = None
  variable_partitioner: Optional[object] = None
  tpu: Optional[str] = ''

  def validate(self):
    if not isinstance(self.loss, str):
      raise ValueError('multi-task loss maps are not supported by the fused step')
    if self.optimizer not in ('adagrad', 'sgd'):
      raise ValueError("optimizer must be 'adagrad' or 'sgd', got %r" % (self.optimizer,))
    if self.num_epochs < 1 or self.steps_per_epoch < 1 or self.validation_steps < 0:
      raise ValueError('num_epochs / steps_per_epoch must be >= 1, validation_steps >= 0')
    if self.loss_weights is not None and not isinstance(self.loss_weights, (int, float)):
      raise ValueError('loss_weights must be None or a scalar')


class ModelFitPipeline(object):
  """`ModelFitPipeline.train_and_validate` (keras/pipeline.py:561-632) over the fused CUDA
  step: `num_epochs` x (`steps_per_epoch` training steps, `validation_steps` evaluation
  batches), checkpoint + resume in `model_dir`, optional early stopping
  (`early_stopping_patience` / `_min_delta` on `best_exporter_metric`, keras/pipeline.py:
  474-496), ReduceLROnPlateau (`automatic_reduce_lr`: factor 0.1, patience = a third of the
  epochs, as build_callbacks does) and the best checkpoint (`export_best_model`).

  tower_fn()            -> ranking_b200.keras.layers.Tower (the model_builder's role)
  train_batches_fn()    -> iterable of (x, y) batches, restarted when exhausted
  valid_batches_fn()    -> the same for validation (may be None)
  """

  def __init__(self, tower_fn, train_batches_fn, valid_batches_fn, hparams,
               process_group=None, log_fn=print):
    hparams.validate()
    self._hparams = hparams
    self._tower_fn = tower_fn
    self._train_fn = train_batches_fn
    self._valid_fn = valid_batches_fn
    self._group = process_group
    self._log = log_fn
    self.history = []

  def build_loss(self):
    from ranking_b200.keras import losses as keras_losses
    kw = {}
    if self._hparams.loss_reduction not in (None, 'auto'):
      kw['reduction'] = self._hparams.loss_reduction
    return keras_losses.get(self._hparams.loss, **kw)

  def build_metrics(self):
    return keras_metrics.MetricGroup.default()

  def _monitor(self, logs):
    key = self._hparams.best_exporter_metric
    if key not in logs:
      raise ValueError('best_exporter_metric %r is not among %s' % (key, sorted(logs)))
    return logs[key]

  def train_and_validate(self, verbose=0):
    from ranking_b200 import train as train_lib
    hp = self._hparams
    tower = self._tower_fn()
    trainer = train_lib.RankingTrainer(tower, self.build_loss(), optimizer=hp.optimizer,
                                       learning_rate=hp.learning_rate,
                                       process_group=self._group)
    os.makedirs(hp.model_dir, exist_ok=True)
    ckpt = os.path.join(hp.model_dir, 'ckpt.pt')
    best_dir = os.path.join(hp.model_dir, 'best_checkpoint')
    step = load_checkpoint(trainer, ckpt) if os.path.exists(ckpt) else 0
    first_epoch = step // hp.steps_per_epoch
    higher = hp.best_exporter_metric_higher_better
    best, bad_epochs, lr_bad = None, 0, 0
    train_it = iter(self._train_fn())
    dev = trainer.device
    for epoch in range(first_epoch, hp.num_epochs):
      loss_sum, n = 0.0, 0
      while step < (epoch + 1) * hp.steps_per_epoch:
        try:
          x, y = next(train_it)
        except StopIteration:
          train_it = iter(self._train_fn())
          x, y = next(train_it)
        loss = trainer.train_step(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True))
        step += 1
        loss_sum += float(loss)
        n += 1
      logs = {'epoch': epoch + 1, 'step': step, 'train_loss': loss_sum / max(n, 1)}
      if self._valid_fn is not None and hp.validation_steps > 0:
        import itertools
        vb = list(itertools.islice(iter(self._valid_fn()), hp.validation_steps))
        logs.update(evaluate(trainer, vb, self.build_metrics()))
        vloss = 0.0
        loss_obj = trainer.loss
        for x, y in vb:
          xd, yd = x.to(dev), y.to(dev)
          vloss += float(loss_obj(yd, trainer.predict(xd, mask=(yd >= 0)).clone()))
        logs['loss'] = vloss / max(len(vb), 1)
      else:
        logs['loss'] = logs['train_loss']
      self.history.append(logs)
      if verbose:
        self._log(str(logs))
      save_checkpoint(trainer, ckpt, step)
      cur = self._monitor(logs)
      improved = best is None or (cur > best + hp.early_stopping_min_delta if higher
                                  else cur < best - hp.early_stopping_min_delta)
      if improved:
        best, bad_epochs, lr_bad = cur, 0, 0
        if hp.export_best_model:
          os.makedirs(best_dir, exist_ok=True)
          save_checkpoint(trainer, os.path.join(best_dir, 'ckpt.pt'), step)
      else:
        bad_epochs += 1
        lr_bad += 1
        if hp.automatic_reduce_lr and lr_bad >= max(1, hp.num_epochs // 3):
          trainer.lr *= 0.1        # ReduceLROnPlateau(factor=0.1), keras/pipeline.py:465-472
          lr_bad = 0
        if hp.early_stopping_patience > 0 and bad_epochs >= hp.early_stopping_patience:
          break
    self.trainer = trainer
    return self.history
