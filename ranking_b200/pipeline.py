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
