"""Data-parallel plumbing of the hot path.

Lists are independent, the scorer parameters are shared: the batch is sharded by
lists across ranks (one process per GPU) and the ONLY data-path collective per
training step is one all-reduce(SUM) of the flat fp32 scorer gradient, with the
1/num_replicas factor of the reference (extension/task.py:256-262) folded into
the optimizer kernel.  Metric means all-reduce their (sum v*w, sum w) state.
Works on any backend: NCCL over NVLink on B200 boxes, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def is_distributed():
  return dist.is_available() and dist.is_initialized()


def world_size(group=None):
  return dist.get_world_size(group) if is_distributed() else 1


def rank(group=None):
  return dist.get_rank(group) if is_distributed() else 0


def all_reduce_sum_(tensor, group=None):
  """In-place SUM all-reduce of a flat buffer (no-op for a single process)."""
  if world_size(group) > 1:
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
  return tensor


def broadcast_(tensor, src=0, group=None):
  """Makes every replica start from rank `src`'s parameters."""
  if world_size(group) > 1:
    dist.broadcast(tensor, src=src, group=group)
  return tensor


def replica_grad_scale(group=None):
  """Factor applied to the summed gradient: 1 / num_replicas_in_sync."""
  return 1.0 / world_size(group)


def shard_lists(num_lists, group=None):
  """Contiguous, equal shard of the batch dimension for this rank (the reference
  semantics need equal local batches for SUM_OVER_BATCH_SIZE to equal the global
  loss)."""
  w, r = world_size(group), rank(group)
  if num_lists % w != 0:
    raise ValueError('batch of %d lists does not split evenly over %d replicas' %
                     (num_lists, w))
  per = num_lists // w
  return slice(r * per, (r + 1) * per)
