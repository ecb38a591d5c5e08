"""Data-parallel semantics on CPU with world_size 2 over gloo.

Two ranks each take half of the lists, compute the local Keras-reduced loss and
its gradient (CPU oracle standing in for the kernels), all-reduce(SUM) the flat
gradient through ranking_b200.dp and apply the 1/num_replicas factor: the result
must equal the single-process gradient over the full batch, and the Adagrad step
must leave both replicas with identical parameters.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(b, n, d, seed):
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(b, n, d, generator=g, dtype=torch.float64)
  y = torch.randint(0, 5, (b, n), generator=g).double()
  y[:, n - 3:] = -1.0
  return x, y


def _flat_grad(tr):
  return torch.cat([p.grad.reshape(-1) for p in tr.leaves])


def _worker(rank, world, port, loss_key, out):
  sys.path.insert(0, ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from oracle.train_step import OracleTrainer
    from ranking_b200 import dp
    assert dp.world_size() == world and dp.rank() == rank
    x, y = _batch(8, 12, 6, seed=5)
    sl = dp.shard_lists(x.shape[0])
    tr = OracleTrainer(6, [8, 4], loss_key, dtype=torch.float64, seed=3,
                       learning_rate=0.1)
    flat0 = torch.cat([p.detach().reshape(-1) for p in tr.leaves])
    dp.broadcast_(flat0, src=0)
    mask = y[sl] >= 0
    loss = tr.loss(y[sl], tr.forward(x[sl], mask))
    loss.backward()
    g = _flat_grad(tr).clone()
    dp.all_reduce_sum_(g)
    g *= dp.replica_grad_scale()
    # metric state reduction: (sum v*w, sum w) pairs add up across ranks
    state = torch.tensor([float(rank + 1), 2.0])
    dp.all_reduce_sum_(state)
    if rank == 0:
      out.put((g, float(loss), state))
    dist.barrier()
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('loss_key', ['approx_ndcg_loss', 'pairwise_logistic_loss',
                                      'softmax_loss'])
def test_two_rank_gradient_equals_full_batch(loss_key):
  from oracle.train_step import OracleTrainer
  ctx = mp.get_context('spawn')
  out = ctx.Queue()
  port = 29500 + (os.getpid() % 2000)
  procs = [ctx.Process(target=_worker, args=(r, 2, port, loss_key, out))
           for r in range(2)]
  for p in procs:
    p.start()
  g2, loss0, state = out.get(timeout=120)
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  x, y = _batch(8, 12, 6, seed=5)
  tr = OracleTrainer(6, [8, 4], loss_key, dtype=torch.float64, seed=3,
                     learning_rate=0.1)
  loss = tr.loss(y, tr.forward(x, y >= 0))
  loss.backward()
  g1 = _flat_grad(tr)
  torch.testing.assert_close(g2, g1, rtol=1e-10, atol=1e-12)
  assert state.tolist() == [3.0, 4.0]


def test_single_process_helpers_are_noops():
  sys.path.insert(0, ROOT)
  from ranking_b200 import dp
  t = torch.ones(4)
  assert dp.world_size() == 1 and dp.rank() == 0
  assert dp.all_reduce_sum_(t) is t and dp.replica_grad_scale() == 1.0
  assert dp.shard_lists(8) == slice(0, 8)


def _weights_worker(rank, world, port, out):
  sys.path.insert(0, ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from ranking_b200 import dp
    raw = _raw_rows()
    sl = dp.shard_lists(raw.shape[0])
    w = dp.cross_replica_list_weights(raw[sl], 'mrr')
    gathered = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(gathered, w)
    if rank == 0:
      out.put(torch.cat(gathered))
    dist.barrier()
  finally:
    dist.destroy_process_group()


def _raw_rows():
  # {sum w, sum w*gain, sum gain, sum w*rel, sum rel} for 6 lists: two without relevant
  # items (one of them on each rank), one with zero weights
  return torch.tensor([[3., 0., 0., 2.0, 2.], [3., 0., 0., 0.0, 0.], [0., 0., 0., 0.0, 1.],
                       [4., 0., 0., 6.0, 3.], [2., 0., 0., 0.0, 0.], [5., 0., 0., 0.5, 1.]],
                      dtype=torch.float64)


def test_cross_replica_list_weights_equal_single_process():
  """metrics_impl.py:63-119 over the concatenated batch == 2 ranks with the cross-replica
  average (and != the per-replica average when the shards differ)."""
  from oracle import metrics_impl as OM
  raw = _raw_rows()
  # oracle rule on the global batch: weights / relevance rows that reproduce `raw`
  sum_w, wr, r = raw[:, 0], raw[:, 3], raw[:, 4]
  per = torch.where(r != 0, wr / torch.where(r != 0, r, torch.ones_like(r)), torch.zeros_like(r))
  cnt = ((sum_w > 0) & (r > 0)).double().sum()
  avg = per.sum() / cnt
  want = torch.where(sum_w > 0, torch.where(r > 0, per, avg.expand_as(per)), torch.zeros_like(per))
  ctx = mp.get_context('spawn')
  out = ctx.SimpleQueue()
  port = 29500 + (os.getpid() % 2000) + 7
  procs = [ctx.Process(target=_weights_worker, args=(r_, 2, port, out)) for r_ in range(2)]
  for p in procs:
    p.start()
  got = out.get()
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert torch.allclose(got, want, rtol=1e-12)
  del OM
