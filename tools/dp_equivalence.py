"""torchrun helper: a 2-rank data-parallel step must equal the 1-GPU step on the
concatenated batch (parameters after one Adagrad step).  Prints DP_EQUIVALENCE_OK."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

local_rank = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local_rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
import ranking_b200 as tfr  # noqa: E402

world, rank = dist.get_world_size(), dist.get_rank()
b, n, d = 8 * world, 40, 16
g = torch.Generator().manual_seed(7)
x = torch.randn(b, n, d, generator=g).cuda()
y = torch.randint(0, 5, (b, n), generator=g).float().cuda()
y[:, n - 6:] = -1.


def make(collective):
  t = tfr.keras.layers.create_tower([32, 16], 1, activation='relu', use_batch_norm=False,
                                    dropout=0, input_dim=d, seed=11)
  return t, tfr.train.RankingTrainer(t, tfr.keras.losses.get('approx_ndcg_loss'),
                                     optimizer='adagrad', learning_rate=0.1,
                                     collective=collective)


COLLECTIVE = os.environ.get('TFR_COLLECTIVE', 'fused')   # the K7 kernel by default
tower, trainer = make(COLLECTIVE)
tfr.dp.broadcast_(tower.flat.data)
sl = tfr.dp.shard_lists(b)
loss = trainer.train_step(x[sl], y[sl])
torch.cuda.synchronize()
if rank == 0:
  ref_tower, ref_trainer = make('nccl')   # no peer-memory reducer: rank 0 steps alone
  ref_trainer.world = 1
  ref_trainer.group = None
  # single-process reference: no collective (use a trainer whose dp hooks are no-ops)
  import ranking_b200.dp as dpmod
  saved = (dpmod.all_reduce_sum_, dpmod.replica_grad_scale)
  dpmod.all_reduce_sum_ = lambda t, group=None: t
  dpmod.replica_grad_scale = lambda group=None: 1.0
  ref_trainer.train_step(x, y)
  dpmod.all_reduce_sum_, dpmod.replica_grad_scale = saved
  torch.cuda.synchronize()
  err = float((tower.flat.data - ref_tower.flat.data).abs().max() /
              ref_tower.flat.data.abs().max())
  print('max rel param diff', err)
  assert err < 1e-5, err
  print('DP_EQUIVALENCE_OK', COLLECTIVE)
dist.barrier()
dist.destroy_process_group()
