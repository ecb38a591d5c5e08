"""torchrun probe: where does time go in the multi-GPU step loop?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

local_rank = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local_rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
import ranking_b200 as tfr  # noqa: E402
import bench  # noqa: E402

rank = dist.get_rank()
g = torch.zeros(76353, device='cuda')


def timed(name, fn, n):
  torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter()
  e0.record()
  for _ in range(n):
    fn()
  e1.record()
  t_enq = time.perf_counter() - t0
  torch.cuda.synchronize()
  if rank == 0:
    print('%-28s gpu %.3f ms/iter   cpu-enqueue %.3f ms/iter' % (
        name, e0.elapsed_time(e1) / n, t_enq * 1e3 / n), flush=True)


for _ in range(5):
  dist.all_reduce(g)
timed('all_reduce 76K floats', lambda: dist.all_reduce(g), 50)

tower = tfr.keras.layers.create_tower(bench.HIDDEN, 1, activation='relu', use_batch_norm=False,
                                      dropout=0, input_dim=bench.D, seed=1238, precision='tf32x3')
tr = tfr.train.RankingTrainer(tower, tfr.keras.losses.get(bench.LOSS_KEY), learning_rate=0.05)
x, y = bench.make_batch(1234 + rank)
x, y = x.cuda(), y.cuda()
for _ in range(5):
  tr.train_step(x, y)
timed('train_step (with allreduce)', lambda: tr.train_step(x, y), 20)
world = tr.world
tr.world = 1
import ranking_b200.dp as dpm
saved = dpm.all_reduce_sum_
dpm.all_reduce_sum_ = lambda t, group=None: t
timed('train_step (no allreduce)', lambda: tr.train_step(x, y), 20)
dpm.all_reduce_sum_ = saved
timed('train_step (with allreduce) again', lambda: tr.train_step(x, y), 20)
dist.barrier()
dist.destroy_process_group()
