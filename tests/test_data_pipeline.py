"""LibSVM loader (examples/tf_ranking_libsvm.py:137-195 semantics) and the fit /
checkpoint / resume loop."""
import os

import numpy as np
import pytest
import torch

LIBSVM = """2 qid:10 1:0.5 3:1.5 # doc a
0 qid:10 2:2.0
1 qid:7 1:1.0 4:4.0
0 qid:10 4:-1.0
3 qid:7 2:0.25
1 qid:10 1:9.0
"""


def test_load_libsvm_data(tmp_path):
  from ranking_b200 import data
  p = tmp_path / 'train.txt'
  p.write_text(LIBSVM)
  x, y, info = data.load_libsvm_data(str(p), list_size=3, num_features=4)
  assert info == {'num_queries': 2, 'num_docs': 6, 'num_discarded': 1}
  assert x.shape == (2, 3, 4) and y.shape == (2, 3)
  np.testing.assert_array_equal(y, [[2., 0., 0.], [1., 3., -1.]])   # qid 10 first; 4th doc cut
  np.testing.assert_array_equal(x[0], [[0.5, 0., 1.5, 0.], [0., 2., 0., 0.],
                                       [0., 0., 0., -1.]])
  np.testing.assert_array_equal(x[1], [[1., 0., 0., 4.], [0., 0.25, 0., 0.],
                                       [0., 0., 0., 0.]])
  with pytest.raises(AssertionError):
    data.load_libsvm_data(str(p), list_size=3, num_features=3)      # feature 4 unknown


def test_batch_iterator(tmp_path):
  from ranking_b200 import data
  x = np.arange(5 * 2 * 3, dtype=np.float32).reshape(5, 2, 3)
  y = np.arange(10, dtype=np.float32).reshape(5, 2)
  got = list(data.batch_iterator(x, y, 2, pin_memory=False))
  assert len(got) == 2 and got[0][0].shape == (2, 2, 3)
  np.testing.assert_array_equal(got[1][1].numpy(), y[2:4])
  got = list(data.batch_iterator(x, y, 2, drop_remainder=False, pin_memory=False))
  assert len(got) == 3 and got[2][0].shape[0] == 1
  a = [b[1].clone() for b in data.batch_iterator(x, y, 2, shuffle=True, seed=3,
                                                 pin_memory=False)]
  b = [b[1].clone() for b in data.batch_iterator(x, y, 2, shuffle=True, seed=3,
                                                 pin_memory=False)]
  assert all(torch.equal(p, q) for p, q in zip(a, b))


@pytest.mark.gpu
def test_fit_checkpoint_resume(tmp_path):
  """An interrupted run resumed from its checkpoint ends on the same parameters as an
  uninterrupted one (BN state and the Adagrad accumulator are part of the checkpoint)."""
  import ranking_b200 as tfr
  from ranking_b200 import data, pipeline
  g = torch.Generator().manual_seed(0)
  x = torch.randn(24, 10, 8, generator=g).numpy()
  y = torch.randint(0, 3, (24, 10), generator=g).float().numpy()

  def make():
    tower = tfr.keras.layers.create_tower([16, 8], 1, activation='relu',
                                          use_batch_norm=True, input_batch_norm=True,
                                          dropout=0.0, input_dim=8, seed=4)
    return tfr.train.RankingTrainer(tower, tfr.keras.losses.get('softmax_loss'),
                                    optimizer='adagrad', learning_rate=0.05)

  ref = make()
  pipeline.fit(ref, data.batch_iterator(x, y, 4, repeat=True), 9)
  run = make()
  d = str(tmp_path / 'ckpt')
  step, _ = pipeline.fit(run, data.batch_iterator(x, y, 4, repeat=True), 5,
                         checkpoint_dir=d, steps_per_checkpoint=2)
  assert step == 5 and os.path.exists(os.path.join(d, 'ckpt.pt'))
  resumed = make()
  it = data.batch_iterator(x, y, 4, repeat=True)
  for _ in range(5):      # the data iterator is the caller's: skip what was consumed
    next(it)
  step, _ = pipeline.fit(resumed, it, 9, checkpoint_dir=d, log_fn=lambda m: None)
  assert step == 9
  torch.testing.assert_close(resumed.tower.flat, ref.tower.flat, rtol=0, atol=0)
  torch.testing.assert_close(resumed.tower.bn_state, ref.tower.bn_state, rtol=0, atol=0)
  res = pipeline.evaluate(resumed, data.batch_iterator(x, y, 8, drop_remainder=False))
  assert set(res) >= {'metric/ndcg_10', 'metric/mrr', 'metric/arp', 'metric/map',
                      'metric/ordered_pair_accuracy', 'metric/precision', 'metric/dcg'}
