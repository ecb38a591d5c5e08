"""Oracle training step: the reference's per-replica step on CPU.

scorer (create_tower) -> RestoreList -> tfr.keras loss -> autograd ->
tf.keras.optimizers.Adagrad.  Test infrastructure / CPU baseline only
(see oracle/__init__.py); uses the same [B, N, N] formulation as the reference.
"""
import torch

from oracle import keras_losses
from oracle import scorer


class OracleTrainer(object):

  def __init__(self, input_dim, hidden_layer_dims, loss_key='approx_ndcg_loss',
               activation='relu', learning_rate=0.001, epsilon=1e-7,
               initial_accumulator_value=0.1, seed=1238, dtype=torch.float32,
               loss_kwargs=None, group_size=1):
    # group_size > 1: groupwise scoring (model.py:273-421), the tower is the group score
    # function over the concatenated member features
    self.group_size = group_size
    self.params = scorer.init_tower_params(input_dim * group_size, hidden_layer_dims,
                                           group_size, seed=seed, dtype=dtype)
    self.activation = activation
    self.loss = keras_losses.get(loss_key, **(loss_kwargs or {}))
    self.lr = learning_rate
    self.eps = epsilon
    self.leaves = self.params['dense_w'] + self.params['dense_b']
    for p in self.leaves:
      p.requires_grad_()
    self.accum = [torch.full_like(p, initial_accumulator_value)
                  for p in self.leaves]

  def forward(self, x, mask):
    b, n, d = x.shape
    if self.group_size > 1:
      gs = self.group_size
      score_fn = lambda gf: scorer.tower_forward(   # noqa: E731
          gf.reshape(gf.shape[0], gs * d), self.params, activation=self.activation)
      return scorer.groupwise_logits(x, mask, gs, score_fn)
    flat = scorer.tower_forward(x.reshape(b * n, d), self.params,
                                activation=self.activation)
    return scorer.restore_list(flat, mask)

  def train_step(self, x, y_true, sample_weight=None):
    mask = y_true >= 0
    for p in self.leaves:
      p.grad = None
    logits = self.forward(x, mask)
    loss = self.loss(y_true, logits, sample_weight)
    loss.backward()
    with torch.no_grad():
      for p, a in zip(self.leaves, self.accum):
        a.add_(p.grad * p.grad)
        p.sub_(self.lr * p.grad / (a.sqrt() + self.eps))
    return float(loss.detach())
