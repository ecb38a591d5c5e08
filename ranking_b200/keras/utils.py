"""Gain / discount / positive functions of tfr.keras.utils (keras/utils.py:50-135).

They are ordinary callables on torch tensors, so user code written against
`tfr.keras.utils` keeps working.  The CUDA kernels cannot call Python, so the
loss / metric wrappers recognise these functions by identity and select the
matching in-kernel enum; any other callable is evaluated once per batch into a
small table that the kernel reads (`_C.GAIN_TABLE` / `_C.DISC_TABLE`).
"""
import math

import torch


def identity(label):
  """keras/utils.py:50-62."""
  return label


def inverse(rank):
  """keras/utils.py:65-76: divide_no_nan(1, rank)."""
  rank = torch.as_tensor(rank)
  r = rank.to(rank.dtype if rank.is_floating_point() else torch.float32)
  return torch.where(r == 0, torch.zeros_like(r), 1. / r)


def pow_minus_1(label):
  """keras/utils.py:79-91: 2**x - 1."""
  return torch.pow(2., torch.as_tensor(label)) - 1.


def log2_inverse(rank):
  """keras/utils.py:94-107: divide_no_nan(ln 2, log1p(rank))."""
  rank = torch.as_tensor(rank)
  r = rank.to(rank.dtype if rank.is_floating_point() else torch.float32)
  d = torch.log1p(r)
  return torch.where(d == 0, torch.zeros_like(d), math.log(2.) / d)


def log1p_inverse(rank):
  """1 / log1p(rank): the estimator-era default (losses_impl.py:111, losses.py:455)."""
  rank = torch.as_tensor(rank)
  r = rank.to(rank.dtype if rank.is_floating_point() else torch.float32)
  return 1. / torch.log1p(r)


def is_greater_equal_1(label):
  """keras/utils.py:110-121."""
  return torch.as_tensor(label) >= 1.0


def symmetric_log1p(t):
  """keras/utils.py:124-135."""
  t = torch.as_tensor(t)
  return torch.log1p(t * torch.sign(t)) * torch.sign(t)
