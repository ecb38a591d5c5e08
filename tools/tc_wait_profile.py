"""Per-warp-role wait cycles of the persistent tcgen05 GEMM for the bench shapes."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ranking_b200 as tfr
from ranking_b200 import _C

M = 204800
dbg = torch.zeros(148, 12, dtype=torch.int64, device='cuda')
_C.lib.tfr_tc_set_debug(ctypes.c_void_p(dbg.data_ptr()))
NAMES = ['prod_wait_empty', 'prod_total', 'mma_wait_accempty', 'mma_wait_operands',
         'mma_total', 'split_wait_tma', 'epi_wait_accfull', 'epi_total', 'epi_tmem_ld',
         'epi_math', 'epi_wait_stage', 'epi_sts_fence']


def run(name, gm, gn, gk, a_mn, b_mn, passes, split_b, epi, transposed=0, splits=1):
  A = torch.randn((gk, gm) if a_mn else (gm, gk), device='cuda')
  B = torch.randn((gk, gn) if b_mn else (gn, gk), device='cuda')
  Blo = torch.randn_like(B) * 1e-4 if (passes == 3 and not split_b) else None
  bias = torch.randn(gn, device='cuda')
  aux = torch.randn(gm, gn, device='cuda')
  bits = torch.randint(-2 ** 31, 2 ** 31 - 1, ((gn + 31) // 32, gm), dtype=torch.int32,
                       device='cuda')
  stride = gm * gn if splits > 1 else 0
  C = torch.empty(max(splits, 1) * gm * gn, device='cuda')
  ldc = gm if transposed else gn
  for rep in range(3):
    dbg.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _C.check(_C.lib.tfr_tc_gemm(_C.ptr(A), A.shape[1], _C.ptr(B), B.shape[1], _C.ptr(Blo),
                                _C.ptr(C), ldc, gm, gn, gk, a_mn, b_mn, passes, split_b, epi,
                                _C.ptr(bias), _C.ptr(aux), 1, transposed, splits, stride,
                                _C.ptr(bits if epi == 1 else None),
                                _C.ptr(bits if epi == 3 else None), _C.stream()))
    e1.record()
    torch.cuda.synchronize()
  dd = dbg.double()
  # per column: mean over the CTAs that wrote it (in pair mode only the leaders issue MMAs)
  cnt = (dd != 0).sum(0).clamp(min=1)
  d = (dd.sum(0) / cnt).tolist()
  print('%-34s %7.1f us  ' % (name, e0.elapsed_time(e1) * 1e3) +
        '  '.join('%s=%.0fk' % (n, v / 1e3) for n, v in zip(NAMES, d)), flush=True)


for passes in (3,):
  print('passes', passes)
  # forward: B = W^T [out, in] pre-split, K-major (mlp_tc.cu); CTA pairs when enabled
  run('fwd L1 136->256', M, 256, 136, 0, 0, passes, 0, 1)
  run('fwd L2 256->128', M, 128, 256, 0, 0, passes, 0, 1)
  run('fwd L3 128->64', M, 64, 128, 0, 0, passes, 0, 1)
  run('dH1 128->256 mask', M, 256, 128, 0, 0, passes, 0, 2)
  run('dH2 64->128 mask', M, 128, 64, 0, 0, passes, 0, 2)
  run('dH1 128->256 bits', M, 256, 128, 0, 0, passes, 0, 3)
  run('dH2 64->128 bits', M, 128, 64, 0, 0, passes, 0, 3)
  run('dH1 store only', M, 256, 128, 0, 0, passes, 0, 0)
  run('dW1^T 256x136 (swapped)', 256, 136, M, 1, 1, passes, 1 if passes == 3 else 0, 0, 1, 100)
  run('dW2 256x128', 256, 128, M, 1, 1, passes, 1 if passes == 3 else 0, 0, 0, 100)
  run('dW3 128x64', 128, 64, M, 1, 1, passes, 1 if passes == 3 else 0, 0, 0, 100)
