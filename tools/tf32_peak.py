"""cuBLAS TF32 / bf16 GEMM throughput on this box (sets the tensor roofline for kind::tf32)."""
import torch
torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
for dt, name in ((torch.float32, 'tf32'), (torch.bfloat16, 'bf16')):
  a = torch.randn(n, n, device='cuda', dtype=dt)
  b = torch.randn(n, n, device='cuda', dtype=dt)
  for _ in range(3):
    a @ b
  torch.cuda.synchronize()
  best = 1e9
  for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
      a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 5)
  print(name, 'TFLOP/s', 2 * n ** 3 / (best * 1e-3) / 1e12, 'ms', best)
