"""Throughput of the native record decoder (host side) on the bench's list shape:
lists/s and MB/s of serialized ELWC for 1..all host threads."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ranking_b200 import data

N, D, B = 200, 136, 256
rng = np.random.default_rng(0)
recs = []
for _ in range(B):
  n = int(rng.integers(N // 2, N + 1))
  recs.append(data.encode_elwc({'q': [1.0]}, [
      {'f': rng.standard_normal(D).astype(np.float32).tolist(),
       'y': [float(rng.integers(0, 5))]} for _ in range(n)]))
nbytes = sum(map(len, recs))
print('serialized bytes per list', nbytes // B)
spec_e = {'f': (D, 0.0), 'y': (1, -1.0)}
threads = 1
while threads <= (os.cpu_count() or 1):
  out = data.parse_from_example_list(recs, N, {'q': (1, 0.0)}, spec_e, num_threads=threads)
  t0 = time.time()
  for _ in range(5):
    out = data.parse_from_example_list(recs, N, {'q': (1, 0.0)}, spec_e, num_threads=threads,
                                       out=out)       # reused (pre-faulted) buffers
  dt = (time.time() - t0) / 5
  print('threads %3d  %8.0f lists/s  %7.0f MB/s' % (threads, B / dt, nbytes / dt / 1e6))
  threads *= 2
