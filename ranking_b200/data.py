"""Input side of the hot path: LibSVM ranking files -> dense [num_queries, list_size, D]
features + labels (the format of examples/tf_ranking_libsvm.py:137-195), and
pinned-memory batch iteration for `HostBatchPipeline`.
"""
import numpy as np
import torch


def load_libsvm_data(path, list_size, num_features):
  """examples/tf_ranking_libsvm.py:137-195: one line per document,
  `label qid:<id> <feature>:<value> ... [# comment]`; feature ids are 1-based.

  Queries keep the order of first appearance; only the first `list_size` documents
  of a query are kept; missing features are 0; padded slots have label -1.
  Returns (features float32 [Q, list_size, num_features], labels float32 [Q, list_size]).
  """
  qid_to_index = {}
  qid_to_ndoc = []
  feats, labels = [], []
  total_docs = discarded_docs = 0
  with open(path, 'rt') as f:
    for line in f:
      tokens = line.split('#')[0].split()
      if not tokens:
        continue
      assert len(tokens) >= 2, 'Ill-formatted line: {}'.format(line)
      label = float(tokens[0])
      qid = tokens[1]
      idx = qid_to_index.get(qid)
      if idx is None:
        idx = qid_to_index[qid] = len(qid_to_index)
        qid_to_ndoc.append(0)
        feats.append(np.zeros([list_size, num_features], dtype=np.float32))
        labels.append(np.full([list_size], -1., dtype=np.float32))
      total_docs += 1
      doc_idx = qid_to_ndoc[idx]
      qid_to_ndoc[idx] += 1
      if doc_idx >= list_size:      # keep the first `list_size` docs only
        discarded_docs += 1
        continue
      row = feats[idx][doc_idx]
      for kv in tokens[2:]:
        k, v = kv.split(':')
        k = int(k)
        assert 1 <= k <= num_features, 'Key {} not found in features.'.format(k)
        row[k - 1] = float(v)
      labels[idx][doc_idx] = label
  info = {'num_queries': len(qid_to_index), 'num_docs': total_docs,
          'num_discarded': discarded_docs}
  if not feats:
    return (np.zeros([0, list_size, num_features], np.float32),
            np.zeros([0, list_size], np.float32), info)
  return np.stack(feats), np.stack(labels), info


def batch_iterator(features, labels, batch_size, shuffle=False, seed=0,
                   drop_remainder=True, repeat=False, pin_memory=True):
  """Yields (x [B, N, D], y [B, N]) CPU tensors (pinned when CUDA is there) in the
  order tf.data's from_tensor_slices -> shuffle -> batch would (the shuffle here is a
  full permutation per epoch)."""
  x_all = torch.as_tensor(np.ascontiguousarray(features), dtype=torch.float32)
  y_all = torch.as_tensor(np.ascontiguousarray(labels), dtype=torch.float32)
  q = x_all.shape[0]
  pin = pin_memory and torch.cuda.is_available()
  gen = torch.Generator().manual_seed(seed)
  while True:
    order = torch.randperm(q, generator=gen) if shuffle else torch.arange(q)
    for lo in range(0, q, batch_size):
      idx = order[lo:lo + batch_size]
      if idx.numel() < batch_size and drop_remainder:
        break
      x, y = x_all[idx].contiguous(), y_all[idx].contiguous()
      if pin:
        x, y = x.pin_memory(), y.pin_memory()
      yield x, y
    if not repeat:
      return
