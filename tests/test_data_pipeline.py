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


# ------------------------- ELWC / TFRecord input side ------------------------
def _elwc_records():
  from ranking_b200 import data
  recs = [
      data.encode_elwc({'query_length': [3], 'qf': [0.5, 1.5]},
                       [{'utility': [0.0], 'f': [1.0, 2.0, 3.0]},
                        {'utility': [1.0], 'f': [4.0, 5.0, 6.0], 'unigrams': [b'to', b'rank']},
                        {'utility': [2.0]}]),                       # 'f' missing -> default
      data.encode_elwc({'query_length': [2]},                        # 'qf' missing -> default
                       [{'utility': [1.0], 'f': [7.0, 8.0, 9.0]}]),
      data.encode_elwc(None, [{'utility': [0.0], 'f': [1.0, 1.0, 1.0]}] * 5),   # truncated to 4
  ]
  return recs


def test_parse_from_example_list():
  """data.py:133-208 semantics: padding with defaults, truncation, sizes, mask."""
  from ranking_b200 import data
  out = data.parse_from_example_list(
      _elwc_records(), list_size=4,
      context_feature_spec={'query_length': (1, 0.0), 'qf': (2, -7.0)},
      example_feature_spec={'utility': (1, -1.0), 'f': (3, 0.25)})
  assert out['sizes'].tolist() == [3, 1, 5]
  assert out['mask'].tolist() == [[True, True, True, False], [True, False, False, False],
                                  [True, True, True, True]]
  np.testing.assert_array_equal(out['context'].numpy(),
                                [[3., 0.5, 1.5], [2., -7., -7.], [0., -7., -7.]])
  ex = out['examples'].numpy()
  np.testing.assert_array_equal(ex[0], [[0., 1., 2., 3.], [1., 4., 5., 6.],
                                        [2., .25, .25, .25], [-1., .25, .25, .25]])
  np.testing.assert_array_equal(ex[1], [[1., 7., 8., 9.]] + [[-1., .25, .25, .25]] * 3)
  np.testing.assert_array_equal(ex[2], [[0., 1., 1., 1.]] * 4)


def test_parse_from_example_list_errors():
  from ranking_b200 import data
  recs = _elwc_records()
  with pytest.raises(ValueError):      # wrong fixed length
    data.parse_from_example_list(recs, 4, None, {'f': (2, 0.0)})
  with pytest.raises(ValueError):      # bytes feature requested as dense input
    data.parse_from_example_list(recs, 4, None, {'unigrams': (1, 0.0)})
  with pytest.raises(ValueError):      # garbage bytes
    data.parse_from_example_list([b'\xff\xff\xff'], 4, None, {'f': (3, 0.0)})


def test_tfrecord_roundtrip_and_batches(tmp_path):
  from ranking_b200 import data
  path = str(tmp_path / 'train.tfrecord')
  recs = _elwc_records()
  data.write_tfrecords(path, recs)
  assert list(data.read_tfrecords(path)) == recs
  raw = bytearray(open(path, 'rb').read())
  raw[20] ^= 0xff                      # flip a payload byte: the data crc must catch it
  bad = str(tmp_path / 'bad.tfrecord')
  open(bad, 'wb').write(bytes(raw))
  with pytest.raises(ValueError):
    list(data.read_tfrecords(bad))
  batches = list(data.elwc_batches(
      path, batch_size=2, list_size=4, context_feature_spec={'qf': (2, 0.0)},
      example_feature_spec={'f': (3, 0.0), 'utility': (1, -1.0)}, label_feature='utility',
      drop_remainder=False, pin_memory=False))
  assert [tuple(x.shape) for x, _ in batches] == [(2, 4, 5), (1, 4, 5)]
  x, y = batches[0]
  np.testing.assert_array_equal(y.numpy(), [[0., 1., 2., -1.], [1., -1., -1., -1.]])
  np.testing.assert_array_equal(x[0, 1].numpy(), [0.5, 1.5, 4., 5., 6.])   # context first
  np.testing.assert_array_equal(x[1, 0].numpy(), [0., 0., 7., 8., 9.])


def _protobuf_classes():
  """tf.Example / ExampleListWithContext rebuilt with the real protobuf runtime
  (tensorflow/core/example/{example,feature}.proto and
  tensorflow_serving/apis/input.proto field numbers)."""
  from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
  fd = descriptor_pb2.FileDescriptorProto()
  fd.name = 'tfr_b200_test_example.proto'
  fd.package = 'tfrtest'
  fd.syntax = 'proto3'
  T = descriptor_pb2.FieldDescriptorProto

  def msg(name, fields):
    m = fd.message_type.add()
    m.name = name
    for fname, num, typ, label, tname in fields:
      f = m.field.add()
      f.name, f.number, f.type, f.label = fname, num, typ, label
      if tname:
        f.type_name = '.tfrtest.' + tname
    return m

  R, O = T.LABEL_REPEATED, T.LABEL_OPTIONAL
  msg('BytesList', [('value', 1, T.TYPE_BYTES, R, None)])
  msg('FloatList', [('value', 1, T.TYPE_FLOAT, R, None)])
  msg('Int64List', [('value', 1, T.TYPE_INT64, R, None)])
  feat = msg('Feature', [('bytes_list', 1, T.TYPE_MESSAGE, O, 'BytesList'),
                         ('float_list', 2, T.TYPE_MESSAGE, O, 'FloatList'),
                         ('int64_list', 3, T.TYPE_MESSAGE, O, 'Int64List')])
  del feat
  features = msg('Features', [('feature', 1, T.TYPE_MESSAGE, R, 'Features.FeatureEntry')])
  entry = features.nested_type.add()
  entry.name = 'FeatureEntry'
  entry.options.map_entry = True
  for fname, num, typ, tname in (('key', 1, T.TYPE_STRING, None),
                                 ('value', 2, T.TYPE_MESSAGE, 'Feature')):
    f = entry.field.add()
    f.name, f.number, f.type, f.label = fname, num, typ, O
    if tname:
      f.type_name = '.tfrtest.' + tname
  msg('Example', [('features', 1, T.TYPE_MESSAGE, O, 'Features')])
  msg('ExampleListWithContext', [('examples', 1, T.TYPE_MESSAGE, R, 'Example'),
                                 ('context', 2, T.TYPE_MESSAGE, O, 'Example')])
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fd)
  get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName('tfrtest.' + n))
  return get('Example'), get('ExampleListWithContext')


def test_parser_reads_protobuf_runtime_output():
  """Records serialized by the real protobuf runtime (packed repeated fields, map
  entries in its own order) decode to the same tensors as our writer's."""
  pytest.importorskip('google.protobuf')
  from ranking_b200 import data
  Example, ELWC = _protobuf_classes()
  elwc = ELWC()
  elwc.context.features.feature['qf'].float_list.value.extend([0.5, 1.5])
  elwc.context.features.feature['query_length'].int64_list.value.append(3)
  for util, f in ((0.0, [1.0, 2.0, 3.0]), (1.0, [4.0, 5.0, 6.0])):
    ex = elwc.examples.add()
    ex.features.feature['utility'].float_list.value.append(util)
    ex.features.feature['f'].float_list.value.extend(f)
    ex.features.feature['unigrams'].bytes_list.value.append(b'x')
  ex = elwc.examples.add()
  ex.features.feature['utility'].int64_list.value.append(-2)       # int64 -> float cast
  spec_c = {'query_length': (1, 0.0), 'qf': (2, -7.0)}
  spec_e = {'utility': (1, -1.0), 'f': (3, 0.25)}
  got = data.parse_from_example_list([elwc.SerializeToString()], 4, spec_c, spec_e)
  np.testing.assert_array_equal(got['context'].numpy(), [[3., 0.5, 1.5]])
  np.testing.assert_array_equal(
      got['examples'][0].numpy(),
      [[0., 1., 2., 3.], [1., 4., 5., 6.], [-2., .25, .25, .25], [-1., .25, .25, .25]])
  assert got['sizes'].tolist() == [3]
  # and the runtime parses what our writer emits
  mine = data.encode_elwc({'qf': [0.5, 1.5], 'query_length': [3]},
                          [{'utility': [0.0], 'f': [1.0, 2.0, 3.0], 'unigrams': [b'x']}])
  back = ELWC()
  back.ParseFromString(mine)
  assert list(back.context.features.feature['qf'].float_list.value) == [0.5, 1.5]
  assert list(back.context.features.feature['query_length'].int64_list.value) == [3]
  assert list(back.examples[0].features.feature['f'].float_list.value) == [1.0, 2.0, 3.0]
  assert list(back.examples[0].features.feature['unigrams'].bytes_list.value) == [b'x']


def _protobuf_sequence_example():
  from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
  fd = descriptor_pb2.FileDescriptorProto()
  fd.name = 'tfr_b200_test_seq.proto'
  fd.package = 'tfrseq'
  fd.syntax = 'proto3'
  T = descriptor_pb2.FieldDescriptorProto
  R, O = T.LABEL_REPEATED, T.LABEL_OPTIONAL

  def msg(name, fields, parent=None):
    m = (parent.nested_type if parent is not None else fd.message_type).add()
    m.name = name
    for fname, num, typ, label, tname in fields:
      f = m.field.add()
      f.name, f.number, f.type, f.label = fname, num, typ, label
      if tname:
        f.type_name = '.tfrseq.' + tname
    return m

  msg('BytesList', [('value', 1, T.TYPE_BYTES, R, None)])
  msg('FloatList', [('value', 1, T.TYPE_FLOAT, R, None)])
  msg('Int64List', [('value', 1, T.TYPE_INT64, R, None)])
  msg('Feature', [('bytes_list', 1, T.TYPE_MESSAGE, O, 'BytesList'),
                  ('float_list', 2, T.TYPE_MESSAGE, O, 'FloatList'),
                  ('int64_list', 3, T.TYPE_MESSAGE, O, 'Int64List')])
  features = msg('Features', [('feature', 1, T.TYPE_MESSAGE, R, 'Features.FeatureEntry')])
  e = msg('FeatureEntry', [('key', 1, T.TYPE_STRING, O, None),
                           ('value', 2, T.TYPE_MESSAGE, O, 'Feature')], features)
  e.options.map_entry = True
  msg('FeatureList', [('feature', 1, T.TYPE_MESSAGE, R, 'Feature')])
  fls = msg('FeatureLists', [('feature_list', 1, T.TYPE_MESSAGE, R,
                              'FeatureLists.FeatureListEntry')])
  e = msg('FeatureListEntry', [('key', 1, T.TYPE_STRING, O, None),
                               ('value', 2, T.TYPE_MESSAGE, O, 'FeatureList')], fls)
  e.options.map_entry = True
  msg('SequenceExample', [('context', 1, T.TYPE_MESSAGE, O, 'Features'),
                          ('feature_lists', 2, T.TYPE_MESSAGE, O, 'FeatureLists')])
  msg('Example', [('features', 1, T.TYPE_MESSAGE, O, 'Features')])
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fd)
  get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName('tfrseq.' + n))
  return get('SequenceExample'), get('Example')


def test_example_in_example_and_sequence_example_formats():
  """The three record formats of data.py:857-911 decode to the same tensors."""
  pytest.importorskip('google.protobuf')
  from ranking_b200 import data
  context = {'qf': [0.5, 1.5], 'query_length': [3]}
  examples = [{'utility': [0.0], 'f': [1.0, 2.0, 3.0]},
              {'utility': [1.0], 'f': [4.0, 5.0, 6.0]},
              {'utility': [2.0], 'f': [7.0, 8.0, 9.0]}]
  spec_c = {'query_length': (1, 0.0), 'qf': (2, -7.0)}
  spec_e = {'utility': (1, -1.0), 'f': (3, 0.25)}
  want = data.parse_from_example_list([data.encode_elwc(context, examples)], 4, spec_c, spec_e)
  for fmt, rec in ((data.EXAMPLE_IN_EXAMPLE, data.encode_example_in_example(context, examples)),
                   (data.SEQUENCE_EXAMPLE, data.encode_sequence_example(context, examples))):
    got = data.make_parsing_fn(fmt, 4, spec_c, spec_e)([rec])
    for k in ('context', 'examples', 'sizes', 'mask'):
      assert torch.equal(got[k], want[k]), (fmt, k)
  with pytest.raises(ValueError):
    data.make_parsing_fn('libsvm', 4, spec_c, spec_e)
  # truncation + a shorter / missing feature list in a SequenceExample
  rec = data.encode_sequence_example({}, [{'utility': [1.0], 'f': [1., 1., 1.]},
                                          {'utility': [0.0]}, {'utility': [2.0]}],
                                     feature_names=['utility', 'f'])
  got = data.parse_from_sequence_example([rec], 2, None, spec_e)
  assert got['sizes'].tolist() == [3]
  np.testing.assert_array_equal(got['examples'][0].numpy(),
                                [[1., 1., 1., 1.], [0., .25, .25, .25]])
  # records produced by the protobuf runtime
  SequenceExample, Example = _protobuf_sequence_example()
  se = SequenceExample()
  se.context.feature['qf'].float_list.value.extend([0.5, 1.5])
  se.context.feature['query_length'].int64_list.value.append(3)
  for ex in examples:
    se.feature_lists.feature_list['utility'].feature.add().float_list.value.extend(ex['utility'])
    se.feature_lists.feature_list['f'].feature.add().float_list.value.extend(ex['f'])
  got = data.parse_from_sequence_example([se.SerializeToString()], 4, spec_c, spec_e)
  for k in ('context', 'examples', 'sizes', 'mask'):
    assert torch.equal(got[k], want[k]), k
  outer = Example()
  inner_c = Example()
  inner_c.features.feature['qf'].float_list.value.extend([0.5, 1.5])
  inner_c.features.feature['query_length'].int64_list.value.append(3)
  outer.features.feature['serialized_context'].bytes_list.value.append(
      inner_c.SerializeToString())
  for ex in examples:
    inner = Example()
    inner.features.feature['utility'].float_list.value.extend(ex['utility'])
    inner.features.feature['f'].float_list.value.extend(ex['f'])
    outer.features.feature['serialized_examples'].bytes_list.value.append(
        inner.SerializeToString())
  got = data.parse_from_example_in_example([outer.SerializeToString()], 4, spec_c, spec_e)
  for k in ('context', 'examples', 'sizes', 'mask'):
    assert torch.equal(got[k], want[k]), k


def test_prefetcher_order_exceptions_and_close():
  """data.Prefetcher: same batches in the same order, producer exceptions re-raised at the
  consumer, close() stops a producer that is blocked on a full queue."""
  import time
  from ranking_b200 import data as D
  assert list(D.Prefetcher(iter(range(50)), depth=3)) == list(range(50))
  assert list(D.Prefetcher([], depth=1)) == []

  def boom():
    yield 1
    yield 2
    raise RuntimeError('decode failed')

  p = D.Prefetcher(boom(), depth=2)
  got = [next(p), next(p)]
  assert got == [1, 2]
  try:
    next(p)
    assert False, 'expected the producer exception'
  except RuntimeError as e:
    assert 'decode failed' in str(e)

  def endless():
    i = 0
    while True:
      yield i
      i += 1

  p = D.Prefetcher(endless(), depth=2)
  assert next(p) == 0
  p.close()
  time.sleep(0.05)
  assert not p._thread.is_alive()
  try:
    D.Prefetcher(iter([]), depth=0)
    assert False
  except ValueError:
    pass


def test_prefetcher_feeds_elwc_batches(tmp_path):
  """Prefetched ELWC batches equal the synchronous iterator's."""
  import torch
  from ranking_b200 import data as D
  recs = []
  for b in range(10):
    ex = [{'f': [float(b), float(i)], 'label': [float(i % 3)]} for i in range(1 + b % 4)]
    recs.append(D.encode_elwc({'c': [0.5 * b]}, ex))
  path = str(tmp_path / 'elwc.tfrecord')
  D.write_tfrecords(path, recs)
  args = (path, 4, 5, {'c': (1, 0.)}, {'f': (2, 0.), 'label': (1, -1.)}, 'label')
  plain = list(D.elwc_batches(*args, drop_remainder=False, pin_memory=False))
  pre = list(D.Prefetcher(D.elwc_batches(*args, drop_remainder=False, pin_memory=False), depth=2))
  assert len(plain) == len(pre) == 3
  for (x0, y0), (x1, y1) in zip(plain, pre):
    assert torch.equal(x0, x1) and torch.equal(y0, y1)


def test_pipeline_hparams_validation():
  """pipeline.PipelineHparams: the reference's field names and defaults
  (keras/pipeline.py:312-334); unsupported choices raise ValueError."""
  from ranking_b200 import pipeline as P
  hp = P.PipelineHparams(model_dir='/tmp/x', num_epochs=2, steps_per_epoch=5, validation_steps=2,
                         learning_rate=0.01, loss='softmax_loss')
  hp.validate()
  assert hp.steps_per_execution == 10 and hp.best_exporter_metric == 'loss'
  assert hp.early_stopping_patience == 0 and not hp.export_best_model
  import dataclasses
  for bad in (dict(optimizer='adam'), dict(loss={'a': 'softmax_loss'}), dict(num_epochs=0),
              dict(loss_weights={'a': 1.0})):
    try:
      dataclasses.replace(hp, **bad).validate()
      assert False, bad
    except ValueError:
      pass


@pytest.mark.gpu
def test_model_fit_pipeline_train_and_validate(tmp_path):
  """ModelFitPipeline: epochs x (steps, validation), history, checkpoint resume, best
  checkpoint, early stopping."""
  import torch
  import ranking_b200 as tfr
  from ranking_b200 import pipeline as P
  g = torch.Generator().manual_seed(0)
  w_true = torch.randn(8, generator=g)

  def batches():
    gg = torch.Generator().manual_seed(1)
    while True:
      x = torch.randn(16, 12, 8, generator=gg)
      y = ((x @ w_true) > 0.5).float() + ((x @ w_true) > 1.5).float()
      y[:, -2:] = -1.
      yield x, y

  def tower_fn():
    return tfr.keras.layers.create_tower([16, 8], 1, activation='relu', use_batch_norm=False,
                                         dropout=0, input_dim=8, seed=5)

  hp = P.PipelineHparams(model_dir=str(tmp_path), num_epochs=3, steps_per_epoch=20,
                         validation_steps=2, learning_rate=0.1, loss='approx_ndcg_loss',
                         export_best_model=True, best_exporter_metric='metric/ndcg_5',
                         best_exporter_metric_higher_better=True)
  pipe = P.ModelFitPipeline(tower_fn, batches, batches, hp)
  hist = pipe.train_and_validate()
  assert [h['epoch'] for h in hist] == [1, 2, 3] and hist[-1]['step'] == 60
  assert hist[-1]['metric/ndcg_5'] > hist[0]['metric/ndcg_5'] - 0.05
  assert hist[-1]['train_loss'] <= hist[0]['train_loss'] + 0.02   # (it learns; noise-tolerant)
  import os
  assert os.path.exists(os.path.join(str(tmp_path), 'ckpt.pt'))
  assert os.path.exists(os.path.join(str(tmp_path), 'best_checkpoint', 'ckpt.pt'))
  # resume: nothing left to do for the same hparams; two more epochs when asked for five
  pipe2 = P.ModelFitPipeline(tower_fn, batches, batches,
                             __import__('dataclasses').replace(hp, num_epochs=5))
  hist2 = pipe2.train_and_validate()
  assert [h['epoch'] for h in hist2] == [4, 5] and hist2[-1]['step'] == 100
  # early stopping on a metric that cannot improve
  hp3 = __import__('dataclasses').replace(hp, model_dir=str(tmp_path / 'es'), num_epochs=6,
                                          learning_rate=0.0, early_stopping_patience=2,
                                          best_exporter_metric='loss',
                                          best_exporter_metric_higher_better=False,
                                          export_best_model=False)
  hist3 = P.ModelFitPipeline(tower_fn, batches, batches, hp3).train_and_validate()
  assert len(hist3) == 3      # epoch 1 sets the best, epochs 2 and 3 do not improve
