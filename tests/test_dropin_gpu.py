# coding=utf-8
"""GPU test of the reference-facing surface: get_model -> Tester.step(sess, batch) through the
`tensorflow` shim's Session, i.e. the call code/test.py and code/pred_utils.evaluate make
(code/pred_models.py:1761-1790, code/pred_utils.py:415), checked against the oracle."""
import os
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_batch(cfg, feeds, n):
  ns = len(cfg.scene_grids)
  data = dict(
      obs_grid_class=[np.stack([feeds["grid_obs_labels"][j][i] for j in range(ns)]) for i in range(n)],
      pred_grid_class=[np.stack([feeds["grid_pred_labels"][j][i] for j in range(ns)]) for i in range(n)],
      batch_scene_feat=feeds["scene_feat"], batch_obs_scene=feeds["obs_scene"][:, :, None],
      original_batch_size=n)
  for j in range(ns):
    data["obs_grid_target_all_%d" % j] = [feeds["grid_obs_regress"][j][i] for i in range(n)]
  return (tuple(range(n)), types.SimpleNamespace(data=data))


@pytest.mark.parametrize("beam", [False, True])
def test_tester_step_through_session(beam, monkeypatch):
  monkeypatch.syspath_prepend(os.path.join(ROOT, "multiverse_b200", "dropin"))
  for m in ("tensorflow", "pred_models", "multiverse_b200.pred_models"):
    monkeypatch.delitem(sys.modules, m, raising=False)
  import tensorflow as tf
  import pred_models
  from multiverse_b200 import synthetic
  from oracle import multiverse_ref as R
  tf.reset_default_graph()
  over = dict(batch_size=2)
  if beam:
    over.update(use_grids=[False, True], use_beam_search=True, beam_size=6, diverse_beam=True,
                diverse_gamma=0.01, fix_num_timestep=1)
  cfg = synthetic.make_config(**over)
  args = types.SimpleNamespace(**vars(cfg))
  args.modelname, args.use_soft_grid_class, args.use_gt_grid = "m", False, False
  w = synthetic.make_weights(cfg, 77)
  feeds = synthetic.make_feeds(cfg, 2, 77)
  model = pred_models.get_model(args, gpuid=0)
  tf.global_variables_initializer().run()
  for v in tf.global_variables():
    key = v.name.split(":")[0]
    if key in w:
      v.assign(w[key])
  with tf.Session(config=tf.ConfigProto(allow_soft_placement=True)) as sess:
    tester = pred_models.Tester(model, args, sess)
    cls, reg, beam_out = tester.step(sess, make_batch(cfg, feeds, 2))
    assert int(sess.run(model.global_step)) == 0
  ref = R.forward(R.default_config(**over), w, feeds, np.float64)
  rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
  for i in range(2):
    if not cfg.use_grids[i]:
      assert cls[i] == [] and reg[i] == []
      continue
    assert isinstance(cls[i], np.ndarray) and cls[i].shape == ref["grid_pred_decoded"][i].shape
    assert rel(cls[i], ref["grid_pred_decoded"][i]) < 1e-4
    assert rel(reg[i], ref["grid_pred_reg_decoded"][i]) < 1e-4
  if beam:
    lg, ids, lp = beam_out
    assert ids.dtype == np.int32 and np.array_equal(ids, ref["beam_outputs"][1])
    assert lg.shape == (2, 6, 12, 18 * 9) and lp.shape == (2, 6)
  else:
    assert beam_out is None
  # north_star's unit surface: Model.enc_cell / dec_cell on NHWC numpy tensors
  i = 1 if beam else 0
  h, ww = cfg.scene_grids[i]
  rng = np.random.default_rng(0)
  x = rng.standard_normal((2, h, ww, 32)).astype(np.float32)
  c = rng.standard_normal((2, h, ww, 256)).astype(np.float32)
  hh = np.tanh(rng.standard_normal((2, h, ww, 256))).astype(np.float32)
  c1, h1 = model.dec_cell(x, (c, hh), scale=i, kind="class")
  sw = R.scale_weights(R.cast_tree(w, np.float64), i)
  c_ref, h_ref = R.convlstm_cell(x.astype(np.float64), c.astype(np.float64), hh.astype(np.float64), *sw.dec_class)
  assert rel(c1, c_ref) < 3e-5 and rel(h1, h_ref) < 3e-5


def test_rollout_length_follows_fed_pred_length(monkeypatch):
  """raw_rnn stops at the FED pred_length (code/pred_models.py:347, :520); multifuture_inference.py
  feeds max_pred_lengths[idx] (:311), which differs from config.pred_len."""
  monkeypatch.syspath_prepend(os.path.join(ROOT, "multiverse_b200", "dropin"))
  for m in ("tensorflow", "pred_models", "multiverse_b200.pred_models"):
    monkeypatch.delitem(sys.modules, m, raising=False)
  import tensorflow as tf
  import pred_models
  from multiverse_b200 import synthetic
  from oracle import multiverse_ref as R
  tf.reset_default_graph()
  over = dict(batch_size=1, use_grids=[False, True], use_beam_search=True, beam_size=5, diverse_beam=True,
              diverse_gamma=0.01, fix_num_timestep=1)
  cfg = synthetic.make_config(**over)                      # config.pred_len = 12
  args = types.SimpleNamespace(**vars(cfg)); args.modelname = "m"; args.use_soft_grid_class = False
  args.use_gt_grid = False
  w = synthetic.make_weights(cfg, 8); feeds = synthetic.make_feeds(cfg, 1, 8)
  model = pred_models.get_model(args, gpuid=0)
  tf.global_variables_initializer().run()
  for v in tf.global_variables():
    if v.name.split(":")[0] in w:
      v.assign(w[v.name.split(":")[0]])
  fd = model.get_feed_dict(make_batch(cfg, dict(feeds, grid_pred_labels=[np.zeros((1, 12), np.int32)] * 2), 1)[1])
  fd[model.pred_length] = np.array([15], dtype="int32")
  with tf.Session() as sess:
    cls, reg, beam = sess.run([model.grid_pred_decoded[1], model.grid_pred_reg_decoded[1], model.beam_outputs], fd)
  ref = R.forward(R.default_config(pred_len=15, **over), w, feeds, np.float64)
  assert cls.shape == (1, 15, 18, 9, 1) and reg.shape == (1, 15, 18, 9, 2) and beam[1].shape == (1, 5, 15)
  assert np.array_equal(beam[1], ref["beam_outputs"][1])
  assert float(np.abs(reg - ref["grid_pred_reg_decoded"][1]).max() / np.abs(ref["grid_pred_reg_decoded"][1]).max()) < 1e-4


def test_compact_grid_feeds_equal_dense_feeds(monkeypatch):
  """SURVEY.md section 8 row f-1, wired: a pred_utils-style batch that carries the observed trajectories and the
  grid centres is fed as trajectories (Model.get_feed_dict), the engine rebuilds the dense per-cell offsets on the
  device, and every fetched tensor equals the dense-fed run bit for bit."""
  monkeypatch.syspath_prepend(os.path.join(ROOT, "multiverse_b200", "dropin"))
  for m in ("tensorflow", "pred_models", "multiverse_b200.pred_models"):
    monkeypatch.delitem(sys.modules, m, raising=False)
  import tensorflow as tf
  import pred_models
  from multiverse_b200 import synthetic
  tf.reset_default_graph()
  cfg = synthetic.make_config(batch_size=3)
  args = types.SimpleNamespace(**vars(cfg))
  args.modelname, args.use_soft_grid_class, args.use_gt_grid = "m", False, False
  w = synthetic.make_weights(cfg, 5); feeds = synthetic.make_feeds(cfg, 3, 5)
  model = pred_models.get_model(args, gpuid=0)
  tf.global_variables_initializer().run()
  for v in tf.global_variables():
    if v.name.split(":")[0] in w:
      v.assign(w[v.name.split(":")[0]])
  _, batch = make_batch(cfg, feeds, 3)
  batch.data["obs_traj"] = list(feeds["traj64"][:, :cfg.obs_len])
  batch.shared = {"grid_center_%d" % j: c for j, c in enumerate(synthetic.grid_centers(cfg))}
  args.device_grid_feeds = False
  dense_fd = model.get_feed_dict(batch)
  args.device_grid_feeds = True
  compact_fd = model.get_feed_dict(batch)
  assert model.obs_traj in compact_fd and model.grid_obs_regress[0] not in compact_fd
  fetches = [model.grid_pred_decoded[0], model.grid_pred_reg_decoded[0], model.grid_pred_decoded[1],
             model.grid_pred_reg_decoded[1]]
  with tf.Session() as sess:
    a = sess.run(fetches, dense_fd)
    b = sess.run(fetches, compact_fd)
  for x, y in zip(a, b):
    assert np.array_equal(x, y)


def test_fetched_arrays_own_their_pinned_blocks(monkeypatch):
  """Session.run hands out numpy arrays over pooled pinned blocks: a result the caller still holds is never
  overwritten by a later run, and a released block is used again instead of a fresh pinned allocation."""
  monkeypatch.syspath_prepend(os.path.join(ROOT, "multiverse_b200", "dropin"))
  for m in ("tensorflow", "pred_models", "multiverse_b200.pred_models"):
    monkeypatch.delitem(sys.modules, m, raising=False)
  import tensorflow as tf
  import pred_models
  from multiverse_b200 import synthetic
  tf.reset_default_graph()
  cfg = synthetic.make_config(batch_size=2)
  args = types.SimpleNamespace(**vars(cfg)); args.modelname = "m"; args.use_soft_grid_class = False
  args.use_gt_grid = False
  w = synthetic.make_weights(cfg, 5)
  model = pred_models.get_model(args, gpuid=0)
  tf.global_variables_initializer().run()
  for v in tf.global_variables():
    if v.name.split(":")[0] in w:
      v.assign(w[v.name.split(":")[0]])
  fds = []
  for seed in (5, 6):
    feeds = synthetic.make_feeds(cfg, 2, seed)
    fds.append(model.get_feed_dict(make_batch(cfg, feeds, 2)[1]))
  fetch = [model.grid_pred_decoded[0], model.grid_pred_reg_decoded[0]]
  with tf.Session() as sess:
    a = sess.run(fetch, fds[0])
    keep = [x.copy() for x in a]
    b = sess.run(fetch, fds[1])                   # `a` is still alive: must land in other blocks
    assert all(np.array_equal(x, y) for x, y in zip(a, keep))
    assert not np.array_equal(a[0], b[0])
    view = a[1][:, 3:]                            # a view keeps its block alive after `a` is gone
    ref_view = view.copy()
    del a
    c = sess.run(fetch, fds[1])
    assert all(np.array_equal(x, y) for x, y in zip(b, c))
    assert np.array_equal(view, ref_view)
    sizes = {k: len(v) for k, v in model._pinned.items()}
    assert sizes and max(sizes.values()) == 3, sizes      # a's view, b, c
    del view, b, c
    for _ in range(6):                            # nothing held: the same blocks serve every run
      sess.run(fetch, fds[0])
    assert {k: len(v) for k, v in model._pinned.items()} == sizes
