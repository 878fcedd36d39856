# coding=utf-8
"""Executes the UNMODIFIED reference ``/root/reference/code/pred_models.py`` on the eager TF-1.15
stand-in of this directory (``oracle/tf1_eager/tensorflow``).  TEST INFRASTRUCTURE.

What runs is the reference's own code, line by line: ``Model.__init__`` (placeholders :62-115),
``build_forward`` (:123-308), ``grid_decoder`` (:311-471), ``grid_decoder_beam_search`` (:474-806),
``gnn_edge / gnn_node / gnn_mask_edge`` (:808-909), ``grid_emb / hidden2grid`` (:912-959),
``build_loss`` (:961-1040), ``add_div_penalty / gather_helper / wd_cost`` (:1197-1275),
``conv2d / softmax / exp_mask`` (:1333-1401) and ``Trainer.__init__`` (:1639-1717).  Only the
TensorFlow ops underneath are emulated (torch fp64).  This is what pins ``oracle/multiverse_ref.py``
and the committed goldens (tests/test_reference_exec_cpu.py, tests/golden/make_golden.py).

Only usable where ``/root/reference`` exists (this container, not the GPU box).
"""
from __future__ import annotations

import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("MVB_REFERENCE_ROOT", "/root/reference")
REFERENCE_FILE = os.path.join(REFERENCE_ROOT, "code", "pred_models.py")
_CACHE = {}


def available():
  return os.path.exists(REFERENCE_FILE)


def load():
  """(tf stand-in module, reference pred_models module).  ``tensorflow`` in sys.modules is swapped
  only for the duration of the import, so the product's own drop-in shim of the same name (or a
  real TensorFlow) is left untouched for every other test."""
  if "mods" in _CACHE:
    return _CACHE["mods"]
  saved = {k: v for k, v in sys.modules.items() if k == "tensorflow" or k.startswith("tensorflow.")}
  for k in saved:
    del sys.modules[k]
  sys.path.insert(0, HERE)
  try:
    import tensorflow as tf  # the stand-in: oracle/tf1_eager/tensorflow
    assert tf.__version__.endswith("eager-standin"), tf.__file__
    spec = importlib.util.spec_from_file_location("_multiverse_reference_pred_models", REFERENCE_FILE)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)       # executes the reference file unmodified
  finally:
    sys.path.remove(HERE)
    for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
      del sys.modules[k]
    sys.modules.update(saved)
  _CACHE["mods"] = (tf, ref)
  return tf, ref


def reference_config(cfg, tf, **kw):
  """argparse Namespace as code/train.py:25-138 / test.py / multifuture_inference.py:419-452 build
  it and pred_utils.process_args (:70-146) finishes it, from an oracle / synthetic config."""
  d = dict(vars(cfg))
  d.update(modelname="model", is_train=False, keep_prob=1.0, use_soft_grid_class=False,
           soft_grid=0, use_gt_grid=False, use_single_decoder=False, use_teacher_forcing=False,
           train_w_onehot=True, mask_grid_regression=False, grid_loss_weight=1.0,
           grid_reg_loss_weight=0.1, wd=0.001, init_lr=0.2, emb_lr=1.0, optimizer="adadelta",
           learning_rate_decay=0.95, num_epoch_per_decay=2.0, use_cosine_lr=False,
           clip_gradient_norm=10.0, train_num_examples=64, num_epochs=1)
  d.update(kw)
  assert d["activation_func"] in ("tanh", tf.nn.tanh)
  d["activation_func"] = tf.nn.tanh          # pred_utils.py:86-94 maps the string to tf.nn.tanh
  return SimpleNamespace(**d)


def _feed_values(model, config, feeds, is_train):
  """{placeholder index: value} keyed through the reference Model's OWN placeholder attributes, the
  way Model.get_feed_dict (:1042-1194) fills them."""
  n, t, tp = config.batch_size, config.obs_len, config.pred_len
  vals = {}
  put = lambda ph, v: vals.__setitem__(ph.placeholder_index, v)
  put(model.obs_length, np.full([n], t, np.int32))                     # :1057-1063
  put(model.pred_length, np.full([n], tp, np.int32))
  put(model.is_train, bool(is_train))                                  # :1065
  put(model.obs_scene, np.asarray(feeds["obs_scene"], np.int32))       # :1169-1183
  put(model.obs_scene_mask, np.ones([n, t], bool))
  put(model.scene_feat, np.asarray(feeds["scene_feat"], np.float64))   # :1174
  for j, (h, w) in enumerate(config.scene_grids):
    put(model.grid_pred_labels[j], np.zeros([n], np.int32))
    put(model.grid_pred_targets[j], np.zeros([n, 2]))
    put(model.grid_obs_labels[j], np.asarray(feeds["grid_obs_labels"][j], np.int32))   # :1186-1191
    put(model.grid_obs_targets[j], np.zeros([n, t, 2]))
    put(model.grid_obs_regress[j], np.asarray(feeds["grid_obs_regress"][j], np.float64))  # :1147
    if is_train:                                                       # :1149-1152
      put(model.grid_pred_labels_T[j], np.asarray(feeds["grid_pred_labels"][j], np.float64))
      put(model.grid_pred_regress[j], np.asarray(feeds["grid_pred_regress"][j], np.float64))
    else:                                                              # :1153-1162
      put(model.grid_pred_labels_T[j], np.zeros([n, tp]))
      put(model.grid_pred_regress[j], np.zeros([n, tp, h, w, 2]))
  return vals


def build(cfg, weights, feeds, is_train=False, with_trainer=False, **config_kw):
  """Construct the reference Model (and Trainer) eagerly on the given weights and feeds.
  Returns (tf, model, trainer or None, config)."""
  tf, ref = load()
  config = reference_config(cfg, tf, is_train=is_train, **config_kw)
  # pass 1 (probe): create the placeholders only, to learn which attribute is which placeholder
  with tf.building(weights, feed=None):
    probe = ref.Model.__new__(ref.Model)
    try:
      probe.__init__(config, config.modelname)
      raise AssertionError("probe pass was expected to stop at the first op of build_forward")
    except tf._StopBuild:
      pass
  vals = _feed_values(probe, config, feeds, is_train)
  # pass 2: the real construction = the whole forward pass (and loss), executed eagerly
  with tf.building(weights, feed=lambda idx, name, dtype, shape: vals[idx],
                   requires_grad=with_trainer):
    model = ref.get_model(config, 0)
    trainer = ref.Trainer(model, config) if with_trainer else None
  return tf, model, trainer, config


def _np(t):
  return t.numpy() if hasattr(t, "numpy") else t


def forward(cfg, weights, feeds, **config_kw):
  """Same return convention as oracle.multiverse_ref.forward: what Tester.step (:1761-1790) fetches."""
  tf, model, _, config = build(cfg, weights, feeds, is_train=False, **config_kw)
  out = dict(grid_pred_decoded=[_np(t) if not isinstance(t, list) else t for t in model.grid_pred_decoded],
             grid_pred_reg_decoded=[_np(t) if not isinstance(t, list) else t
                                    for t in model.grid_pred_reg_decoded],
             beam_outputs=None, variables=[v.op.name for v in tf.global_variables()])
  if model.beam_outputs is not None:
    out["beam_outputs"] = [_np(t) for t in model.beam_outputs]
  out["scene_convs"] = [_np(t) for t in model.scene_convs]
  return out


def train_step(cfg, weights, feeds, **config_kw):
  """Model(is_train) + Trainer: loss terms, clipped gradients, and the variables after ONE
  ``train_op`` (what Trainer.step :1719-1742 runs)."""
  tf, model, trainer, config = build(cfg, weights, feeds, is_train=True, with_trainer=True, **config_kw)
  tvars = tf.trainable_variables()
  grads = {v.op.name: (None if g is None else g.numpy().copy()) for v, g in zip(tvars, trainer.grads)}
  out = dict(loss=float(model.loss.numpy()), wd_loss=float(model.wd_loss.numpy()),
             pred_grid_loss=[float(l.numpy()) for l in model.pred_grid_loss], grads=grads)
  trainer.train_op.run()
  out["updated"] = {v.op.name: v.numpy().copy() for v in tvars}
  out["global_step"] = int(model.global_step.numpy())
  return out
