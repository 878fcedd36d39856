# coding=utf-8
"""Executes the UNMODIFIED ``/root/reference/SimAug/code/pred_models.py`` on the eager TF-1.15 stand-in of this
directory.  TEST INFRASTRUCTURE.

What runs is the reference's own code: ``Model.__init__`` with ``multiview_train`` (placeholders :191-267, the
augmentation branch :304-310), ``multiview_augmentation`` (:346-541: tiling over the views, ``one_step_attack`` with
``build_tower`` + ``tf.gradients`` w.r.t. the input features, per-view losses, ``multiview_exp`` selection, focal
weights, Beta mixup), ``white_box_attack`` (:60-170: random targets, FGSM / the PGD ``tf.while_loop``, mixup - see
``adversarial()`` below), ``build_tower`` (:544-, incl. the mixed observed class maps of experiment 3) and ``build_loss``
(:1340-, incl. the mixed labels and ``double_weighting``).  Only the TensorFlow ops underneath are emulated (torch
fp64, autograd for ``tf.gradients``); TensorFlow's random streams cannot be reproduced, so the two draws on this path
are injected: the Beta sample (``beta``) and - with ``adv_start_from_clean_prob >= 1`` - no uniform start noise.

Only usable where ``/root/reference`` exists (this container, not the GPU box).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

from . import run_reference as RR

HERE = os.path.dirname(os.path.abspath(__file__))
SIMAUG_FILE = os.path.join(RR.REFERENCE_ROOT, "SimAug", "code", "pred_models.py")
_CACHE = {}


def available():
  return os.path.exists(SIMAUG_FILE)


def load():
  """(tf stand-in, SimAug pred_models module); `tensorflow` in sys.modules is swapped only during the import."""
  if "mods" in _CACHE:
    return _CACHE["mods"]
  tf, _ = RR.load()                      # the same stand-in module object the base harness uses
  saved = {k: v for k, v in sys.modules.items() if k == "tensorflow" or k.startswith("tensorflow.")}
  for k in saved:
    del sys.modules[k]
  sys.modules["tensorflow"] = tf
  try:
    spec = importlib.util.spec_from_file_location("_multiverse_reference_simaug_pred_models", SIMAUG_FILE)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)         # executes the reference file unmodified
  finally:
    for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
      del sys.modules[k]
    sys.modules.update(saved)
  _CACHE["mods"] = (tf, ref)
  return tf, ref


def simaug_config(cfg, tf, m, exp, eps, **kw):
  """The argparse Namespace SimAug/code/train.py builds, from an oracle / synthetic config."""
  c = RR.reference_config(cfg, tf, is_train=True)
  d = dict(vars(c))
  d.update(norm_input=False, norm_feat=False, adv_train=False, multiview_train=True, multiview_max_num=m,
           multiview_exp=exp, adv_epsilon=eps, adv_step_size=eps / 4, adv_num_iter=1, adv_use_fgsm=True,
           adv_start_from_clean_prob=1.0, use_mixup=False, mixup_alpha=1.0, mixup_mix_adv=False,
           multiview_max_weight_for_first=True, multiview_use_adv_for_loss=False, multiview_random=False,
           fl_gamma=2.0, double_weighting=False, standard_aug=False, scene_conv_dim=64)
  d.update(kw)
  from types import SimpleNamespace
  return SimpleNamespace(**d)


def _feed_values(model, config, feeds, extra):
  """{placeholder index: value}, keyed through the SimAug Model's own placeholder attributes the way its
  get_feed_dict (:1440-1560) fills them.  scene_feat is fed as a torch tensor that requires grad, so that
  tf.gradients(loss, <features gathered from it>) has a path."""
  import torch
  n, t, tp, m = config.batch_size, config.obs_len, config.pred_len, config.multiview_max_num
  vals = {}
  put = lambda ph, v: vals.__setitem__(ph.placeholder_index, v)
  put(model.obs_length, np.full([n], t, np.int32))
  put(model.pred_length, np.full([n], tp, np.int32))
  put(model.is_train, True)
  put(model.obs_scene, np.asarray(feeds["obs_scene"], np.int32))
  put(model.scene_feat, torch.tensor(np.asarray(feeds["scene_feat"], np.float64), requires_grad=True))
  if extra is not None:
    put(model.obs_scene_extra, np.asarray(extra["obs_scene_extra"], np.int32))
  for j, (h, w) in enumerate(config.scene_grids):
    put(model.grid_obs_labels[j], np.asarray(feeds["grid_obs_labels"][j], np.int32))
    put(model.grid_obs_regress[j], np.asarray(feeds["grid_obs_regress"][j], np.float64))
    put(model.grid_pred_labels_T[j], np.asarray(feeds["grid_pred_labels"][j], np.float64))
    put(model.grid_pred_regress[j], np.asarray(feeds["grid_pred_regress"][j], np.float64))
    if extra is None:
      continue
    # the other views (:250-267): labels of their own; the regression arrays are tiled from the main view
    obs_l = extra["grid_obs_labels_extra"][j]
    pred_l = extra["grid_pred_labels_extra"][j]
    put(model.grid_obs_labels_extra[j], np.zeros([n, m, t], np.int32) if obs_l is None else np.asarray(obs_l, np.int32))
    put(model.grid_pred_labels_T_extra[j], np.zeros([n, m, tp]) if pred_l is None else np.asarray(pred_l, np.float64))
    put(model.grid_pred_regress_extra[j], np.zeros([n, m, tp, h, w, 2]))
    put(model.grid_obs_regress_extra[j], np.zeros([n, m, t, h, w, 2]))
  return vals


def multiview(cfg, weights, feeds, extra, m, exp, eps, beta, with_trainer=False, int_draws=(), **config_kw):
  """Builds the SimAug Model in training mode with multiview_train on the given weights and feeds - which executes
  multiview_augmentation, the training tower on the augmented features and build_loss - and returns
  dict(adv_final [N*T,SH,SW,SC] (the return value of multiview_augmentation, recorded by a wrapper), beta_weight,
  focal_loss_weight / selected_extra_indices (experiment 3), loss, losses [cls, reg], grads {name: array} and the
  stand-in's variable list)."""
  tf, ref = load()
  config = simaug_config(cfg, tf, m, exp, eps, **config_kw)
  rec = {}
  orig = ref.Model.multiview_augmentation

  def recording(self, obs_scene):
    out = orig(self, obs_scene)
    rec["adv_final"] = out
    return out

  ref.Model.multiview_augmentation = recording
  try:
    with tf.building(weights, feed=None):
      probe = ref.Model.__new__(ref.Model)
      try:
        probe.__init__(config, config.modelname)
        raise AssertionError("probe pass was expected to stop at the first op after the placeholders")
      except tf._StopBuild:
        pass
    vals = _feed_values(probe, config, feeds, extra)
    with tf.building(weights, feed=lambda idx, name, dtype, shape: vals[idx], requires_grad=True) as state:
      state.beta_sample = beta
      draws = [np.asarray(d) for d in int_draws]

      def uniform_hook(shape, minval, maxval, dtype):
        # float draws: get_start_adv (:353) draws its start noise before it looks at adv_start_from_clean_prob; with
        # the probability at 1 the noise is never used - zeros.  Integer draws (experiment 2's view indices, :474-481):
        # taken in order from `int_draws`.
        if dtype is not None and "int" in str(dtype):
          v = draws.pop(0)
          assert tuple(v.shape) == tuple(shape) and v.min() >= minval and v.max() < maxval
          return v
        assert config.adv_start_from_clean_prob >= 1.0, "start noise would be used: not reproducible"
        return np.zeros(shape)
      state.uniform_hook = uniform_hook
      model = ref.get_model(config, 0)
      out = dict(adv_final=RR._np(rec["adv_final"]), beta_weight=float(RR._np(model.beta_weight)),
                 loss=float(RR._np(model.loss)), losses=[float(RR._np(l)) for l in model.pred_grid_loss])
      if exp == 3:
        out["focal_loss_weight"] = RR._np(model.focal_loss_weight)
        out["selected_extra_indices"] = RR._np(model.selected_extra_indices)
      if with_trainer:
        import torch
        names = [v for v in tf.global_variables() if getattr(v, "trainable", False)]
        gs = torch.autograd.grad(tf._raw(model.loss), [tf._raw(v) for v in names], allow_unused=True)
        out["grads"] = {v.op.name: (np.zeros(tuple(tf._raw(v).shape)) if g is None else g.numpy())
                        for v, g in zip(names, gs)}
  finally:
    ref.Model.multiview_augmentation = orig
  return out


def adversarial(cfg, weights, feeds, eps, target_offsets, fgsm=True, step_size=None, num_iter=1, mixup_beta=None,
                **config_kw):
  """Builds the SimAug Model in training mode with adv_train (white_box_attack, :60-170, called at :289-301) and
  returns dict(adv_final [N*T,SH,SW,SC], target_label [N,Tp], loss, losses).  The random target offsets
  (create_random_target's tf.random_uniform ints, :66-72) are injected (`target_offsets` int [N,Tp] in [1, h*w));
  the start is the clean input (adv_start_from_clean_prob = 1); `mixup_beta` switches use_mixup on with that draw."""
  tf, ref = load()
  config = simaug_config(cfg, tf, 1, 1, eps, multiview_train=False, adv_train=True, adv_use_fgsm=bool(fgsm),
                         adv_step_size=step_size if step_size is not None else eps / 4, adv_num_iter=num_iter,
                         use_mixup=mixup_beta is not None, **config_kw)
  rec = {}
  orig = ref.white_box_attack

  def recording(*a, **kw):
    out = orig(*a, **kw)
    rec["adv_final"], rec["target_label"] = out
    return out

  ref.white_box_attack = recording
  try:
    with tf.building(weights, feed=None):
      probe = ref.Model.__new__(ref.Model)
      try:
        probe.__init__(config, config.modelname)
        raise AssertionError("probe pass was expected to stop at the first op after the placeholders")
      except tf._StopBuild:
        pass
    vals = _feed_values(probe, config, feeds, None)
    with tf.building(weights, feed=lambda idx, name, dtype, shape: vals[idx], requires_grad=True) as state:
      state.beta_sample = mixup_beta
      draws = [np.asarray(target_offsets)]

      def uniform_hook(shape, minval, maxval, dtype):
        if dtype is not None and "int" in str(dtype):
          v = draws.pop(0)
          assert tuple(v.shape) == tuple(shape) and v.min() >= minval and v.max() < maxval
          return v
        return np.zeros(shape)            # start noise: drawn (:76-79) but unused with adv_start_from_clean_prob = 1
      state.uniform_hook = uniform_hook
      model = ref.get_model(config, 0)
      out = dict(adv_final=RR._np(rec["adv_final"]), target_label=RR._np(rec["target_label"]),
                 loss=float(RR._np(model.loss)), losses=[float(RR._np(l)) for l in model.pred_grid_loss])
  finally:
    ref.white_box_attack = orig
  return out
