# coding=utf-8
"""SimAug's white-box attack on the scene input, on the B200 engine (first part).

Mirrors ``white_box_attack`` of ``SimAug/code/pred_models.py:60-170`` - same argument meaning, same config
attributes (``adv_epsilon, adv_step_size, adv_num_iter, adv_start_from_clean_prob, adv_use_fgsm, use_mixup,
mixup_alpha, mixup_mix_adv, use_grids, scene_grids``) - with the reference's ``model_func`` + ``tf.gradients(
classification_loss, adv_input)`` replaced by ``TrainEngine.loss_and_grads(dscene_out=...)``: the ordinary BPTT of
this library with the scene CNN's input gradient switched on (``mvb_scene_conv_bwd`` ``din`` of the first
convolution).  The update itself is ``mvb_adv_step`` / ``mvb_mix``.  Random draws (start noise, random target
offsets, the Beta mixup weight) come from a numpy Generator - TensorFlow's random streams cannot be reproduced.

Not built yet: SimAug's multi-view mixup (``multiview_augmentation``, :346-541) and its ``build_tower`` Model
variant (:544), which reuse the same kernels and this input gradient.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def create_random_target(label, max_class, rng):
  """:66-72: label + uniform{1..max_class-1} modulo max_class - always a WRONG class."""
  off = rng.integers(1, max_class, size=label.shape)
  return ((label.astype(np.int64) + off) % max_class).astype(np.int32)


def scene_input_grad(engine, feeds, target_label, scale_idx):
  """d sum(sparse CE(target)) / d scene_feat up to a positive factor (the attack only uses its sign):
  SimAug/code/pred_models.py:96-115.  feeds: a TrainEngine feed dict whose scene_feat [F,SH,SW,SC] the gradient
  is taken for; target_label int32 [N,Tp] replaces grid_pred_labels of the attacked scale."""
  f = dict(feeds)
  f["grid_pred_labels"] = list(feeds["grid_pred_labels"])
  f["grid_pred_labels"][scale_idx] = target_label
  g = torch.zeros_like(feeds["scene_feat"], dtype=torch.float32)
  engine.loss_and_grads(f, dscene_out=g, cls_weight=1.0, reg_weight=0.0)
  return g


def white_box_attack(engine, feeds, label, config, rng=None, norm_feat=False):
  """SimAug/code/pred_models.py:60-170.  `feeds["scene_feat"]` is the clean input feature (one frame per (n,t)
  row like the reference's [N*T_obs,SH,SW,SC], or unique frames + obs_scene indices: the gradient then sums over
  the rows that share a frame).  label: numpy int [N,T_pred] of the attacked scale.  Returns (adv_final fp32 tensor
  shaped like scene_feat, target_label numpy int32 [N,T_pred])."""
  rng = rng or np.random.default_rng()
  assert not norm_feat, "norm_feat (softmax of the start point) is not implemented"
  scale_idx = list(config.use_grids).index(True)                      # :128
  h, w = config.scene_grids[scale_idx]
  dev = engine.device
  x = feeds["scene_feat"].float().contiguous()
  target_label = create_random_target(np.asarray(label), int(h * w), rng)      # :132-133
  tl = torch.from_numpy(target_label).to(dev)
  eps = float(config.adv_epsilon)

  def get_start_adv():                                                # :75-89
    if config.adv_start_from_clean_prob >= 1.0:
      return x.clone()
    noise = torch.from_numpy(rng.uniform(-eps, eps, size=tuple(x.shape)).astype(np.float32)).to(dev)
    if config.adv_start_from_clean_prob > 0:
      noise = noise * float(rng.uniform() > config.adv_start_from_clean_prob)
    return x + noise

  def one_step_attack(adv):                                           # :91-124
    g = scene_input_grad(engine, dict(feeds, scene_feat=adv), tl, scale_idx)
    out = torch.empty_like(adv)
    ops.adv_step(x, adv, g, out, eps, eps if config.adv_use_fgsm else float(config.adv_step_size))
    return out

  adv = one_step_attack(get_start_adv())
  if not config.adv_use_fgsm:                                          # PGD: adv_num_iter steps in all (:145-155)
    for _ in range(int(config.adv_num_iter) - 1):
      adv = one_step_attack(adv)
  if getattr(config, "use_mixup", False):                             # :157-170
    weight = float(rng.beta(config.mixup_alpha, config.mixup_alpha))
    out = torch.empty_like(adv)
    if getattr(config, "mixup_mix_adv", False):
      assert config.adv_use_fgsm and config.adv_start_from_clean_prob < 1.0
      adv2 = one_step_attack(get_start_adv())
      ops.mix(adv2, adv, out, weight)
    else:
      ops.mix(x, adv, out, weight)
    adv = out
  return adv, target_label
