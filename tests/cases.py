# coding=utf-8
"""Seeded unit-case inputs shared by tests/golden/make_golden.py and the parity tests.

Inputs are regenerated from the seed (the 12 MB ConvLSTM kernels are not committed); every golden
file stores a checksum of the regenerated inputs so a drifting generator is detected, plus the
oracle's fp64 outputs."""
from __future__ import annotations

import math

import numpy as np

CELL_CASES = {
    # name: (ns, h, w, cx, x_scale, seed)
    "dec_cx32": (2, 6, 5, 32, 1.0, 11),
    "enc_class_cx64": (3, 5, 7, 64, 1.0, 12),
    "enc_reg_cx2": (2, 7, 4, 2, 600.0, 13),      # raw pixel offsets: large magnitude
    "tile_edge": (5, 9, 5, 32, 1.0, 14),          # 300 halo rows: crosses a 128-row M tile
}


def checksum(*arrays):
  return float(sum(float(np.sum(np.asarray(a, dtype=np.float64))) for a in arrays))


def cell_case(name):
  ns, h, w, cx, xs, seed = CELL_CASES[name]
  rng = np.random.default_rng(seed)
  ch = 256
  lim = math.sqrt(6.0 / (9 * (cx + ch) + 9 * 4 * ch))
  kernel = rng.uniform(-lim, lim, size=(3, 3, cx + ch, 4 * ch)).astype(np.float32)
  biases = (rng.standard_normal(4 * ch) * 0.1).astype(np.float32)
  x = (rng.standard_normal((ns, h, w, cx)) * xs).astype(np.float32)
  hh = np.tanh(rng.standard_normal((ns, h, w, ch))).astype(np.float32)
  c = rng.standard_normal((ns, h, w, ch)).astype(np.float32)
  return dict(x=x, h=hh, c=c, kernel=kernel, biases=biases)


def gnn_case(seed=21, ns=3, h=5, w=4):
  rng = np.random.default_rng(seed)
  return dict(h=np.tanh(rng.standard_normal((ns, h, w, 256))).astype(np.float32),
              scene=np.tanh(rng.standard_normal((ns, h, w, 64))).astype(np.float32))


def head_case(seed=31, ns=3, h=6, w=5, e=32):
  rng = np.random.default_rng(seed)
  return dict(h=np.tanh(rng.standard_normal((ns, h, w, 256))).astype(np.float32),
              Wo1=(rng.standard_normal((3, 3, 256, 1)) * 0.1).astype(np.float32),
              Wo2=(rng.standard_normal((3, 3, 256, 2)) * 0.1).astype(np.float32),
              We1=(rng.standard_normal((3, 3, 1, e)) * 0.5).astype(np.float32),
              We2=(rng.standard_normal((3, 3, 2, e)) * 0.5).astype(np.float32),
              be=(rng.standard_normal(e) * 0.2).astype(np.float32))


def beam_case(seed=41, n=3, b=5, v=30):
  rng = np.random.default_rng(seed)
  logits = (rng.standard_normal((n, b, v)) * 2).astype(np.float32)
  logits[0, 1, 7] = logits[0, 1, 3]          # exact tie inside a row (rank / top-k tie-break)
  score = (-np.abs(rng.standard_normal((n, b)))).astype(np.float32)
  return dict(logits=logits, score=score)


def scene_case(seed=51, f=3, sh=12, sw=10, sc=11):
  rng = np.random.default_rng(seed)
  seg = rng.integers(0, sc, size=(f, sh, sw))
  feat = np.eye(sc, dtype=np.float32)[seg]
  return dict(scene_feat=feat,
              W1=(rng.standard_normal((3, 3, sc, 64)) * 0.2).astype(np.float32),
              b1=(rng.standard_normal(64) * 0.1).astype(np.float32),
              W2=(rng.standard_normal((3, 3, 64, 64)) * 0.1).astype(np.float32),
              b2=(rng.standard_normal(64) * 0.1).astype(np.float32),
              obs_scene=rng.integers(0, f, size=(4, 8)).astype(np.int32))


ROLLOUTS = {
    # name: config overrides (oracle.default_config), seed
    "greedy_two_scale": (dict(batch_size=2), 0),
    "beam_k20_diverse": (dict(batch_size=2, use_grids=[True, False], use_beam_search=True,
                              beam_size=20, diverse_beam=True, diverse_gamma=0.01,
                              fix_num_timestep=1), 1),
    "beam_k5_plain": (dict(batch_size=2, use_grids=[False, True], use_beam_search=True,
                           beam_size=5, diverse_beam=False, fix_num_timestep=0), 2),
    "greedy_native_18x32": (dict(batch_size=2, scene_h=36, scene_w=64, use_grids=[True, False]), 3),
}


# At-size rollouts (the shapes bench.py actually runs: CTA-pair cell kernel, x-fold + row_map, fan-out with K = 20).
# Their goldens hold REDUCED statistics of the fp64 oracle run (tests/golden/make_golden_atsize.py), not the tensors.
ROLLOUTS_ATSIZE = {
    "beam_k20_n16": (dict(batch_size=16, use_grids=[True, False], use_beam_search=True, beam_size=20,
                          diverse_beam=True, diverse_gamma=0.01, fix_num_timestep=1), 21),
    "greedy_two_scale_n64": (dict(batch_size=64), 22),
    "beam_k20_native_18x32_n4": (dict(batch_size=4, scene_h=36, scene_w=64, use_grids=[True, False],
                                      use_beam_search=True, beam_size=20, diverse_beam=True, diverse_gamma=0.01,
                                      fix_num_timestep=1), 23),
}


def rollout_stats(cfg, out):
  """Size-reduced view of a forward() result (numpy arrays shaped like the reference fetches): what the at-size
  goldens store and what the GPU run is reduced to before the comparison."""
  import numpy as np
  st = {}
  n, tp = cfg.batch_size, cfg.pred_len
  for i, (h, w) in enumerate(cfg.scene_grids):
    if not cfg.use_grids[i]:
      continue
    v = h * w
    lg = np.asarray(out["grid_pred_decoded"][i], np.float64).reshape(n, tp, v)
    reg = np.asarray(out["grid_pred_reg_decoded"][i], np.float64).reshape(n, tp, v, 2)
    am = lg.argmax(-1)
    srt = np.sort(lg, -1)
    cells = (np.arange(8) * 79 + 3) % v
    st["argmax_%d" % i] = am.astype(np.int32)
    st["margin_%d" % i] = srt[..., -1] - srt[..., -2]
    st["lg_max_%d" % i] = srt[..., -1]
    st["lg_mean_%d" % i] = lg.mean(-1)
    st["lg_at_%d" % i] = lg[..., cells]
    st["reg_at_argmax_%d" % i] = np.take_along_axis(reg, am[..., None, None], 2)[:, :, 0]
    st["reg_mean_%d" % i] = reg.mean(2)
    st["reg_at_%d" % i] = reg[:, :, cells]
  if out.get("beam_outputs") is not None:
    blg, ids, lp = out["beam_outputs"]
    blg = np.asarray(blg, np.float64)
    st["beam_ids"] = np.asarray(ids, np.int32)
    st["beam_logprobs"] = np.asarray(lp, np.float64)
    st["beam_lg_max"] = blg.max(-1)
    st["beam_lg_mean"] = blg.mean(-1)
  return st


def simaug_case():
  """Seeded inputs of the SimAug multi-view golden (tests/golden/make_golden_simaug.py): 2 samples, 3 other views,
  one scale (18x9), soft scene features in (-1, 1).  Returns (synthetic config, weights, feeds, extra-view feeds,
  spec)."""
  from multiverse_b200 import synthetic
  n, m = 2, 3
  # gnn_scene_in_greedy=False: SimAug's gnn_edge feeds the scene features to the attention only in the beam decoder
  conf = dict(batch_size=n, use_grids=[False, True], grid_loss_weight=1.0, grid_reg_loss_weight=0.1, wd=0.001,
              gnn_scene_in_greedy=False)
  cfg = synthetic.make_config(clip_gradient_norm=10.0, **conf)
  w = synthetic.make_weights(cfg, 41)
  f = synthetic.make_feeds(cfg, n, 41, with_pred=True)
  rng = np.random.default_rng(9)
  f["scene_feat"] = np.clip(f["scene_feat"] * 0.8 + rng.uniform(-0.1, 0.1, f["scene_feat"].shape), -1, 1).astype(np.float32)
  hw = 18 * 9
  extra = dict(grid_pred_labels_extra=[None, rng.integers(0, hw, size=(n, m, cfg.pred_len)).astype(np.int32)],
               grid_obs_labels_extra=[None, rng.integers(0, hw, size=(n, m, cfg.obs_len)).astype(np.int32)],
               obs_scene_extra=rng.integers(0, f["scene_feat"].shape[0], size=(n, m, cfg.obs_len)).astype(np.int32))
  return cfg, w, f, extra, dict(n=n, m=m, eps=0.1, beta_draw=0.3, config=conf)


def grad_sample_stride(size):
  """Stride of the gradient samples stored in tests/golden/simaug_multiview.npz (<= 2048 entries per variable)."""
  return max(1, -(-size // 2048))


ADV_SAMPLE_STRIDE = 7      # every 7th element of the augmented features is stored in simaug_multiview.npz
