# coding=utf-8
"""GPU probe: whole forward (greedy two-scale, then beam) vs the numpy oracle."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiverse_b200.engine import ConvRNNEngine
from oracle import multiverse_ref as R

dev = torch.device("cuda:0")


def to_dev(feeds):
  return dict(scene_feat=torch.from_numpy(feeds["scene_feat"]).to(dev),
              obs_scene=torch.from_numpy(feeds["obs_scene"]).to(dev),
              grid_obs_labels=[torch.from_numpy(a).to(dev) for a in feeds["grid_obs_labels"]],
              grid_obs_regress=[torch.from_numpy(a).to(dev) for a in feeds["grid_obs_regress"]])


def rel(a, b):
  return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def greedy(n, seed, planes):
  cfg = R.default_config(batch_size=n)
  w = R.make_weights(cfg, seed)
  f = R.make_inputs(cfg, seed)
  ref = R.forward(cfg, w, f, np.float64)
  eng = ConvRNNEngine(cfg, {k: torch.from_numpy(v) for k, v in w.items()}, dev, planes)
  out = eng.forward(to_dev(f))
  torch.cuda.synchronize()
  for i in range(2):
    a = out["grid_pred_decoded"][i].cpu().numpy().astype(np.float64)
    b = ref["grid_pred_decoded"][i]
    ia = a.reshape(n, 12, -1).argmax(-1); ib = b.reshape(n, 12, -1).argmax(-1)
    s = np.sort(b.reshape(n, 12, -1), -1); margin = (s[..., -1] - s[..., -2]).min()
    ra = out["grid_pred_reg_decoded"][i].cpu().numpy().astype(np.float64)
    rb = ref["grid_pred_reg_decoded"][i]
    print("greedy n=%d P=%d scale %d: logits rel %.3e  argmax equal %s (min margin %.2e)  reg rel %.3e"
          % (n, planes, i, rel(a, b), bool((ia == ib).all()), margin, rel(ra, rb)), flush=True)


def beam(n, k, seed, planes, diverse):
  cfg = R.default_config(batch_size=n, use_grids=[True, False], use_beam_search=True, beam_size=k,
                         diverse_beam=diverse, diverse_gamma=0.01, fix_num_timestep=1)
  w = R.make_weights(cfg, seed)
  f = R.make_inputs(cfg, seed)
  ref = R.forward(cfg, w, f, np.float64)
  eng = ConvRNNEngine(cfg, {kk: torch.from_numpy(v) for kk, v in w.items()}, dev, planes)
  out = eng.forward(to_dev(f))
  torch.cuda.synchronize()
  lg, ids, lp = [t.cpu().numpy() for t in out["beam_outputs"]]
  rl, rids, rlp = ref["beam_outputs"]
  print("beam n=%d K=%d diverse=%s P=%d: ids equal %s (%d/%d)  logits rel %.3e  logprob abs %.3e  reg rel %.3e"
        % (n, k, diverse, planes, bool((ids == rids).all()), int((ids == rids).sum()), ids.size,
           rel(lg.astype(np.float64), rl), float(np.abs(lp - rlp).max()),
           rel(out["grid_pred_reg_decoded"][0].cpu().numpy().astype(np.float64), ref["grid_pred_reg_decoded"][0])), flush=True)


if __name__ == "__main__":
  greedy(2, 0, 2)
  greedy(3, 1, 2)
  beam(2, 5, 0, 2, False)
  beam(2, 20, 1, 2, True)
