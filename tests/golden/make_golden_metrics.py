# coding=utf-8
"""Golden vectors of the multi-future evaluation metrics, produced by THE REFERENCE'S OWN numpy functions imported
from /root/reference/code (they need no TensorFlow): get_min (multifuture_eval_trajs.py:16-21) inside the loop body of
:41-78, and softmax / get_hw_prob / compute_nll / xys_to_indexes (multifuture_eval_trajs_prob.py:20-60) inside the loop
body of :170-197.  Run in the container that has /root/reference:  python tests/golden/make_golden_metrics.py"""
import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np

REF = os.path.join(os.environ.get("MVB_REFERENCE_ROOT", "/root/reference"), "code")
OUT = os.path.dirname(os.path.abspath(__file__))


def load(name):
  spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, name + ".py"))
  m = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(m)
  return m


def main():
  ev = load("multifuture_eval_trajs")
  evp = load("multifuture_eval_trajs_prob")
  rng = np.random.default_rng(7)
  # ---- minADE / minFDE ---------------------------------------------------------------------------------------
  n, k, tp, g, tg = 6, 20, 14, 4, 12
  pred = (rng.uniform(0, 1920, size=(n, k, tp, 2))).astype(np.float32)
  gt = (pred[:, rng.integers(0, k, size=g)].transpose(0, 1, 2, 3)[:, :, :tg] +
        rng.normal(0, 30, size=(n, g, tg, 2))).astype(np.float32)
  gt_len = rng.choice([0, 5, 9, 12], size=(n, g)).astype(np.int32)
  gt_len[0, 0] = 12
  pred[1, 3] = pred[1, 7]                        # two identical predictions: the first index must win
  ade_err = np.zeros((n, g, tg)); ade_idx = -np.ones((n, g), np.int32)
  fde = np.zeros((n, g)); fde_idx = -np.ones((n, g), np.int32)
  for i in range(n):
    prediction = [pred[i, kk] for kk in range(k)]                      # what the pickle holds per trajectory
    for j in range(g):
      if gt_len[i, j] == 0:
        continue
      gt_traj = np.array([list(map(float, p)) for p in gt[i, j, :gt_len[i, j]]])   # python floats, like the .p files
      pred_len = len(gt_traj)
      this_ade_errors, this_fde_errors = [], []
      for pred_out in prediction:                                     # multifuture_eval_trajs.py:61-67
        diff = gt_traj - pred_out[:pred_len]
        diff = diff**2
        diff = np.sqrt(np.sum(diff, axis=1))
        this_ade_errors.append(diff.tolist())
        this_fde_errors.append([diff[-1]])
      min_ade_errors, min_ade_traj_idx = ev.get_min(this_ade_errors)
      min_fde_errors, min_fde_traj_idx = ev.get_min(this_fde_errors)
      ade_err[i, j, :pred_len] = min_ade_errors; ade_idx[i, j] = min_ade_traj_idx
      fde[i, j] = min_fde_errors[0]; fde_idx[i, j] = min_fde_traj_idx
  # ---- NLL ------------------------------------------------------------------------------------------------------
  args = SimpleNamespace(scene_h=18, scene_w=32, video_h=1080, video_w=1920)
  args.w_gap = args.video_w * 1.0 / args.scene_w; args.h_gap = args.video_h * 1.0 / args.scene_h
  n2, v, tpb = 3, args.scene_h * args.scene_w, 6
  beams = (rng.standard_normal((n2, 1, k, tpb, v)) * 3).astype(np.float32)
  logprobs = (-np.abs(rng.standard_normal((n2, 1, k))) * 4).astype(np.float32)
  time_list = [0, 1, 2, 3, 4]
  gxy = rng.uniform([0, 0], [1920, 1080], size=(n2, g, tpb, 2))
  gxy[0, 0, 0] = [0.0, 0.0]                       # ceil(0) = 0 -> cell 0 (the reference's special case)
  g_len = rng.choice([0, 2, 4, 6], size=(n2, g)); g_len[:, 0] = 6
  nll = np.zeros((n2, len(time_list))); cnt = np.zeros((n2, len(time_list)), np.int32)
  gt_idx = -np.ones((n2, len(time_list), g), np.int32)
  for i in range(n2):
    lp = evp.softmax(np.squeeze(logprobs[i]))                         # :177-179
    bm = evp.softmax(np.squeeze(beams[i]), axis=-1)
    grid_probs = [evp.get_hw_prob(bm, lp, t) for t in time_list]
    for jj, timestep in enumerate(time_list):
      gt_xys, which = [], []
      for f in range(g):
        if g_len[i, f] <= timestep:
          continue
        gt_xys.append(list(gxy[i, f, timestep])); which.append(f)
      if not gt_xys:
        continue
      idxs = evp.xys_to_indexes(np.asarray(gt_xys), args)
      nll[i, jj] = evp.compute_nll(grid_probs[jj], idxs)
      cnt[i, jj] = len(idxs)
      gt_idx[i, jj, which] = idxs
  np.savez_compressed(os.path.join(OUT, "metrics.npz"), pred=pred, gt=gt, gt_len=gt_len, ade_err=ade_err,
                      ade_idx=ade_idx, fde=fde, fde_idx=fde_idx, beams=beams[:, 0], logprobs=logprobs[:, 0],
                      steps=np.asarray(time_list, np.int32), gt_xy=gxy, gt_idx=gt_idx, nll=nll, count=cnt,
                      grid=np.asarray([args.scene_h, args.scene_w, args.video_h, args.video_w]))
  print("wrote metrics.npz: ADE mean %.3f FDE mean %.3f NLL %s" % (
      ade_err[gt_len > 0].sum() / gt_len.sum(), fde[gt_len > 0].mean(), nll.mean(0)))


if __name__ == "__main__":
  main()
