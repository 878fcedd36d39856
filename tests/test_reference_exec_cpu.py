# coding=utf-8
"""Pins the oracle on an EXECUTION of the reference's own graph code.

``oracle/tf1_eager`` imports the unmodified ``/root/reference/code/pred_models.py`` against an
eager, torch-fp64-backed stand-in for the TensorFlow-1.15 symbols it uses and runs
``Model.__init__ / build_forward / build_loss`` and ``Trainer.__init__`` as written.  These tests
assert that this run equals ``oracle/multiverse_ref.py`` (fp64, <=1e-12; ids identical) - the
wiring of code/pred_models.py:123-308, 311-471, 474-806, 808-909, 961-1040, 1197-1251, 1636-1717 is
therefore executed reference code, not a restatement; only the per-op TF semantics underneath stay
restated (and torch-anchored in test_oracle_cpu.py).

Needs /root/reference (this container); skipped on the GPU box, where the committed goldens
generated from the same execution (tests/golden/make_golden.py) carry the pin.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import multiverse_ref as R  # noqa: E402
from oracle.tf1_eager import run_reference as X  # noqa: E402

pytestmark = pytest.mark.skipif(not X.available(), reason="/root/reference is not mounted here")

TOL = 1e-12


def rel(a, b):
  return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-300))


def check_forward(cfg, w, f):
  out = X.forward(cfg, w, f)
  ref = R.forward(cfg, w, f, np.float64)
  assert set(out["variables"]) - {"global_step"} == set(w.keys())     # TF variable names, §8a
  for i in range(len(cfg.scene_grids)):
    if not cfg.use_grids[i]:
      assert out["grid_pred_decoded"][i] == [] and out["grid_pred_reg_decoded"][i] == []   # :170-171
      continue
    assert out["grid_pred_decoded"][i].shape == ref["grid_pred_decoded"][i].shape
    assert rel(out["grid_pred_decoded"][i], ref["grid_pred_decoded"][i]) < TOL
    assert rel(out["grid_pred_reg_decoded"][i], ref["grid_pred_reg_decoded"][i]) < TOL
    assert rel(out["scene_convs"][i], ref["scene_convs"][i]) < TOL
  if cfg.use_beam_search:
    lg, ids, lp = out["beam_outputs"]
    assert ids.dtype == np.int32 and np.array_equal(ids, ref["beam_outputs"][1])
    assert rel(lg, ref["beam_outputs"][0]) < TOL
    assert np.abs(lp - ref["beam_outputs"][2]).max() < 1e-11
  else:
    assert out["beam_outputs"] is None
  return out, ref


def test_reference_beam_k5_plain_equals_oracle_and_golden():
  """Coarse 18x9 grid, K=5 plain beam (no penalty, fix_num_timestep=0): the reference's own
  grid_decoder_beam_search + back-trace, and the committed golden made from it."""
  over, seed = cases.ROLLOUTS["beam_k5_plain"]
  cfg = R.default_config(**over)
  w, f = R.make_weights(cfg, seed), R.make_inputs(cfg, seed)
  out, _ = check_forward(cfg, w, f)
  g = np.load(os.path.join(ROOT, "tests", "golden", "rollout_beam_k5_plain.npz"))
  assert str(g["source"]) == "reference_exec"
  assert np.array_equal(g["beam_ids"], out["beam_outputs"][1])
  assert np.abs(g["beam_logprobs"] - out["beam_outputs"][2]).max() < 1e-11
  assert np.abs(g["logits_1"] - out["grid_pred_decoded"][1]).max() < 1e-6     # stored as fp32


def test_reference_beam_k20_diverse_equals_oracle():
  """K=20 diverse beam (gamma 0.01, first step's scores zeroed) - the multifuture_inference.py
  configuration (TESTING.md:84-93) - on the coarse grid with 3 trajectories."""
  cfg = R.default_config(batch_size=3, use_grids=[False, True], use_beam_search=True, beam_size=20,
                         diverse_beam=True, diverse_gamma=0.01, fix_num_timestep=1)
  check_forward(cfg, R.make_weights(cfg, 7), R.make_inputs(cfg, 7))


def test_reference_greedy_two_scale_equals_oracle():
  """Both scales, greedy class decoder with graph attention + regression decoder (test.py path)."""
  cfg = R.default_config(batch_size=2, scene_h=24, scene_w=16)
  assert cfg.scene_grids == [(12, 8), (6, 4)]
  check_forward(cfg, R.make_weights(cfg, 8), R.make_inputs(cfg, 8))


def test_reference_ragged_pred_length_and_no_gnn():
  """use_gnn off (the reference then hands the raw state to the cell)."""
  cfg = R.default_config(batch_size=2, use_grids=[False, True], use_gnn=False)
  check_forward(cfg, R.make_weights(cfg, 9), R.make_inputs(cfg, 9))


def test_reference_training_step_equals_oracle():
  """Model.build_loss + Trainer.__init__ executed: total / class / Huber / wd losses, the clipped
  gradient of every trainable variable (tf.gradients -> clip_by_value +-10, :1698-1705) and one
  Adadelta train_op (:1672,:1716) against the oracle's torch-autograd restatement and the
  closed-form update the CUDA optimizer kernel implements."""
  from oracle import multiverse_ref_torch as RT
  kw = dict(grid_loss_weight=1.0, grid_reg_loss_weight=0.1, wd=0.001)
  cfg = R.default_config(batch_size=2, use_grids=[False, True], **kw)
  w, f = R.make_weights(cfg, 10), R.make_inputs(cfg, 10)
  got = X.train_step(cfg, w, f, **kw)
  tot, losses, wd, grads = RT.loss_and_grads(cfg, w, f)
  assert abs(got["loss"] - tot) < 1e-11 * abs(tot)
  assert abs(got["wd_loss"] - wd) < 1e-12 * wd
  assert np.abs(np.array(got["pred_grid_loss"]) - np.array(losses)).max() < 1e-11
  assert set(got["grads"]) == set(w.keys())
  lr = 0.2 * 1.0 * 0.95 ** 0        # init_lr * emb_lr * decay^(floor(step/decay_steps)), step 0
  for k, g in grads.items():
    gc = np.clip(g, -10.0, 10.0)
    assert got["grads"][k] is not None, k
    assert np.abs(got["grads"][k] - gc).max() <= 1e-10 * max(np.abs(gc).max(), 1e-30), k
    acc = 0.05 * gc * gc                                  # rho=.95, zero slots, eps=1e-8
    upd = np.sqrt(1e-8) / np.sqrt(acc + 1e-8) * gc
    want = w[k].astype(np.float64) - lr * upd
    assert np.abs(got["updated"][k] - want).max() < 1e-12, k
  assert got["global_step"] == 1
