# coding=utf-8
"""At-size goldens: the oracle (fp64; torch port of oracle/multiverse_ref.py, which the small-size tests pin to the
numpy oracle and to the executed reference) run ONCE here at the batch sizes the benchmark's kernel variants need
(CTA-pair cell kernel from 54 sample rows of 36x18, K = 20 fan-out, x-fold + row_map), reduced to statistics
(tests/cases.py rollout_stats) so that the committed files stay small.  Takes ~10 minutes on 8 cores:
    python tests/golden/make_golden_atsize.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import multiverse_ref as R  # noqa: E402
from oracle import multiverse_ref_torch as RT  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main(only=None):
  torch.set_num_threads(os.cpu_count())
  for name, (over, seed) in cases.ROLLOUTS_ATSIZE.items():
    if only and name not in only:
      continue
    cfg = R.default_config(**over)
    w = R.make_weights(cfg, seed); f = R.make_inputs(cfg, seed)
    wt = {k: torch.from_numpy(np.ascontiguousarray(v)).double() for k, v in w.items()}
    t0 = time.time()
    with torch.no_grad():
      out = RT._forward(cfg, wt, f, torch.float64)
    res = dict(grid_pred_decoded=[t.numpy() if torch.is_tensor(t) else t for t in out["grid_pred_decoded"]],
               grid_pred_reg_decoded=[t.numpy() if torch.is_tensor(t) else t for t in out["grid_pred_reg_decoded"]],
               beam_outputs=out["beam_outputs"])
    st = cases.rollout_stats(cfg, res)
    if "beam_margins" in out:
      st["beam_margins"] = out["beam_margins"]
    st["checksum"] = cases.checksum(*w.values()) + cases.checksum(f["scene_feat"], f["traj"])
    np.savez_compressed(os.path.join(OUT, "atsize_" + name + ".npz"), **st)
    print("wrote atsize_%s in %.0f s" % (name, time.time() - t0), {k: np.asarray(v).shape for k, v in st.items()}, flush=True)


if __name__ == "__main__":
  main(sys.argv[1:])
