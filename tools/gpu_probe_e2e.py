# coding=utf-8
"""GPU probe: host-side phases of one Session.run at a small batch (where e2e trails the device-resident value)."""
import os, sys, time, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multiverse_b200", "dropin"))
from multiverse_b200 import synthetic
import tensorflow as tf, pred_models as pm
import multiverse_b200.pred_models as impl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = synthetic.make_config(batch_size=n, use_grids=[True, False], use_beam_search=True, beam_size=20, diverse_beam=True,
                            diverse_gamma=0.01, fix_num_timestep=1)
w = synthetic.make_weights(cfg); f = synthetic.make_feeds(cfg, n)
args = types.SimpleNamespace(**vars(cfg)); args.modelname, args.use_soft_grid_class, args.use_gt_grid, args.is_train = "p", False, False, False
model = pm.get_model(args, gpuid=0)
tf.global_variables_initializer().run()
for v in tf.global_variables():
  k = v.name.split(":")[0]
  if k in w: v.assign(w[k])
sess = tf.Session()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
fd = {model.scene_feat: pin(f["scene_feat"]), model.obs_scene: pin(f["obs_scene"]), model.obs_length: np.full((n,), 8, "int32"),
      model.pred_length: np.full((n,), 12, "int32"), model.is_train: False,
      model.obs_traj: np.ascontiguousarray(f["traj64"][:, :8]), model.grid_obs_labels[0]: pin(f["grid_obs_labels"][0]),
      model.grid_centers[0]: np.asarray(synthetic.grid_centers(cfg)[0], np.float64)}
fetches = [model.grid_pred_decoded[0], model.grid_pred_reg_decoded[0], model.beam_outputs]
if os.environ.get("PROBE_SMALL_FETCH"): fetches = [model.grid_pred_reg_decoded[0]]
for _ in range(4): sess.run(fetches, fd)
T = {}
def timed(name, fn):
  def g(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(*a, **k); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    T.setdefault(name, []).append((t1 - t0, t2 - t0)); return r
  return g
model._device_feeds = timed("device_feeds", model._device_feeds)
eng = model._engine
eng.forward_graph = timed("forward_graph", eng.forward_graph)
eng.forward = timed("forward", eng.forward)
_empty = torch.empty
def empty_timed(*a, **k):
  t0 = time.perf_counter(); r = _empty(*a, **k)
  if k.get("pin_memory") and r.numel() > 1e5: T.setdefault("pinned_empty %s" % (tuple(r.shape),), []).append((time.perf_counter() - t0,) * 2)
  return r
torch.empty = empty_timed
_contig = torch.Tensor.contiguous
def contig_timed(self, *a, **k):
  t0 = time.perf_counter(); r = _contig(self, *a, **k)
  if self.is_cuda and self.numel() > 1e6: T.setdefault("contiguous %s same=%s" % (tuple(r.shape), r is self), []).append((time.perf_counter() - t0,) * 2)
  return r
torch.Tensor.contiguous = contig_timed
t0 = time.perf_counter()
for _ in range(10): sess.run(fetches, fd)
tot = (time.perf_counter() - t0) / 10
print("n=%d total per run %.2f ms" % (n, tot * 1e3))
for k, v in T.items():
  print("  %-14s host %.2f ms, until device idle %.2f ms" % (k, 1e3 * np.mean([a for a, _ in v]), 1e3 * np.mean([b for _, b in v])))

# device-resident forward of the same feeds, eager and graph, no fetch
torch.empty = _empty; torch.Tensor.contiguous = _contig
T.clear()
with torch.cuda.device(eng.device):
  feeds = model._device_feeds(fd)
  for _ in range(5): eng.forward(feeds, 12)
  if os.environ.get("MVB_CUDA_GRAPH") != "0":
    for _ in range(5): eng.forward_graph(feeds, 12)
for k, v in T.items():
  print("  resident %-14s host %.2f ms, until device idle %.2f ms" % (k, 1e3 * np.mean([a for a, _ in v][-3:]), 1e3 * np.mean([b for _, b in v][-3:])))
