# coding=utf-8
"""2-GPU data-parallel equivalence (run under torchrun, see tests/test_ddp_gpu.py): the
all-reduced, 1/G-scaled gradients and the updated weights of G ranks x N/G trajectories equal a
single rank's full-batch step (losses are means over equal shards, the weight
decay term is batch independent)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiverse_b200 import synthetic
from multiverse_b200.train_engine import TrainEngine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 4 * world
kw = dict(use_grids=[False, True], is_train=True, grid_loss_weight=1.0, grid_reg_loss_weight=0.1, wd=0.001,
          clip_gradient_norm=10.0)
w = synthetic.make_weights(synthetic.make_config(batch_size=N, **kw), 3)
f = synthetic.make_feeds(synthetic.make_config(batch_size=N, **kw), N, 3, with_pred=True)
g = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
def feeds_of(r, wsize):
  sh = synthetic.shard_feeds(f, r, wsize)       # the product's definition of a shard (bench.py uses the same)
  return {k: ([g(a) for a in v] if isinstance(v, list) else g(v)) for k, v in sh.items() if k != "traj"}
n_loc = N // world
eng = TrainEngine(synthetic.make_config(batch_size=n_loc, **kw), {k: torch.from_numpy(v) for k, v in w.items()}, dev, 2)
losses, _ = eng.train_step(feeds_of(rank, world), 0.2, dist)
ok = True
if rank == 0:
  full = TrainEngine(synthetic.make_config(batch_size=N, **kw), {k: torch.from_numpy(v) for k, v in w.items()}, dev, 2)
  l_full, _ = full.train_step(feeds_of(0, 1), 0.2, None)
  e_loss = float((losses - l_full).abs().max() / l_full.abs().max())
  e_grad = float((eng.flat_grad / world - full.flat_grad).abs().max() / full.flat_grad.abs().max())
  e_w = max(float((eng.params[k] - full.params[k]).abs().max()) for k in eng.names)
  moved = max(float((full.params[k].cpu() - torch.from_numpy(w[k])).abs().max()) for k in eng.names)
  print("DDP_CHECK loss_rel=%.3e grad_rel=%.3e weight_abs=%.3e (update magnitude %.3e)" % (e_loss, e_grad, e_w, moved), flush=True)
  ok = e_loss < 1e-4 and e_grad < 5e-4 and e_w < 1e-3 * moved + 1e-7
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
