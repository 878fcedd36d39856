# coding=utf-8
"""One training step (workload c5 shape) inside a cudaProfilerStart/Stop range."""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multiverse_b200 import synthetic
from multiverse_b200.train_engine import TrainEngine
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=128); a = ap.parse_args()
wl = bench.WORKLOADS["c5"]
cfg = synthetic.make_config(batch_size=a.batch, **wl["cfg"])
dev = torch.device("cuda:0")
f = synthetic.make_feeds(cfg, a.batch, with_pred=True)
g = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
feeds = {k: ([g(x) for x in v] if isinstance(v, list) else g(v)) for k, v in f.items() if k != "traj"}
eng = TrainEngine(cfg, {k: torch.from_numpy(v) for k, v in synthetic.make_weights(cfg).items()}, dev, 2)
for _ in range(2):
  eng.train_step(feeds, 0.2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.train_step(feeds, 0.2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
