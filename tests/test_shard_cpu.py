# coding=utf-8
"""World-size-2 gloo test (CPU) of the multi-GPU host logic: trajectory sharding of the feeds
(bench.py / SURVEY.md §8e: contiguous split, per-shard scene-frame re-indexing, no data-path
collective) and the max-over-ranks timing reduction, through multiverse_b200.synthetic.shard_feeds - the function
bench.py's inference and training arms and tests/ddp_check.py call.  The per-shard compute is replaced by a
deterministic row-wise function, so the test checks exactly what the sharding must guarantee:
concatenating the shards' outputs reproduces the full-batch output, row for row."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multiverse_b200 import synthetic


shard_feeds = synthetic.shard_feeds      # the function bench.py and tests/ddp_check.py shard with


def rowwise_digest(f):
  """A per-trajectory function of exactly the tensors a trajectory's rollout may depend on."""
  frames = f["scene_feat"][f["obs_scene"]]                      # [n,T,SH,SW,SC]
  d = frames.reshape(frames.shape[0], -1).sum(1)
  for lab, reg in zip(f["grid_obs_labels"], f["grid_obs_regress"]):
    d = d + lab.sum(1) + reg.reshape(reg.shape[0], -1).sum(1)
  return d


def _worker(rank, world, port, ret):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  cfg = synthetic.make_config(batch_size=8)
  full = synthetic.make_feeds(cfg, 8, seed=11)
  mine = rowwise_digest(shard_feeds(full, rank, world))
  gathered = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
  dist.all_gather(gathered, torch.from_numpy(mine.astype(np.float64)))
  ms = torch.tensor([10.0 + rank], dtype=torch.float64)      # bench.py: max over ranks
  dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  dist.barrier()
  if rank == 0:
    ret["cat"] = torch.cat(gathered).numpy()
    ret["full"] = rowwise_digest(full)
    ret["ms"] = float(ms)
  dist.destroy_process_group()


def test_two_rank_sharding_matches_full_batch():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  with mp.Manager() as mgr:
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert np.array_equal(ret["cat"], ret["full"])
    assert ret["ms"] == 11.0


def test_shard_frame_indices_are_local():
  cfg = synthetic.make_config(batch_size=6)
  full = synthetic.make_feeds(cfg, 6, seed=2)
  for r in range(3):
    sh = shard_feeds(full, r, 3)
    assert sh["obs_scene"].min() == 0 and sh["obs_scene"].max() == sh["scene_feat"].shape[0] - 1


def test_shard_recompacts_shared_and_unordered_frames():
  """Trajectories that share segmentation frames, in any order (the reference reuses one frame for many steps and
  rows, code/preprocess.py:390-400): a shard carries only the frames it indexes, and every row still sees its own."""
  cfg = synthetic.make_config(batch_size=8)
  full = synthetic.make_feeds(cfg, 8, seed=5)
  rng = np.random.default_rng(0)
  full["obs_scene"] = rng.integers(0, 8, size=full["obs_scene"].shape).astype(np.int32)
  want = rowwise_digest(full)
  for world in (2, 4):
    got = np.concatenate([rowwise_digest(shard_feeds(full, r, world)) for r in range(world)])
    assert np.array_equal(got, want)
    for r in range(world):
      sh = shard_feeds(full, r, world)
      assert sh["scene_feat"].shape[0] == len(np.unique(full["obs_scene"][r * (8 // world):(r + 1) * (8 // world)]))
