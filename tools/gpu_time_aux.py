# coding=utf-8
"""CUDA-event timings of the HBM-bound auxiliary kernels at a given number of sample rows."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiverse_b200 import ops
dev = torch.device("cuda:0")
ns, h, w, b = int(sys.argv[1]) if len(sys.argv) > 1 else 2560, 36, 18, 20
h32 = ops.alloc_state(ns, h, w, dev); h32.normal_()
sm = torch.randn(ns // b, h, w, 64, device=dev)
xh = ops.alloc_xh(ns, h, w, 288, 2, dev)
Wo = torch.randn(3, 3, 256, 1, device=dev) * 0.1
logits = torch.empty(ns, h * w, device=dev); ids = torch.empty(ns, dtype=torch.int32, device=dev)
sc = torch.zeros(ns // b, b, device=dev); so = torch.zeros_like(sc)
bi = torch.empty(ns // b, b, dtype=torch.int32, device=dev); bp = torch.empty_like(bi); rm = torch.empty(ns, dtype=torch.int32, device=dev)
def t(fn, n=10):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n
g = t(lambda: ops.gnn_attend_fwd(h32, sm, xh, h, w, ns, beam=b))
hd = t(lambda: ops.head_class_fwd(h32, Wo, logits, ids, None, None, None, h, w, ns))
bs = t(lambda: ops.beam_step(logits.view(ns // b, b, h * w), sc, so, bi, bp, rm, ns // b, b, h * w, False, False, True, 0.01))
gb = ns * h * w * (256 + 64 / b) * 4 / 1e9 + ns * h * w * 256 * 2 * 2 / 1e9
print("rows %d: gnn %.3f ms (%.0f GB/s of 6571), head %.3f ms (%.0f GB/s), beam_step %.3f ms" %
      (ns, g, gb / g * 1e3, hd, ns * h * w * 1024 / 1e9 / hd * 1e3, bs))
