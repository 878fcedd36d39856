# coding=utf-8
"""GPU probe: time of the graph-attention and head kernels at the beam decoder's shape (10 240 rows of 36x18)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multiverse_b200 import ops
dev = torch.device("cuda:0")
ns, h, w, beam = int(sys.argv[1]) if len(sys.argv) > 1 else 10240, 36, 18, 20
S = (h + 1) * (w + 1)
h32 = torch.randn(ns * S, 256, device=dev) * 0.3
scene = torch.randn(ns // beam, h * w, 64, device=dev)
rm = torch.randint(0, ns, (ns,), device=dev, dtype=torch.int32)
def timeit(fn, n=5):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
  for i in range(n):
    e[i].record(); fn()
  e[n].record(); torch.cuda.synchronize()
  return min(e[i].elapsed_time(e[i + 1]) for i in range(n))
for planes in (ops.PLANES_F16F8, 2):
  xh = ops.alloc_xh(ns, h, w, 256 + 64, planes, dev)
  t = timeit(lambda: ops.gnn_attend_fwd(h32, scene, xh, h, w, ns, beam=beam, row_map=rm))
  rd, wr = ns * h * w * (256 + 64 / beam) * 4, ns * h * w * 256 * 4
  print("gnn planes=%d: %.3f ms  (%.0f GB/s algorithmic: h32 + scene read, operand planes written)" % (planes, t, (rd + wr) / t / 1e6))
Wo = torch.randn(3, 3, 256, 1, device=dev) * 0.05
We = torch.randn(3, 3, 1, 32, device=dev) * 0.3; be = torch.randn(32, device=dev) * 0.1
logits = torch.empty(ns, h * w, device=dev); ids = torch.empty(ns, dtype=torch.int32, device=dev)
t = timeit(lambda: ops.head_class_fwd(h32, Wo, logits, ids, None, None, None, h, w, ns, planes=ops.PLANES_F16F8))
print("head_class (logits only, as in the beam loop): %.3f ms (%.0f GB/s)" % (t, ns * h * w * 257 * 4 / t / 1e6))
xh = ops.alloc_xh(ns, h, w, 256 + 64, ops.PLANES_F16F8, dev)
t = timeit(lambda: ops.head_class_fwd(h32, Wo, logits, ids, We, be, xh, h, w, ns))
print("head_class + embedded one-hot feedback: %.3f ms" % t)
Wo2 = torch.randn(3, 3, 256, 2, device=dev) * 0.05
We2 = torch.randn(3, 3, 2, 32, device=dev) * 0.3
off = torch.empty(ns, h * w, 2, device=dev)
t = timeit(lambda: ops.head_reg_fwd(h32, Wo2, off, We2, be, xh, h, w, ns))
print("head_reg + embedded feedback: %.3f ms (%.0f GB/s)" % (t, ns * h * w * 258 * 4 / t / 1e6))
