# coding=utf-8
"""Probe: sustained dense rate of cuBLASLt int8 (torch._int_mm) beside bf16 on this B200, same
method as MEASURED_PEAKS.json (8192^3, back to back for ~3 s).  Decides whether the sliced-int8
cell (tcgen05 kind::i8) can beat the 3-pass bf16 cell under the 1 kW power cap."""
import json, time, torch
dev = torch.device("cuda:0")
n = 8192
def sustained(fn, secs=3.0):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
  t0 = time.time(); k = 0
  e0.record()
  while time.time() - t0 < secs:
    for _ in range(10): fn()
    k += 10
    torch.cuda.synchronize()
  e1.record(); torch.cuda.synchronize()
  return 2.0 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
a = torch.randn(n, n, device=dev, dtype=torch.bfloat16); b = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
ai = torch.randint(-128, 127, (n, n), device=dev, dtype=torch.int8); bi = torch.randint(-128, 127, (n, n), device=dev, dtype=torch.int8).t().contiguous().t()
out = {}
out["bf16_tflops_sustained"] = sustained(lambda: torch.matmul(a, b))
try:
  out["int8_tops_sustained"] = sustained(lambda: torch._int_mm(ai, bi))
except Exception as e:
  out["int8_error"] = repr(e)
try:
  a8 = a.to(torch.float8_e4m3fn); b8 = b.t().contiguous().to(torch.float8_e4m3fn).t()
  s = torch.tensor(1.0, device=dev)
  out["fp8_tflops_sustained"] = sustained(lambda: torch._scaled_mm(a8, b8, scale_a=s, scale_b=s, out_dtype=torch.bfloat16))
except Exception as e:
  out["fp8_error"] = repr(e)
print(json.dumps(out))
