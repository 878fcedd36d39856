# coding=utf-8
"""GPU probe: cost of the pinned allocation and of the device->host copy of a beam-logit sized tensor."""
import time, torch
dev = torch.device("cuda:0")
t = torch.randn(64, 20, 12, 648, device=dev)
side = torch.cuda.Stream(device=dev)
busy = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for it in range(6):
  torch.cuda.synchronize()
  for _ in range(20): busy @ busy          # ~15 ms of queued device work, like a graph segment in flight
  t0 = time.perf_counter()
  side.wait_event(torch.cuda.current_stream(dev).record_event())
  with torch.cuda.stream(side):
    dst = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    t1 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dst.copy_(t, non_blocking=True); e1.record()
  t2 = time.perf_counter()
  side.synchronize()
  t3 = time.perf_counter()
  a = dst.numpy()
  print("iter %d: pinned alloc %.2f ms, copy call %.2f ms, wait %.2f ms, copy on device %.2f ms (%.1f GB/s), is_pinned %s"
        % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), e0.elapsed_time(e1), t.numel() * 4 / e0.elapsed_time(e1) / 1e6, dst.is_pinned()))
  if it >= 3: del a, dst      # first iterations keep the previous block alive until rebinding (like a caller holding results)
