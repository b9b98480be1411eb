"""Residency diagnostics: how many streams really run concurrently, per-stream latency vs batch size."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import gmix_b200
from gmix_b200 import synth

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
counts = [int(x) for x in sys.argv[2:]] or [1, 148, 1036]
c = gmix_b200.Context(0)
for n in counts:
    streams = [synth.synthetic_text_chunk(i, size) for i in range(n)]
    c.compress_batch(streams[:1])
    t = time.time(); out = c.compress_batch(streams); dt = time.time() - t
    u = c.get_usage(n).astype(np.int64)
    start, end, sm = u[:, 5], u[:, 6], u[:, 4]
    t0 = start.min()
    ev = sorted([(s - t0, 1) for s in start] + [(e - t0, -1) for e in end])
    cur = mx = 0
    for _, d in ev:
        cur += d; mx = max(mx, cur)
    per_sm = np.bincount(sm, minlength=148)
    dur = (end - start)
    print(f"{n} x {size}: kernel {c.last_kernel_ms:.0f} ms -> {n*size/c.last_kernel_ms/1e3:.3f} MB/s; resident(arg) {c.resident_streams}; "
          f"max concurrent {mx}; streams per SM min/max {per_sm.min()}/{per_sm.max()}; per-stream ms mean {dur.mean()/1e3:.0f} "
          f"({dur.mean()/size:.1f} us/byte); sparse max {u[:,0].max()} sets max {u[:,1].max()}", flush=True)
