"""Phase breakdown of the stream kernel on the GPU (cycle counters of thread 0, lap-timer style)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import gmix_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nstreams = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = open("tests/data/english.dic", "rb").read()
c = gmix_b200.Context(0)
c.set_profile(True)
from gmix_b200 import synth
streams = [synth.synthetic_text_chunk(i, n) for i in range(nstreams)]
t = time.time()
out = c.compress_batch(streams)
dt = time.time() - t
prof = c.get_profile(nstreams).astype(np.float64)
tot = prof.sum(axis=1)
print(f"{nstreams} x {n} B: wall {dt:.3f}s kernel {c.last_kernel_ms:.1f} ms resident {c.resident_streams} "
      f"-> {nstreams*n/c.last_kernel_ms/1e3:.3f} MB/s, {c.last_kernel_ms*1e3/n:.1f} us/byte/stream")
print('rows with data:', int((tot > 0).sum()), 'of', len(tot))
mean = prof[tot > 0].mean(axis=0)
for name, v in zip(c.PROFILE_SLOTS, mean):
    print(f"  {name:14s} {v/n:12.0f} cyc/byte  {100*v/mean.sum():5.1f}%")
print(f"  total          {mean.sum()/n:12.0f} cyc/byte")
