"""Throughput vs resident streams per SM (same total work per resident stream). per_sm = 0: one full wave of whatever
the library makes resident."""
import sys, time
sys.path.insert(0, ".")
import gmix_b200
from gmix_b200 import synth
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
c = gmix_b200.Context(0)
sms = c.sm_count
chunks = {}
for per_sm in [int(x) for x in (sys.argv[2:] or "1 2 3 4 5 6 7 8".split())]:
    if per_sm == 0:
        c.configure(size, 0)
        n = c.max_resident_streams
    else:
        n = sms * per_sm
        c.configure(size, n)
    for i in range(n):
        if i not in chunks:
            chunks[i] = synth.synthetic_text_chunk(i, size)
    c.compress_batch([chunks[i] for i in range(n)])
    print(f"{per_sm}/SM (retried {c.retried_streams}): {n} x {size}: kernel {c.last_kernel_ms:.0f} ms -> {n*size/c.last_kernel_ms/1e3:.3f} MB/s ({c.last_kernel_ms*1e3/size:.1f} us/byte/stream)", flush=True)
