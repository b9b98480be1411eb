"""Throughput vs resident streams per SM (same total work per resident stream)."""
import sys, time
sys.path.insert(0, ".")
import gmix_b200
from gmix_b200 import synth
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
c = gmix_b200.Context(0)
chunks = [synth.synthetic_text_chunk(i, size) for i in range(1184)]
for per_sm in [int(x) for x in (sys.argv[2:] or "1 2 3 4 5 6 7 8".split())]:
    n = 148 * per_sm
    c.configure(size, n)
    c.compress_batch(chunks[:n])
    print(f"{per_sm}/SM (retried {c.retried_streams}): {n} x {size}: kernel {c.last_kernel_ms:.0f} ms -> {n*size/c.last_kernel_ms/1e3:.3f} MB/s ({c.last_kernel_ms*1e3/size:.1f} us/byte/stream)", flush=True)
