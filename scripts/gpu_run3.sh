set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np
import gmix_b200
from gmix_b200 import synth
c = gmix_b200.Context(0)
for n, size in ((148, 4096), (1036, 16384), (1036, 65536)):
    streams = [synth.synthetic_text_chunk(i, size) for i in range(n)]
    t = time.time(); out = c.compress_batch(streams); dt = time.time() - t
    u = c.get_usage(n)
    print(f"{n} x {size}: wall {dt:.2f}s kernel {c.last_kernel_ms:.0f} ms resident {c.resident_streams} arena {c.arena_bytes>>20} MiB "
          f"-> {n*size/c.last_kernel_ms/1e3:.3f} MB/s; retried {c.retried_streams}; usage max {u.max(axis=0).tolist()} mean {u.mean(axis=0).astype(int).tolist()}", flush=True)
    if size == 4096:
        back = c.decompress_batch(out[:8])
        assert back == streams[:8]
PY
python scripts/gpu_profile.py 8192 1
python scripts/gpu_profile.py 4096 1036
