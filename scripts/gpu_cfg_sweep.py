"""Throughput of kernel configurations (role split of the stream CTA) on LEN-byte synthetic-text streams, after a
byte-parity check of that configuration against the golden vectors.
  python scripts/gpu_cfg_sweep.py LEN [N_STREAMS|0 = one full wave] [cfg ...]"""
import sys
sys.path.insert(0, ".")
import gmix_b200
from gmix_b200 import synth
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nstreams = int(sys.argv[2]) if len(sys.argv) > 2 else 0
c = gmix_b200.Context(0)
cfgs = [int(x) for x in sys.argv[3:]] or list(range(len(c.kernel_configs())))
names = ["text1k", "repetitive", "random1200", "text_mid", "synth_chunk0_4k"]
ins = [open(f"tests/golden/{n}.in", "rb").read() for n in names]
want = [open(f"tests/golden/{n}.gmix", "rb").read() for n in names]
chunks = {}
for k in cfgs:
    wb, wl, minb, serial, ws = c.kernel_configs()[k]
    c.set_kernel_config(k)
    got = c.compress_batch(ins)
    ok = got == want and c.decompress_batch(got) == ins
    c.configure(size, 0)
    n = nstreams or c.max_resident_streams
    for i in range(n):
        if i not in chunks:
            chunks[i] = synth.synthetic_text_chunk(i, size)
    batch = [chunks[i] for i in range(n)]
    comp = c.compress_batch(batch)
    ms = c.last_kernel_ms
    back = c.decompress_batch(comp)
    dms = c.last_kernel_ms
    print(f"cfg {k} ({'serial' if serial else 'roles'} bit {wb}w, lstm {wl}w, {minb}/SM{', resident weights' if ws else ''}) parity {'OK' if ok else 'FAILED'} "
          f"roundtrip {'OK' if back == batch else 'FAILED'}: {n} x {size}: compress {ms:.0f} ms -> {n*size/ms/1e3:.3f} MB/s ({size/ms:.2f} KB/s per stream), "
          f"decompress {dms:.0f} ms -> {n*size/dms/1e3:.3f} MB/s", flush=True)
