set -x
nproc; lscpu | head -20; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
g++ -std=c++17 -O2 -ffp-contract=off -pthread tests/native/dmath_check.cpp -o /tmp/dmath_check && /tmp/dmath_check 97
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python -c "
import time, gmix_b200
c = gmix_b200.Context(0)
t=time.time(); print('full math sweep', c.selftest_math(1), time.time()-t, 's')
d = open('tests/data/english.dic','rb').read()
for n in (4096, 16384):
    t=time.time(); out = c.compress_batch([d[:n]]); dt=time.time()-t
    print('single stream', n, '->', len(out[0]), 'bytes', dt, 's kernel_ms', c.last_kernel_ms, 'arena MiB', c.arena_bytes>>20)
streams=[d[i*3000:i*3000+2048] for i in range(128)]
t=time.time(); out = c.compress_batch(streams); dt=time.time()-t
print('128 x 2KiB', dt, 's kernel_ms', c.last_kernel_ms, 'resident', c.resident_streams, 'MB/s', 128*2048/c.last_kernel_ms/1e3)
" 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()"
