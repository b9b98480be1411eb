"""GPU: one lock-step generation batch for ncu (no checks, no golden part).
  python scripts/gpu_gen_profile.py <mode 0|1|2> [n_prompts] [out_bytes]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmix_b200  # noqa: E402
from gmix_b200 import synth  # noqa: E402

mode = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
out_bytes = int(sys.argv[3]) if len(sys.argv) > 3 else 8
ctx = gmix_b200.Context(0)
corpus = bytes(synth.enwik_shaped_corpus(16400 + 64 * n + 4096))
sh, lo = ctx.train_checkpoint(corpus[:16400])
m = gmix_b200.Model(ctx, sh, lo, max_new_bytes=64 + out_bytes + 64)
prompts = [corpus[16400 + 64 * i: 16400 + 64 * i + 64] for i in range(n)]
ru = np.random.default_rng(5).random(n * out_bytes * 8, dtype=np.float32)
ctx.set_generation_mode(mode)
out = ctx.generate_batch(m, prompts, out_bytes, 1.0, ru, out_bytes * 8)
print(f"mode {mode} ran {ctx.last_generation_mode}: {n} x {out_bytes} B, kernel {ctx.last_kernel_ms:.1f} ms")
