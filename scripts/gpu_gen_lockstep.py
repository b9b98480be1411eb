"""GPU: the batched gate product kernels (exact / tcgen05) and lock-step generation against per-stream generation.
  python scripts/gpu_gen_lockstep.py [n_prompts] [out_bytes]
Prints: the gate self test, then for the golden 600-byte checkpoint and a GPU-trained 64 KiB one:
kernel ms of the three generation modes, exact == per-stream, and how far the tensor-core samples follow the exact ones."""
import gzip
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmix_b200  # noqa: E402
from gmix_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def divergence(a, b):
    same = sum(x == y for x, y in zip(a, b))
    first = []
    for x, y in zip(a, b):
        if x != y:
            first.append(next(i for i in range(len(x)) if x[i] != y[i]))
    return same, (float(np.mean(first)) if first else None)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
    out_bytes = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    ctx = gmix_b200.Context(0)
    for n_slots in (300, 1184):
        bad, e_tc, e_seq, mag = ctx.selftest_gate(n_slots, 7)
        print(f"gate self test {n_slots} streams: exact mismatches {bad}, tensor-core max |err| {e_tc:.3e}, sequential fp32 max |err| {e_seq:.3e}, max |value| {mag:.3f}", flush=True)
    sh = gzip.open(os.path.join(GOLD, "ckpt600.short.gz")).read()
    lo = gzip.open(os.path.join(GOLD, "ckpt600.long.gz")).read()
    prompt = open(os.path.join(GOLD, "ckpt600_prompt.txt"), "rb").read()
    want = open(os.path.join(GOLD, "ckpt600_gen_48_1.0.out"), "rb").read()
    m = gmix_b200.Model(ctx, sh, lo, max_new_bytes=2048)
    for mode in (0, 1, 2):
        ctx.set_generation_mode(mode)
        out = ctx.generate_batch(m, [prompt] * 3 + [prompt[:9] + b"\n"], 48, 1.0)
        print(f"golden checkpoint, mode {mode} (ran {ctx.last_generation_mode}): reference bytes reproduced {[o == want for o in out[:3]]}", flush=True)
    m.close()
    # a bigger model and a full wave of distinct prompts with distinct draws
    corpus = bytes(synth.enwik_shaped_corpus(65536 + 64 * n + 4096))
    t0 = time.time()
    sh, lo = ctx.train_checkpoint(corpus[:65536])
    print(f"trained 64 KiB checkpoint in {time.time() - t0:.1f} s", flush=True)
    m = gmix_b200.Model(ctx, sh, lo, max_new_bytes=64 + out_bytes + 64)
    prompts = [corpus[65536 + 64 * i: 65536 + 64 * i + 64] for i in range(n)]
    rng = np.random.default_rng(5)
    ru = rng.random(n * out_bytes * 8, dtype=np.float32)
    outs = {}
    for mode in (0, 1, 2):
        ctx.set_generation_mode(mode)
        ctx.generate_batch(m, prompts[:8], 8, 1.0, ru, out_bytes * 8)   # warm-up: arenas, function attributes
        t0 = time.time()
        outs[mode] = ctx.generate_batch(m, prompts, out_bytes, 1.0, ru, out_bytes * 8)
        wall = time.time() - t0
        print(f"mode {mode} (ran {ctx.last_generation_mode}): {n} x {out_bytes} B, kernel {ctx.last_kernel_ms:.1f} ms, wall {wall:.2f} s, "
              f"{n * out_bytes / (ctx.last_kernel_ms / 1e3) / 1e6:.3f} MB/s of samples, launches so far {ctx.kernel_launches}", flush=True)
    print("exact lock-step == per-stream:", outs[1] == outs[0])
    same, first = divergence(outs[0], outs[2])
    print(f"tensor-core lock-step: {same} of {n} streams byte-identical to the exact samples; mean first differing byte of the others: {first}")


if __name__ == "__main__":
    main()
