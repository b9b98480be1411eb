"""BASELINE.json configs[3] and configs[4] at full size on one B200, with the unmodified reference (oracle/_ref, CPU)
checking a sample in the background. Not part of bench.py's default run (each takes minutes); results go to stdout as
one JSON object per config.

  python scripts/full_configs.py [enwik] [generate]

enwik    : 100 000 000 B of the enwik-shaped corpus as 100 streams of 1 000 000 B, compressed as one batch, all streams
           decompressed again and compared; streams 0 and 57 also compressed by `oracle/_ref/gmix -c` meanwhile.
generate : checkpoint = Predict/Perceive/Learn over the first 1 000 000 B of the corpus ON THE GPU, written in the
           reference's format; 8192 prompts = 64-byte windows at offsets 1 000 000 + 4096 k; 1024 bytes each,
           temperature 1.0, the reference's draws. `oracle/_ref/gmix -g` loads the GPU-written checkpoint and generates
           for prompts 0 and 4097: the bytes must be identical.
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "gmix")


def main():
    what = sys.argv[1:] or ["enwik", "generate"]
    from gmix_b200 import synth
    import gmix_b200
    n_streams, stream_bytes = 100, 1000000
    t0 = time.time()
    corpus = synth.enwik_shaped_corpus(n_streams * stream_bytes if "enwik" in what else 1000000 + 4096 * 8192 + 64)
    print(f"corpus of {len(corpus)} B generated in {time.time() - t0:.0f} s", file=sys.stderr, flush=True)
    ctx = gmix_b200.Context(0)
    work = tempfile.mkdtemp(prefix="gmix_full_")

    if "enwik" in what:
        streams = [corpus[i * stream_bytes:(i + 1) * stream_bytes] for i in range(n_streams)]
        sample = [0, 57]
        procs = []
        if os.path.exists(REF):
            for i in sample:
                open(os.path.join(work, f"s{i}.in"), "wb").write(streams[i])
                procs.append(subprocess.Popen([REF, "-c", f"s{i}.in", f"s{i}.gmix"], cwd=work, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        t0 = time.time()
        comp = ctx.compress_batch(streams)
        t_c, k_c = time.time() - t0, ctx.last_kernel_ms
        resident, arena = ctx.resident_streams, ctx.arena_bytes
        t0 = time.time()
        back = ctx.decompress_batch(comp)
        t_d, k_d = time.time() - t0, ctx.last_kernel_ms
        lossless = back == streams
        ref_equal = None
        if procs:
            for p in procs:
                p.wait()
            ref_equal = all(open(os.path.join(work, f"s{i}.gmix"), "rb").read() == comp[i] for i in sample)
        total = n_streams * stream_bytes
        print(json.dumps({"config": "configs[4]: 100 MB enwik-shaped corpus as 100 x 1 MB streams, 1 B200", "streams": n_streams,
                          "compress_MBps_e2e": total / t_c / 1e6, "compress_kernel_ms": k_c, "decompress_MBps_e2e": total / t_d / 1e6,
                          "decompress_kernel_ms": k_d, "bits_per_byte": 8.0 * sum(len(c) for c in comp) / total,
                          "resident_streams": resident, "arena_mib_per_stream": arena >> 20, "retried_streams": ctx.retried_streams,
                          "all_streams_lossless": lossless, "reference_sample": sample if procs else None,
                          "reference_sample_byte_identical": ref_equal}), flush=True)
        if not lossless or ref_equal is False:
            raise SystemExit("configs[4]: parity failure")

    if "generate" in what:
        T, n_prompts, G = 1000000, 8192, 1024
        t0 = time.time()
        sh, lo = ctx.train_checkpoint(corpus[:T])
        t_train = time.time() - t0
        open(os.path.join(work, "ck.short"), "wb").write(sh)
        open(os.path.join(work, "ck.long"), "wb").write(lo)
        sample = [0, 4097]
        prompts = [corpus[T + 4096 * k:T + 4096 * k + 64] for k in range(n_prompts)]
        procs = []
        if os.path.exists(REF):
            for k in sample:
                open(os.path.join(work, f"p{k}.txt"), "wb").write(prompts[k])
                procs.append(subprocess.Popen([REF, "-g", "ck", f"p{k}.txt", f"g{k}.out", str(G), "1.0"], cwd=work, stdout=subprocess.DEVNULL,
                                              stderr=subprocess.DEVNULL))
        t0 = time.time()
        model = gmix_b200.Model(ctx, sh, lo, max_new_bytes=64 + G)
        t_load = time.time() - t0
        ctx.generate_batch(model, prompts, 1)      # warm-up with the timed batch shape (sizes the overlay arenas)
        t0 = time.time()
        out = ctx.generate_batch(model, prompts, G, 1.0)
        t_g, k_g = time.time() - t0, ctx.last_kernel_ms
        ref_equal = None
        if procs:
            for p in procs:
                p.wait()
            ref_equal = all(open(os.path.join(work, f"g{k}.out"), "rb").read() == out[k] for k in sample)
        print(json.dumps({"config": "configs[3]: 8192 prompts x 1 KB from a checkpoint trained on 1 MB, learning disabled while sampling",
                          "checkpoint": {"trained_on_gpu_s": t_train, "short_bytes": len(sh), "long_bytes": len(lo), "load_s": t_load,
                                         "arena_mib_per_stream": model.arena_bytes >> 20},
                          "prompts": n_prompts, "generated_bytes_per_s_e2e": n_prompts * G / t_g, "kernel_ms": k_g,
                          "resident_streams": ctx.resident_streams, "overlay_arena_mib_per_stream": ctx.arena_bytes >> 20, "distinct_outputs": len(set(out)),
                          "reference_sample": sample if procs else None, "reference_loads_gpu_checkpoint_and_generates_identical_bytes": ref_equal}),
              flush=True)
        model.close()
        if ref_equal is False:
            raise SystemExit("configs[3]: parity failure")
    ctx.close()


if __name__ == "__main__":
    main()
