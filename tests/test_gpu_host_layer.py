"""GPU tests of the host layer above the C ABI: the Predictor facade (Predict/Perceive/Learn per bit,
reference src/predictor.h:20-38) and the gmixb200 runner (reference src/runner/runner.cpp CLI)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
RUNNER = os.path.join(ROOT, "gmix_b200", "lib", "gmixb200")


def _discretize(p):
    return int(np.float32(1.0) + np.float32(65534.0) * np.float32(p))


def test_predictor_facade_matches_oracle_per_bit(gpu_ctx, oracle):
    import gmix_b200
    data = open(os.path.join(GOLD, "short124.in"), "rb").read()[:48]      # < 125 B: analysis stays off
    _, oprobs, op16 = oracle.compress(data, trace=True)
    pred = gmix_b200.Predictor(gpu_ctx, 64)
    got = []
    for byte in data:
        for j in range(7, -1, -1):
            bit = (byte >> j) & 1
            got.append(pred.predict())
            pred.perceive(bit)
            pred.learn()
    pred.close()
    got = np.asarray(got, dtype=np.float32)
    bad = np.nonzero(got.view(np.uint32) != oprobs.view(np.uint32))[0]
    assert bad.size == 0, f"first differing bit {bad[0]}: gpu {got[bad[0]]!r} oracle {oprobs[bad[0]]!r}"
    assert [_discretize(p) for p in got] == op16.tolist()


def test_predictor_without_learn_matches_oracle(gpu_ctx, oracle):
    """Generation-style use (reference runner-utils.cpp:187-215): learn over a prompt, then Perceive + Predict
    without Learn. Compared bit for bit with the oracle's Predictor facade driven the same way."""
    import gmix_b200
    prompt = open(os.path.join(GOLD, "text1k.in"), "rb").read()[:120]
    tail_bits = [1, 0, 0, 1, 1, 1, 0, 1, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 1, 0, 1, 1]
    lib = oracle.lib
    o = lib.gmo_new()
    pred = gmix_b200.Predictor(gpu_ctx, 256)
    got, want = [], []
    for byte in prompt:
        for j in range(7, -1, -1):
            bit = (byte >> j) & 1
            got.append(pred.predict()); want.append(lib.gmo_predict(o))
            pred.perceive(bit); lib.gmo_perceive(o, bit)
            pred.learn(); lib.gmo_learn(o)
    for bit in tail_bits:                      # learning disabled
        got.append(pred.predict()); want.append(lib.gmo_predict(o))
        pred.perceive(bit); lib.gmo_perceive(o, bit)
    pred.close()
    lib.gmo_free(o)
    got, want = np.asarray(got, dtype=np.float32), np.asarray(want, dtype=np.float32)
    bad = np.nonzero(got.view(np.uint32) != want.view(np.uint32))[0]
    assert bad.size == 0, f"first differing bit {bad[0]} ({'prompt' if bad[0] < 960 else 'no-learn tail'})"
    assert len(set(got[960:].tolist())) > 1


@pytest.mark.parametrize("name", ["text1k", "one_byte", "empty"])
def test_runner_compress_is_reference_compatible(tmp_path, name):
    out, back = str(tmp_path / "o.gmix"), str(tmp_path / "o.back")
    src = os.path.join(GOLD, name + ".in")
    subprocess.run([RUNNER, "-c", src, out], check=True)
    assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".gmix"), "rb").read()
    subprocess.run([RUNNER, "-d", os.path.join(GOLD, name + ".gmix"), back], check=True)   # a reference-written stream
    assert open(back, "rb").read() == open(src, "rb").read()


def test_runner_chunked_roundtrip_and_facade_mode(tmp_path):
    src = os.path.join(GOLD, "text_mid.in")
    cont, back = str(tmp_path / "c.gmxb"), str(tmp_path / "c.back")
    subprocess.run([RUNNER, "-C", "700", src, cont], check=True)
    subprocess.run([RUNNER, "-D", cont, back], check=True)
    assert open(back, "rb").read() == open(src, "rb").read()
    # the Predictor-facade path writes the same bytes as the batch kernel / the reference
    out = str(tmp_path / "p.gmix")
    subprocess.run([RUNNER, "-p", os.path.join(GOLD, "short124.in"), out], check=True)
    assert open(out, "rb").read() == open(os.path.join(GOLD, "short124.gmix"), "rb").read()


def test_runner_checkpoint_generate_and_train_cli(tmp_path):
    """The reference CLI shapes with a checkpoint (runner.cpp:14-104): -c/-d <ckpt>, -g, -t; fixtures written by the
    unmodified reference (tests/golden/make_golden_ckpt.py)."""
    import gzip
    import ckpt_layout
    d = tmp_path
    ck = str(d / "ckpt600")
    short_ref = gzip.open(os.path.join(GOLD, "ckpt600.short.gz")).read()
    long_ref = gzip.open(os.path.join(GOLD, "ckpt600.long.gz")).read()
    open(ck + ".short", "wb").write(short_ref)
    open(ck + ".long", "wb").write(long_ref)
    text = open(os.path.join(GOLD, "text1k.in"), "rb").read()
    (d / "a.in").write_bytes(text[:600])
    (d / "b.in").write_bytes(text[600:])
    subprocess.run([RUNNER, "-c", ck, str(d / "b.in"), str(d / "b.gmix")], check=True)
    assert (d / "b.gmix").read_bytes() == open(os.path.join(GOLD, "ckpt600_b.gmix"), "rb").read()
    subprocess.run([RUNNER, "-d", ck, str(d / "b.gmix"), str(d / "b.back")], check=True)
    assert (d / "b.back").read_bytes() == text[600:]
    subprocess.run([RUNNER, "-g", ck, os.path.join(GOLD, "ckpt600_prompt.txt"), str(d / "gen.out"), "48", "1.0"], check=True)
    assert (d / "gen.out").read_bytes() == open(os.path.join(GOLD, "ckpt600_gen_48_1.0.out"), "rb").read()
    # batched generation from the C++ host: one prompt per line, every sample is what its own `gmix -g` process writes;
    # "exact" = lock-step with the batched gate product in the reference's arithmetic, "stream" = one CTA per prompt
    prompt = open(os.path.join(GOLD, "ckpt600_prompt.txt"), "rb").read()
    want = open(os.path.join(GOLD, "ckpt600_gen_48_1.0.out"), "rb").read()
    if prompt.endswith(b"\n") and prompt.count(b"\n") == 1:
        (d / "prompts.txt").write_bytes(prompt * 3)
        for mode in ("exact", "stream", "tensor"):
            r = subprocess.run([RUNNER, "-G", ck, str(d / "prompts.txt"), str(d / f"gen3_{mode}.out"), "48", "1.0", mode], check=True, capture_output=True, text=True)
            assert f"mode {mode}" in r.stdout, r.stdout
            got = (d / f"gen3_{mode}.out").read_bytes()
            assert len(got) == 3 * 48 and got[:48] == got[48:96] == got[96:]
            if mode != "tensor":
                assert got == want * 3, mode
    subprocess.run([RUNNER, "-t", str(d / "a.in"), str(d / "b.in")], check=True, cwd=str(d))
    assert (d / "data" / "trained_checkpoint.long").read_bytes() == long_ref
    diff = ckpt_layout.differing_sections((d / "data" / "trained_checkpoint.short").read_bytes(), short_ref)
    # the fixture was trained with analysis off (ref_driver), `-t` enables it like the reference's RunTraining does
    # (runner-utils.cpp:268): inactive predictions are then zeroed every bit instead of keeping their last value
    assert set(diff) <= ckpt_layout.SCRATCH | {"stm.predictions"}, diff


def test_runner_chunked_mode_shards_over_all_visible_gpus(tmp_path, gpu_ctx):
    """`gmixb200 -C` = the C++ multi-GPU host (host/multi_gpu.h): one context + host thread per visible GPU, byte-balanced
    contiguous stream ranges, one ncclAllGather of {size, FNV-1a 64} per stream (verified by the runner itself). The
    container must hold exactly the streams the single-GPU batch call writes, whatever the number of GPUs."""
    import struct
    import torch
    from gmix_b200 import synth
    corpus = synth.enwik_shaped_corpus(9 * 3000 + 1234)           # 10 streams, the last one short
    src = tmp_path / "corpus.in"
    src.write_bytes(corpus)
    cont, back = str(tmp_path / "c.gmxb"), str(tmp_path / "c.back")
    r = subprocess.run([RUNNER, "-C", "3000", str(src), cont], capture_output=True, text=True, timeout=240, env=dict(os.environ, GMIXB200_VERBOSE="1"))
    assert r.returncode == 0, r.stdout + r.stderr
    ngpu = torch.cuda.device_count()
    assert f"10 streams on {ngpu} GPU(s)" in r.stdout and "ncclAllGather" in r.stdout, r.stdout
    chunks = [corpus[i:i + 3000] for i in range(0, len(corpus), 3000)]
    want = gpu_ctx.compress_batch(chunks)
    blob = open(cont, "rb").read()
    assert blob[:4] == b"GMXB" and struct.unpack("<I", blob[4:8])[0] == len(chunks)
    sizes = struct.unpack(f"<{len(chunks)}Q", blob[8:8 + 8 * len(chunks)])
    assert list(sizes) == [len(w) for w in want]
    assert blob[8 + 8 * len(chunks):] == b"".join(want)
    subprocess.run([RUNNER, "-D", cont, back], check=True, timeout=240)
    assert open(back, "rb").read() == corpus
    if ngpu > 1:                                                    # and the same bytes when limited to one GPU
        env = dict(os.environ, GMIXB200_GPUS="1")
        subprocess.run([RUNNER, "-C", "3000", str(src), cont + "1"], check=True, env=env)
        assert open(cont + "1", "rb").read() == blob


def test_training_mode_reproduces_the_reference_metrics(tmp_path):
    """`gmixb200 -t train test` against `gmix -t` of the unmodified reference (fixtures: tests/golden/make_golden_ckpt.py
    make_training_golden): the periodic test-file scores on a copy of the predictor (analysis/training.tsv, every number),
    the coded training stream (data/tmp) and the final long-term memory."""
    import hashlib
    import json
    want = json.load(open(os.path.join(GOLD, "train_text1k_short124.json")))
    r = subprocess.run([RUNNER, "-t", os.path.join(GOLD, "text1k.in"), os.path.join(GOLD, "short124.in")], cwd=str(tmp_path), capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "analysis" / "training.tsv").read_text() == open(os.path.join(GOLD, "train_text1k_short124.training.tsv")).read()
    assert (tmp_path / "data" / "tmp").read_bytes() == open(os.path.join(GOLD, "train_text1k_short124.tmp"), "rb").read()
    blob = (tmp_path / "data" / "trained_checkpoint.long").read_bytes()
    assert len(blob) == want["long_bytes"] and hashlib.md5(blob).hexdigest() == want["long_md5"]


def test_runner_writes_the_reference_analysis_files(tmp_path):
    """`gmixb200 -c` from scratch writes analysis/entropy.tsv and analysis/memory.tsv as `gmix -c` does (Predictor::EnableAnalysis /
    RunAnalysis, predictor.cpp:422-504), checked against the unmodified reference run on the same input: memory.tsv byte for byte
    (PPMd's used memory and the match history change per row), entropy.tsv field by field - identical text, except that a field
    may differ by one unit of its last printed digit (1e-5): log2 is CUDA's (<= 1 ulp) where the reference calls glibc's."""
    ref = os.path.join(ROOT, "oracle", "_ref", "gmix")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    src = os.path.join(GOLD, "text1k.in")
    ours, theirs = tmp_path / "ours", tmp_path / "ref"
    ours.mkdir(), theirs.mkdir()
    subprocess.run([RUNNER, "-c", src, "out.gmix"], check=True, cwd=str(ours), stdout=subprocess.DEVNULL)
    subprocess.run([ref, "-c", src, "out.gmix"], check=True, cwd=str(theirs), stdout=subprocess.DEVNULL)
    assert (ours / "out.gmix").read_bytes() == (theirs / "out.gmix").read_bytes()
    assert (ours / "analysis" / "memory.tsv").read_bytes() == (theirs / "analysis" / "memory.tsv").read_bytes()
    a = (ours / "analysis" / "entropy.tsv").read_text().split("\n")
    b = (theirs / "analysis" / "entropy.tsv").read_text().split("\n")
    assert a[0] == b[0] and len(a) == len(b) and len(a) > 1000
    off = 0
    for la, lb in zip(a[1:], b[1:]):
        fa, fb = la.split("\t"), lb.split("\t")
        assert len(fa) == len(fb) and fa[0] == fb[0]
        for x, y in zip(fa[1:], fb[1:]):
            if x != y:
                off += 1
                assert abs(float(x) - float(y)) < 1.5e-5, (fa[0], x, y)
    assert off <= 5, off
    # GMIXB200_ANALYSIS=0: same stream, no files
    quiet = tmp_path / "quiet"
    quiet.mkdir()
    subprocess.run([RUNNER, "-c", src, "out.gmix"], check=True, cwd=str(quiet), stdout=subprocess.DEVNULL, env=dict(os.environ, GMIXB200_ANALYSIS="0"))
    assert (quiet / "out.gmix").read_bytes() == (theirs / "out.gmix").read_bytes() and not (quiet / "analysis").exists()
