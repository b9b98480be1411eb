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
