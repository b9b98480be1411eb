"""Pins the CPU oracle (oracle/gmix_oracle.cpp) against the golden vectors generated from the
UNMODIFIED reference (tests/golden/make_golden.py), and against the reference binary itself when
oracle/_ref is present (this container)."""
import hashlib
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
ROOT = os.path.dirname(HERE)
CASES = sorted(f[:-3] for f in os.listdir(GOLD) if f.endswith(".in"))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def _load(name):
    return (open(os.path.join(GOLD, name + ".in"), "rb").read(), open(os.path.join(GOLD, name + ".gmix"), "rb").read(),
            np.fromfile(os.path.join(GOLD, name + ".p16"), dtype=np.uint16))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_streams_and_probabilities(oracle, name):
    data, want, p16 = _load(name)
    got, _, op16 = oracle.compress(data, trace=True)
    assert got == want
    assert np.array_equal(op16.astype(np.uint16), p16)
    assert int(op16.max(initial=0)) < 65536


@pytest.mark.parametrize("name", CASES)
def test_oracle_decompress_roundtrip(oracle, name):
    data, comp, _ = _load(name)
    assert oracle.decompress(comp) == data


def test_known_answer_english_dic_16k(oracle):
    known = json.load(open(os.path.join(GOLD, "known_answers.json")))["english_dic_16k"]
    data = open(os.path.join(ROOT, "tests", "data", "english.dic"), "rb").read()[:known["input_bytes"]]
    out = oracle.compress(data)
    assert len(out) == known["output_bytes"]
    assert hashlib.md5(out).hexdigest() == known["md5"]


def test_header_is_five_byte_big_endian_length(oracle):
    out = oracle.compress(b"x" * 300)
    assert out[:5] == (300).to_bytes(5, "big")   # runner-utils.cpp:22-27


@pytest.mark.skipif(not os.path.exists(REF), reason="reference build (oracle/_ref) not present")
def test_oracle_matches_live_reference_full_blackboard_trace():
    """Every one of the 90 predictions, the active set and the 33 mixer outputs, bit for bit."""
    data = open(os.path.join(ROOT, "tests", "data", "english.dic"), "rb").read()[3000:3450]
    ora = os.path.join(ROOT, "oracle", "_build", "gmix_oracle")
    with tempfile.TemporaryDirectory() as d:
        inp = os.path.join(d, "in")
        open(inp, "wb").write(data)
        subprocess.run([REF, "trace", inp, os.path.join(d, "ref.tr"), "3"], check=True, capture_output=True)
        subprocess.run([ora, "trace", inp, os.path.join(d, "ora.tr"), "3"], check=True, capture_output=True)
        assert open(os.path.join(d, "ref.tr"), "rb").read() == open(os.path.join(d, "ora.tr"), "rb").read()
