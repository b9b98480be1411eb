"""Dictionary preprocessing (reference src/preprocess/dictionary.cpp + src/runner/dictionary-prep.cpp):
gmix_b200/host/dictionary.h is host-only and must be byte-compatible with the reference's tool in BOTH directions. Checker: the
unmodified reference tool built by oracle/Makefile (oracle/_ref/dictionary-prep) on crafted and random inputs, live; where it is
absent, the committed vectors tests/golden/dict_prep_* (made by tests/golden/make_golden_dict.py from that same tool)."""
import os
import random
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
DIC = os.path.join(HERE, "data", "english.dic")
REF = os.path.join(ROOT, "oracle", "_ref", "dictionary-prep")


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("dict") / "dictionary-prep")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "gmix_b200", "host", "dictionary_prep.cpp")], check=True)
    return exe


def run(exe, mode, data, tmp_path, dic=DIC):
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(data)
    subprocess.run([exe, mode, dic, str(src), str(dst)], check=True, stdout=subprocess.DEVNULL)
    return dst.read_bytes()


def crafted():
    rng = random.Random(20)
    words = [w for w in open(DIC, "rb").read().split() if w.isalpha()]
    text = bytearray()
    for _ in range(3000):
        w = rng.choice(words)
        style = rng.randrange(8)
        if style == 0: w = w.capitalize()
        elif style == 1: w = w.upper()
        elif style == 2: w = w.upper() + rng.choice(words)[:3]            # UPPER run directly followed by lower case
        elif style == 3: w = w + rng.choice(words)                          # long compound: suffix / prefix search
        elif style == 4: w = w[:1].upper() + w[1:2].upper() + w[2:]         # two capitals, then lower case
        text += w + rng.choice([b" ", b", ", b".\n", b" &quot;", b"&quot; ", b"&quo t", b"&&quot;", b" \x07\x06\x0c\x08@ ", b" \xc3\xa9 ", b"1", b"-"])
    return bytes(text)


CASES = {
    "empty": b"", "one_letter": b"a", "one_capital": b"A", "one_other": b"\x0c", "quote_only": b"&quot;", "word_at_end": b"the quick Brown FOX",
    "upper_then_lower": b"HELLOworld HELLO world ABc", "long": b"x" * 300 + b" " + b"internationalization" * 3,
    "markers": bytes(range(256)) * 2, "text1k": open(os.path.join(GOLD, "text1k.in"), "rb").read(), "crafted": None, "random": None,
}


def case_data(name):
    if name == "crafted":
        return crafted()
    if name == "random":
        return bytes(random.Random(5).randrange(256) for _ in range(20000))
    return CASES[name]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/dictionary-prep not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(CASES))
def test_encode_and_decode_match_the_reference_tool(tool, tmp_path, name):
    data = case_data(name)
    ours, theirs = run(tool, "-e", data, tmp_path), run(REF, "-e", data, tmp_path)
    assert ours == theirs
    # decoding: their encoding with our decoder, and ARBITRARY bytes (the raw input) through both decoders
    assert run(tool, "-d", theirs, tmp_path) == run(REF, "-d", theirs, tmp_path)
    assert run(tool, "-d", data, tmp_path) == run(REF, "-d", data, tmp_path)


@pytest.mark.parametrize("name", ["text1k", "crafted", "word_at_end", "upper_then_lower"])
def test_round_trip(tool, tmp_path, name):
    data = case_data(name)
    assert run(tool, "-d", run(tool, "-e", data, tmp_path), tmp_path) == data


def test_committed_vectors(tool, tmp_path):
    data = open(os.path.join(GOLD, "dict_prep_crafted.txt"), "rb").read()
    want = open(os.path.join(GOLD, "dict_prep_crafted.enc"), "rb").read()
    assert run(tool, "-e", data, tmp_path) == want
    assert run(tool, "-d", want, tmp_path) == data


def test_small_dictionary_and_code_lengths(tool, tmp_path):
    """1-, 2- and 3-byte codes: rank < 80, < 3920, beyond (dictionary.cpp:45-68)."""
    words = [w for w in open(DIC, "rb").read().split() if w.isalpha()]
    data = b" ".join([words[3], words[79], words[80], words[3919], words[3920], words[44000], words[-1]]) + b"\n"
    enc = run(tool, "-e", data, tmp_path)
    assert len(enc) == (1 + 1 + 2 + 2 + 3 + 3 + 3) + 7
    assert run(tool, "-d", enc, tmp_path) == data
    if os.path.exists(REF):
        assert enc == run(REF, "-e", data, tmp_path)
