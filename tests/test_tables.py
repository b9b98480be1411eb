"""Data constants that must equal the reference's: the Nonstationary automaton (a literal table in
src/contexts/nonstationary.cpp) and the RunMap rule (src/contexts/run-map.cpp:3-21)."""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _inc(path):
    txt = open(path).read()
    txt = "\n".join(l for l in txt.splitlines() if not l.startswith("//"))
    return bytes(int(x) for x in re.findall(r"\d+", txt))


def test_nonstationary_tables_match_reference_dump():
    gold = open(os.path.join(HERE, "golden", "state_tables.bin"), "rb").read()
    assert len(gold) == 1024
    for p in ("oracle/nonstationary.inc", "gmix_b200/csrc/nonstationary.inc"):
        assert _inc(os.path.join(ROOT, p)) == gold[:512], p


def test_run_map_rule_matches_reference_dump():
    gold = open(os.path.join(HERE, "golden", "state_tables.bin"), "rb").read()[512:]
    for state in range(256):
        for bit in range(2):
            if bit == 0:
                nxt = state + 1 if state < 127 else 1 if state >= 128 else state
            else:
                nxt = 128 if state < 128 else state + 1 if state < 255 else state
            assert gold[state * 2 + bit] == nxt
