"""Deterministic synthetic inputs of the shapes BASELINE.json names.

Word source = dictionary/english.dic (one lower-case word per line; shipped as tests/data/english.dic
because the GPU box only sees this repository). PRNG = splitmix64.
  * synthetic_text_chunk(i, size): Zipf-like word stream, punctuation, sentence capitalisation,
    line breaks after >= 72 columns, truncated to exactly `size` bytes. Seed 0x676D697800000000 + i.
  * enwik_shaped_corpus(total): the same word stream wrapped in <page>/<title>/<id>/<text> markup
    with [[links]], '''bold''' and == headings ==; article lengths log-uniform 500 B..20 KB.
"""
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_WORDS = None
MASK = (1 << 64) - 1


def words():
    global _WORDS
    if _WORDS is None:
        with open(os.path.join(_ROOT, "tests", "data", "english.dic"), "rb") as f:
            _WORDS = [w for w in f.read().split(b"\n") if w]
    return _WORDS


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def uniform(self):
        return (self.next() >> 11) * (1.0 / (1 << 53))


def _word(rng, ws):
    u = rng.uniform()
    return ws[int(len(ws) * u ** 2.5)]


def synthetic_text_chunk(i, size=65536):
    ws = words()
    rng = SplitMix64(0x676D697800000000 + i)
    out = bytearray()
    col = 0
    cap = True
    while len(out) < size:
        w = _word(rng, ws)
        if cap:
            w = w[:1].upper() + w[1:]
            cap = False
        out += w
        col += len(w)
        r = rng.uniform()
        if r < 1.0 / 15:
            out += b". "
            col += 2
            cap = True
        elif r < 1.0 / 15 + 1.0 / 12:
            out += b", "
            col += 2
        else:
            out += b" "
            col += 1
        if col >= 72:
            out += b"\n"
            col = 0
    return bytes(out[:size])


def enwik_shaped_corpus(total, seed=0x656E77696B000000):
    ws = words()
    rng = SplitMix64(seed)
    out = bytearray()
    page = 0
    while len(out) < total:
        page += 1
        target = int(500 * (40.0 ** rng.uniform()))  # log-uniform 500 B .. 20 KB
        title = _word(rng, ws).capitalize() + b" " + _word(rng, ws).capitalize()
        out += b"<page>\n  <title>" + title + b"</title>\n  <id>" + str(page * 7 + 11).encode() + b"</id>\n  <text>"
        start = len(out)
        col = 0
        while len(out) - start < target:
            r = rng.uniform()
            w = _word(rng, ws)
            if r < 0.04:
                w = b"[[" + w + b" " + _word(rng, ws) + b"]]"
            elif r < 0.06:
                w = b"'''" + w + b"'''"
            elif r < 0.07:
                w = b"\n== " + w.capitalize() + b" ==\n"
                col = 0
            out += w
            col += len(w)
            r2 = rng.uniform()
            out += b". " if r2 < 1.0 / 15 else b", " if r2 < 1.0 / 15 + 1.0 / 12 else b" "
            if col >= 72:
                out += b"\n"
                col = 0
        out += b"</text>\n</page>\n"
    return bytes(out[:total])
