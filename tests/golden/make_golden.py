"""Generates the golden vectors in this directory by running the UNMODIFIED reference
(oracle/_ref/ref_driver, built from /root/reference by oracle/Makefile with the canonical strict
flags). Run here (the container that has /root/reference); the GPU box only reads the fixtures.

  <name>.in    input bytes
  <name>.gmix  exactly what `gmix -c` writes (5-byte header + coder bytes)
  <name>.p16   per input bit, the coder's 16-bit probability (u16 little endian) = Discretize(Predict())
  known_answers.json  sizes/md5 of larger reference runs (english.dic prefixes)
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
DIC = open(os.path.join(ROOT, "tests", "data", "english.dic"), "rb").read()


def cases():
    from gmix_b200.synth import synthetic_text_chunk
    rng = np.random.RandomState(1234)
    rep = b"".join(b"abcabcabd" * 3 + bytes([i % 7, 255 - (i % 5)]) for i in range(120))
    return {
        "one_byte": b"a",
        "empty": b"",
        "short124": DIC[:124],
        "text1k": DIC[:1024],
        "text_mid": DIC[100000:102500],
        "random1200": rng.randint(0, 256, 1200, dtype=np.uint8).tobytes(),
        "zeros900": bytes(900),
        "repetitive": rep,
        "synth_chunk0_4k": synthetic_text_chunk(0, 4096),
    }


def run_ref(data, want_trace=True):
    with tempfile.TemporaryDirectory() as d:
        inp, out, tr = os.path.join(d, "in"), os.path.join(d, "out"), os.path.join(d, "tr")
        open(inp, "wb").write(data)
        subprocess.run([REF, "compress", inp, out], check=True, capture_output=True)
        comp = open(out, "rb").read()
        p16 = None
        if want_trace:
            subprocess.run([REF, "trace", inp, tr, "1"], check=True, capture_output=True)
            rec = np.fromfile(tr, dtype=np.uint32).reshape(-1, 2)
            p16 = rec[:, 1].astype(np.uint16)
        return comp, p16


def main():
    for name, data in cases().items():
        comp, p16 = run_ref(data)
        open(os.path.join(HERE, name + ".in"), "wb").write(data)
        open(os.path.join(HERE, name + ".gmix"), "wb").write(comp)
        p16.tofile(os.path.join(HERE, name + ".p16"))
        print(name, len(data), "->", len(comp))
    known = {}
    for label, n in (("english_dic_16k", 16384), ("english_dic_64k", 65536), ("english_dic_full", len(DIC))):
        if "--full" not in sys.argv and n > 16384:
            continue
        comp, _ = run_ref(DIC[:n], want_trace=False)
        known[label] = {"input_bytes": n, "output_bytes": len(comp), "md5": hashlib.md5(comp).hexdigest()}
        print(label, known[label])
    path = os.path.join(HERE, "known_answers.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(known)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
