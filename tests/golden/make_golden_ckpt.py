"""Generates the checkpoint / generation fixtures by running the UNMODIFIED reference (oracle/_ref/{gmix,ref_driver},
built from /root/reference by oracle/Makefile). Run here; the GPU box only reads the fixtures.

  ckpt600.short.gz / ckpt600.long.gz  what Predictor::WriteCheckpoint writes after Predict/Perceive/Learn over
                                      text1k.in[:600] (ref_driver train), gzip-compressed
  ckpt600_b.gmix                      `gmix -c ckpt600 b.in out` for b.in = text1k.in[600:]
  ckpt600_gen_<size>_<temp>.out       `gmix -g ckpt600 prompt out <size> <temp>` for the prompt in ckpt600_prompt.txt
"""
import gzip
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
REF_GMIX = os.path.join(ROOT, "oracle", "_ref", "gmix")
GENS = [(48, "1.0"), (40, "0.5")]
PROMPT = b"the quick brown fox jumps over\n"


def main():
    text = open(os.path.join(HERE, "text1k.in"), "rb").read()
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "a.in"), "wb").write(text[:600])
        open(os.path.join(d, "b.in"), "wb").write(text[600:])
        open(os.path.join(d, "prompt.txt"), "wb").write(PROMPT)
        subprocess.run([REF_DRIVER, "train", os.path.join(d, "a.in"), os.path.join(d, "ckpt600")], check=True, capture_output=True)
        for ext in (".short", ".long"):
            with open(os.path.join(d, "ckpt600" + ext), "rb") as f, gzip.GzipFile(os.path.join(HERE, "ckpt600" + ext + ".gz"), "wb", 9, mtime=0) as g:
                shutil.copyfileobj(f, g)
        subprocess.run([REF_GMIX, "-c", "ckpt600", "b.in", "b.gmix"], cwd=d, check=True, capture_output=True)
        shutil.copy(os.path.join(d, "b.gmix"), os.path.join(HERE, "ckpt600_b.gmix"))
        open(os.path.join(HERE, "ckpt600_prompt.txt"), "wb").write(PROMPT)
        for size, temp in GENS:
            subprocess.run([REF_GMIX, "-g", "ckpt600", "prompt.txt", "gen.out", str(size), temp], cwd=d, check=True, capture_output=True)
            shutil.copy(os.path.join(d, "gen.out"), os.path.join(HERE, f"ckpt600_gen_{size}_{temp}.out"))
            print(size, temp, open(os.path.join(d, "gen.out"), "rb").read())


if __name__ == "__main__":
    main()


def make_training_golden():
    """`gmix -t text1k.in short124.in` (RunTraining runner-utils.cpp:223-322, ~200 s): analysis/training.tsv, data/tmp and
    the md5 of data/trained_checkpoint.long -> train_text1k_short124.*"""
    import hashlib
    import json
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([REF_GMIX, "-t", os.path.join(HERE, "text1k.in"), os.path.join(HERE, "short124.in")], cwd=d, check=True, capture_output=True)
        shutil.copy(os.path.join(d, "analysis", "training.tsv"), os.path.join(HERE, "train_text1k_short124.training.tsv"))
        shutil.copy(os.path.join(d, "data", "tmp"), os.path.join(HERE, "train_text1k_short124.tmp"))
        blob = open(os.path.join(d, "data", "trained_checkpoint.long"), "rb").read()
        json.dump({"long_md5": hashlib.md5(blob).hexdigest(), "long_bytes": len(blob),
                   "made_by": "oracle/_ref/gmix -t tests/golden/text1k.in tests/golden/short124.in (unmodified reference, strict build)"},
                  open(os.path.join(HERE, "train_text1k_short124.json"), "w"), indent=1)
