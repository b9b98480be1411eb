"""Writes tests/golden/dict_prep_crafted.{txt,enc}: a crafted text and its encoding by the UNMODIFIED reference tool
(oracle/_ref/dictionary-prep, built by oracle/Makefile from /root/reference/src/preprocess/dictionary.cpp)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import test_dictionary_prep as t  # noqa: E402

data = t.crafted()[:6000]
src, dst = os.path.join(HERE, "dict_prep_crafted.txt"), os.path.join(HERE, "dict_prep_crafted.enc")
open(src, "wb").write(data)
subprocess.run([t.REF, "-e", t.DIC, src, dst], check=True)
print(len(data), "->", os.path.getsize(dst))
