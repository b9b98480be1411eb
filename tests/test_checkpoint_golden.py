"""CPU: the committed checkpoint fixtures (written by the UNMODIFIED reference, tests/golden/make_golden_ckpt.py) drive
the product's kernel source under the SIMT emulator: load, continue compressing, generate. Unlike
test_checkpoint_emu.py this needs no reference build, so it also pins the format on machines without /root/reference."""
import gzip
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    d = tmp_path_factory.mktemp("ckptgold")
    exe = str(d / "emu_main")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-o", exe, os.path.join(HERE, "emu", "emu_main.cpp")], check=True)
    for ext in (".short", ".long"):
        (d / ("ckpt600" + ext)).write_bytes(gzip.open(os.path.join(GOLD, "ckpt600" + ext + ".gz")).read())
    (d / "b.in").write_bytes(open(os.path.join(GOLD, "text1k.in"), "rb").read()[600:])
    return d, exe


def test_fixture_checkpoint_resumes_to_the_reference_stream(work):
    d, exe = work
    subprocess.run([exe, "resume", str(d / "ckpt600"), str(d / "b.in"), str(d / "b.gmix")], check=True, stderr=subprocess.DEVNULL)
    assert (d / "b.gmix").read_bytes() == open(os.path.join(GOLD, "ckpt600_b.gmix"), "rb").read()


def test_fixture_checkpoint_recodes_to_itself(work):
    d, exe = work
    subprocess.run([exe, "recode", str(d / "ckpt600"), str(d / "rec")], check=True, stderr=subprocess.DEVNULL)
    assert (d / "rec.short").read_bytes() == (d / "ckpt600.short").read_bytes()
    assert (d / "rec.long").read_bytes() == (d / "ckpt600.long").read_bytes()


def test_fixture_generation(work):
    d, exe = work
    subprocess.run([exe, "generate", str(d / "ckpt600"), os.path.join(GOLD, "ckpt600_prompt.txt"), str(d / "gen.out"), "40", "0.5"],
                   check=True, stderr=subprocess.DEVNULL)
    assert (d / "gen.out").read_bytes() == open(os.path.join(GOLD, "ckpt600_gen_40_0.5.out"), "rb").read()


def test_malformed_checkpoints_are_rejected_not_crashed(work):
    """Truncated or padded files fail Parse with a message (the C ABI turns that into GMX_E_ARG); nothing is read out of bounds."""
    d, exe = work
    short, long_ = (d / "ckpt600.short").read_bytes(), (d / "ckpt600.long").read_bytes()
    cases = {"short_cut": (short[:len(short) // 2], long_), "long_cut": (short, long_[:1000]), "short_pad": (short + b"\0" * 7, long_),
             "empty": (b"", b""), "long_garbage_count": (short, b"\xff\xff\xff\x7f" + long_[4:])}
    for name, (s, l) in cases.items():
        (d / (name + ".short")).write_bytes(s)
        (d / (name + ".long")).write_bytes(l)
        r = subprocess.run([exe, "recode", str(d / name), str(d / (name + "_out"))], capture_output=True, text=True)
        assert r.returncode == 1 and "Parse:" in r.stderr, (name, r.returncode, r.stderr[-200:])


def test_dense_table_layout_with_checkpoints(work):
    """Long streams get dense tables instead of the shared sparse map (layout.h); forced here on short inputs
    (EMU_DENSE): compress from scratch, load a reference checkpoint, continue, and write one back."""
    d, exe = work
    env = dict(os.environ, EMU_DENSE="1")
    subprocess.run([exe, "compress", os.path.join(GOLD, "short124.in"), str(d / "d.gmix")], check=True, env=env, stderr=subprocess.DEVNULL)
    assert (d / "d.gmix").read_bytes() == open(os.path.join(GOLD, "short124.gmix"), "rb").read()
    subprocess.run([exe, "resume", str(d / "ckpt600"), str(d / "b.in"), str(d / "bd.gmix")], check=True, env=env, stderr=subprocess.DEVNULL)
    assert (d / "bd.gmix").read_bytes() == open(os.path.join(GOLD, "ckpt600_b.gmix"), "rb").read()
    a = open(os.path.join(GOLD, "text1k.in"), "rb").read()[:600]
    (d / "a.in").write_bytes(a)
    subprocess.run([exe, "train", str(d / "a.in"), str(d / "dense_ck")], check=True, env=env, stderr=subprocess.DEVNULL)
    assert (d / "dense_ck.long").read_bytes() == (d / "ckpt600.long").read_bytes()
