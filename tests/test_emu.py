"""Runs the product's CUDA stream kernel source under the CPU SIMT emulator (tests/emu/cuda_emu.h)
and checks it against the golden vectors: this is how kernel logic is debugged without a GPU. The
emulator executes the very same .cuh files nvcc compiles; it is test infrastructure, not a product
path."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emu") / "emu_main")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-o", exe, os.path.join(HERE, "emu", "emu_main.cpp")], check=True)
    return exe


@pytest.mark.parametrize("name", ["one_byte", "empty", "short124", "text1k", "repetitive"])
def test_emulated_kernel_matches_reference_stream(emu, tmp_path, name):
    out = str(tmp_path / "out")
    subprocess.run([emu, "compress", os.path.join(GOLD, name + ".in"), out], check=True)
    assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".gmix"), "rb").read()
    back = str(tmp_path / "back")
    subprocess.run([emu, "decompress", out, back], check=True)
    assert open(back, "rb").read() == open(os.path.join(GOLD, name + ".in"), "rb").read()


@pytest.fixture(scope="module")
def emu_deferred(tmp_path_factory):
    """Same kernel source, cp.async emulated at its LATEST legal moment (the copy runs when the issuing thread waits
    for it, tests/emu/cuda_emu.h) instead of the earliest: together with the default build this brackets the
    timing freedom the hardware has."""
    exe = str(tmp_path_factory.mktemp("emu_defer") / "emu_main")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-DGMX_EMU_DEFER_CP", "-o", exe, os.path.join(HERE, "emu", "emu_main.cpp")], check=True)
    return exe


@pytest.mark.parametrize("order", ["", "reverse", "shuffle5"])
def test_kernel_is_insensitive_to_copy_timing_and_thread_interleaving(emu_deferred, tmp_path, order):
    env = dict(os.environ)
    if order:
        env["EMU_ORDER"] = order     # threads between two barriers scheduled in another legal order
    for name in ("text1k", "repetitive"):
        out = str(tmp_path / (name + ".out"))
        subprocess.run([emu_deferred, "compress", os.path.join(GOLD, name + ".in"), out], check=True, env=env, stderr=subprocess.DEVNULL)
        assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".gmix"), "rb").read(), (name, order)
    back = str(tmp_path / "back")
    subprocess.run([emu_deferred, "decompress", os.path.join(GOLD, "text1k.gmix"), back], check=True, env=env, stderr=subprocess.DEVNULL)
    assert open(back, "rb").read() == open(os.path.join(GOLD, "text1k.in"), "rb").read()
