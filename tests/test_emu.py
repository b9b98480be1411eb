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


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_driver")), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["short124", "text1k"])     # analysis off (< 125 B) and on
def test_full_blackboard_trace_matches_the_live_reference(emu, tmp_path, name):
    """Per bit: Predict()'s probability, the coder's 16-bit probability, all 90 stretched predictions, the active-model
    mask and the 24 + 8 + 1 mixer outputs of the product kernel are bit-identical to the unmodified reference's
    (oracle/_ref/ref_driver trace level 2): this pins the model graph and the order of every model, not just the bytes."""
    import numpy as np
    src = os.path.join(GOLD, name + ".in")
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_driver"), "trace", src, str(tmp_path / "ref.tr"), "2"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([emu, "compress", src, str(tmp_path / "o"), str(tmp_path / "emu.tr"), str(tmp_path / "emu.ptr")], check=True, stderr=subprocess.DEVNULL)
    ref = np.fromfile(str(tmp_path / "ref.tr"), dtype=np.uint32).reshape(-1, 128)
    got = np.fromfile(str(tmp_path / "emu.tr"), dtype=np.uint32).reshape(-1, 128)
    assert ref.shape == got.shape == (8 * os.path.getsize(src), 128)
    bad = np.argwhere(ref != got)
    assert bad.size == 0, f"first difference at bit {bad[0][0]}, word {bad[0][1]}"


def test_second_stream_of_a_persistent_cta_reuses_the_arena(emu, tmp_path):
    """A batch with more streams than arenas makes a CTA compress a second stream in the arena and shared memory the
    first one left behind: the second stream's bytes must not depend on that."""
    out = str(tmp_path / "second.gmix")
    subprocess.run([emu, "compress2", os.path.join(GOLD, "text_mid.in"), os.path.join(GOLD, "text1k.in"), out], check=True, stderr=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(GOLD, "text1k.gmix"), "rb").read()


@pytest.mark.parametrize("name", ["short124", "text1k"])     # analysis off (< 125 B) and on
def test_stepping_kernel_behind_the_predictor_facade(emu, tmp_path, name):
    """StepKernel (one launch per Predictor::Predict / Learn, the three roles run one after the other) + the host coder
    reproduce the reference's stream."""
    out = str(tmp_path / "steps.gmix")
    subprocess.run([emu, "steps", os.path.join(GOLD, name + ".in"), out], check=True, stderr=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".gmix"), "rb").read()


def test_serial_compress_configuration(tmp_path):
    """Configuration 0 of the library: compress without the role pipeline (all threads walk the phases together, the byte
    models still hand their path nodes over through packet 0)."""
    exe = str(tmp_path / "emu_main")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-DEMU_SERIAL=1", "-o", exe, os.path.join(HERE, "emu", "emu_main.cpp")], check=True)
    for name in ("text1k", "random1200", "short124", "one_byte", "empty"):
        out = str(tmp_path / (name + ".out"))
        subprocess.run([exe, "compress", os.path.join(GOLD, name + ".in"), out], check=True, stderr=subprocess.DEVNULL)
        assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".gmix"), "rb").read(), name


def test_hybrid_compress_configuration(tmp_path):
    """Configuration 10 of the library: the PPMd warp prepares the next byte under the bit path of this one, the LSTM phases run
    with all threads (HybridCompress) - same bytes in every legal interleaving of the two sides (thread order between barriers,
    cp.async at its latest legal moment), and through the parts / resume protocols that start mid-stream."""
    exe = str(tmp_path / "emu_main")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-DEMU_WB=3", "-DEMU_WL=0", "-DEMU_SERIAL=1", "-DGMX_EMU_DEFER_CP", "-o", exe,
                    os.path.join(HERE, "emu", "emu_main.cpp")], check=True)
    for order in ("", "reverse", "shuffle5"):
        env = dict(os.environ)
        if order:
            env["EMU_ORDER"] = order
        for name in ("text1k", "random1200", "short124", "one_byte", "empty") if not order else ("text1k", "repetitive"):
            out = str(tmp_path / (name + ".out"))
            subprocess.run([exe, "compress", os.path.join(GOLD, name + ".in"), out], check=True, stderr=subprocess.DEVNULL, env=env)
            assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".gmix"), "rb").read(), (name, order)
    out = str(tmp_path / "p.out")
    subprocess.run([exe, "parts", os.path.join(GOLD, "text1k.in"), out, "517"], check=True, stderr=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(GOLD, "text1k.gmix"), "rb").read()


def test_latency_configuration(tmp_path):
    """Configuration 9 of the library: one CTA per SM, roles 4 / 4 / 1, the dense gate weights resident in (emulated) shared
    memory and refreshed by bulk copies after every Adam step, straight-line mixer network (LAT), whole-byte prefetch."""
    exe = str(tmp_path / "emu_main")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-DEMU_WB=4", "-DEMU_WL=4", "-DEMU_WS=1", "-DEMU_MINB=1", "-o", exe,
                    os.path.join(HERE, "emu", "emu_main.cpp")], check=True)
    for name in ("text1k", "random1200", "repetitive"):
        out = str(tmp_path / (name + ".out"))
        subprocess.run([exe, "compress", os.path.join(GOLD, name + ".in"), out], check=True, stderr=subprocess.DEVNULL)
        assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".gmix"), "rb").read(), name
    back = str(tmp_path / "back")
    subprocess.run([exe, "decompress", os.path.join(GOLD, "text1k.gmix"), back], check=True, stderr=subprocess.DEVNULL)
    assert open(back, "rb").read() == open(os.path.join(GOLD, "text1k.in"), "rb").read()


@pytest.mark.parametrize("roles", [(1, 2), (1, 1), (2, 2), (3, 0)])   # (3, 0): the two-role variant, PPMd ahead of everything else
def test_other_role_splits_compute_the_same_bytes(tmp_path, roles):
    """The kernel configurations of the library (kernels.h) differ only in how many warps the bit role and the LSTM role
    get: every split must compute the reference's bytes, in both pipeline protocols (ahead: compress, lockstep: decompress)."""
    wb, wl = roles
    exe = str(tmp_path / "emu_main")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", f"-DEMU_WB={wb}", f"-DEMU_WL={wl}", "-o", exe,
                    os.path.join(HERE, "emu", "emu_main.cpp")], check=True)
    for name in ("text1k", "random1200"):
        out = str(tmp_path / (name + ".out"))
        subprocess.run([exe, "compress", os.path.join(GOLD, name + ".in"), out], check=True, stderr=subprocess.DEVNULL)
        assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".gmix"), "rb").read(), (roles, name)
    back = str(tmp_path / "back")
    subprocess.run([exe, "decompress", os.path.join(GOLD, "text1k.gmix"), back], check=True, stderr=subprocess.DEVNULL)
    assert open(back, "rb").read() == open(os.path.join(GOLD, "text1k.in"), "rb").read()


def test_stream_coded_in_two_parts_with_checkpoints(emu, tmp_path):
    """Encoder/Decoder::Write/ReadCheckpoint + Predictor::Write/ReadCheckpoint mid-stream (the reference's restart tests,
    tester.cpp:329-356) at kernel level: part 1 without flush, the stream's state through the reference's file format,
    part 2 from it with the saved coder state; the concatenation is the whole-stream output."""
    for name, split in (("text1k", 517), ("random1200", 1), ("random1200", 1199)):
        out = str(tmp_path / "p.out")
        subprocess.run([emu, "parts", os.path.join(GOLD, name + ".in"), out, str(split)], check=True, stderr=subprocess.DEVNULL)
        assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".gmix"), "rb").read(), (name, split)
    back = str(tmp_path / "u.out")
    subprocess.run([emu, "unparts", os.path.join(GOLD, "text1k.gmix"), back, "333"], check=True, stderr=subprocess.DEVNULL)
    assert open(back, "rb").read() == open(os.path.join(GOLD, "text1k.in"), "rb").read()
