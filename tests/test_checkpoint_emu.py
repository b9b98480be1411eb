"""Checkpoint format (reference src/predictor.cpp:389-420, src/memory/*.cpp) and generation
(runner-utils.cpp:158-221), checked on the CPU: the product's kernel source runs under the SIMT emulator
(tests/emu) and the UNMODIFIED reference (oracle/_ref, built from /root/reference by oracle/Makefile) is the
checker. Covered: files written by the reference load unchanged and continue byte-identically
(`gmix -c/-d <ckpt>`), files we write are loaded by the reference and continue byte-identically, `.long` is
byte-identical, `.short` differs only in documented scratch fields, Parse -> Serialize is the identity, and
`gmix -g` samples the same bytes."""
import os
import subprocess

import pytest

import ckpt_layout

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_DRIVER = os.path.join(REF_DIR, "ref_driver")
REF_GMIX = os.path.join(REF_DIR, "gmix")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_GMIX), reason="oracle/_ref not built (needs /root/reference)")


def run_ref(cwd, *args):
    subprocess.run([REF_GMIX, *args], cwd=cwd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    d = tmp_path_factory.mktemp("ckpt")
    exe = str(d / "emu_main")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-o", exe, os.path.join(HERE, "emu", "emu_main.cpp")], check=True)
    text = open(os.path.join(GOLD, "text1k.in"), "rb").read()
    (d / "a.in").write_bytes(text[:600])      # six BPTT passes, PPMd tree, a few thousand mixer sets
    (d / "b.in").write_bytes(text[600:])
    (d / "empty.in").write_bytes(b"")
    (d / "prompt.txt").write_bytes(b"the quick brown fox jumps over\n")
    subprocess.run([REF_DRIVER, "train", str(d / "a.in"), str(d / "ref_a")], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([REF_DRIVER, "train", str(d / "empty.in"), str(d / "ref_e")], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([exe, "train", str(d / "a.in"), str(d / "emu_a")], check=True, stderr=subprocess.DEVNULL)
    return d, exe


def emu(exe, *args):
    subprocess.run([exe, *[str(a) for a in args]], check=True, stderr=subprocess.DEVNULL)


def test_parse_then_serialize_reproduces_reference_files(work):
    d, exe = work
    for name in ("ref_a", "ref_e"):
        emu(exe, "recode", d / name, d / (name + "_rec"))
        for ext in (".short", ".long"):
            assert (d / (name + "_rec" + ext)).read_bytes() == (d / (name + ext)).read_bytes(), name + ext


def test_reference_checkpoint_loads_and_compress_continues_identically(work):
    d, exe = work
    run_ref(d, "-c", "ref_a", "b.in", "ref_b.gmix")
    emu(exe, "resume", d / "ref_a", d / "b.in", d / "emu_b.gmix")
    assert (d / "emu_b.gmix").read_bytes() == (d / "ref_b.gmix").read_bytes()
    emu(exe, "expand", d / "ref_a", d / "ref_b.gmix", d / "emu_b.back")      # gmix -d <ckpt>
    assert (d / "emu_b.back").read_bytes() == (d / "b.in").read_bytes()


def test_checkpoint_of_a_fresh_predictor(work):
    d, exe = work
    run_ref(d, "-c", "ref_e", "a.in", "ref_a_from_e.gmix")
    emu(exe, "resume", d / "ref_e", d / "a.in", d / "emu_a_from_e.gmix")
    assert (d / "emu_a_from_e.gmix").read_bytes() == (d / "ref_a_from_e.gmix").read_bytes()
    emu(exe, "train", d / "empty.in", d / "emu_e")
    assert (d / "emu_e.long").read_bytes() == (d / "ref_e.long").read_bytes()
    diff = ckpt_layout.differing_sections((d / "emu_e.short").read_bytes(), (d / "ref_e.short").read_bytes())
    # FoundState is a null pointer in a fresh model; the reference serialises it (and AuxUnit, saved_pc) as `0 - HeapStart`
    assert set(diff) <= {"ppmd.aux_unit", "ppmd.found_state", "ppmd.saved_pc"}, diff


def test_written_checkpoint_matches_reference_and_is_resumed_by_it(work):
    d, exe = work
    assert (d / "emu_a.long").read_bytes() == (d / "ref_a.long").read_bytes()
    diff = ckpt_layout.differing_sections((d / "emu_a.short").read_bytes(), (d / "ref_a.short").read_bytes())
    assert set(diff) <= ckpt_layout.SCRATCH, [x for x in diff if x not in ckpt_layout.SCRATCH]
    run_ref(d, "-c", "ref_a", "b.in", "ref_b.gmix")
    run_ref(d, "-c", "emu_a", "b.in", "ref_b_from_ours.gmix")
    assert (d / "ref_b_from_ours.gmix").read_bytes() == (d / "ref_b.gmix").read_bytes()


def test_training_continues_from_a_checkpoint(work):
    d, exe = work
    (d / "ab.in").write_bytes((d / "a.in").read_bytes() + (d / "b.in").read_bytes())
    subprocess.run([REF_DRIVER, "train", str(d / "ab.in"), str(d / "ref_ab")], check=True, stdout=subprocess.DEVNULL)
    emu(exe, "train", d / "b.in", d / "emu_ab", d / "ref_a")
    assert (d / "emu_ab.long").read_bytes() == (d / "ref_ab.long").read_bytes()
    diff = ckpt_layout.differing_sections((d / "emu_ab.short").read_bytes(), (d / "ref_ab.short").read_bytes())
    assert set(diff) <= ckpt_layout.SCRATCH, diff


@pytest.mark.parametrize("size,temp", [(48, "1.0"), (40, "0.5"), (24, "0.0001")])
def test_generation_samples_the_reference_bytes(work, size, temp):
    d, exe = work
    run_ref(d, "-g", "ref_a", "prompt.txt", f"ref_gen_{size}.out", str(size), temp)
    emu(exe, "generate", d / "ref_a", d / "prompt.txt", d / f"emu_gen_{size}.out", size, temp)
    got, want = (d / f"emu_gen_{size}.out").read_bytes(), (d / f"ref_gen_{size}.out").read_bytes()
    assert len(want) == size and got == want
    # and in overlay mode (what the batched generation call runs): the model's tables stay read-only and shared, the
    # stream keeps its changes in an overlay map / local pool / private copies of the small state
    # (both PPMd backings of an overlay arena: the model's window copied whole, and the segmented one big models get)
    # ... and as LOCK-STEP generation (host.cu RunLockstepGenerate): the prompt launch stops in front of the first sampled byte's
    # gate product, then one exact batched gate product (gate_gemm.cuh GateDotsExactKernel) + one GenStepKernel launch per byte
    for extra in ({}, {"EMU_SEGMENTED": "1"}, {"EMU_LOCKSTEP": "1"}, {"EMU_LOCKSTEP": "1", "EMU_SEGMENTED": "1"}):
        env = dict(os.environ, EMU_OVERLAY="1", **extra)
        subprocess.run([exe, "generate", str(d / "ref_a"), str(d / "prompt.txt"), str(d / f"ov_gen_{size}.out"), str(size), temp], check=True,
                       stderr=subprocess.DEVNULL, env=env)
        assert (d / f"ov_gen_{size}.out").read_bytes() == want, extra
