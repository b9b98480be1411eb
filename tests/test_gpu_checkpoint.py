"""GPU: checkpoints in the reference's format and batched generation, through the C ABI, against fixtures produced by
the UNMODIFIED reference (tests/golden/make_golden_ckpt.py): a reference-written checkpoint loads unchanged and
`gmix -c/-d <ckpt>` / `gmix -g` reproduce the reference's bytes; the checkpoint the GPU writes has the reference's
`.long` bytes and a `.short` that differs only in documented scratch fields; the Predictor facade reads and writes
the same files."""
import gzip
import os

import pytest

import ckpt_layout

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
ROOT = os.path.dirname(HERE)
TEXT = open(os.path.join(GOLD, "text1k.in"), "rb").read()
A, B = TEXT[:600], TEXT[600:]
PROMPT = open(os.path.join(GOLD, "ckpt600_prompt.txt"), "rb").read()


@pytest.fixture(scope="module")
def ckpt():
    return gzip.open(os.path.join(GOLD, "ckpt600.short.gz")).read(), gzip.open(os.path.join(GOLD, "ckpt600.long.gz")).read()


@pytest.fixture()
def model(gpu_ctx, ckpt):
    import gmix_b200
    m = gmix_b200.Model(gpu_ctx, ckpt[0], ckpt[1], max_new_bytes=2048)
    yield m
    m.close()


def test_reference_checkpoint_loads_and_compress_decompress_continue_identically(gpu_ctx, model):
    assert model.trained_bytes == 600
    want = open(os.path.join(GOLD, "ckpt600_b.gmix"), "rb").read()
    streams = [B, B[:100], B, b""]
    comp = gpu_ctx.compress_batch_from(model, streams)
    assert comp[0] == want and comp[2] == want          # every stream is an independent clone of the checkpoint
    assert gpu_ctx.decompress_batch_from(model, comp) == streams
    # and the context still serves from-scratch streams afterwards
    assert gpu_ctx.compress_batch([TEXT]) == [open(os.path.join(GOLD, "text1k.gmix"), "rb").read()]


@pytest.mark.parametrize("size,temp", [(48, 1.0), (40, 0.5)])
def test_batched_generation_samples_the_reference_bytes(gpu_ctx, model, size, temp):
    want = open(os.path.join(GOLD, f"ckpt600_gen_{size}_{temp}.out"), "rb").read()
    prompts = [PROMPT] * 5 + [PROMPT[:9] + b"\n"]
    out = gpu_ctx.generate_batch(model, prompts, size, temp)
    assert all(o == want for o in out[:5])
    assert len(out[5]) == size and out[5] != want        # another prompt, another continuation
    # learning is off while sampling, so the same call again gives the same bytes (LongTermMemory untouched, tester.cpp:358-366)
    assert gpu_ctx.generate_batch(model, prompts[:2], size, temp) == out[:2]


@pytest.mark.parametrize("size,temp", [(48, 1.0), (40, 0.5)])
def test_lockstep_generation_with_the_exact_batched_gate_product_samples_the_reference_bytes(gpu_ctx, model, size, temp):
    """GMX_GEN_LOCKSTEP_EXACT: prompt launch, then one batched gate product (reference arithmetic, all streams at once) + one
    GenStepKernel launch per sampled byte - the same bytes as `gmix -g` and as the per-stream kernel."""
    want = open(os.path.join(GOLD, f"ckpt600_gen_{size}_{temp}.out"), "rb").read()
    prompts = [PROMPT] * 5 + [PROMPT[:9] + b"\n"] + [PROMPT] * 140     # 146 streams: two 128-stream tiles, 19 exact-kernel groups
    per_stream = gpu_ctx.generate_batch(model, prompts, size, temp)
    gpu_ctx.set_generation_mode(gpu_ctx.GEN_LOCKSTEP_EXACT)
    try:
        out = gpu_ctx.generate_batch(model, prompts, size, temp)
        assert gpu_ctx.last_generation_mode == gpu_ctx.GEN_LOCKSTEP_EXACT
    finally:
        gpu_ctx.set_generation_mode(gpu_ctx.GEN_PER_STREAM)
    assert out == per_stream
    assert all(o == want for o in out[:5]) and out[5] != want and out[-1] == want


def test_batched_gate_product_kernels(gpu_ctx):
    """gate_gemm.cuh on seeded random operands: the exact kernel is bit-identical to a host loop in the reference's order; the
    tcgen05 kernel (3xTF32, fp32 accumulation in TMEM) is within the stated tolerance of the fp64 sum - 8x the rounding error the
    reference's own sequential fp32 arithmetic shows on the same operands, and 1e-5 of the largest value."""
    for n_slots, seed in ((300, 3), (128, 4), (1, 5)):
        bad, e_tc, e_seq, mag = gpu_ctx.selftest_gate(n_slots, seed)
        assert bad == 0
        assert e_tc <= max(8 * e_seq, 1e-5 * mag), (n_slots, e_tc, e_seq, mag)


def test_tensor_core_generation_follows_the_exact_samples(gpu_ctx, model):
    """GMX_GEN_LOCKSTEP_TENSOR is opt-in and not bit-exact (summation order): its samples must be valid, reproducible, and stay
    with the exact samples except where a draw falls within rounding distance of a probability."""
    prompts = [PROMPT] * 3 + [PROMPT[:9] + b"\n", PROMPT[:20]]
    exact = gpu_ctx.generate_batch(model, prompts, 48, 1.0)
    gpu_ctx.set_generation_mode(gpu_ctx.GEN_LOCKSTEP_TENSOR)
    try:
        out = gpu_ctx.generate_batch(model, prompts, 48, 1.0)
        assert gpu_ctx.last_generation_mode == gpu_ctx.GEN_LOCKSTEP_TENSOR
        again = gpu_ctx.generate_batch(model, prompts, 48, 1.0)
    finally:
        gpu_ctx.set_generation_mode(gpu_ctx.GEN_PER_STREAM)
    assert [len(o) for o in out] == [48] * len(prompts) and out == again
    assert out[0] == out[1] == out[2]
    assert sum(o == e for o, e in zip(out, exact)) >= 3, (out, exact)


def test_lockstep_modes_fall_back_to_per_stream_when_a_prompt_reaches_a_bptt_pass(gpu_ctx, model):
    """The gate matrix is only shared while no stream runs BPTT: a 120-byte prompt crosses the 100-byte horizon."""
    long_prompt = (PROMPT * 8)[:120]
    per_stream = gpu_ctx.generate_batch(model, [long_prompt], 16, 1.0)
    gpu_ctx.set_generation_mode(gpu_ctx.GEN_LOCKSTEP_EXACT)
    try:
        out = gpu_ctx.generate_batch(model, [long_prompt], 16, 1.0)
        assert gpu_ctx.last_generation_mode == gpu_ctx.GEN_PER_STREAM
    finally:
        gpu_ctx.set_generation_mode(gpu_ctx.GEN_PER_STREAM)
    assert out == per_stream


def test_written_checkpoint_matches_the_reference_files(gpu_ctx, ckpt):
    sh, lo = gpu_ctx.train_checkpoint(A)
    assert lo == ckpt[1]
    diff = ckpt_layout.differing_sections(sh, ckpt[0])
    assert set(diff) <= ckpt_layout.SCRATCH, [x for x in diff if x not in ckpt_layout.SCRATCH]
    # training continues from a checkpoint: A then B equals A+B in one go
    import gmix_b200
    m = gmix_b200.Model(gpu_ctx, sh, lo, max_new_bytes=len(B))
    sh2, lo2 = gpu_ctx.train_checkpoint(B, m)
    m.close()
    sh3, lo3 = gpu_ctx.train_checkpoint(TEXT)
    assert lo2 == lo3
    assert set(ckpt_layout.differing_sections(sh2, sh3)) <= ckpt_layout.SCRATCH


def test_predictor_facade_reads_and_writes_checkpoints(gpu_ctx, ckpt):
    import gmix_b200
    import numpy as np
    p16 = np.fromfile(os.path.join(GOLD, "text1k.p16"), dtype=np.uint16)   # the reference's whole-stream run
    p = gmix_b200.Predictor(gpu_ctx, 4096)
    p.enable_analysis(True)                               # what `gmix -c <ckpt>` does for this length (runner-utils.cpp:47)
    p.read_checkpoint(*ckpt)                              # Predictor::ReadCheckpoint
    got = []
    for byte in TEXT[600:640]:
        for j in range(7, -1, -1):
            got.append(int(1 + 65534 * np.float32(p.predict())))
            p.perceive((byte >> j) & 1)
            p.learn()
    assert got == [int(x) for x in p16[600 * 8:640 * 8]]
    sh, lo = p.write_checkpoint()                         # Predictor::WriteCheckpoint
    p.close()
    sh_ref, lo_ref = gpu_ctx.train_checkpoint(TEXT[:640])
    assert lo == lo_ref
    # the facade ran with analysis on (inactive predictions are zeroed each Predict, predictor.cpp:362-365), the
    # training pass with analysis off (they keep their last value): ShortTermMemory::predictions legitimately differs
    assert set(ckpt_layout.differing_sections(sh, sh_ref)) <= ckpt_layout.SCRATCH | {"stm.predictions"}


def test_training_on_incompressible_data_takes_the_worst_case_arena(gpu_ctx, tmp_path):
    """Random bytes overflow the text-sized arena; gmx_train_checkpoint re-runs the stream in one worst-case arena and
    the checkpoint still matches what the unmodified reference writes for the same bytes (oracle/_ref, when present)."""
    import subprocess
    import numpy as np
    import gmix_b200
    data = np.random.RandomState(11).randint(0, 256, 3000, dtype=np.uint8).tobytes()
    before = gpu_ctx.retried_streams
    sh, lo = gpu_ctx.train_checkpoint(data)
    assert gpu_ctx.retried_streams == before + 1
    m = gmix_b200.Model(gpu_ctx, sh, lo, max_new_bytes=600, roomy=True)      # loads back and continues
    assert m.trained_bytes == 3000
    comp = gpu_ctx.compress_batch_from(m, [data[:500]])
    assert gpu_ctx.decompress_batch_from(m, comp) == [data[:500]]
    m.close()
    assert gpu_ctx.compress_batch([B]) == gpu_ctx.compress_batch([B])          # the context is back to from-scratch streams
    ref = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "ref_driver")
    if os.path.exists(ref):
        (tmp_path / "r.in").write_bytes(data)
        subprocess.run([ref, "train", str(tmp_path / "r.in"), str(tmp_path / "ref")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert (tmp_path / "ref.long").read_bytes() == lo
        diff = ckpt_layout.differing_sections(sh, (tmp_path / "ref.short").read_bytes())
        assert set(diff) <= ckpt_layout.SCRATCH, diff


def test_reference_restart_tests_through_the_facade(tmp_path):
    """`gmixb200 -T`: the reference's own test program (runner/tester.cpp:323-378) re-run against this build through the
    Predictor facade (Predict/Perceive/Learn/Copy/Write/ReadCheckpoint) and the host coder with its checkpoints."""
    import subprocess
    runner = os.path.join(ROOT, "gmix_b200", "lib", "gmixb200")
    src = tmp_path / "in.bin"
    src.write_bytes(open(os.path.join(GOLD, "text1k.in"), "rb").read()[:420])
    r = subprocess.run([runner, "-T", str(src), str(tmp_path / "work")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "Tests passed." in r.stdout, r.stdout + r.stderr


def test_stream_coded_in_parts_at_batch_speed(gpu_ctx, oracle):
    """gmx_compress_part / gmx_decompress_part: the same restart scenario without the per-bit facade: first part with the
    header and no flush, predictor checkpoint + coder state, second part from them; and the decoder likewise."""
    import gmix_b200
    from gmix_b200 import synth
    data = synth.synthetic_text_chunk(11, 6000)
    whole = oracle.compress(data)            # analysis is path-invisible here; the runner would enable it
    split = 2500
    a, coder, ck = gpu_ctx.compress_part(data[:split], header_total=len(data))
    model = gmix_b200.Model(gpu_ctx, ck[0], ck[1], max_new_bytes=len(data) - split + 16)
    b, _, ck_end = gpu_ctx.compress_part(data[split:], model=model, coder=coder, last=True)
    assert a + b == whole
    # decode the first 1000 bytes, checkpoint, decode the rest
    first, used, dcoder, dck = gpu_ctx.decompress_part(whole, 1000)
    assert first == data[:1000]
    m2 = gmix_b200.Model(gpu_ctx, dck[0], dck[1], max_new_bytes=len(data) - 1000 + 16)
    rest, _, _, _ = gpu_ctx.decompress_part(whole[used:], len(data) - 1000, model=m2, coder=dcoder, want_checkpoint=False)
    assert rest == data[1000:]
    # the predictor state after compress and after decompress of the same bytes is the same stream state
    full_ck = gpu_ctx.train_checkpoint(data)
    assert ck_end[1] == full_ck[1]            # .long byte-identical
    model.close(); m2.close()
