#!/usr/bin/env python
"""Benchmark of the gmix per-bit path on B200 (BASELINE.json metric: aggregate compress MB/s).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code, all host cores

Workload = BASELINE.json configs[1]: independent 64 KiB synthetic-text chunks (gmix_b200/synth.py,
SURVEY.md section 8d), every chunk compressed from scratch as its own stream, one CTA per stream.
A *step* is one pass of the hot path over one batch of `--chunks` chunks per GPU. The default batch is
one full wave of resident streams (8 CTAs x 148 SMs = 1184 chunks = 77.6 MB of input), so that a default
run (3 warm-up + 2 timed steps + 2 end-to-end steps) finishes within minutes; successive steps walk
through the 4096-chunk set of configs[1] (`--chunks 4096` runs the whole set in one step). Weak scaling: every
rank compresses its own `--chunks` chunks (chunk ids are disjoint across ranks), no collective on the
data path; after every step one NCCL all_gather collects {compressed size, FNV-1a checksum} per stream.

JSON line keys follow the driver contract; `value` is device-resident throughput (inputs already in
HBM), `e2e` is the same metric through the host-pointer C-ABI call (pinned host buffers, H2D and D2H
inside the timed region). `roofline` is the HBM roofline of the stream kernel with SURVEY.md 8(d)'s
algorithmic bytes (4.29e5 B of model-state traffic per input byte).

Further keys of the line (each a leg outside the timed region of `value`; `--no-legs` skips them all):
  parity        streams of the TIMED batch compared byte for byte with the unmodified reference CLI (the same wave of
                `gmix -c` processes that gives `cpu_baseline`)
  decompress    configs[2]: the last step's streams decompressed on the GPU, every stream compared with its input
  generate      configs[3] (reduced scale, stated): batched generation from a checkpoint written on the GPU
  config1_4096  configs[1] as BASELINE.json words it: all 4096 chunks in ONE gmx_compress_batch call (several waves)
  config0       configs[0]: dictionary/english.dic as a single stream, md5 of the output checked against the reference's
  config4       configs[4]: the enwik-shaped corpus cut into equal streams, sharded over the N ranks (strong scaling: the
                total is fixed), one NCCL all_gather of sizes + checksums
The reference arm (`--impl reference`) times the unmodified reference CLI on the same configuration: every step is one
wave of one FULL chunk per host core (a bounded sample of the step's chunk list, stated in cpu_baseline.sample).
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_INPUT_BYTE = 4.29e5   # SURVEY.md section 8(d); derivation in DESIGN.md


def ncu_traffic_per_input_byte():
    """dram__bytes_read.sum + dram__bytes_write.sum per input byte from the committed `ncu` capture of the stream kernel
    (profiles/traffic.json, written from the raw page of the capture named there), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    t = json.load(open(p))
    return (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["input_bytes"], t["capture"]
METRIC = "aggregate compress MB/s"
UNIT = "MB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", type=int, default=int(os.environ.get("GMIX_BENCH_CHUNKS", "1184")), help="chunks per GPU per step")
    ap.add_argument("--chunk-bytes", type=int, default=65536)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-decompress", action="store_true", help="skip the configs[2] leg (GPU decompress of the last step's streams)")
    ap.add_argument("--no-legs", action="store_true", help="skip every leg outside the timed region (parity, decompress, generate, config0/1/4)")
    ap.add_argument("--kernel-config", type=int, default=-1, help="kernel configuration (role split) of the timed steps; -1 = library default")
    ap.add_argument("--c4-streams", type=int, default=100, help="config4 leg: number of streams")
    ap.add_argument("--c4-bytes", type=int, default=131072, help="config4 leg: bytes per stream (BASELINE.json: 1000000; truncated to keep the run short, stated in the line)")
    ap.add_argument("--verify", type=int, default=1, help="streams per rank checked by a GPU decompress round trip after timing")
    ap.add_argument("--no-generate", action="store_true", help="skip the configs[3] leg (batched generation from a checkpoint)")
    ap.add_argument("--gen-train-bytes", type=int, default=65536, help="bytes of the enwik-shaped corpus the generation checkpoint is trained on")
    ap.add_argument("--gen-prompts", type=int, default=8192, help="prompts per GPU for the generation leg (BASELINE.json configs[3]: 8192; 0 = one wave of resident streams)")
    ap.add_argument("--gen-bytes", type=int, default=1024)
    ap.add_argument("--workload", default="chunks", choices=["chunks", "enwik"],
                    help="chunks = configs[1] (default, the metric's configuration); enwik = configs[4]: --chunks streams of --chunk-bytes "
                         "bytes cut from the enwik-shaped corpus (100 x 1000000 in BASELINE.json)")
    return ap.parse_args()


def make_config(args, world):
    """`config` of the JSON line; identical for both arms (the reference arm samples it, see cpu_baseline.sample)."""
    if args.workload == "chunks":
        wl = "configs[1]: independent 64 KiB synthetic-text chunks compressed from scratch, one CTA per stream"
    else:
        wl = "configs[4]: enwik-shaped corpus cut into independent streams, sharded over the GPUs, NCCL gather of sizes+checksums"
    return {"workload": wl, "chunk_bytes": args.chunk_bytes, "chunks_per_step": args.chunks, "chunks_per_step_all_gpus": args.chunks * world,
            "chunk_set": "step k of rank r compresses chunks r*4096 + (k*chunks_per_step + i) mod 4096 of the synthetic-text set (gmix_b200/synth.py)",
            "l2": "256 MiB flush write between timed steps; per-stream arenas exceed L2",
            "parallelism": f"streams sharded over {world} GPU(s), all_gather of sizes+checksums per step"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ---------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        self.proc = subprocess.Popen([exe, "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                mhz.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(mhz)}


# ---------------------------------------------------------------------------------------------------
def chunk_ids(rank, step, per_step, total_set=4096):
    """Chunk ids of this rank for one step: every rank owns its own 4096-chunk set (disjoint ids) and successive
    steps walk through it (wrapping around)."""
    span = max(total_set, per_step)
    return [rank * span + (step * per_step + i) % span for i in range(per_step)]


_CHUNK_CACHE = {}


_WORKLOAD = "chunks"
_CORPUS = {}


def _gen_chunk(args):
    from gmix_b200 import synth
    i, size = args
    if _WORKLOAD == "enwik":       # stream i = bytes [i*size, (i+1)*size) of the enwik-shaped corpus
        need = (i + 1) * size
        have = _CORPUS.get("data", b"")
        if len(have) < need:
            _CORPUS["data"] = have = synth.enwik_shaped_corpus(max(need, _CORPUS.get("want", 0)))
        return have[i * size:(i + 1) * size]
    return synth.synthetic_text_chunk(i, size)


def pregenerate(ids, size, procs):
    """Generate the synthetic chunks with a process pool (pure-Python generator, ~26 ms per 64 KiB chunk).
    Must run before CUDA is initialised in this process (fork)."""
    todo = sorted({i for i in ids if (i, size) not in _CHUNK_CACHE})
    if not todo:
        return
    if procs > 1 and len(todo) >= 64:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            for i, c in zip(todo, pool.map(_gen_chunk, [(i, size) for i in todo], chunksize=16)):
                _CHUNK_CACHE[(i, size)] = c
    else:
        for i in todo:
            _CHUNK_CACHE[(i, size)] = _gen_chunk((i, size))


def make_chunks(ids, size):
    pregenerate(ids, size, 1)
    return [_CHUNK_CACHE[(i, size)] for i in ids]


def fnv1a(b):
    h = 0xcbf29ce484222325
    for x in b:
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation, one `gmix -c` process per host core
def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "gmix")
    if os.path.exists(p):
        return p, "reference"
    p2 = os.path.join(ROOT, "oracle", "_build", "gmix_oracle")
    if not os.path.exists(p2):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    return p2, "port"


def cpu_wave(binary, chunks, workdir, cores, keep_outputs=False):
    """Compress `chunks` with one process per core (at most `cores` at a time); returns wall seconds (and the
    compressed streams when keep_outputs)."""
    paths = []
    for i, c in enumerate(chunks):
        p = os.path.join(workdir, f"c{i}.in")
        with open(p, "wb") as f:
            f.write(c)
        paths.append(p)
    t0 = time.perf_counter()
    running, nxt = [], 0
    while nxt < len(paths) or running:
        while nxt < len(paths) and len(running) < cores:
            running.append(subprocess.Popen([binary, "-c", paths[nxt], paths[nxt] + ".gmix"], stdout=subprocess.DEVNULL,
                                            stderr=subprocess.DEVNULL, cwd=workdir))
            nxt += 1
        for p in list(running):
            if p.poll() is not None:
                if p.returncode != 0:
                    raise RuntimeError(f"{binary} exited with {p.returncode}")
                running.remove(p)
        time.sleep(0.005)
    dt = time.perf_counter() - t0
    if keep_outputs:
        return dt, [open(p + ".gmix", "rb").read() for p in paths]
    return dt


def first_diff_bit(a, b):
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            return 8 * i + (7 - ((x ^ y).bit_length() - 1))
    return 8 * min(len(a), len(b)) if len(a) != len(b) else None


def cpu_baseline_and_parity(ids, size, gpu_streams, cores=None):
    """One wave of the unmodified reference CLI over FULL chunks of the timed batch, one `gmix -c` process per host core:
    the wall time is the CPU baseline, the bytes are the parity check of the GPU's streams for the same chunks."""
    binary, kind = reference_binary()
    cores = cores or os.cpu_count() or 1
    k = min(cores, len(ids))
    chunks = make_chunks(ids[:k], size)
    with tempfile.TemporaryDirectory(prefix="gmix_cpu_") as wd:
        os.makedirs(os.path.join(wd, "analysis"), exist_ok=True)
        dt, ref = cpu_wave(binary, chunks, wd, cores, keep_outputs=True)
    base = {"value": k * size / dt / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{k} full {size}-byte chunks of the timed batch (chunk ids {ids[0]}..), one `gmix -c` process per core in one wave, "
                      f"{dt:.1f} s wall incl. process start-up (0.7 s of ~{dt:.0f})"}
    same = [g == r for g, r in zip(gpu_streams[:k], ref)]
    fd = None
    for g, r in zip(gpu_streams[:k], ref):
        if g != r:
            fd = first_diff_bit(g, r)
            break
    gb, rb = sum(len(g) for g in gpu_streams[:k]), sum(len(r) for r in ref)
    parity = {"checked": k, "identical": sum(same), "bits_per_byte_delta": 8.0 * (gb - rb) / (k * size), "first_diff_bit": fd,
              "against": f"oracle/_ref/gmix -c ({kind}) on the same chunks of the timed batch"}
    return base, parity


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    binary, kind = reference_binary()
    cores = os.cpu_count() or 1
    # a step = one wave of one FULL chunk per host core: the first `cores` chunks of the step's chunk list
    size = args.chunk_bytes
    times = []
    with tempfile.TemporaryDirectory(prefix="gmix_ref_") as wd:
        os.makedirs(os.path.join(wd, "analysis"), exist_ok=True)
        for step in range(args.warmup + args.steps):
            chunks = make_chunks(chunk_ids(0, step, args.chunks)[:cores], size)
            dt = cpu_wave(binary, chunks, wd, cores)
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = args.steps * cores * size / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+u8 (bit-exact integer/fp32 model state)", "data": "synthetic",
        "config": make_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"per step the first {cores} chunks of the step's {args.chunks}-chunk list, FULL {size}-byte chunks, one "
                                   f"`gmix -c` process per host core in one wave (strict -O2 build), process start-up inside the timed region"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    global _WORKLOAD
    world0 = int(os.environ.get("WORLD_SIZE", "1"))
    rank0 = int(os.environ.get("RANK", "0"))
    _WORKLOAD = args.workload
    all_ids = [i for st in range(args.warmup + args.steps) for i in chunk_ids(rank0, st, args.chunks, 4096 if args.workload == "chunks" else args.chunks)]
    if args.workload == "enwik":   # one sequential generator: build the corpus once, then slice
        _CORPUS["want"] = (max(all_ids) + 1) * args.chunk_bytes
        pregenerate(all_ids, args.chunk_bytes, 1)
    else:
        pregenerate(all_ids, args.chunk_bytes, max(1, min(32, (os.cpu_count() or 1) // world0)))
    import numpy as np
    import torch
    import torch.distributed as dist
    import gmix_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gmix_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = gmix_b200.Context(local)
    if args.kernel_config >= 0:
        ctx.set_kernel_config(args.kernel_config)
    if args.no_legs:
        args.no_decompress = args.no_generate = True
    stream = torch.cuda.current_stream()
    ctx.set_cuda_stream(stream.cuda_stream)
    n, size = args.chunks, args.chunk_bytes
    ctx.configure(size, 0)
    cap = gmix_b200.compress_bound(size)
    in_off = torch.arange(n + 1, dtype=torch.int64) * size
    out_off = torch.arange(n + 1, dtype=torch.int64) * cap
    d_in_off, d_out_off = in_off.to(dev), out_off.to(dev)
    d_out = torch.zeros(n * cap + 16, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_sum = torch.zeros(n, dtype=torch.int64, device=dev)
    gathered = [torch.zeros(2 * n, dtype=torch.int64, device=dev) for _ in range(world)] if world > 1 else None
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    total_set = 4096 if args.workload == "chunks" else n      # enwik: rank r owns streams [r*n, (r+1)*n) of one corpus, every step

    def step_inputs(step):
        data = b"".join(make_chunks(chunk_ids(rank, step, n, total_set), size))
        return torch.frombuffer(bytearray(data), dtype=torch.uint8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kernel_ms, launches0 = [], ctx.kernel_launches

    def device_step(d_in):
        ctx.compress_batch_device(d_in.data_ptr(), d_in_off.data_ptr(), n, d_out.data_ptr(), d_out_off.data_ptr(),
                                  d_len.data_ptr(), d_status.data_ptr(), size)
        kernel_ms.append(ctx.last_kernel_ms)
        ctx.checksum_device(d_out.data_ptr(), d_out_off.data_ptr(), d_len.data_ptr(), n, d_sum.data_ptr())
        if world > 1:
            dist.all_gather(gathered, torch.cat([d_len, d_sum]))

    # ---- device-resident measurement --------------------------------------------------------------
    total_steps = args.warmup + args.steps
    inputs = [step_inputs(s).to(dev) for s in range(total_steps)]
    for s in range(args.warmup):
        device_step(inputs[s])
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    kernel_ms.clear()
    launches_before = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for s in range(args.warmup, total_steps):
        l2_flush.fill_(s & 0xff)          # flush L2 between timed steps (the arenas alone exceed L2 as well)
        device_step(inputs[s])
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = ctx.kernel_launches - launches_before
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    kms = torch.tensor([sum(kernel_ms) / max(len(kernel_ms), 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms_avg = float(ms.item()), float(kms.item())
    bad = int((d_status != 0).sum().item())
    if bad:
        raise SystemExit(f"bench.py: {bad} streams failed on rank {rank}: {d_status[d_status != 0][:8].tolist()}")
    comp_bytes = int(d_len.sum().item())
    value = world * n * size * args.steps / (ms_total / 1e3) / 1e6
    resident, arena_mib = ctx.resident_streams, ctx.arena_bytes >> 20
    kc = ctx.kernel_configs()[ctx.kernel_config]
    kcfg_desc = (f"config {ctx.kernel_config}: {32 * (kc[0] + kc[1] + 1)} threads per stream CTA, {kc[2]} CTAs/SM, " +
                 ("hybrid order: all threads run the LSTM phases, the PPMd warp prepares the next byte under this byte's bit path"
                  if kc[3] and kc[1] == 0 else "all threads walk the phases together" if kc[3]
                  else f"pipelined roles: {kc[0]} bit warps + {kc[1]} LSTM warps + 1 PPMd warp"))
    kernel_name = f"gmx::StreamKernel<{kc[0]}, {kc[1]}, MODE_COMPRESS, {kc[2]}, false, {'true' if kc[3] else 'false'}>"

    # ---- parity of the TIMED batch against the unmodified reference + CPU baseline (one wave of full chunks) --------
    cpu_base, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "chunks":
        kpar = min(os.cpu_count() or 1, n)
        lens = d_len[:kpar].cpu().tolist()
        gpu_streams = [bytes(d_out[i * cap:i * cap + lens[i]].cpu().numpy()) for i in range(kpar)]
        cpu_base, parity = cpu_baseline_and_parity(chunk_ids(rank, total_steps - 1, n, total_set), size, gpu_streams)
        if parity["identical"] != parity["checked"]:
            print(f"bench.py: WARNING parity {parity}", file=sys.stderr)

    # ---- parity spot check outside the timed region: GPU decompress of the last step's streams ------
    if args.verify:
        k = min(args.verify, n)
        lens = d_len[:k].cpu().tolist()
        comp = [bytes(d_out[i * cap:i * cap + lens[i]].cpu().numpy()) for i in range(k)]
        ctx.set_cuda_stream(0)
        back = ctx.decompress_batch(comp)
        ctx.set_cuda_stream(stream.cuda_stream)
        want = make_chunks(chunk_ids(rank, total_steps - 1, n, total_set)[:k], size)
        assert back == want, "GPU decompress of a GPU-compressed stream does not reproduce the input"

    # ---- configs[2]: GPU decompress of every stream of the last step, device resident, checked ---------
    decomp = None
    if not args.no_decompress:
        d_back = torch.zeros(n * size + 16, dtype=torch.uint8, device=dev)
        d_back_len = torch.zeros(n, dtype=torch.int64, device=dev)
        d_back_off = in_off.to(dev)                     # capacity = the original stream length
        # the compressed streams sit at out_off[i] with length d_len[i]; decompress reads in_off[i+1]-in_off[i]
        # bytes per stream, so pack them contiguously first
        lens = d_len.clone()
        pack_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        pack_off[1:] = torch.cumsum(lens, 0)
        idx = torch.arange(int(pack_off[-1].item()), device=dev)
        which = torch.searchsorted(pack_off[1:], idx, right=True)
        d_pack = d_out[out_off.to(dev)[which] + (idx - pack_off[which])]
        barrier()
        dv0, dv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dv0.record(stream)
        ctx.decompress_batch_device(d_pack.data_ptr(), pack_off.data_ptr(), n, d_back.data_ptr(), d_back_off.data_ptr(),
                                    d_back_len.data_ptr(), d_status.data_ptr(), size)
        dv1.record(stream)
        barrier()
        dms = torch.tensor([dv0.elapsed_time(dv1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dms, op=dist.ReduceOp.MAX)
        lossless = bool(torch.equal(d_back[:n * size], inputs[total_steps - 1][:n * size])) and int((d_status != 0).sum().item()) == 0
        if not lossless:
            raise SystemExit(f"bench.py: GPU decompress of the GPU-compressed streams is not lossless on rank {rank}")
        decomp = {"value": world * n * size / (float(dms.item()) / 1e3) / 1e6, "unit": UNIT, "lossless_streams": n * world,
                  "workload": "configs[2]: decompression of the same chunk set, every stream compared with its input"}

    # ---- configs[3]: batched generation from a checkpoint, learning disabled while sampling ------------------
    generate = None
    if not args.no_generate and args.workload == "chunks":
        from gmix_b200 import synth
        ctx.set_cuda_stream(0)
        npr = args.gen_prompts or ctx.resident_streams
        T, G = args.gen_train_bytes, args.gen_bytes
        corpus = synth.enwik_shaped_corpus(T + 4096 * npr + 64)
        t0 = time.perf_counter()
        ck_short, ck_long = ctx.train_checkpoint(corpus[:T])            # Predict/Perceive/Learn + Predictor::WriteCheckpoint on the GPU
        train_s = time.perf_counter() - t0
        model = gmix_b200.Model(ctx, ck_short, ck_long, max_new_bytes=64 + G)
        prompts = [corpus[T + 4096 * k:T + 4096 * k + 64] for k in range(npr)]
        prompts[1] = prompts[0]                                          # same prompt, same draws -> must give the same bytes
        warm = prompts[:min(npr, 2 * ctx.max_resident_streams)]
        ctx.generate_batch(model, warm, G)                               # warm-up with the timed call's stream length and kernel configuration (sizes the overlay arenas)
        barrier()
        t0 = time.perf_counter()
        out = ctx.generate_batch(model, prompts, G, 1.0)
        barrier()
        gen_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        gen_kernel_ms = ctx.last_kernel_ms
        if world > 1:
            dist.all_reduce(gen_s, op=dist.ReduceOp.MAX)
        if out[0] != out[1] or len(set(out)) < max(2, npr // 2) or any(len(o) != G for o in out):
            raise SystemExit("bench.py: generation is not deterministic per prompt / not prompt dependent")
        generate = {"value": world * npr * G / float(gen_s.item()), "unit": "generated bytes/s", "prompts": npr * world, "prompt_bytes": 64,
                    "bytes_per_prompt": G, "temperature": 1.0, "kernel_ms": gen_kernel_ms,
                    "checkpoint": f"written on the GPU after {T} B of the enwik-shaped corpus ({train_s:.1f} s), reference format "
                                  f"(.short {len(ck_short)} B, .long {len(ck_long)} B), loaded back through gmx_model_load",
                    "arena_mib_per_stream": model.arena_bytes >> 20,
                    "resident_prompts": ctx.resident_streams,
                    "workload": f"configs[3]: {npr} prompts x {G} B per GPU from a checkpoint trained on {T} B (BASELINE.json: 1 MB; stated reduction), every "
                                "stream an overlay of the shared model (no clone of its tables); host buffers, H2D of prompts + draws and D2H of samples "
                                "inside the timed region"}
        # the same batch in the two LOCK-STEP modes (gmix_b200.h GMX_GEN_*): all streams advance one sampled byte per launch and
        # the LSTM gate products of a byte step are one batched kernel - with the reference's arithmetic (same bytes), or on the
        # tensor cores (tcgen05 kind::tf32, 3xTF32; summation order differs, so the divergence from the exact samples is reported)
        for mode, key in ((ctx.GEN_LOCKSTEP_EXACT, "lockstep_exact"), (ctx.GEN_LOCKSTEP_TENSOR, "lockstep_tensor")):
            ctx.set_generation_mode(mode)
            ctx.generate_batch(model, warm, G)                           # (lock-step keeps twice as many arenas: re-sized here, not in the timed call)
            barrier()
            t0 = time.perf_counter()
            out2 = ctx.generate_batch(model, prompts, G, 1.0)
            barrier()
            s2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(s2, op=dist.ReduceOp.MAX)
            ran = ctx.last_generation_mode
            same = sum(a == b for a, b in zip(out, out2))
            firsts = [next(i for i in range(G) if a[i] != b[i]) for a, b in zip(out, out2) if a != b]
            generate[key] = {"value": world * npr * G / float(s2.item()), "unit": "generated bytes/s", "kernel_ms": ctx.last_kernel_ms, "mode_ran": ran,
                             "streams_identical_to_per_stream": same, "streams": npr,
                             "mean_first_differing_byte": (sum(firsts) / len(firsts)) if firsts else None}
            if mode == ctx.GEN_LOCKSTEP_EXACT and ran == mode and same != npr:
                raise SystemExit("bench.py: lock-step generation with the exact gate product differs from per-stream generation")
        ctx.set_generation_mode(ctx.GEN_PER_STREAM)
        model.close()
        ctx.set_cuda_stream(stream.cuda_stream)

    # ---- configs[1] as worded: all 4096 chunks in ONE host-pointer call (several waves of resident streams) -----------
    config1_4096 = None
    if not args.no_legs and args.workload == "chunks":
        ctx.set_cuda_stream(0)
        n4 = 4096
        big = make_chunks(chunk_ids(rank, 0, n4), size)
        h_big = torch.frombuffer(bytearray(b"".join(big)), dtype=torch.uint8).pin_memory()
        del big
        big_in_off = (torch.arange(n4 + 1, dtype=torch.int64) * size).numpy()
        big_out_off = (torch.arange(n4 + 1, dtype=torch.int64) * cap).numpy()
        h_big_out = torch.zeros(n4 * cap + 16, dtype=torch.uint8).pin_memory()
        h_big_len = torch.zeros(n4, dtype=torch.int64).pin_memory()
        h_big_status = torch.zeros(n4, dtype=torch.int32).pin_memory()
        barrier()
        t0 = time.perf_counter()
        rc = ctx.lib.gmx_compress_batch(ctx.h, h_big.data_ptr(), big_in_off.ctypes.data, n4, h_big_out.data_ptr(), big_out_off.ctypes.data,
                                        h_big_len.data_ptr(), h_big_status.data_ptr())
        ctx._check(rc, "gmx_compress_batch")
        comp_total = int(h_big_len.sum().item())
        barrier()
        dt4 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        k4 = torch.tensor([ctx.last_kernel_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt4, op=dist.ReduceOp.MAX)
            dist.all_reduce(k4, op=dist.ReduceOp.MAX)
        config1_4096 = {"e2e": world * n4 * size / float(dt4.item()) / 1e6, "value": world * n4 * size / (float(k4.item()) / 1e3) / 1e6, "unit": UNIT,
                        "chunks_per_gpu": n4, "waves": -(-n4 // ctx.resident_streams), "bits_per_byte": 8.0 * comp_total / (n4 * size),
                        "workload": "configs[1]: 4096 independent 64 KiB chunks in one gmx_compress_batch call per GPU (host buffers; `value` = kernel time only)"}
        del h_big, h_big_out
        ctx.set_cuda_stream(stream.cuda_stream)

    # ---- configs[0]: dictionary/english.dic as a single stream (rank 0), md5 against the reference's ----------------------
    config0 = None
    if not args.no_legs and rank == 0:
        import hashlib
        ctx.set_cuda_stream(0)
        dic = open(os.path.join(ROOT, "tests", "data", "english.dic"), "rb").read()
        known = json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))["english_dic_full"]
        t0 = time.perf_counter()
        comp0 = ctx.compress_batch([dic])[0]
        dt0 = time.perf_counter() - t0
        config0 = {"value": len(dic) / dt0 / 1e6, "unit": UNIT, "seconds": dt0, "input_bytes": len(dic), "output_bytes": len(comp0),
                   "md5_matches_reference": hashlib.md5(comp0).hexdigest() == known["md5"] and len(comp0) == known["output_bytes"],
                   "workload": "configs[0]: dictionary/english.dic as ONE stream (one CTA: single-stream latency), host buffers"}
        ctx.set_cuda_stream(stream.cuda_stream)

    # ---- configs[4]: enwik-shaped corpus in equal streams sharded over the ranks (strong scaling), NCCL gather -----------
    config4 = None
    if not args.no_legs:
        from gmix_b200 import shard, synth
        ctx.set_cuda_stream(0)
        ns4, sz4 = args.c4_streams, args.c4_bytes
        corpus4 = synth.enwik_shaped_corpus(ns4 * sz4)
        lo, hi = shard.shard_ranges([sz4] * ns4, world)[rank]
        mine = [corpus4[i * sz4:(i + 1) * sz4] for i in range(lo, hi)]
        barrier()
        t0 = time.perf_counter()
        comp4 = ctx.compress_batch(mine) if mine else []
        rec = torch.zeros(2 * ns4, dtype=torch.int64, device=dev)
        for j, cst in enumerate(comp4):
            rec[2 * (lo + j)] = len(cst)
            rec[2 * (lo + j) + 1] = shard.fnv1a64(cst) - (1 << 64) if shard.fnv1a64(cst) >= (1 << 63) else shard.fnv1a64(cst)
        if world > 1:
            dist.all_reduce(rec, op=dist.ReduceOp.SUM)      # disjoint slots: a gather of {size, checksum} per stream
        barrier()
        dt4s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt4s, op=dist.ReduceOp.MAX)
        sizes = rec[0::2].cpu().tolist()
        config4 = {"value": ns4 * sz4 / float(dt4s.item()) / 1e6, "unit": UNIT, "scaling": "strong", "streams": ns4, "stream_bytes": sz4,
                   "streams_per_gpu": [b - a for a, b in shard.shard_ranges([sz4] * ns4, world)], "bits_per_byte": 8.0 * sum(sizes) / (ns4 * sz4),
                   "all_streams_reported": all(x > 5 for x in sizes),
                   "workload": f"configs[4]: {ns4} x {sz4} B of the enwik-shaped corpus (BASELINE.json: 100 x 1 000 000 B; truncated to bound the run), "
                               f"contiguous stream ranges per rank, one NCCL reduction gathers sizes + FNV-1a checksums; host buffers"}
        ctx.set_cuda_stream(stream.cuda_stream)

    # ---- end-to-end measurement through the host-pointer C ABI ---------------------------------------
    e2e = None
    if not args.no_e2e:
        h_in = [step_inputs(s).pin_memory() for s in range(total_steps)]
        h_out = torch.zeros(n * cap + 16, dtype=torch.uint8).pin_memory()
        h_len = torch.zeros(n, dtype=torch.int64).pin_memory()
        h_status = torch.zeros(n, dtype=torch.int32).pin_memory()
        np_in_off, np_out_off = in_off.numpy(), out_off.numpy()

        def host_step(s):
            rc = ctx.lib.gmx_compress_batch(ctx.h, h_in[s].data_ptr(), np_in_off.ctypes.data, n, h_out.data_ptr(),
                                            np_out_off.ctypes.data, h_len.data_ptr(), h_status.data_ptr())
            ctx._check(rc, "gmx_compress_batch")
            return int(h_len.sum().item())           # the step's result is read on the host

        e2e_steps = max(1, min(args.steps, 2))
        host_step(0)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.warmup, args.warmup + e2e_steps):
            host_step(s)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n * size * e2e_steps / float(dt.item()) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": n * size + 2 * 8 * (n + 1), "d2h_bytes_per_step": n * cap + 12 * n,
               "steps": e2e_steps, "timing": "host wall clock around gmx_compress_batch (pinned buffers), max over ranks"}

    if rank == 0:
        pk, pk_kind = peaks()
        achieved = n * size * ALGO_BYTES_PER_INPUT_BYTE / (kernel_ms_avg / 1e3) / 1e9
        tpb, tsrc = ncu_traffic_per_input_byte()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+u8 (bit-exact integer/fp32 model state)", "data": "synthetic",
            "config": make_config(args, world),
            "runtime": {"resident_streams_per_gpu": resident, "arena_mib_per_stream": arena_mib, "kernel_config": kcfg_desc},
            "bits_per_byte": 8.0 * comp_bytes / (n * size),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"], "traffic": tpb * n * size if tpb else None,
                         "traffic_source": f"{tpb:.0f} B of DRAM traffic per input byte in {tsrc}, scaled to this launch's input bytes" if tpb else None,
                         "peak_kind": pk_kind,
                         "kernel": kernel_name, "kernel_ms": kernel_ms_avg,
                         "algorithmic_bytes_per_input_byte": ALGO_BYTES_PER_INPUT_BYTE},
            "e2e": e2e, "parity": parity, "decompress": decomp, "generate": generate, "config1_4096": config1_4096, "config0": config0,
            "config4": config4, "retried_streams": ctx.retried_streams, "gpu_launches": gpu_launches, "clocks": clocks,
        }
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
