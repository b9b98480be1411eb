"""TEST INFRASTRUCTURE ONLY: runs the unmodified reference CLI (oracle/_ref/gmix, built by oracle/Makefile from
/root/reference with the canonical strict flags) over many streams, one process per host core. Falls back to the CPU
restatement (oracle/_build/gmix_oracle, same CLI) when the reference binary did not travel."""
import os
import subprocess
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def binary():
    p = os.path.join(ROOT, "oracle", "_ref", "gmix")
    if os.path.exists(p):
        return p, "reference"
    p2 = os.path.join(ROOT, "oracle", "_build", "gmix_oracle")
    if not os.path.exists(p2):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    return p2, "port"


def run_many(mode, streams, cores=None):
    """mode '-c' or '-d'; returns (list of output bytes, wall seconds). One `gmix <mode> in out` per stream, at most
    `cores` at a time."""
    exe, _ = binary()
    cores = cores or os.cpu_count() or 1
    outs = [None] * len(streams)
    with tempfile.TemporaryDirectory(prefix="gmix_refcli_") as wd:
        os.makedirs(os.path.join(wd, "analysis"), exist_ok=True)   # the reference writes analysis/*.tsv relative to cwd
        for i, s in enumerate(streams):
            with open(os.path.join(wd, f"s{i}.in"), "wb") as f:
                f.write(s)
        t0 = time.perf_counter()
        running, nxt = [], 0
        while nxt < len(streams) or running:
            while nxt < len(streams) and len(running) < cores:
                p = subprocess.Popen([exe, mode, os.path.join(wd, f"s{nxt}.in"), os.path.join(wd, f"s{nxt}.out")],
                                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=wd)
                running.append((p, nxt))
                nxt += 1
            for p, i in list(running):
                if p.poll() is not None:
                    if p.returncode != 0:
                        raise RuntimeError(f"{exe} {mode} exited with {p.returncode} on stream {i}")
                    running.remove((p, i))
            time.sleep(0.005)
        dt = time.perf_counter() - t0
        for i in range(len(streams)):
            with open(os.path.join(wd, f"s{i}.out"), "rb") as f:
                outs[i] = f.read()
    return outs, dt
