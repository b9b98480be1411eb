"""Joins an `ncu --page source --csv --print-source sass` dump of the stream kernel with the line table of the cubin
it was captured from (nvdisasm --print-line-info), and prints stall samples / executed instructions per source
function and per source line. The instruction order of the two listings is identical, so they are joined by index.

  python scripts/sass_join.py gpurun_out/prof_X_sass.csv.gz gmix_b200/lib/obj/kernel_compress.o [top_n]
"""
import collections
import csv
import gzip
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gmix_b200", "csrc")


def disassemble(obj):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], check=True, capture_output=True, text=True).stdout
    out, cur, fn = [], None, None
    for line in txt.split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            fn, cur = m.group(1), None
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out.append((fn, m.group(2).strip(), cur))
    return out


def functions_of(path):
    out = []
    for i, l in enumerate(open(path), 1):
        m = re.match(r"^\s*(?:static\s+)?(?:GMX_DEV|GMX_HD|__global__)[^;]*?\b(\w+)\s*\(", l)
        if m and not l.strip().startswith("//"):
            out.append((i, m.group(1)))
    return out


def main():
    prof, obj = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    dis = disassemble(obj)
    r = csv.reader(gzip.open(prof, "rt"))
    kernel = next(r)[1]
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    rows = [x for x in r if x and x[0].startswith("0x")]
    dis = [d for d in dis if d[0] and kernel.split("<")[0].split("::")[-1] in d[0]][:len(rows)] if len(dis) != len(rows) else dis
    assert len(dis) == len(rows), (len(dis), len(rows), "profile and object are from different builds")
    fmap = {f: functions_of(os.path.join(CSRC, f)) for f in ("stream_kernel.cuh", "ppmd.cuh", "dmath.cuh")}

    def fn_of(loc):
        if not loc:
            return "?"
        best = "?"
        for s, n in fmap.get(loc[0], []):
            if s > loc[1]:
                break
            best = n
        return loc[0].split(".")[0] + ":" + best

    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    S, EX, BAR = idx["# Samples"], idx["Instructions Executed"], idx["stall_barrier"]
    by_fn = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    by_line = collections.defaultdict(lambda: [0, 0, 0])
    ctx = None
    for d, x in zip(dis, rows):
        loc = d[2]
        if loc and loc[0] in ("stream_kernel.cuh", "ppmd.cuh"):
            ctx = loc                       # inlined dmath helpers are charged to the line that called them
        a = by_fn[fn_of(loc)]
        a[0] += int(x[S]); a[1] += int(x[EX])
        for c in stalls:
            a[2][c] += int(x[idx[c]])
        b = by_line[ctx]
        b[0] += int(x[S]) - int(x[BAR]); b[1] += int(x[idx["stall_long_sb"]]); b[2] += int(x[EX])
    tot = sum(a[0] for a in by_fn.values()); tex = sum(a[1] for a in by_fn.values())
    print(f"{kernel}\n{len(rows)} SASS instructions, {tot} stall samples, {tex} warp instructions executed\n")
    print(f"{'function':34s} {'samples%':>8s} {'exec%':>6s}  top stall reasons")
    for f, a in sorted(by_fn.items(), key=lambda kv: -kv[1][0])[:top]:
        why = ", ".join(f"{k[6:]} {100 * v / max(a[0], 1):.0f}%" for k, v in a[2].most_common(3))
        print(f"{f:34s} {100 * a[0] / tot:8.1f} {100 * a[1] / tex:6.1f}  {why}")
    nb = sum(b[0] for b in by_line.values())
    print(f"\nsource lines by non-barrier stall samples ({nb} samples)")
    src = {f: open(os.path.join(CSRC, f)).read().split("\n") for f in ("stream_kernel.cuh", "ppmd.cuh")}
    for loc, b in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        text = src[loc[0]][loc[1] - 1].strip()[:90] if loc else ""
        print(f"{100 * b[0] / nb:6.2f}%  long_sb {100 * b[1] / max(b[0], 1):3.0f}%  exec {100 * b[2] / tex:5.2f}%  {loc[0] + ':' + str(loc[1]) if loc else '?':24s} {text}")


if __name__ == "__main__":
    main()
