# A/B of kernel build variants (gmix_b200/build.py: build_variant): throughput of one full wave of 8 KiB streams,
# preceded by a byte-parity check of the variant against the golden vectors.
for v in default "$@"; do
  if [ "$v" = default ]; then unset GMIX_B200_LIB; else export GMIX_B200_LIB=$PWD/gmix_b200/lib/variants/$v/libgmix_b200.so; fi
  echo "== $v"
  python - <<'PY'
import sys; sys.path.insert(0, ".")
import gmix_b200
c = gmix_b200.Context(0)
names = ["text1k", "repetitive", "random1200", "text_mid"]
ins = [open(f"tests/golden/{n}.in", "rb").read() for n in names]
got = c.compress_batch(ins)
ok = all(g == open(f"tests/golden/{n}.gmix", "rb").read() for g, n in zip(got, names)) and c.decompress_batch(got) == ins
print("parity", "OK" if ok else "FAILED")
PY
  python scripts/gpu_sweep.py 8192 8 2>&1 | tail -1
done
unset GMIX_B200_LIB
