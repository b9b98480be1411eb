"""Quick byte-parity check of the loaded library (GMIX_B200_LIB selects a variant) against the golden vectors."""
import sys
sys.path.insert(0, ".")
import gmix_b200
c = gmix_b200.Context(0)
names = ["text1k", "repetitive", "random1200", "text_mid", "synth_chunk0_4k"]
ins = [open(f"tests/golden/{n}.in", "rb").read() for n in names]
got = c.compress_batch(ins)
ok = all(g == open(f"tests/golden/{n}.gmix", "rb").read() for g, n in zip(got, names)) and c.decompress_batch(got) == ins
print("parity", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
