"""gmix_b200 — B200-native (sm_100a) implementation of byronknoll/gmix's per-bit predict/update path.

The package is a thin host-side mirror of the reference's runner interface over the C ABI in
include/gmix_b200.h (libgmix_b200.so). There is no CPU fallback: importing works anywhere, but every
compute entry point needs the CUDA library and a B200.
"""
from .api import Context, GmixError, Model, Predictor, compress_bound, library_path, load_library, reference_rand_u  # noqa: F401
