"""ctypes binding of the CPU oracle (oracle/gmix_oracle.h). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "_build", "libgmix_oracle.so")
REF_DRIVER = os.path.join(ORACLE_DIR, "_ref", "ref_driver")


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.gmo_new.restype = C.c_void_p
        lib.gmo_free.argtypes = [C.c_void_p]
        lib.gmo_set_analysis.argtypes = [C.c_void_p, C.c_int]
        lib.gmo_predict.argtypes = [C.c_void_p]
        lib.gmo_predict.restype = C.c_float
        lib.gmo_perceive.argtypes = [C.c_void_p, C.c_int]
        lib.gmo_learn.argtypes = [C.c_void_p]
        lib.gmo_peek.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        sig = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        lib.gmo_compress.argtypes = sig
        lib.gmo_decompress.argtypes = sig
        lib.gmo_compress_trace.argtypes = sig + [C.c_void_p, C.c_void_p]

    def compress(self, data, trace=False):
        n = len(data)
        src = np.frombuffer(bytes(data), dtype=np.uint8).copy() if n else np.zeros(1, np.uint8)
        cap = n + n // 8 + 64
        out = np.zeros(cap, dtype=np.uint8)
        out_len = C.c_uint64(0)
        if trace:
            probs = np.zeros(max(8 * n, 1), dtype=np.float32)
            p16 = np.zeros(max(8 * n, 1), dtype=np.uint32)
            rc = self.lib.gmo_compress_trace(src.ctypes.data, n, out.ctypes.data, cap, C.byref(out_len), probs.ctypes.data, p16.ctypes.data)
            assert rc == 0
            return out[:out_len.value].tobytes(), probs[:8 * n], p16[:8 * n]
        rc = self.lib.gmo_compress(src.ctypes.data, n, out.ctypes.data, cap, C.byref(out_len))
        assert rc == 0
        return out[:out_len.value].tobytes()

    def decompress(self, data):
        n = len(data)
        src = np.frombuffer(bytes(data), dtype=np.uint8).copy() if n else np.zeros(1, np.uint8)
        cap = int.from_bytes(bytes(data[:5]), "big") + 8
        out = np.zeros(cap, dtype=np.uint8)
        out_len = C.c_uint64(0)
        rc = self.lib.gmo_decompress(src.ctypes.data, n, out.ctypes.data, cap, C.byref(out_len))
        assert rc == 0
        return out[:out_len.value].tobytes()


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)


def load():
    if not os.path.exists(SO):
        build()
    return Oracle(C.CDLL(SO))
