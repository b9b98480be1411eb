"""ctypes binding of include/gmix_b200.h plus a small host-side mirror of the reference's runner
interface (reference src/runner/runner-utils.h:24-40: RunCompression / RunDecompression), batched
over independent streams."""
import ctypes as C
import os

import numpy as np

_LIB = None


class GmixError(RuntimeError):
    pass


def library_path():
    # GMIX_B200_LIB: development only, an alternative build of the same library (kernel A/B measurements)
    return os.environ.get("GMIX_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libgmix_b200.so")


def load_library():
    """Load libgmix_b200.so or fail loudly — there is no CPU path behind this package."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise GmixError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(gmix_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    u8p, u64p, u32p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_float)
    lib.gmx_version.restype = C.c_char_p
    lib.gmx_global_error.restype = C.c_char_p
    lib.gmx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.gmx_destroy.argtypes = [C.c_void_p]
    lib.gmx_last_error.argtypes = [C.c_void_p]
    lib.gmx_last_error.restype = C.c_char_p
    lib.gmx_set_cuda_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.gmx_configure.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
    lib.gmx_compress_bound.argtypes = [C.c_uint64]
    lib.gmx_compress_bound.restype = C.c_uint64
    batch = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gmx_compress_batch.argtypes = batch
    lib.gmx_decompress_batch.argtypes = batch
    lib.gmx_compress_batch_device.argtypes = batch + [C.c_uint64]
    lib.gmx_decompress_batch_device.argtypes = batch + [C.c_uint64]
    lib.gmx_checksum_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.gmx_compress_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, u64p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gmx_resident_streams.argtypes = [C.c_void_p]
    lib.gmx_resident_streams.restype = C.c_uint32
    lib.gmx_pred_copy.argtypes = [C.c_void_p, C.c_void_p]
    lib.gmx_compress_part.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p,
                                      C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_void_p]
    lib.gmx_decompress_part.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p,
                                        C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.gmx_set_kernel_config.argtypes = [C.c_void_p, C.c_int]
    lib.gmx_get_kernel_config.argtypes = [C.c_void_p]
    lib.gmx_kernel_config_info.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 5
    lib.gmx_arena_count.argtypes = [C.c_void_p]
    lib.gmx_arena_count.restype = C.c_uint32
    lib.gmx_arena_bytes.argtypes = [C.c_void_p]
    lib.gmx_arena_bytes.restype = C.c_uint64
    lib.gmx_retried_streams.argtypes = [C.c_void_p]
    lib.gmx_retried_streams.restype = C.c_uint64
    lib.gmx_kernel_launches.argtypes = [C.c_void_p]
    lib.gmx_kernel_launches.restype = C.c_uint64
    lib.gmx_last_kernel_ms.argtypes = [C.c_void_p]
    lib.gmx_last_kernel_ms.restype = C.c_double
    lib.gmx_device_sm_count.argtypes = [C.c_void_p]
    lib.gmx_set_profile.argtypes = [C.c_void_p, C.c_int]
    lib.gmx_get_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.gmx_pred_new.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.gmx_pred_free.argtypes = [C.c_void_p]
    lib.gmx_pred_enable_analysis.argtypes = [C.c_void_p, C.c_int]
    lib.gmx_pred_predict.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.gmx_pred_perceive.argtypes = [C.c_void_p, C.c_int]
    lib.gmx_pred_learn.argtypes = [C.c_void_p]
    lib.gmx_get_usage.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.gmx_model_load.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_void_p)]
    lib.gmx_model_free.argtypes = [C.c_void_p]
    lib.gmx_model_arena_bytes.argtypes = [C.c_void_p]
    lib.gmx_model_arena_bytes.restype = C.c_uint64
    lib.gmx_model_trained_bytes.argtypes = [C.c_void_p]
    lib.gmx_model_trained_bytes.restype = C.c_uint64
    lib.gmx_compress_batch_from.argtypes = [C.c_void_p, C.c_void_p] + batch[1:]
    lib.gmx_decompress_batch_from.argtypes = [C.c_void_p, C.c_void_p] + batch[1:]
    lib.gmx_generate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p,
                                       C.c_uint64, C.c_void_p, C.c_void_p]
    lib.gmx_generate_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p,
                                              C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    lib.gmx_reference_rand_u.argtypes = [C.c_void_p, C.c_uint64]
    lib.gmx_reference_rand_u.restype = None
    blobs = [C.POINTER(C.c_void_p), u64p, C.POINTER(C.c_void_p), u64p]
    lib.gmx_train_checkpoint.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64] + blobs
    lib.gmx_pred_write_checkpoint.argtypes = [C.c_void_p] + blobs
    lib.gmx_pred_read_checkpoint.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.gmx_selftest_math.argtypes = [C.c_void_p, C.c_uint32, u64p, u32p]
    lib.gmx_set_generation_mode.argtypes = [C.c_void_p, C.c_int]
    lib.gmx_compress_analysis.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, u64p, C.c_uint32, C.c_void_p, C.c_uint32, u32p]
    lib.gmx_last_generation_mode.argtypes = [C.c_void_p]
    lib.gmx_selftest_gate.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, u64p, C.POINTER(C.c_double)]
    _LIB = lib
    return lib


def compress_bound(n):
    return int(load_library().gmx_compress_bound(n))


def _pack(streams):
    """list of bytes-like -> (uint8 array, uint64 offsets)."""
    lens = [len(s) for s in streams]
    off = np.zeros(len(streams) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer(b"".join(bytes(s) for s in streams), dtype=np.uint8).copy() if sum(lens) else np.zeros(1, np.uint8)
    return buf, off


def reference_rand_u(n):
    """The n draws rand()/RAND_MAX one `gmix -g` process uses for sampling (after the Predictor constructor)."""
    out = np.zeros(max(n, 1), dtype=np.float32)
    load_library().gmx_reference_rand_u(out.ctypes.data, n)
    return out[:n]


def _blobs(fn, *args):
    sp, lp = C.c_void_p(), C.c_void_p()
    sl, ll = C.c_uint64(0), C.c_uint64(0)
    rc = fn(*args, C.byref(sp), C.byref(sl), C.byref(lp), C.byref(ll))
    if rc != 0:
        return rc, None, None
    return 0, C.string_at(sp.value, sl.value), C.string_at(lp.value, ll.value)


class Model:
    """A loaded checkpoint (reference Predictor::ReadCheckpoint, predictor.cpp:406-420): `.short` + `.long` bytes as the
    reference writes them, parked on the GPU; batch calls start every stream as a clone of it."""

    def __init__(self, ctx, short_blob, long_blob, max_new_bytes, roomy=False):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._check(ctx.lib.gmx_model_load(ctx.h, short_blob, len(short_blob), long_blob, len(long_blob), max_new_bytes, int(roomy), C.byref(h)),
                   "gmx_model_load")
        self.h = h
        ctx._children.append(self)

    @classmethod
    def from_files(cls, ctx, prefix, max_new_bytes, roomy=False):
        return cls(ctx, open(prefix + ".short", "rb").read(), open(prefix + ".long", "rb").read(), max_new_bytes, roomy)

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:   # a closed Context has already freed its models
            self.ctx.lib.gmx_model_free(self.h)
        self.h = None

    @property
    def arena_bytes(self):
        return int(self.ctx.lib.gmx_model_arena_bytes(self.h))

    @property
    def trained_bytes(self):
        return int(self.ctx.lib.gmx_model_trained_bytes(self.h))


class Predictor:
    """One stream stepped bit by bit: the reference's Predictor interface (src/predictor.h:20-38)."""

    def __init__(self, ctx, max_stream_len):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._check(ctx.lib.gmx_pred_new(ctx.h, max_stream_len, C.byref(h)), "gmx_pred_new")
        self.h = h
        ctx._children.append(self)

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.gmx_pred_free(self.h)
        self.h = None

    def enable_analysis(self, on=True):
        self.ctx._check(self.ctx.lib.gmx_pred_enable_analysis(self.h, int(on)), "gmx_pred_enable_analysis")

    def predict(self):
        p = C.c_float()
        self.ctx._check(self.ctx.lib.gmx_pred_predict(self.h, C.byref(p)), "gmx_pred_predict")
        return p.value

    def perceive(self, bit):
        self.ctx._check(self.ctx.lib.gmx_pred_perceive(self.h, int(bit)), "gmx_pred_perceive")

    def learn(self):
        self.ctx._check(self.ctx.lib.gmx_pred_learn(self.h), "gmx_pred_learn")

    def copy_from(self, other):
        """Predictor::Copy (predictor.cpp:42-48)."""
        self.ctx._check(self.ctx.lib.gmx_pred_copy(self.h, other.h), "gmx_pred_copy")

    def write_checkpoint(self):
        """Predictor::WriteCheckpoint: returns (.short bytes, .long bytes)."""
        rc, sh, lo = _blobs(self.ctx.lib.gmx_pred_write_checkpoint, self.h)
        self.ctx._check(rc, "gmx_pred_write_checkpoint")
        return sh, lo

    def read_checkpoint(self, short_blob, long_blob):
        self.ctx._check(self.ctx.lib.gmx_pred_read_checkpoint(self.h, short_blob, len(short_blob), long_blob, len(long_blob)),
                        "gmx_pred_read_checkpoint")


class Context:
    """One GPU. Mirrors the reference runner at batch granularity: compress_batch(streams) returns, for
    every stream, exactly the bytes `gmix -c` writes for it (5-byte big-endian length + coder bytes)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.gmx_create(device, C.byref(h))
        if rc != 0:
            raise GmixError(f"gmx_create({device}) failed ({rc}): {self.lib.gmx_global_error().decode()}")
        self.h = h
        self.device = device
        self._children = []   # live Model / Predictor handles: they point into this ctx and are freed with it

    def close(self):
        if getattr(self, "h", None):
            for ch in self._children:
                ch.close()
            self._children = []
            self.lib.gmx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise GmixError(f"{what} failed ({rc}): {self.lib.gmx_last_error(self.h).decode()}")

    def set_cuda_stream(self, stream_ptr):
        self._check(self.lib.gmx_set_cuda_stream(self.h, C.c_void_p(stream_ptr)), "gmx_set_cuda_stream")

    def configure(self, max_stream_len, max_resident=0):
        self._check(self.lib.gmx_configure(self.h, max_stream_len, max_resident), "gmx_configure")

    # ---- host-buffer API (the call a user makes) ----
    def _run_host(self, fn, what, streams, caps, model=None):
        n = len(streams)
        if n == 0:
            return []
        if model is not None:
            plain = fn
            fn = lambda h, *a: plain(h, model.h, *a)
        buf, off = _pack(streams)
        ooff = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(np.asarray(caps, dtype=np.uint64), out=ooff[1:])
        out = np.zeros(int(ooff[-1]) + 1, dtype=np.uint8)
        out_len = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.uint32)
        rc = fn(self.h, buf.ctypes.data, off.ctypes.data, n, out.ctypes.data, ooff.ctypes.data, out_len.ctypes.data, status.ctypes.data)
        self.last_status = status
        self._check(rc, what)
        return [out[int(ooff[i]):int(ooff[i]) + int(out_len[i])].tobytes() for i in range(n)]

    def compress_batch(self, streams):
        return self._run_host(self.lib.gmx_compress_batch, "gmx_compress_batch", streams, [compress_bound(len(s)) for s in streams])

    def decompress_batch(self, streams):
        caps = [int.from_bytes(bytes(s[:5]), "big") + 8 if len(s) >= 5 else 8 for s in streams]
        return self._run_host(self.lib.gmx_decompress_batch, "gmx_decompress_batch", streams, caps)

    def compress_batch_from(self, model, streams):
        """`gmix -c <ckpt> in out` for every stream."""
        return self._run_host(self.lib.gmx_compress_batch_from, "gmx_compress_batch_from", streams, [compress_bound(len(s)) for s in streams], model)

    def decompress_batch_from(self, model, streams):
        caps = [int.from_bytes(bytes(s[:5]), "big") + 8 if len(s) >= 5 else 8 for s in streams]
        return self._run_host(self.lib.gmx_decompress_batch_from, "gmx_decompress_batch_from", streams, caps, model)

    def generate_batch(self, model, prompts, out_bytes, temperature=1.0, rand_u=None, rand_stride=0):
        """`gmix -g <ckpt> prompt out out_bytes temperature` for every prompt; rand_u defaults to the reference's draws,
        shared by all prompts (what separate gmix processes do)."""
        n = len(prompts)
        if n == 0:
            return []
        if rand_u is None:
            rand_u = reference_rand_u(out_bytes * 8)
        rand_u = np.ascontiguousarray(rand_u, dtype=np.float32)
        need = (n - 1) * rand_stride + out_bytes * 8   # what gmx_generate_batch reads (host.cu)
        if rand_u.size < need:
            raise GmixError(f"rand_u holds {rand_u.size} draws, {need} are needed ((n - 1) * rand_stride + out_bytes * 8)")
        buf, off = _pack(prompts)
        out = np.zeros(n * out_bytes + 1, dtype=np.uint8)
        status = np.zeros(n, dtype=np.uint32)
        rc = self.lib.gmx_generate_batch(self.h, model.h, buf.ctypes.data, off.ctypes.data, n, out_bytes, temperature, rand_u.ctypes.data,
                                         rand_stride, out.ctypes.data, status.ctypes.data)
        self.last_status = status
        self._check(rc, "gmx_generate_batch")
        return [out[i * out_bytes:(i + 1) * out_bytes].tobytes() for i in range(n)]

    def train_checkpoint(self, data, model=None):
        """Predict/Perceive/Learn over data, then Predictor::WriteCheckpoint: returns (.short bytes, .long bytes)."""
        src = np.frombuffer(bytes(data), dtype=np.uint8).copy() if len(data) else np.zeros(1, np.uint8)
        rc, sh, lo = _blobs(self.lib.gmx_train_checkpoint, self.h, model.h if model is not None else None, src.ctypes.data, len(data))
        self._check(rc, "gmx_train_checkpoint")
        return sh, lo

    def compress_part(self, data, model=None, coder=None, header_total=None, last=False, analysis=0, want_checkpoint=True):
        """One part of a stream coded in parts (Encoder::Write/ReadCheckpoint + Predictor::Write/ReadCheckpoint). coder = (x1, x2)
        the part starts with (None = fresh); header_total = length the 5-byte header announces (None = no header).
        Returns (bytes appended to the stream, (x1, x2) after the part, (.short, .long) checkpoint or None)."""
        n = len(data)
        src = np.frombuffer(bytes(data), dtype=np.uint8).copy() if n else np.zeros(1, np.uint8)
        cap = compress_bound(n)
        out = np.zeros(cap, dtype=np.uint8)
        out_len = C.c_uint64(0)
        cin = np.array(list(coder) + [0], dtype=np.uint32) if coder is not None else None
        cout = np.zeros(3, dtype=np.uint32)
        sp, sl, lp, ll = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
        rc = self.lib.gmx_compress_part(self.h, model.h if model is not None else None, cin.ctypes.data if cin is not None else None,
                                        int(header_total is not None), int(header_total or 0), int(last), int(analysis), src.ctypes.data, n,
                                        out.ctypes.data, cap, C.byref(out_len), cout.ctypes.data,
                                        C.byref(sp) if want_checkpoint else None, C.byref(sl), C.byref(lp), C.byref(ll), None)
        self._check(rc, "gmx_compress_part")
        ck = (C.string_at(sp.value, sl.value), C.string_at(lp.value, ll.value)) if want_checkpoint else None
        return out[:out_len.value].tobytes(), (int(cout[0]), int(cout[1])), ck

    def decompress_part(self, coded, out_bytes, model=None, coder=None, analysis=0, want_checkpoint=True):
        """One part of a stream decoded in parts. coder = (x1, x2, x) (None = first part: header + 4 bytes are read from `coded`).
        Returns (bytes, coded bytes consumed, (x1, x2, x) after the part, checkpoint or None)."""
        n = len(coded)
        src = np.frombuffer(bytes(coded), dtype=np.uint8).copy() if n else np.zeros(1, np.uint8)
        out = np.zeros(out_bytes + 16, dtype=np.uint8)
        used = C.c_uint64(0)
        cin = np.array(list(coder), dtype=np.uint32) if coder is not None else None
        cout = np.zeros(3, dtype=np.uint32)
        sp, sl, lp, ll = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
        rc = self.lib.gmx_decompress_part(self.h, model.h if model is not None else None, cin.ctypes.data if cin is not None else None, int(analysis),
                                          src.ctypes.data, n, out_bytes, out.ctypes.data, C.byref(used), cout.ctypes.data,
                                          C.byref(sp) if want_checkpoint else None, C.byref(sl), C.byref(lp), C.byref(ll))
        self._check(rc, "gmx_decompress_part")
        ck = (C.string_at(sp.value, sl.value), C.string_at(lp.value, ll.value)) if want_checkpoint else None
        return out[:out_bytes].tobytes(), int(used.value), tuple(int(x) for x in cout), ck

    def compress(self, data):
        return self.compress_batch([data])[0]

    def decompress(self, data):
        return self.decompress_batch([data])[0]

    def compress_trace(self, data, blackboard=False):
        """Single stream; returns (compressed bytes, probs[8n] float32, p16[8n] uint32, blackboard or None)."""
        n = len(data)
        src = np.frombuffer(bytes(data), dtype=np.uint8).copy() if n else np.zeros(1, np.uint8)
        cap = compress_bound(n)
        out = np.zeros(cap, dtype=np.uint8)
        out_len = C.c_uint64(0)
        probs = np.zeros(max(8 * n, 1), dtype=np.float32)
        p16 = np.zeros(max(8 * n, 1), dtype=np.uint32)
        bb = np.zeros((max(8 * n, 1), 126), dtype=np.float32) if blackboard else None
        rc = self.lib.gmx_compress_trace(self.h, src.ctypes.data, n, out.ctypes.data, cap, C.byref(out_len), probs.ctypes.data,
                                         p16.ctypes.data, bb.ctypes.data if blackboard else None)
        self._check(rc, "gmx_compress_trace")
        return out[:out_len.value].tobytes(), probs[:8 * n], p16[:8 * n], bb

    # ---- device-pointer API (inputs already resident in HBM) ----
    def compress_batch_device(self, d_in, d_in_off, n, d_out, d_out_off, d_out_len, d_status, max_stream_len):
        self._check(self.lib.gmx_compress_batch_device(self.h, d_in, d_in_off, n, d_out, d_out_off, d_out_len, d_status, max_stream_len),
                    "gmx_compress_batch_device")

    def decompress_batch_device(self, d_in, d_in_off, n, d_out, d_out_off, d_out_len, d_status, max_stream_len):
        self._check(self.lib.gmx_decompress_batch_device(self.h, d_in, d_in_off, n, d_out, d_out_off, d_out_len, d_status, max_stream_len),
                    "gmx_decompress_batch_device")

    def checksum_device(self, d_data, d_off, d_len, n, d_sum):
        self._check(self.lib.gmx_checksum_device(self.h, d_data, d_off, d_len, n, d_sum), "gmx_checksum_device")

    def set_kernel_config(self, cfg):
        self._check(self.lib.gmx_set_kernel_config(self.h, int(cfg)), "gmx_set_kernel_config")

    @property
    def kernel_config(self):
        return int(self.lib.gmx_get_kernel_config(self.h))

    def kernel_configs(self):
        """[(bit warps, LSTM warps, CTAs per SM, serial compress, resident gate weights)] of every kernel configuration."""
        out = []
        for k in range(self.lib.gmx_kernel_config_count()):
            v = [C.c_int() for _ in range(5)]
            self.lib.gmx_kernel_config_info(k, *[C.byref(x) for x in v])
            out.append(tuple(x.value for x in v))
        return out

    # ---- introspection ----
    @property
    def resident_streams(self):
        return int(self.lib.gmx_resident_streams(self.h))

    @property
    def max_resident_streams(self):
        return int(self.lib.gmx_arena_count(self.h))

    @property
    def arena_bytes(self):
        return int(self.lib.gmx_arena_bytes(self.h))

    @property
    def retried_streams(self):
        return int(self.lib.gmx_retried_streams(self.h))

    @property
    def kernel_launches(self):
        return int(self.lib.gmx_kernel_launches(self.h))

    @property
    def last_kernel_ms(self):
        return float(self.lib.gmx_last_kernel_ms(self.h))

    @property
    def sm_count(self):
        return int(self.lib.gmx_device_sm_count(self.h))

    # cycle counters of the PROF kernel variant, per role (stream_kernel.cuh): bit role 0-15, LSTM role 16-23, PPMd role 24-31
    PROFILE_SLOTS = ("bit:bookkeeping", "bit:wait_packet", "bit:boundary_ctx", "bit:lookups", "bit:gate_select", "bit:set_swap", "bit:l0_dot",
                     "bit:l0_chain", "bit:l1", "bit:final", "bit:predict_tail", "bit:learn_scalars+tables", "bit:weight_update", "bit:trace",
                     "init", "-",
                     "lstm:wait", "lstm:fwd_gates", "lstm:fwd_rest", "lstm:out_step", "lstm:bptt_epochs", "lstm:bptt_grads", "-", "-",
                     "ppmd:wait", "ppmd:model", "ppmd:normalise+nodes", "-", "-", "-", "-", "-")

    def set_profile(self, on=True):
        self._check(self.lib.gmx_set_profile(self.h, int(on)), "gmx_set_profile")

    def get_profile(self, max_streams=1):
        buf = np.zeros((max_streams, len(self.PROFILE_SLOTS)), dtype=np.uint64)
        n = self.lib.gmx_get_profile(self.h, buf.ctypes.data, max_streams)
        if n < 0:
            self._check(n, "gmx_get_profile")
        return buf[:n]

    def get_usage(self, max_streams):
        """[n, 8] uint32: sparse-table entries, mixer weight sets, PPMd unit bytes, history bytes, SM id,
        start us, end us, 0 per stream."""
        buf = np.zeros((max_streams, 8), dtype=np.uint32)
        n = self.lib.gmx_get_usage(self.h, buf.ctypes.data, max_streams)
        if n < 0:
            self._check(n, "gmx_get_usage")
        return buf[:n]

    ANALYSIS_ROW = np.dtype([("bits_seen", np.uint64), ("neg_entropy", np.float64, (33,)), ("ppmd_used", np.uint64), ("history", np.uint64)])

    def compress_analysis(self, data, sample_frequency=None):
        """One stream compressed with the reference's analysis output (Predictor::EnableAnalysis / RunAnalysis): returns
        (compressed bytes, rows) with one ANALYSIS_ROW per sample_frequency bits (default: 8 * len / 1000 as `gmix -c` uses)."""
        data = bytes(data)
        freq = int(sample_frequency if sample_frequency is not None else 8 * len(data) // 1000)
        if freq <= 0:
            raise GmixError("sample_frequency must be positive (the reference enables analysis from 125 input bytes on)")
        cap = compress_bound(len(data))
        out = np.zeros(cap + 1, dtype=np.uint8)
        src = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, dtype=np.uint8)
        max_rows = (8 * len(data)) // freq + 1
        rows = np.zeros(max_rows, dtype=self.ANALYSIS_ROW)
        n_out, n_rows = C.c_uint64(0), C.c_uint32(0)
        self._check(self.lib.gmx_compress_analysis(self.h, src.ctypes.data, len(data), out.ctypes.data, cap, C.byref(n_out), freq, rows.ctypes.data,
                                                   max_rows, C.byref(n_rows)), "gmx_compress_analysis")
        return out[:n_out.value].tobytes(), rows[:n_rows.value]

    GEN_PER_STREAM, GEN_LOCKSTEP_EXACT, GEN_LOCKSTEP_TENSOR = 0, 1, 2

    def set_generation_mode(self, mode):
        """How generate_batch runs the sampling phase (gmix_b200.h: GMX_GEN_*): per stream, lock-step with the exact batched
        gate product, or lock-step with the tensor-core (tcgen05) gate product."""
        self._check(self.lib.gmx_set_generation_mode(self.h, int(mode)), "gmx_set_generation_mode")

    @property
    def last_generation_mode(self):
        return int(self.lib.gmx_last_generation_mode(self.h))

    def selftest_gate(self, n_slots=300, seed=1):
        """Batched gate product kernels on seeded random operands: (entries of the exact kernel that differ bitwise from the
        host loop in the reference's order, max |tensor-core - fp64|, max |sequential fp32 - fp64|, max |value|)."""
        bad = C.c_uint64(0)
        err = (C.c_double * 3)()
        self._check(self.lib.gmx_selftest_gate(self.h, n_slots, seed, C.byref(bad), err), "gmx_selftest_gate")
        return int(bad.value), float(err[0]), float(err[1]), float(err[2])

    def selftest_math(self, stride=1):
        mism = (C.c_uint64 * 3)()
        first = (C.c_uint32 * 3)()
        self._check(self.lib.gmx_selftest_math(self.h, stride, mism, first), "gmx_selftest_math")
        return {name: (int(mism[i]), int(first[i])) for i, name in enumerate(("expf", "logf", "tanhf"))}
