"""ctypes binding of libkexcuda.so (include/kexcuda.h) and the host-side
mirror of the compiled-binary contract of the reference
(crt/crt.c:372-467; src/KMC/Program/Backends/C.hs:72-83): a `CompiledProgram`
is what `kexc compile` would have produced, `run()` is `./bin < in > out`.

There is no CPU fallback: if the CUDA library is missing or fails to load the
import of this module's entry points raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# KEX_LIB: an alternative build of the same library (kernel experiments: scripts/build_exp.py)
LIB_PATH = os.environ.get("KEX_LIB") or os.path.join(HERE, "libkexcuda.so")

KEX_OK, KEX_ERR_OUT_CAP, KEX_ERR_UNSUPPORTED, KEX_RETRY_EXACT = 0, -3, -4, 1
ACCEPT, REJECT = 0, 1

_lib = None


class KexError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libkexcuda: %s (%d)" % (msg, code))
        self.code = code


class KexInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "nphases", "nstates", "nclasses", "nregs", "nactions", "max_out_per_byte", "chunk_bytes", "monoid_kernels",
        "emit_kernel", "exact_tiles")]


EXPORTS = ["kex_load", "kex_free", "kex_info", "kex_run_device", "kex_run_host", "kex_shard_summarize",
           "kex_seam_bytes", "kex_shard_walk", "kex_stitch_live", "kex_shard_emit", "kex_final_action", "kex_out_bound",
           "kex_select_phase", "kex_stream_begin", "kex_stream_feed", "kex_stream_end", "kex_last_launch_count",
           "kex_set_timing", "kex_last_kernel_ms", "kex_strerror", "kex_last_cuda_error", "kex_set_shard_tail"]


def lib():
    """Load libkexcuda.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            "%s is missing: build it with `python -m kleenexlang_b200.build` "
            "(there is no CPU fallback for the transducer path)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, u32, u8p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_void_p
    L.kex_load.argtypes = [ctypes.c_char_p, sz, ctypes.c_int, ctypes.POINTER(vp)]
    L.kex_load.restype = ctypes.c_int
    L.kex_free.argtypes = [vp]
    L.kex_free.restype = None
    L.kex_info.argtypes = [vp, u32, ctypes.POINTER(KexInfo)]
    L.kex_run_device.argtypes = [vp, u8p, sz, u8p, sz, ctypes.POINTER(sz), ctypes.POINTER(ctypes.c_int),
                                 ctypes.POINTER(sz), vp]
    L.kex_run_host.argtypes = [vp, ctypes.c_char_p, sz, u8p, sz, ctypes.POINTER(sz), ctypes.POINTER(ctypes.c_int),
                               ctypes.POINTER(sz)]
    L.kex_shard_summarize.argtypes = [vp, u8p, sz, ctypes.POINTER(ctypes.c_uint16), vp]
    L.kex_shard_walk.argtypes = [vp, u32, ctypes.POINTER(u32), ctypes.POINTER(sz), ctypes.POINTER(ctypes.c_uint8), vp]
    L.kex_shard_emit.argtypes = [vp, u32, sz, u8p, sz, ctypes.POINTER(sz), vp]
    L.kex_seam_bytes.argtypes = [vp]
    L.kex_seam_bytes.restype = sz
    L.kex_stitch_live.argtypes = [vp, ctypes.c_char_p, sz, u32, ctypes.POINTER(u32)]
    L.kex_final_action.argtypes = [vp, u32, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(u32),
                                   ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(sz)]
    L.kex_select_phase.argtypes = [vp, u32]
    L.kex_stream_begin.argtypes = [vp]
    L.kex_stream_feed.argtypes = [vp, ctypes.c_char_p, sz, u8p, sz, ctypes.POINTER(sz)]
    L.kex_stream_end.argtypes = [vp, u8p, sz, ctypes.POINTER(sz), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(sz)]
    L.kex_out_bound.argtypes = [vp, sz]
    L.kex_out_bound.restype = sz
    L.kex_last_launch_count.argtypes = [vp]
    L.kex_last_launch_count.restype = u32
    L.kex_set_timing.argtypes = [vp, ctypes.c_int]
    L.kex_set_shard_tail.argtypes = [vp, ctypes.c_int]
    L.kex_last_kernel_ms.argtypes = [vp, u32]
    L.kex_last_kernel_ms.restype = ctypes.c_float
    L.kex_strerror.argtypes = [ctypes.c_int]
    L.kex_strerror.restype = ctypes.c_char_p
    L.kex_last_cuda_error.argtypes = [vp]
    L.kex_last_cuda_error.restype = ctypes.c_char_p
    _lib = L
    return L


class CompiledProgram:
    """A loaded kexprog blob: the CUDA counterpart of a binary produced by
    `kexc compile` (src/kexc.hs:29-50)."""

    def __init__(self, blob: bytes, device: int = 0):
        self._L = lib()
        self._h = ctypes.c_void_p()
        self.blob = blob
        rc = self._L.kex_load(blob, len(blob), device, ctypes.byref(self._h))
        if rc != KEX_OK:
            raise KexError(rc, self._L.kex_strerror(rc).decode())
        self.device = device

    @classmethod
    def from_kex(cls, src: str, opt: int = 3, device: int = 0):
        from .kexprog import compile_kex
        return cls(compile_kex(src, opt), device)

    @classmethod
    def from_file(cls, path: str, opt: int = 3, device: int = 0):
        with open(path, encoding="utf-8") as f:
            return cls.from_kex(f.read(), opt, device)

    def close(self):
        if self._h:
            self._L.kex_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != KEX_OK:
            msg = self._L.kex_strerror(rc).decode()
            if rc == -2:
                msg += ": " + self._L.kex_last_cuda_error(self._h).decode()
            raise KexError(rc, msg)

    def info(self, phase=0):
        i = KexInfo()
        self._check(self._L.kex_info(self._h, phase, ctypes.byref(i)))
        return {n: getattr(i, n) for n, _ in KexInfo._fields_}

    def out_bound(self, n):
        return self._L.kex_out_bound(self._h, n)

    def select_phase(self, phase=0):
        """`--phase N` of the compiled binary: later runs evaluate only phase N
        (1-based); 0 = the whole pipeline."""
        self._check(self._L.kex_select_phase(self._h, phase))

    def set_timing(self, on=True):
        self._L.kex_set_timing(self._h, 1 if on else 0)

    def kernel_ms(self):
        return [self._L.kex_last_kernel_ms(self._h, i) for i in range(4)]

    def launch_count(self):
        return self._L.kex_last_launch_count(self._h)

    def run(self, data: bytes, out_cap=None):
        """`./bin < data` : returns (status, output bytes, count).  status 0 =
        accept, 1 = reject with `count` = the reference's
        "Match error at input symbol <count>!" (C.hs:79-81)."""
        n = len(data)
        cap = out_cap if out_cap is not None else max(4 * n + 4096, 4096)
        while True:
            buf = ctypes.create_string_buffer(cap)
            ol, st, fc = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_size_t()
            rc = self._L.kex_run_host(self._h, data, n, ctypes.cast(buf, ctypes.c_void_p), cap, ctypes.byref(ol),
                                      ctypes.byref(st), ctypes.byref(fc))
            if rc == KEX_ERR_OUT_CAP and out_cap is None:
                cap = ol.value + 4096
                continue
            self._check(rc)
            return st.value, buf.raw[:ol.value], fc.value

    def run_stream(self, blocks, write, block_out_cap=None):
        """`./bin < in > out` with bounded memory: `blocks` yields the input in
        order, `write(bytes)` receives the output.  Like the reference's
        runtime only whole 16 KiB flushes are written until the run accepts
        (crt/crt.c:107-159,217-227).  Returns (status, count)."""
        L = self._L
        self._check(L.kex_stream_begin(self._h))
        ol, st, fc = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_size_t()
        held = bytearray()

        def push(chunk, final_accept=False):
            held.extend(chunk)
            keep = 0 if final_accept else len(held) % 16384
            if len(held) > keep:
                write(bytes(held[:len(held) - keep]))
                del held[:len(held) - keep]

        cap = block_out_cap or 4096
        buf = ctypes.create_string_buffer(cap)

        def call(fn):
            # KEX_ERR_OUT_CAP leaves the stream untouched: repeat with the size it asks for
            nonlocal cap, buf
            while True:
                rc = fn(ctypes.cast(buf, ctypes.c_void_p), cap)
                if rc != KEX_ERR_OUT_CAP:
                    self._check(rc)
                    return
                cap = ol.value + 4096
                buf = ctypes.create_string_buffer(cap)

        for blk in blocks:
            if not blk:
                continue
            if not block_out_cap and cap < 3 * len(blk):
                cap = 3 * len(blk)
                buf = ctypes.create_string_buffer(cap)
            call(lambda b, c: L.kex_stream_feed(self._h, blk, len(blk), b, c, ctypes.byref(ol)))
            push(ctypes.string_at(buf, ol.value))
        call(lambda b, c: L.kex_stream_end(self._h, b, c, ctypes.byref(ol), ctypes.byref(st), ctypes.byref(fc)))
        push(ctypes.string_at(buf, ol.value), final_accept=st.value == ACCEPT)
        return st.value, fc.value

    def run_device(self, d_in: int, n: int, d_out: int, out_cap: int, stream: int = 0):
        """Input and output are device pointers (ints).  Returns
        (status, out_len, count)."""
        ol, st, fc = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_size_t()
        rc = self._L.kex_run_device(self._h, d_in, n, d_out, out_cap, ctypes.byref(ol), ctypes.byref(st),
                                    ctypes.byref(fc), stream)
        if rc == KEX_ERR_OUT_CAP:
            raise KexError(rc, "output buffer too small: need %d bytes" % ol.value)
        self._check(rc)
        return st.value, ol.value, fc.value

    # ---- sharded evaluation (one shard per GPU)
    def shard_summarize(self, d_in: int, n: int, stream: int = 0):
        q1 = self.info()["nstates"] + 1
        m = (ctypes.c_uint16 * q1)()
        self._check(self._L.kex_shard_summarize(self._h, d_in, n, m, stream))
        return list(m)

    def seam_bytes(self):
        return self._L.kex_seam_bytes(self._h)

    def shard_walk(self, start_state: int, stream: int = 0):
        """-> (end state, first failing position or None, seam summary bytes)."""
        end, fail = ctypes.c_uint32(), ctypes.c_size_t()
        seam = (ctypes.c_uint8 * self.seam_bytes())()
        self._check(self._L.kex_shard_walk(self._h, start_state, ctypes.byref(end), ctypes.byref(fail), seam, stream))
        fp = fail.value
        return end.value, (None if fp == ctypes.c_size_t(-1).value else fp), bytes(seam)

    def stitch_live(self, seams, final_code: int):
        """Seam summaries of all shards (in order) + the code of the end-of-input
        action -> seam code at the end of every shard."""
        n = len(seams)
        codes = (ctypes.c_uint32 * n)()
        self._check(self._L.kex_stitch_live(self._h, b"".join(seams), n, final_code, codes))
        return list(codes)

    def set_shard_tail(self, on=True):
        """Opt in to the G-mode tail evaluation for the sharded entry points
        (include/kexcuda.h kex_set_shard_tail): `shard_emit` may then return
        None = every rank repeats walk -> exchange -> stitch -> emit."""
        self._check(self._L.kex_set_shard_tail(self._h, 1 if on else 0))

    def shard_emit(self, seam_code: int, n_eff: int, d_out: int, out_cap: int, stream: int = 0):
        """-> output length, or None when the shard steps must be repeated
        (KEX_RETRY_EXACT, only after `set_shard_tail`)."""
        ol = ctypes.c_size_t()
        rc = self._L.kex_shard_emit(self._h, seam_code, n_eff, d_out, out_cap, ctypes.byref(ol), stream)
        if rc == KEX_ERR_OUT_CAP:
            raise KexError(rc, "output buffer too small: need %d bytes" % ol.value)
        if rc == KEX_RETRY_EXACT:
            return None
        self._check(rc)
        return ol.value

    def final_action(self, state: int):
        """-> (accepting, seam code of the end-of-input action, literal tail)."""
        acc, code = ctypes.c_int(), ctypes.c_uint32()
        tail, tl = ctypes.c_void_p(), ctypes.c_size_t()
        self._check(self._L.kex_final_action(self._h, state, ctypes.byref(acc), ctypes.byref(code),
                                             ctypes.byref(tail), ctypes.byref(tl)))
        data = ctypes.string_at(tail.value, tl.value) if tl.value else b""
        return bool(acc.value), code.value, data
