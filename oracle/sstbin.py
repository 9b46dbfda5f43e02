"""ORACLE / TEST INFRASTRUCTURE -- not part of the product path.

Serialises an SST for oracle/kex_oracle.c and binds the C oracle with ctypes.
Assignments are written in the execution order the reference fixes
(`orderAssignments`, src/KMC/SSTCompiler.hs:85-99)."""
import ctypes
import os
import struct
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libkexoracle.so")
SRC = os.path.join(HERE, "kex_oracle.c")


def build_lib(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    return LIB


def _atoms(w):
    out = [len(w)]
    for a in w:
        if a[0] == "v":
            out += [0, a[1]]
        elif a[0] == "c":
            bs = bytes(a[1])
            pad = bs + b"\0" * ((-len(bs)) % 4)
            out += [1, len(bs)] + list(struct.unpack("<%dI" % (len(pad) // 4), pad))
        elif a[0] == "t":
            out += [3] + list(struct.unpack("<64I", bytes(a[1])))
        else:
            out += [2]
    return out


def serialize_sst(sst) -> bytes:
    from kleenexlang_b200.frontend.il import order_assignments
    from kleenexlang_b200.frontend import byteset as BS
    words = [0x5453534B, sst.nstates, sst.nvars, sst.initial]
    for q in range(sst.nstates):
        fin = sst.final.get(q)
        words.append(1 if fin is not None else 0)
        words += _atoms(fin if fin is not None else ())
        es = sorted(sst.edges.get(q, ()), key=lambda e: BS.to_ranges(e[0]))
        words.append(len(es))
        for p, upd, q2 in es:
            words += [(p >> (32 * k)) & 0xFFFFFFFF for k in range(8)]
            order = order_assignments(upd)
            words += [q2, len(order)]
            for v, w in order:
                words.append(v)
                words += _atoms(w)
    return struct.pack("<%dI" % len(words), *words)


_lib = None


def oracle_run(ssts, data: bytes):
    """Run a pipeline of SSTs through the C oracle.  Returns
    (status, output, count) with the reference binary's semantics: on reject
    the output holds only whole 16 KiB flushes and a failing phase hands its
    truncated stream to the next one (crt/crt.c:414-455)."""
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_lib())
        _lib.kex_oracle_run.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t,
                                        ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t),
                                        ctypes.POINTER(ctypes.c_size_t)]
        _lib.kex_oracle_run_stream.argtypes = _lib.kex_oracle_run.argtypes
        _lib.kex_oracle_act.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_char_p,
                                        ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    status, count = 0, 0
    ssts = list(ssts)
    for k, s in enumerate(ssts):
        if hasattr(s, "nregs"):
            # action interpreter phase (frontend/actions.py); a stage that rejected in its
            # transducer phase stays rejected: the stream up to the failing symbol is interpreted
            # and the result keeps whole 16 KiB flushes only (one truncation, on the stage's output)
            buf = ctypes.create_string_buffer(max(1, len(data)))
            ol = ctypes.c_size_t()
            rc = _lib.kex_oracle_act(data, len(data), s.nregs, buf, len(data), ctypes.byref(ol))
            if rc < 0:
                raise RuntimeError("oracle: malformed action stream")
            data = buf.raw[:ol.value]
            if status:
                data = data[:len(data) // 16384 * 16384]
            continue
        blob = s if isinstance(s, (bytes, bytearray)) else serialize_sst(s)
        cap = max(4096, 4 * len(data) + 4096)
        while True:
            buf = ctypes.create_string_buffer(cap)
            ol, cnt = ctypes.c_size_t(), ctypes.c_size_t()
            feeds_interpreter = k + 1 < len(ssts) and hasattr(ssts[k + 1], "nregs")
            run = _lib.kex_oracle_run_stream if feeds_interpreter else _lib.kex_oracle_run
            rc = run(blob, len(blob), data, len(data), buf, cap, ctypes.byref(ol), ctypes.byref(cnt))
            if rc == -3:
                cap = ol.value + 4096
                continue
            if rc < 0:
                raise RuntimeError("oracle: malformed program")
            break
        status, count, data = rc, cnt.value, buf.raw[:ol.value]
    return status, data, count
