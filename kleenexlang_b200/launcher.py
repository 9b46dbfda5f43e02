"""The process contract of a binary produced by `kexc compile`
(crt/crt.c:372-467), served by libkexcuda.so:

    ./bin < input > output      exit 0, or exit 1 and
                                "Match error at input symbol <count>!" on stderr
    ./bin -i                    compilation info, exit 2   (crt/crt.c:15,384-386)
    ./bin -h                    usage, exit 1              (crt/crt.c:14,326-332)
    ./bin -t                    also prints "time (ms): N" to stderr (crt/crt.c:457-464)
    ./bin -p N | --phase N      runs only phase N          (crt/crt.c:380-399)

`bench/runningtime.sh:32` (`$p -t`) and `test/test_compiled/runtest.sh:25`
(`echo ... | ./bin`) work unmodified against the generated launcher.
"""
import os
import struct
import sys
import time

RETC_PRINT_USAGE = 1
RETC_PRINT_INFO = 2


def _phase_headers(blob):
    magic, ver, n, total = struct.unpack_from("<IIII", blob, 0)
    out = []
    for i in range(n):
        off, ln = struct.unpack_from("<II", blob, 16 + 8 * i)
        if struct.unpack_from("<I", blob, off)[0] == 0x4158454B:      # action-interpreter phase
            out.append({"action_registers": struct.unpack_from("<I", blob, off + 8)[0]})
            continue
        h = struct.unpack_from("<10I", blob, off)
        out.append({"states": h[2], "classes": h[3], "registers": h[4], "actions": h[5],
                    "monoid_section": struct.unpack_from("<I", blob, off + 80)[0] > 0})
    return out


def describe(blob, meta=None):
    """Text of `-i` / `--srcout`: what `printCompilationInfo` reports for a C
    binary (C.hs:49-52,575-584), for a kexprog blob."""
    lines = ["Compiler info:", "  back end: CUDA (libkexcuda.so, sm_100a)"]
    for k in ("source", "md5", "options"):
        if meta and k in meta:
            lines.append("  %s: %s" % (k, meta[k]))
    ph = _phase_headers(blob)
    lines.append("  phases: %d" % len(ph))
    for i, h in enumerate(ph):
        if "action_registers" in h:
            lines.append("  phase %d: action interpreter over the stream of phase %d, %d registers" % (
                i + 1, i, h["action_registers"]))
            continue
        lines.append("  phase %d: %d SST states, %d byte classes, %d registers, %d actions, monoid tables: %s" % (
            i + 1, h["states"], h["classes"], h["registers"], h["actions"], "yes" if h["monoid_section"] else "no"))
    return "\n".join(lines) + "\n"


def _read_meta(blob_path):
    meta = {}
    try:
        for line in open(blob_path + ".info"):
            k, _, v = line.partition(": ")
            meta[k.strip()] = v.strip()
    except OSError:
        pass
    return meta


def usage(prog):
    return ("Normal usage: %s < infile > outfile\n"
            "- \"%s -i\": Print compilation info\n"
            "- \"%s -t\": Runs normally, but prints timing to stderr\n"
            "- \"%s -p N\": Runs only phase N\n" % (prog, prog, prog, prog))


def main(blob_path, argv):
    prog = argv[0]
    do_timing, phase = False, 0
    args = argv[1:]
    i = 0
    while i < len(args):
        a = args[i]
        if a == "-i":
            sys.stdout.write(describe(open(blob_path, "rb").read(), _read_meta(blob_path)))
            return RETC_PRINT_INFO
        if a == "-t":
            do_timing = True
        elif a in ("-p", "--phase") and i + 1 < len(args):
            i += 1
            phase = int(args[i])
        elif a.startswith("--phase="):
            phase = int(a.split("=", 1)[1])
        elif a.startswith("-p") and len(a) > 2:
            phase = int(a[2:])
        else:               # -h and anything unknown
            sys.stdout.write(usage(prog))
            return RETC_PRINT_USAGE
        i += 1
    t0 = time.time()
    from .runtime import CompiledProgram, KexError, KEX_ERR_UNSUPPORTED
    try:
        return _run(blob_path, phase, do_timing, t0, CompiledProgram, KexError, KEX_ERR_UNSUPPORTED)
    except KexError as e:                       # one line and exit 1, like the runtime's fatal errors
        sys.stdout.buffer.flush()
        sys.stderr.write("%s\n" % e)
        return 1


def _run(blob_path, phase, do_timing, t0, CompiledProgram, KexError, KEX_ERR_UNSUPPORTED):
    cp = CompiledProgram(open(blob_path, "rb").read())
    if phase:
        try:
            cp.select_phase(phase)
        except KexError:
            sys.stderr.write("Invalid phase: %d given\n" % phase)      # C.hs:59-69
            return 1
    # inputs larger than one block are streamed block by block with bounded memory
    # (kex_stream_*); KEX_NO_STREAM=1 reads the whole input first
    block = int(os.environ.get("KEX_STREAM_BLOCK_MIB", "256")) << 20
    first = sys.stdin.buffer.read(block)
    more = sys.stdin.buffer.read(1) if len(first) == block else b""
    streamable = bool(more) and not phase and not os.environ.get("KEX_NO_STREAM")
    if streamable:
        try:
            cp._check(cp._L.kex_stream_begin(cp._h))
        except KexError:
            streamable = False                  # multi-phase program or tables beyond the v3 kernels
    if streamable:
        # Until the first output byte exists the fed blocks are kept: when the library reports that the
        # input cannot be evaluated block by block (a register stays live across blocks: a second block
        # would have to wait, KEX_ERR_UNSUPPORTED) the rest of stdin is read and everything is evaluated
        # at once, as the reference binary would.
        fed, wrote = [], [0]

        def blocks():
            fed.append(first)
            yield first
            head = more
            while True:
                blk = head + sys.stdin.buffer.read(block - len(head))
                head = b""
                if not blk:
                    return
                if not wrote[0]:
                    fed.append(blk)
                else:
                    del fed[:]
                yield blk

        def write(b):
            wrote[0] += len(b)
            sys.stdout.buffer.write(b)

        try:
            status, count = cp.run_stream(blocks(), write)
        except KexError as e:
            if e.code != KEX_ERR_UNSUPPORTED or wrote[0]:
                raise
            data = b"".join(fed) + sys.stdin.buffer.read()
            cp.close()
            cp = CompiledProgram(open(blob_path, "rb").read())
            status, out, count = cp.run(data)
            sys.stdout.buffer.write(out)
    else:
        data = first + more + (sys.stdin.buffer.read() if more else b"")
        status, out, count = cp.run(data)
        sys.stdout.buffer.write(out)
    sys.stdout.buffer.flush()
    if status != 0:
        sys.stderr.write("Match error at input symbol %d!\n" % count)
        return 1
    if do_timing:
        sys.stderr.write("time (ms): %d\n" % int((time.time() - t0) * 1000))
    return 0
