"""Parity of the CUDA path (through the C ABI of libkexcuda.so) with the CPU
oracle: reference golden vectors, bundled samples, seeded synthetic inputs,
reject paths, ragged sizes, the sharded entry points, and size-independent
properties at larger sizes.  Bit-exact: this is byte/integer work."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import load_vectors, vec_matches, program_source, sample, ROOT, GOLDEN
from kleenexlang_b200 import workloads
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200.kexprog import compile_kex, UnsupportedProgram
from oracle.sstbin import oracle_run

pytestmark = pytest.mark.gpu
VECS = [v for v in load_vectors() if not v["uses_registers"]]
REF = os.path.join(ROOT, "oracle", "_ref")
_cache = {}


def gpu_prog(src, opt=3, with_fast=True):
    from kleenexlang_b200.runtime import CompiledProgram
    key = (src, opt, with_fast)
    if key not in _cache:
        _cache[key] = (CompiledProgram(compile_kex(src, opt, with_fast)), build_ssts(src, opt))
    return _cache[key]


@pytest.mark.parametrize("v", VECS, ids=[v["name"] for v in VECS])
def test_golden_vectors(v):
    try:
        prog, ssts = gpu_prog(v["program"])
    except UnsupportedProgram:
        pytest.skip("exceeds device register limit")
    st, out, cnt = prog.run(v["input"])
    assert st == 0 and vec_matches(v, out)
    assert (st, out, cnt if st else 0) == (lambda r: (r[0], r[1], r[2] if r[0] else 0))(oracle_run(ssts, v["input"]))


SAMPLES = [("csv2json", "csv_sample.csv"), ("iso_datetime_to_json", "datetime_sample.txt"),
           ("thousand_sep", "numbers_sample.txt"), ("add-commas", "numbers_sample.txt"),
           ("apache_log", "apache_sample.log")]


@pytest.mark.parametrize("name,fixture", SAMPLES)
def test_bundled_samples(name, fixture):
    prog, ssts = gpu_prog(program_source(name))
    d = sample(fixture)
    st, out, _ = prog.run(d)
    est, eout, _ = oracle_run(ssts, d)
    assert (st, out) == (est, eout)


def test_csv2json_pinned_output():
    """The CUDA path against the committed csv2json output that was derived
    from the reference's Ragel equivalent, not from the restated front end
    (scripts/make_csv2json_pin.py), and against that restatement on 4 MiB."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from make_csv2json_pin import csv2json_rl
    prog, _ = gpu_prog(program_source("csv2json"))
    want = open(os.path.join(GOLDEN, "csv2json_sample.out"), "rb").read()
    assert prog.run(sample("csv_sample.csv"))[:2] == (0, want)
    big = workloads.gen_csv(4 << 20, seed=22).tobytes()
    assert prog.run(big)[:2] == (0, csv2json_rl(big))


PROGS = ["csv2json", "iso_datetime_to_json", "thousand_sep", "add-commas", "fastq2fasta"]


def _check(prog, ssts, d):
    st, out, cnt = prog.run(d)
    est, eout, ecnt = oracle_run(ssts, d)
    assert st == est, (st, est, cnt, ecnt)
    if st:
        assert cnt == ecnt
    assert len(out) == len(eout)
    assert out == eout


@pytest.mark.parametrize("name", PROGS)
@pytest.mark.parametrize("opt", [0, 3])
def test_synthetic_vs_oracle(name, opt):
    prog, ssts = gpu_prog(program_source(name), opt)
    d = workloads.GENERATORS[name](3 << 20, seed=21).tobytes()
    _check(prog, ssts, d)


@pytest.mark.parametrize("name", PROGS)
def test_generic_kernels_vs_oracle(name):
    """Blob without the monoid section: the generic kernels (any program up to
    32 registers) must stay bit-exact too."""
    prog, ssts = gpu_prog(program_source(name), 3, with_fast=False)
    assert prog.info()["monoid_kernels"] == 0
    d = workloads.GENERATORS[name](1 << 20, seed=22).tobytes()
    _check(prog, ssts, d)
    _check(prog, ssts, d[:300000] + b"\x01" + d[300001:])


@pytest.mark.parametrize("name", PROGS)
def test_monoid_kernels_selected(name):
    prog, ssts = gpu_prog(program_source(name))
    assert prog.info()["monoid_kernels"] == 1
    # twice: the second run uses the staging window sized from the first
    d = workloads.GENERATORS[name](5 << 20, seed=23).tobytes()
    _check(prog, ssts, d)
    _check(prog, ssts, d)


def test_output_capacity_error():
    from kleenexlang_b200.runtime import KexError
    prog, ssts = gpu_prog(program_source("csv2json"))
    d = workloads.gen_csv(200000, seed=5).tobytes()
    with pytest.raises(KexError) as ei:
        prog.run(d, out_cap=len(d))
    assert ei.value.code == -3
    _check(prog, ssts, d)


@pytest.mark.parametrize("name", PROGS)
def test_ragged_sizes(name):
    prog, ssts = gpu_prog(program_source(name))
    d = workloads.GENERATORS[name](70000, seed=4).tobytes()
    for n in (0, 1, 2, 31, 32, 33, 4095, 4096, 4097, 8191, 8192, 12289, 65536, len(d)):
        _check(prog, ssts, d[:n])


@pytest.mark.parametrize("name", PROGS)
def test_reject_paths(name):
    prog, ssts = gpu_prog(program_source(name))
    d = workloads.GENERATORS[name](400000, seed=9).tobytes()
    for pos in (0, 1, 4095, 4096, 100000, 262145, len(d) - 1):
        bad = d[:pos] + b"\x01" + d[pos + 1:]
        _check(prog, ssts, bad)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "csv2json")), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("name", PROGS)
def test_vs_reference_binary(name):
    prog, _ = gpu_prog(program_source(name))
    d = workloads.GENERATORS[name](8 << 20, seed=33).tobytes()
    r = subprocess.run([os.path.join(REF, name)], input=d, capture_output=True)
    st, out, _ = prog.run(d)
    assert st == r.returncode
    assert hashlib.sha256(out).hexdigest() == hashlib.sha256(r.stdout).hexdigest()


@pytest.mark.parametrize("name", PROGS)
def test_host_pipeline(name, monkeypatch):
    """kex_run_host cuts large inputs into sub-waves that are copied, evaluated
    as consecutive shards and copied back on three streams; with 1 MiB sub-waves
    a 5 MiB input exercises every seam (accept, reject in the first / a middle /
    the last sub-wave, output capacity error)."""
    from kleenexlang_b200.runtime import KexError
    prog, ssts = gpu_prog(program_source(name))
    monkeypatch.setenv("KEX_HOST_WAVE_MIB", "1")
    d = workloads.GENERATORS[name](5 << 20, seed=41).tobytes()
    _check(prog, ssts, d)
    for pos in (5, (1 << 20) - 1, 1 << 20, (2 << 20) + 12345, len(d) - 2):
        _check(prog, ssts, d[:pos] + b"\x01" + d[pos + 1:])
    with pytest.raises(KexError) as ei:
        prog.run(d, out_cap=len(d) // 2)
    assert ei.value.code == -3
    monkeypatch.setenv("KEX_NO_HOST_PIPELINE", "1")
    _check(prog, ssts, d)
    # tapered schedule (quarter and half sub-waves first and last) needs >= 6 sub-waves of >= 4 MiB
    monkeypatch.delenv("KEX_NO_HOST_PIPELINE")
    monkeypatch.setenv("KEX_HOST_WAVE_MIB", "4")
    big = workloads.GENERATORS[name](27 << 20, seed=42).tobytes()
    _check(prog, ssts, big[:len(big) - 12345])


def test_pipeline_program():
    src = 'start: p >> a >> b\np := ~/abc/ "a"\na := /./ "b"\nb := /ab/ "c"\n   | ~/[^ab]/ "lol"\n'
    prog, ssts = gpu_prog(src)
    assert prog.info()["nphases"] == 3
    assert prog.run(b"abc") == (0, b"abc", 0)
    _check(prog, ssts, b"abd")


@pytest.mark.parametrize("name", ["csv2json", "iso_datetime_to_json", "thousand_sep"])
def test_sharded_entry_points(name):
    """Three shards cut at arbitrary byte positions, stitched with the state
    and fate maps exactly as the multi-GPU launcher does."""
    import torch
    from kleenexlang_b200.runtime import CompiledProgram
    from kleenexlang_b200.sharding import stitch_states, stitch_live
    blob = compile_kex(program_source(name))
    ssts = build_ssts(program_source(name))
    d = workloads.GENERATORS[name](600000, seed=2).tobytes()
    cuts = [0, 199993, 400016, len(d)]       # multiples of 16 not required between ranks; each shard is its own buffer
    shards = [d[cuts[i]:cuts[i + 1]] for i in range(3)]
    progs = [CompiledProgram(blob) for _ in shards]
    bufs = [torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda() for s in shards]
    maps = [p.shard_summarize(b.data_ptr(), b.numel()) for p, b in zip(progs, bufs)]
    init = 0
    info = progs[0].info()
    starts = stitch_states(maps, build_ssts(program_source(name))[0].initial)
    walks = [p.shard_walk(s) for p, s in zip(progs, starts)]
    assert all(w[1] is None for w in walks)
    acc, code, tail = progs[-1].final_action(walks[-1][0])
    assert acc
    lives = stitch_live(progs[0], [w[2] for w in walks], code)
    from kleenexlang_b200.sharding import stitch_codes
    assert lives == stitch_codes([list(w[2]) for w in walks], code)
    outs = []
    for p, b, live in zip(progs, bufs, lives):
        o = torch.empty(6 * b.numel() + 64, dtype=torch.uint8, device="cuda")
        n = p.shard_emit(live, b.numel(), o.data_ptr(), o.numel())
        outs.append(bytes(o[:n].cpu().numpy()))
    got = b"".join(outs) + tail
    assert (0, got) == oracle_run(ssts, d)[:2]


def test_large_properties_csv2json():
    """256 MiB: tiling the input tiles the output (record* grammar), and the
    output length is input + 127 bytes per row."""
    import torch
    prog, ssts = gpu_prog(program_source("csv2json"))
    block = workloads.gen_csv(4 << 20, seed=77)
    reps = 64
    rows = int((block == 10).sum())
    dev_block = torch.from_numpy(block).cuda()
    big = dev_block.repeat(reps)
    out = torch.empty(int(big.numel() * 2.2), dtype=torch.uint8, device="cuda")
    st, olen, _ = prog.run_device(big.data_ptr(), big.numel(), out.data_ptr(), out.numel())
    assert st == 0 and olen == big.numel() + 127 * rows * reps
    est, eout, _ = oracle_run(ssts, block.tobytes())
    per = len(eout)
    assert olen == per * reps
    o = out[:olen].view(reps, per)
    ref = torch.frombuffer(bytearray(eout), dtype=torch.uint8).cuda()
    assert bool((o == ref[None, :]).all())


@pytest.mark.parametrize("name,block_mib,reps", [("iso_datetime_to_json", 2, 64), ("fastq2fasta", 4, 64),
                                                 ("thousand_sep", 4, 32), ("add-commas", 4, 32)])
def test_large_properties_tiled(name, block_mib, reps):
    """BASELINE configs at sizes the oracle cannot reach in seconds: the grammars
    are record*, so tiling a block of whole records tiles the output (checked
    bit-exact against the oracle's output for one block), on the device-resident
    path and through the host pipeline."""
    import torch
    prog, ssts = gpu_prog(program_source(name))
    block = workloads.GENERATORS[name](block_mib << 20, seed=91)
    est, eout, _ = oracle_run(ssts, block.tobytes())
    assert est == 0
    per = len(eout)
    big = torch.from_numpy(block).cuda().repeat(reps)
    out = torch.empty(per * reps + 4096, dtype=torch.uint8, device="cuda")
    st, olen, _ = prog.run_device(big.data_ptr(), big.numel(), out.data_ptr(), out.numel())
    assert st == 0 and olen == per * reps
    ref = torch.frombuffer(bytearray(eout), dtype=torch.uint8).cuda()
    assert bool((out[:olen].view(reps, per) == ref[None, :]).all())
    # the same bytes through kex_run_host (sub-wave pipeline over three streams)
    st2, hout, _ = prog.run(bytes(big[: 8 * len(block)].cpu().numpy()))
    assert st2 == 0 and hout == eout * 8


@pytest.mark.parametrize("name", PROGS)
@pytest.mark.parametrize("knob", [{"KEX_V3_NOSPEC": "1"}, {"KEX_V3_STAGE": "2048", "KEX_V3_RECCAP": "8"},
                                  {"KEX_V3_NOLIT": "1"}, {"KEX_NO_V3": "1"}, {"KEX_NO_V4": "1", "KEX_V3_NORMW": "1"},
                                  {"KEX_NO_V4": "1", "KEX_V3_WORKERS": "2"}])
def test_v3_paths_forced(name, knob, monkeypatch):
    """The rarely taken paths of the v3 kernels, forced: exact live sets for
    every tile (no guessing), tiles that do not fit the staging window / record
    slots (byte stores to global), and the CTA-tile monoid kernels that take
    over when a program exceeds the v3 table limits."""
    from kleenexlang_b200.runtime import CompiledProgram
    for k, v in knob.items():
        monkeypatch.setenv(k, v)
    prog = CompiledProgram(compile_kex(program_source(name)))          # KEX_NO_V3 is read at load time
    ssts = build_ssts(program_source(name))
    assert prog.info()["chunk_bytes"] == (4096 if "KEX_NO_V3" in knob else 1024)
    d = workloads.GENERATORS[name](2 << 20, seed=61).tobytes()
    _check(prog, ssts, d)
    _check(prog, ssts, d[:700001])
    _check(prog, ssts, d[:1000000] + b"\x01" + d[1000001:])


V4_PROGS = ["csv2json", "fastq2fasta"]      # iso_datetime: 37 x 14 entries do not fit per-lane copies -> k3_emit


@pytest.mark.parametrize("name", PROGS)
@pytest.mark.parametrize("knob", [{}, {"KEX_V4_EXACT": "1"}, {"KEX_V4_STAGE": "256", "KEX_V4_RECCAP": "8"}, {"KEX_NO_V4": "1"},
                                  {"KEX_V3_WORKERS": "3"}, {"KEX_V4_LOG6": "1"}, {"KEX_V4_NOTAIL": "1"}, {"KEX_V4_TAIL_TEST": "1"},
                                  {"KEX_V4_NORMW": "1"}])
def test_v4_paths_forced(name, knob, monkeypatch):
    """The G-mode emit kernel (kex_v4.cuh) and its rarely taken paths, forced:
    every tile evaluated exactly from the tables in global memory
    (KEX_V4_EXACT), tiles that do not fit the staging window / record slots
    (byte stores to global), k3_emit instead (KEX_NO_V4), few worker warps per
    CTA (many groups per CTA: the chained scan and the deferred stage-out).
    Inputs long enough for G to be learnt from the run itself (>= 2 chunk
    boundaries), twice (the second run uses the learnt table and the window
    sized from the first -- and, from then on, the tail evaluation: k3_seams and
    the live-set tree over the last 256 tiles only; KEX_V4_NOTAIL keeps them
    over everything, KEX_V4_TAIL_TEST makes the kernel report a broken induction
    so that the host repeats the run exactly), truncated, and rejecting."""
    from kleenexlang_b200.runtime import CompiledProgram
    for k, v in knob.items():
        monkeypatch.setenv(k, v)
    prog = CompiledProgram(compile_kex(program_source(name)))          # KEX_NO_V4 is read at load time
    ssts = build_ssts(program_source(name))
    assert prog.info()["chunk_bytes"] == 1024
    d = workloads.GENERATORS[name](3 << 20, seed=62).tobytes()
    _check(prog, ssts, d)
    _check(prog, ssts, d)
    _check(prog, ssts, d[:700001])
    _check(prog, ssts, d[:5000])
    _check(prog, ssts, d[:1000000] + b"\x01" + d[1000001:])
    prog.close()


@pytest.mark.parametrize("name", V4_PROGS)
def test_v4_gmode_taken(name, monkeypatch, capfd):
    """csv2json and fastq2fasta have state-determined live sets (and tables that fit):
    once G is learnt from the run, (nearly) every tile goes through the G-mode
    passes -- the library reports how many tiles it evaluated exactly."""
    from kleenexlang_b200.runtime import CompiledProgram
    prog = CompiledProgram(compile_kex(program_source(name)))
    ssts = build_ssts(program_source(name))
    d = workloads.GENERATORS[name](4 << 20, seed=63).tobytes()
    _check(prog, ssts, d)
    inf = prog.info()
    assert inf["emit_kernel"] == 4
    assert inf["exact_tiles"] <= 4, inf
    prog.close()


def test_v4_not_state_determined():
    """thousand_sep: the live sets follow the digit count, not the state; after
    a few runs with most tiles evaluated exactly the phase goes back to k3_emit
    (output identical all along)."""
    from kleenexlang_b200.runtime import CompiledProgram
    prog = CompiledProgram(compile_kex(program_source("thousand_sep")))
    ssts = build_ssts(program_source("thousand_sep"))
    d = workloads.GENERATORS["thousand_sep"](2 << 20, seed=64).tobytes()
    for _ in range(5):
        _check(prog, ssts, d)
    assert prog.info()["emit_kernel"] == 3
    prog.close()


@pytest.mark.parametrize("nolit", [False, True])
def test_one_byte_literals(nolit, monkeypatch):
    """An action that replaces its input byte by a different one-byte literal:
    the write pass of k3_emit stores it directly (no template record); with
    KEX_V3_NOLIT the same emission goes through a template record."""
    from kleenexlang_b200.runtime import CompiledProgram
    if nolit:
        monkeypatch.setenv("KEX_V3_NOLIT", "1")
    src = 'main := (~/a/ "x" | /b/ | ~/c/ "yz" | /\\n/)*\n'
    prog = CompiledProgram(compile_kex(src))
    ssts = build_ssts(src)
    assert prog.info()["chunk_bytes"] == 1024
    rng = np.random.default_rng(7)
    d = bytes(rng.choice(np.frombuffer(b"aabbbc\n", dtype=np.uint8), size=1 << 20))
    _check(prog, ssts, d)
    _check(prog, ssts, d[:333333] + b"q" + d[333334:])
    st, out, _ = prog.run(b"abcab\n")
    assert (st, out) == (0, b"xbyzxb\n")


def _random_program(rng):
    """A small well-formed Kleenex program: a starred choice of terms built from
    byte classes over {a..e, newline}, suppression, output literals and
    bounded repetition (ordered choice makes overlapping alternatives legal)."""
    atoms = ["/a/", "/b/", "/[ab]/", "/[c-e]/", "/[a-e]/", "/\\n/", "/ab/", "/a+/", "/[bc]*d/", "/e{2,3}/"]
    lits = ['"x"', '"yz"', '"<"', '">\\n"', '"--"', '"0123456789"']

    def term(depth):
        k = rng.integers(0, 7)
        if k == 0:
            return atoms[rng.integers(len(atoms))]
        if k == 1:
            return "~" + atoms[rng.integers(len(atoms))]
        if k == 2:
            return lits[rng.integers(len(lits))] + " " + atoms[rng.integers(len(atoms))]
        if k == 3:
            return atoms[rng.integers(len(atoms))] + " " + lits[rng.integers(len(lits))]
        if k == 4 and depth < 2:
            return "(" + term(depth + 1) + " | " + term(depth + 1) + ")"
        if k == 5 and depth < 2:
            return "(" + term(depth + 1) + " " + term(depth + 1) + ")"
        return "~" + atoms[rng.integers(len(atoms))] + " " + lits[rng.integers(len(lits))]

    n = int(rng.integers(2, 5))
    return "main := (" + " | ".join(term(0) for _ in range(n)) + ")*\n"


def test_random_programs_vs_oracle():
    """Seeded random programs x random inputs (accepting prefixes and rejects):
    exercises table construction, kernel selection and every kernel family on
    shapes the bundled programs do not have."""
    from kleenexlang_b200.runtime import CompiledProgram, KexError
    rng = np.random.default_rng(20261017)
    alphabet = np.frombuffer(b"aaabbbcde\n", dtype=np.uint8)
    done = accepted = 0
    for _ in range(60):
        src = _random_program(rng)
        try:
            ssts = build_ssts(src)
            blob = compile_kex(src)
        except Exception:
            continue                                   # not well-formed / exceeds a front-end limit
        try:
            prog = CompiledProgram(blob)
        except KexError:
            continue
        for n in (0, 1, 37, 5000, 70001):
            d = bytes(rng.choice(alphabet, size=n))
            got, exp = prog.run(d), oracle_run(ssts, d)
            assert got[:2] == exp[:2] and (got[0] == 0 or got[2] == exp[2]), (src, n)
            if exp[0]:
                # the consumed prefix, tiled: mostly accepting runs with real output
                pre = d[:exp[2]]
                if pre:
                    big = pre * max(1, 40000 // len(pre))
                    got, exp = prog.run(big), oracle_run(ssts, big)
                    assert got[:2] == exp[:2] and (got[0] == 0 or got[2] == exp[2]), (src, n, "tiled prefix")
                    accepted += exp[0] == 0
        done += 1
        prog.close()
    assert done >= 20 and accepted >= 20


@pytest.mark.parametrize("name", ["csv2json", "iso_datetime_to_json", "fastq2fasta", "thousand_sep"])
def test_block_streaming(name):
    """kex_stream_*: the input is fed in blocks (any sizes), each block's output
    comes back one block later, whole 16 KiB flushes only until the run accepts;
    equal to the whole-input run on accepting and rejecting inputs."""
    prog, ssts = gpu_prog(program_source(name))
    d = workloads.GENERATORS[name](3 << 20, seed=71).tobytes()
    rng = np.random.default_rng(3)

    def blocks_of(data):
        pos = 0
        while pos < len(data):
            n = int(rng.integers(1, 700000))
            yield data[pos:pos + n]
            pos += n

    for data in (d, d[:1500000] + b"\x01" + d[1500001:], d[:5] + b"\x01" + d[6:], d[:len(d) - 7], d[:100]):
        out = bytearray()
        st, cnt = prog.run_stream(blocks_of(data), out.extend)
        est, eout, ecnt = oracle_run(ssts, data)
        assert st == est and bytes(out) == eout and (st == 0 or cnt == ecnt), (name, len(data))
    out = bytearray()
    assert prog.run_stream(iter(()), out.extend)[0] == oracle_run(ssts, b"")[0]


def test_block_streaming_unsupported_program():
    from kleenexlang_b200.runtime import KexError
    src = 'start: p >> a\np := /[ab]/*\na := /a/ "x" | /b/\n'
    prog, _ = gpu_prog(src)
    with pytest.raises(KexError) as ei:
        prog.run_stream(iter([b"ab"]), lambda b: None)
    assert ei.value.code == -4


@pytest.mark.parametrize("name", ["csv2json", "fastq2fasta"])
@pytest.mark.parametrize("force_retry", [False, True])
def test_sharded_tail_evaluation(name, force_retry, monkeypatch):
    """kex_set_shard_tail: three shards of 2 MiB cut at arbitrary positions,
    evaluated twice -- the first round learns G (exact live sets everywhere),
    the second uses the tail evaluation: the seam summary is the constant map
    to G[start state] and kex_shard_emit verifies it.  With KEX_V4_TAIL_TEST the
    kernel reports a broken induction: every shard repeats walk / stitch / emit
    (KEX_RETRY_EXACT) and the library evaluates exactly from then on."""
    import torch
    from kleenexlang_b200.runtime import CompiledProgram
    from kleenexlang_b200.sharding import stitch_states, stitch_live
    if force_retry:
        monkeypatch.setenv("KEX_V4_TAIL_TEST", "1")
    blob = compile_kex(program_source(name))
    ssts = build_ssts(program_source(name))
    d = workloads.GENERATORS[name](6 << 20, seed=4).tobytes()
    cuts = [0, 2 * 1048576 + 77777, 4 * 1048576 + 1234, len(d)]
    shards = [d[cuts[i]:cuts[i + 1]] for i in range(3)]
    progs = [CompiledProgram(blob) for _ in shards]
    for p in progs:
        p.set_shard_tail(True)
    bufs = [torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda() for s in shards]
    outs_buf = [torch.empty(6 * b.numel() + 64, dtype=torch.uint8, device="cuda") for b in bufs]
    expect = oracle_run(ssts, d)
    retried = 0
    for rnd in range(3):
        maps = [p.shard_summarize(b.data_ptr(), b.numel()) for p, b in zip(progs, bufs)]
        starts = stitch_states(maps, ssts[0].initial)
        while True:
            walks = [p.shard_walk(s) for p, s in zip(progs, starts)]
            assert all(w[1] is None for w in walks)
            acc, code, tail = progs[-1].final_action(walks[-1][0])
            lives = stitch_live(progs[0], [w[2] for w in walks], code)
            lens = [p.shard_emit(live, b.numel(), o.data_ptr(), o.numel()) for p, b, o, live in zip(progs, bufs, outs_buf, lives)]
            if all(n is not None for n in lens):
                break
            retried += 1
            assert retried < 4
        got = b"".join(bytes(o[:n].cpu().numpy()) for o, n in zip(outs_buf, lens)) + tail
        assert (0, got) == expect[:2], rnd
    assert (retried > 0) == (force_retry and progs[0].info()["emit_kernel"] == 4)


# ---- the reference's register-free benchmark programs (tests/golden/reference_bench_vectors.json:
# random walks through each program's own grammar, expected = lockstep simulation)
def _bench_vectors():
    import base64, json
    vs = json.load(open(os.path.join(GOLDEN, "reference_bench_vectors.json")))
    for v in vs:
        v["input"], v["output"] = base64.b64decode(v["input"]), base64.b64decode(v["output"])
    return vs


BENCH = _bench_vectors()


@pytest.mark.parametrize("name", sorted({v["name"] for v in BENCH}))
def test_reference_bench_programs(name):
    """All 41 of bench/kleenex/src/*.kex without register actions, whichever kernel family their tables
    select (monoid v4 / v3 / CTA-tile / generic) and whichever --opt fallback compile_kex takes."""
    from kleenexlang_b200.runtime import CompiledProgram
    vs = [v for v in BENCH if v["name"] == name]
    prog = CompiledProgram(compile_kex(vs[0]["program"]))
    for v in vs:
        st, out, _ = prog.run(v["input"])
        assert (st, out) == (0, v["output"]), (name, len(v["input"]))
    # many tiles: the longest vector repeated (accepted where the grammar is a loop, else a reject
    # somewhere inside), against the C oracle on the SSTs
    big = max(vs, key=lambda v: len(v["input"]))["input"] * 40
    ssts = build_ssts(vs[0]["program"], 3)
    est, eout, ecnt = oracle_run(ssts, big)
    st, out, cnt = prog.run(big)
    assert (st, out) == (est, eout)
    if st and len(ssts) == 1:
        assert cnt == ecnt
    prog.close()


RE_CASES = [("(a|b)*c[0-9]+", [b"abbac2016", b"c7", b"bbbbc00", b"abx", b"c", b"", b"ab" * 5000 + b"c" + b"7" * 3000]),
            ("(?:ab|a)(?:bc|c)?x{2,3}", [b"abxx", b"abcxxx", b"acxx", b"abxxxxx", b"abx", b"abcxxxxxx"]),
            ("[a-z]+(,[a-z]+)*\\n", [b"ab,c,def\n", b"q\n", b"ab,,c\n", b"ab", b",".join([b"kleenex"] * 4000) + b"\n"])]


@pytest.mark.parametrize("re_src,inputs", RE_CASES)
@pytest.mark.parametrize("sb", [False, True])
def test_regex_coder_on_device(re_src, inputs, sb):
    """`kexc compile x.re` (compileCoder, Commands.hs:246-275): the device writes the bit-coded parse; the
    coder's tables (AppendTblI) are lowered to constants per byte.  Checked against the C oracle running
    the SST with its table atoms."""
    from kleenexlang_b200.runtime import CompiledProgram
    from kleenexlang_b200.kexprog import compile_re
    from kleenexlang_b200.frontend.driver import build_coder_ssts
    coder = build_coder_ssts(re_src, 3, lookahead=False, suppress_bits=sb)
    prog = CompiledProgram(compile_re(re_src, 3, suppress_bits=sb))
    for d in inputs:
        est, eout, ecnt = oracle_run(coder, d)
        st, out, cnt = prog.run(d)
        assert (st, out) == (est, eout)
        if st:
            assert cnt == ecnt
    prog.close()


@pytest.mark.parametrize("name", ["csv2json", "iso_datetime_to_json", "thousand_sep"])
def test_reference_phases_on_device(name):
    """`kexc compile --phases=reference`: oracle phase + action phase per stage as in the reference's
    default build; `-p 1` gives the reference's code (oracle SST with table atoms under the C oracle),
    the whole pipeline gives what the direct SST gives."""
    from kleenexlang_b200.runtime import CompiledProgram
    from kleenexlang_b200.kexprog import compile_reference_phases
    from kleenexlang_b200.frontend.driver import build_oracle_action_pipeline
    src = program_source(name)
    ref = build_oracle_action_pipeline(src, 3, lookahead=False, suppress_bits=True)
    prog = CompiledProgram(compile_reference_phases(src, 3, suppress_bits=True))
    direct, ssts = gpu_prog(src)
    d = workloads.GENERATORS[name](300000, seed=31).tobytes()
    assert prog.run(d)[:2] == direct.run(d)[:2]
    # a reject inside the oracle phase: its code is cut to whole 16 KiB flushes and the action phase runs on
    # that (crt/crt.c:414-455) -- the two-phase pipeline under the C oracle is the expectation, not the direct SST
    for data in (d, d[:100000] + b"\x01" + d[100001:], b""):
        assert prog.run(data)[:2] == oracle_run(ref, data)[:2]
    est, code, _ = oracle_run(ref[:1], d)
    prog.select_phase(1)
    assert prog.run(d)[:2] == (est, code)
    prog.select_phase(2)
    assert prog.run(code)[:2] == direct.run(d)[:2]
    prog.select_phase(0)
    prog.close()


@pytest.mark.parametrize("v", VECS, ids=[v["name"] for v in VECS])
def test_oracle_code_golden_vectors(v):
    from kleenexlang_b200.runtime import CompiledProgram
    from kleenexlang_b200.kexprog import compile_reference_phases
    try:
        prog = CompiledProgram(compile_reference_phases(v["program"], 3, suppress_bits=True))
    except UnsupportedProgram:
        pytest.skip("exceeds device register limit")
    st, out, _ = prog.run(v["input"])
    assert st == 0 and vec_matches(v, out)
    prog.close()
