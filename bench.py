#!/usr/bin/env python
"""bench.py -- input GiB/s of the compiled-transducer hot path (default: csv2json.kex).

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K ...   (reference C binary on host cores)
  python bench.py --config {csv2json,add-commas,iso_datetime,fastq2fasta} [--scaling strong]

A step is one pass of the hot path over the whole synthetic input that is
already resident in HBM (default: BASELINE.json config "csv2json.kex on 16 GiB
synthetic CSV, 1 GPU"; 16 GiB per GPU at N>1, i.e. weak scaling).  At N>1 the
input is ONE stream cut at arbitrary (not record-aligned) byte offsets, one
shard per rank.  After the timed region every rank compares its output on the
device with the reference C binary's output of the 64 MiB block the stream is
tiled from (`config.verified`).  One JSON line on stdout.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GIB = float(1 << 30)
# BASELINE.json configs: program file, workload generator, size, scaling, and dram bytes per input
# byte of one emit launch from the committed `ncu --set full` capture (None: not captured).
CONFIGS = {
    "csv2json": {"program": "csv2json", "gen": "csv2json", "gib": 16.0, "scaling": "weak",
                 "what": "csv2json.kex on %.2f GiB synthetic CSV per GPU (gen_csv.pl distribution)",
                 # profiles/r02_ncu_k4_full_16gib.txt: k4_emit dram read + write per launch / input bytes
                 "emit_traffic_per_in": (19.540329 + 33.476041) / 17.179869},
    "add-commas": {"program": "add-commas", "gen": "add-commas", "gib": 4.0, "scaling": "weak",
                   "what": "add-commas.kex (README example) on %.2f GiB random digits per GPU (gen_numbers.pl, avglen 1000)",
                   "emit_traffic_per_in": None},
    "iso_datetime": {"program": "iso_datetime_to_json", "gen": "iso_datetime_to_json", "gib": 32.0, "scaling": "strong",
                     "what": "iso_datetime_to_json.kex on %.2f GiB of timestamps (gen_datetime.pl distribution)",
                     "emit_traffic_per_in": None},
    "fastq2fasta": {"program": "fastq2fasta", "gen": "fastq2fasta", "gib": 32.0, "scaling": "weak",
                    "what": "fastq2fasta.kex on %.2f GiB synthetic FASTQ per GPU (256 GiB over 8 GPUs)",
                    "emit_traffic_per_in": None},
}
PROGRAM = "csv2json"          # set from --config in main()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons; started before the warm-up so that it
    is already polling when the timed region begins, `mark()` brackets the region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None
        self.lo = 0
        self.hi = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def mark(self, end=False):
        if end:
            self.hi = len(self.rows)
        else:
            self.lo = len(self.rows)

    def stop(self):
        if self.proc:
            self.proc.terminate()
        rows = self.rows[self.lo:self.hi] or self.rows[max(0, self.lo - 1):(self.hi or 0) + 1]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------ reference arm
def ref_binary(variant=""):
    """oracle/_ref/<program>: `--act=false` (one process per instance); variant ".act": the
    reference's default mode `--act=true` (oracle + action program, two processes per instance)."""
    p = os.path.join(ROOT, "oracle", "_ref", PROGRAM + variant)
    return p if os.path.exists(p) else None


def gen_block(gen_name, nbytes=64 << 20, seed=100):
    """A seeded block of whole records whose length is a multiple of 16 (so that
    multiples of it are 16-byte aligned offsets into the tiled input)."""
    import numpy as np
    from kleenexlang_b200 import workloads
    block = workloads.GENERATORS[gen_name](nbytes, seed=seed)
    per_record = 4 if gen_name == "fastq2fasta" else 1
    ends = np.flatnonzero(block == 10)[per_record - 1::per_record] + 1
    good = ends[ends % 16 == 0]
    assert len(good), "no record boundary at a multiple of 16"
    return np.ascontiguousarray(block[:int(good[-1])])


def reference_output(data: bytes):
    """Output of the reference's compiled C binary (oracle/_ref) for `data`; the
    oracle port when the binary is absent.  -> (bytes, description)"""
    binp = ref_binary()
    if binp:
        r = subprocess.run([binp], input=data, capture_output=True)
        assert r.returncode == 0, "reference binary rejected the synthetic input"
        return r.stdout, "oracle/_ref/%s (emitted C + verbatim crt.c)" % PROGRAM
    from kleenexlang_b200.frontend.driver import build_ssts
    from oracle.sstbin import oracle_run
    src = open(os.path.join(ROOT, "programs", PROGRAM + ".kex"), encoding="utf-8").read()
    st, out, _ = oracle_run(build_ssts(src), data)
    assert st == 0
    return out, "oracle/kex_oracle.c port"


VARIANT_FLAGS = {"": "--act=false --la=false (1 process)", ".la": "--act=false --la=true (1 process)",
                 ".act": "--act=true --sb=true --la=false (oracle + action program: 2 processes)",
                 ".default": "--act=true --sb=true --la=true (the reference's default flags: 2 processes)"}


def make_reference_inputs(gen_name, sample_bytes, instances):
    """One record-aligned input file per instance in /dev/shm (written once, reused by every step)."""
    block = gen_block(gen_name, min(sample_bytes, 64 << 20), seed=1234)
    reps = max(1, sample_bytes // len(block))
    files = []
    for i in range(instances):
        f = "/dev/shm/kexbench_%d_%d.in" % (os.getpid(), i)
        with open(f, "wb") as fh:
            for _ in range(reps):
                fh.write(block.tobytes())
        files.append(f)
    return files, reps * len(block) * instances


def run_reference_once(files, variant=""):
    """Times one concurrent pass of the reference's compiled C binary
    (oracle/_ref, emitted C + verbatim crt.c, `cc -O3 -D FLAG_WORDALIGNED`) over
    the files, stdout to /dev/null as bench/runningtime.sh:101 does."""
    binp = ref_binary(variant)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([binp], stdin=open(f, "rb"), stdout=subprocess.DEVNULL) for f in files]
    rcs = [p.wait() for p in procs]
    dt = time.perf_counter() - t0
    assert all(rc == 0 for rc in rcs), "reference binary rejected the synthetic input"
    return dt


def time_reference(gen_name, sample_bytes, instances, variant=""):
    files, total = make_reference_inputs(gen_name, sample_bytes, instances)
    try:
        return run_reference_once(files, variant), total
    finally:
        for f in files:
            os.unlink(f)


def time_oracle_port(gen_name, sample_bytes):
    from kleenexlang_b200.frontend.driver import build_ssts
    from oracle.sstbin import oracle_run, serialize_sst
    src = open(os.path.join(ROOT, "programs", PROGRAM + ".kex"), encoding="utf-8").read()
    blobs = [serialize_sst(s) for s in build_ssts(src)]
    data = gen_block(gen_name, sample_bytes, seed=1234).tobytes()
    t0 = time.perf_counter()
    st, _, _ = oracle_run(blobs, data)
    dt = time.perf_counter() - t0
    assert st == 0
    return dt, len(data)


def cpu_baseline(gen_name, sample_bytes=1 << 30):
    if ref_binary():
        dt, nb = time_reference(gen_name, sample_bytes, 1)
        variants = {VARIANT_FLAGS[""]: nb / GIB / dt}
        for v in (".la", ".act", ".default"):
            if ref_binary(v):
                dt2, nb2 = time_reference(gen_name, sample_bytes // 4, 1, v)
                variants[VARIANT_FLAGS[v]] = nb2 / GIB / dt2
        return {"value": nb / GIB / dt, "unit": "GiB/s", "cores": 1, "kind": "reference", "variants": variants,
                "sample": "%d MiB synthetic input, 1 process of oracle/_ref/%s (emitted C + verbatim crt.c, "
                          "cc -O3 -D FLAG_WORDALIGNED, --opt 3 --la=false --act=false), stdin from /dev/shm, "
                          "stdout /dev/null" % (nb >> 20, PROGRAM)}
    dt, nb = time_oracle_port(gen_name, 64 << 20)
    return {"value": nb / GIB / dt, "unit": "GiB/s", "cores": 1, "kind": "port",
            "sample": "%d MiB synthetic input through oracle/kex_oracle.c (interpreting C restatement)" % (nb >> 20)}


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = "reference" if ref_binary() else "port"
    # bounded sample: at most ~8 GiB of /dev/shm over all processes, 64-256 MiB each
    per = max(64 << 20, min(256 << 20, (8 << 30) // cores))
    vals = []
    variants = {}
    files, total = make_reference_inputs(cfg["gen"], per, cores) if kind == "reference" else ([], 0)
    try:
        for i in range(args.warmup + args.steps):
            if kind == "reference":
                dt, nb = run_reference_once(files), total
            else:
                dt, nb = time_oracle_port(cfg["gen"], 32 << 20)
            if i >= args.warmup:
                vals.append((dt, nb))
        # the other flag sets of the reference's C back end over the same files: one warm-up pass and up
        # to three timed ones each (the two-process variants run one instance per two cores); the line's
        # value is the fastest flag set
        if kind == "reference":
            for v in (".la", ".act", ".default"):
                if ref_binary(v):
                    fs = files if v == ".la" else files[:max(1, cores // 2)]
                    run_reference_once(fs, v)                                   # warm-up pass
                    k = max(1, min(args.steps, 3))
                    dt = sum(run_reference_once(fs, v) for _ in range(k)) / k
                    variants[VARIANT_FLAGS[v]] = (total * len(fs) / len(files)) / GIB / dt
    finally:
        for f in files:
            os.unlink(f)
    tot_t = sum(v[0] for v in vals)
    tot_b = sum(v[1] for v in vals)
    value = tot_b / GIB / tot_t
    variants[VARIANT_FLAGS[""]] = value
    best = max(variants, key=variants.get)
    flags_used = best if kind == "reference" else "port"
    value = variants[best]
    used = cores if kind == "reference" else 1
    line = {"impl": "reference", "metric": "input GiB/s on %s.kex" % PROGRAM, "value": value, "unit": "GiB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * (tot_b / len(vals) / GIB) / value, "higher_is_better": True,
            "scaling": args.scaling or cfg["scaling"],
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s.kex on synthetic input (the reference generator's distribution); each step = "
                                   "%d MiB per process" % (PROGRAM, per >> 20 if kind == "reference" else 32)},
            "cpu_baseline": {"value": value, "unit": "GiB/s", "cores": used, "kind": kind, "variants": variants,
                             "sample": "concurrent processes of the reference C binary (--opt 3) on all %d host cores, "
                                       "%d MiB per instance and step; value = the fastest flag set: %s" % (
                                           used, per >> 20, flags_used)
                             if kind == "reference" else "oracle/kex_oracle.c port, single thread, 32 MiB per step"},
            "e2e": {"value": value, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(index):
    """Best effort: run this rank (and allocate its pinned host buffers) on the
    NUMA node its GPU hangs off, so the end-to-end leg does not cross sockets."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(index), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def shard_cuts(total_n, world, block, block_len):
    """Byte offsets that cut the stream into one shard per rank: near-equal
    parts, every inner cut moved off the record boundary (the cuts of a real
    stream fall anywhere)."""
    offs = [0]
    for r in range(1, world):
        o = r * (total_n // world) + 7919 * r + 13
        while block[(o - 1) % block_len] == 10:      # never right after a newline
            o += 1
        offs.append(o)
    offs.append(total_n)
    return offs


def fill_periodic(torch, d_block, start, n):
    """Device buffer holding bytes [start, start+n) of the stream `block block block ...`."""
    B = d_block.numel()
    buf = torch.empty(n + 64, dtype=torch.uint8, device="cuda")[:n]
    s = start % B
    head = min(n, B - s)
    buf[:head] = d_block[s:s + head]
    k = (n - head) // B
    if k:
        buf[head:head + k * B].view(k, B)[:] = d_block
    rest = n - head - k * B
    if rest:
        buf[head + k * B:] = d_block[:rest]
    return buf


def equals_periodic(torch, d_out, length, pos, d_ref):
    """d_out[:length] == bytes [pos, pos+length) of the stream `ref ref ref ...` (device compare)."""
    P = d_ref.numel()
    i = 0
    while i < length:
        s = (pos + i) % P
        m = min(P - s, length - i)
        if s == 0 and m == P:
            k = (length - i) // P                     # whole copies at once (the comparison allocates k * P bytes)
            k = min(k, max(1, (1 << 31) // P))
            if not bool((d_out[i:i + k * P].view(k, P) == d_ref[None, :]).all()):
                return False
            i += k * P
            continue
        if not bool(torch.equal(d_out[i:i + m], d_ref[s:s + m])):
            return False
        i += m
    return True


# ------------------------------------------------------------------ our arm
def main():
    global PROGRAM
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="csv2json", choices=sorted(CONFIGS))
    ap.add_argument("--gib", type=float, default=None, help="input GiB (per GPU when weak, in total when strong)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--e2e-gib", type=float, default=4.0, help="host buffer size of one kex_run_host call in the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    PROGRAM = cfg["program"]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return

    # libraries (NCCL's version banner) may write to stdout: keep fd 1 for the ONE JSON line
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    from kleenexlang_b200.kexprog import compile_kex
    from kleenexlang_b200.runtime import CompiledProgram
    from kleenexlang_b200.sharding import stitch_states, stitch_live

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    src = open(os.path.join(ROOT, "programs", PROGRAM + ".kex"), encoding="utf-8").read()
    prog = CompiledProgram(compile_kex(src), device=local)
    prog.set_timing(True)
    info = prog.info()
    scaling = args.scaling or cfg["scaling"]
    gib = args.gib if args.gib is not None else cfg["gib"]

    # ---- the stream: a seeded 64 MiB block of whole records (the same on every rank), tiled; the
    # reference binary's output of that block is what every rank's output is compared with
    block = gen_block(cfg["gen"])
    B = len(block)
    ref_out, ref_how = reference_output(block.tobytes())
    P = len(ref_out)
    ref_sha = hashlib.sha256(ref_out).hexdigest()
    ratio = P / B
    per_rank_reps = max(1, int(gib * GIB) // B) if scaling == "weak" else max(1, int(gib * GIB) // B // world)
    total_reps = per_rank_reps * world
    total_n = total_reps * B
    offs = shard_cuts(total_n, world, block, B)
    my_off, n = offs[rank], offs[rank + 1] - offs[rank]
    d_block = torch.from_numpy(block).cuda()
    d_ref = torch.frombuffer(bytearray(ref_out), dtype=torch.uint8).cuda()
    d_in = fill_periodic(torch, d_block, my_off, n)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- waves: when input + output of a rank do not fit in HBM together, the rank evaluates its
    # (resident) input in waves of whole blocks into one reused output buffer
    free_b, _ = torch.cuda.mem_get_info()
    budget = int(free_b * 0.85)
    n_waves = 1
    if world == 1:
        while (total_reps + n_waves - 1) // n_waves * P + (1 << 26) > budget:
            n_waves += 1
    wave_reps = (total_reps + n_waves - 1) // n_waves if world == 1 else 0
    waves = []
    if world == 1:
        r0 = 0
        while r0 < total_reps:
            k = min(wave_reps, total_reps - r0)
            waves.append((r0, k))
            r0 += k
    out_cap = (wave_reps * P if world == 1 else int(n * ratio)) + (1 << 22)
    d_out = torch.empty(out_cap, dtype=torch.uint8, device="cuda")

    def step_single():
        launches = 0
        for r0, k in waves:
            st, olen, _ = prog.run_device(d_in.data_ptr() + r0 * B, k * B, d_out.data_ptr(), d_out.numel(), stream)
            assert st == 0 and olen == k * P, (st, olen, k * P)
            launches += prog.launch_count()
            kk = prog.kernel_ms()
            for i in range(4):
                kms_step[i] += kk[i]
        return launches

    q1, nseam = info["nstates"] + 1, prog.seam_bytes()
    from kleenexlang_b200.frontend.driver import build_ssts
    init_state = build_ssts(src)[0].initial
    shard_out = [0, b""]
    kms_step = [0.0, 0.0, 0.0, 0.0]

    if world > 1:
        prog.set_shard_tail(True)       # G-mode tail evaluation of the shards (include/kexcuda.h kex_set_shard_tail)
    retries = [0]

    def step_sharded():
        # state maps locally, all-gather them, seams locally, all-gather the seam
        # summaries + end-of-input code, emit locally; output stays sharded in rank order
        m = prog.shard_summarize(d_in.data_ptr(), n, stream)
        t = torch.tensor(m, dtype=torch.int32, device="cuda")
        allm = torch.empty(world * len(m), dtype=torch.int32, device="cuda")
        dist.all_gather_into_tensor(allm, t)
        allm = allm.view(world, len(m)).tolist()
        starts = stitch_states(allm, init_state)
        while True:
            end, fail, seam = prog.shard_walk(starts[rank], stream)
            assert fail is None
            acc, code, tail = prog.final_action(end)
            t2 = torch.tensor(list(seam) + [code if acc else 0], dtype=torch.int32, device="cuda")
            allf = torch.empty(world * (nseam + 1), dtype=torch.int32, device="cuda")
            dist.all_gather_into_tensor(allf, t2)
            fl = allf.view(world, nseam + 1).tolist()
            lives = stitch_live(prog, [bytes(f[:nseam]) for f in fl], fl[-1][nseam])
            olen = prog.shard_emit(lives[rank], n, d_out.data_ptr(), d_out.numel(), stream)
            # tail evaluation: a rank whose shard broke the assumption behind its seam summary makes
            # every rank repeat the walk / exchange / emit (it evaluates exactly from then on)
            again = torch.tensor([1 if olen is None else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(again, op=dist.ReduceOp.MAX)
            if not int(again.item()):
                break
            retries[0] += 1
        shard_out[0], shard_out[1] = olen, (tail if rank == world - 1 else b"")
        kk = prog.kernel_ms()
        for i in range(3):
            kms_step[i] += kk[i]
        return prog.launch_count()      # cumulative since shard_summarize

    step = step_single if world == 1 else step_sharded

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms_step[:] = [0.0, 0.0, 0.0, 0.0]
    launches = 0
    ev0.record()
    for _ in range(args.steps):
        launches += step()
    ev1.record()
    torch.cuda.synchronize()
    sampler.mark(end=True)
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    kms = [x / args.steps for x in kms_step]
    clocks = sampler.stop() if rank == 0 else None
    my_out = (total_reps * P) if world == 1 else shard_out[0]
    if world > 1:
        t = torch.tensor([ms] + kms[:3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0].item())
        kms = [float(x) for x in t[1:].tolist()] + [0.0]
    ms_per_step = ms / args.steps
    value = total_n / GIB / (ms_per_step / 1000.0)

    # ---- bit-exactness at full size, outside the timed region: the stream is the block tiled and the
    # grammar is record*, so the output stream is the reference binary's output of the block, tiled;
    # every rank compares the part it wrote (its position follows from the lengths of the ranks before it)
    if world == 1:
        ok = True
        for r0, k in waves:
            st, olen, _ = prog.run_device(d_in.data_ptr() + r0 * B, k * B, d_out.data_ptr(), d_out.numel(), stream)
            ok = ok and st == 0 and olen == k * P and equals_periodic(torch, d_out, olen, 0, d_ref)
        cuts_note = "1 shard%s" % ("" if n_waves == 1 else ", %d waves of whole blocks" % n_waves)
    else:
        lens = torch.zeros(world, dtype=torch.int64, device="cuda")
        lens[rank] = shard_out[0]
        dist.all_reduce(lens)
        lens = lens.tolist()
        pos = sum(lens[:rank])
        ok = equals_periodic(torch, d_out, shard_out[0], pos, d_ref)
        if rank == world - 1:
            tail = shard_out[1]
            end = pos + shard_out[0]
            ok = ok and end + len(tail) == total_reps * P and bytes(ref_out[P - len(tail):] if tail else b"") == tail
        okt = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(okt.item())
        cuts_note = "%d shards cut at stream offsets %s (not record-aligned), per-rank output bytes %s" % (
            world, offs[1:-1], lens)
    assert ok, "output differs from the reference binary's output of the tiled block"
    verified = ("every rank's output compared on the device, after the timed region, with the output of %s for the "
                "64 MiB block (sha256 %s, %d -> %d bytes) tiled %d times; %s" % (ref_how, ref_sha, B, P, total_reps, cuts_note))

    # ---- end to end through the C ABI with host buffers (H2D + run + D2H inside)
    e2e = None
    if not args.no_e2e:
        import ctypes
        e2e_in = min(args.e2e_gib * GIB, 12 * GIB / (1.0 + ratio), n)
        wave_reps_h = max(1, int(e2e_in) // B)
        wave_n = wave_reps_h * B
        n_calls = max(1, n // wave_n)
        h_in = torch.from_numpy(np.tile(block, wave_reps_h)).pin_memory()
        wave_out = wave_reps_h * P
        h_out = torch.empty(wave_out + 4096, dtype=torch.uint8).pin_memory()
        L = prog._L
        ol, stt, fc = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_size_t()

        def e2e_step():
            for _ in range(n_calls):
                rc = L.kex_run_host(prog._h, ctypes.c_char_p(h_in.data_ptr()), wave_n, h_out.data_ptr(), h_out.numel(),
                                    ctypes.byref(ol), ctypes.byref(stt), ctypes.byref(fc))
                assert rc == 0 and stt.value == 0 and ol.value == wave_out, (rc, stt.value, ol.value)

        e2e_steps = min(args.steps, 2)
        prog.set_timing(False)
        L.kex_run_host(prog._h, ctypes.c_char_p(h_in.data_ptr()), wave_n, h_out.data_ptr(), h_out.numel(),
                       ctypes.byref(ol), ctypes.byref(stt), ctypes.byref(fc))      # warm the staging buffers
        e2e_ok = hashlib.sha256(bytes(h_out[:P].numpy())).hexdigest() == ref_sha
        assert e2e_ok, "host-path output differs from the reference binary's"
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * n_calls * wave_n * e2e_steps / GIB / dt, "unit": "GiB/s",
               "h2d_bytes_per_step": n_calls * wave_n, "d2h_bytes_per_step": n_calls * wave_out,
               "note": "kex_run_host over pinned host buffers, %d calls of %.2f GiB per step and rank, %d steps; inside "
                       "a call 64 MiB sub-waves are copied in, evaluated and copied out on three streams; first "
                       "block of the host output sha256-equal to the reference binary's" % (n_calls, wave_n / GIB, e2e_steps)}

    if rank == 0:
        peak, how = peaks()
        kern = {4: "k4_emit", 3: "k3_emit", 2: "k_emit_fast"}.get(prog.info().get("emit_kernel", 0), "k_emit")
        line = {"metric": "input GiB/s on %s.kex" % PROGRAM, "value": value, "unit": "GiB/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": (cfg["what"] % ((total_n / world if scaling == "weak" else total_n) / GIB)) +
                                       ", 64 MiB seeded block tiled in HBM" +
                                       (", strong scaling: %.2f GiB in total" % (total_n / GIB) if scaling == "strong" else ""),
                           "input_bytes_total": total_n, "input_bytes_rank0": n, "output_bytes_total": total_reps * P,
                           "program": "programs/%s.kex --opt 3 --la=false --act=false" % PROGRAM,
                           "sst": {"states": info["nstates"], "classes": info["nclasses"], "registers": info["nregs"]},
                           "l2": "inputs (%.1f GiB per rank) far exceed the 126 MB L2; no explicit flush" % (n / GIB),
                           "verified": verified,
                           "parallelism": ("1 shard per GPU cut anywhere in the stream, 2 all-gathers of seam summaries + a "
                                           "1-word all-reduce (tail evaluation retry flag; retries on rank 0: %d); rank 0 bound "
                                           "to NUMA node %s" % (retries[0], numa)) if world > 1 else "1 GPU"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches}
        emit_ms = kms[2]
        if emit_ms > 0:
            algo = float(n + my_out)                   # rank 0's shard: every input byte read once, every output byte written once
            ach = algo / (emit_ms / 1000.0) / 1e9
            tpi = cfg["emit_traffic_per_in"]
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": (tpi * n / max(1, n_waves)) if tpi else None,
                                "kernel": kern, "peak_source": "%s (MEASURED_PEAKS.json hbm_gbs)" % how,
                                "algorithmic_bytes_per_launch": algo / max(1, n_waves),
                                "launches_per_step": max(1, n_waves),
                                "kernel_ms": {"forward_monoid": kms[0], "seams": kms[1], "emit": emit_ms,
                                              "all_device_work": kms[3] if world == 1 else None},
                                "note": "per GPU; kernel times are the max over ranks" if world > 1 else "1 GPU",
                                "pipeline_frac": (total_n + total_reps * P) / world / (ms_per_step / 1000.0) / 1e9 / peak}
        if not args.no_cpu_baseline and world == 1:      # a reported baseline, at N=1 only
            line["cpu_baseline"] = cpu_baseline(cfg["gen"])
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
