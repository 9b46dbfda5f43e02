#!/usr/bin/env python
"""bench.py -- input GiB/s of the compiled-transducer hot path on csv2json.kex.

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K ...   (reference C binary on host cores)

A step is one pass of the hot path over the whole synthetic CSV input that is
already resident in HBM (BASELINE.json config "csv2json.kex on 16 GiB synthetic
CSV, 1 GPU"; per GPU at N>1, i.e. weak scaling).  One JSON line on stdout.
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
PROGRAM = "csv2json"
OUT_PER_IN = 1.954            # +127 B per ~133 B row
ALGO_BYTES_PER_IN = 1.0 + OUT_PER_IN
# dram__bytes_read.sum + dram__bytes_write.sum of one k3_emit launch from the
# committed `ncu --set full` capture (profiles/r01_ncu_v3_full_16gib.txt: 19.557 GB
# read + 33.385 GB written for 17.180 GB of input), per input byte of that launch.
EMIT_TRAFFIC_PER_IN = (19.557043 + 33.384793) / 17.179869


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
def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", PROGRAM)
    return p if os.path.exists(p) else None


def make_reference_inputs(sample_bytes, instances):
    """One record-aligned CSV file per instance in /dev/shm (written once, reused by every step)."""
    from kleenexlang_b200 import workloads
    block = workloads.gen_csv(min(sample_bytes, 64 << 20), seed=1234)
    reps = max(1, sample_bytes // len(block))
    files = []
    for i in range(instances):
        f = "/dev/shm/kexbench_%d_%d.csv" % (os.getpid(), i)
        with open(f, "wb") as fh:
            for _ in range(reps):
                fh.write(block.tobytes())
        files.append(f)
    return files, reps * len(block) * instances


def run_reference_once(files):
    """Times one concurrent pass of the reference's compiled C binary
    (oracle/_ref, emitted C + verbatim crt.c, `cc -O3 -D FLAG_WORDALIGNED`) over
    the files, stdout to /dev/null as bench/runningtime.sh:101 does."""
    binp = ref_binary()
    t0 = time.perf_counter()
    procs = [subprocess.Popen([binp], stdin=open(f, "rb"), stdout=subprocess.DEVNULL) for f in files]
    rcs = [p.wait() for p in procs]
    dt = time.perf_counter() - t0
    assert all(rc == 0 for rc in rcs), "reference binary rejected the synthetic CSV"
    return dt


def time_reference(sample_bytes, instances):
    files, total = make_reference_inputs(sample_bytes, instances)
    try:
        return run_reference_once(files), total
    finally:
        for f in files:
            os.unlink(f)


def time_oracle_port(sample_bytes):
    from kleenexlang_b200 import workloads
    from kleenexlang_b200.frontend.driver import build_ssts
    from oracle.sstbin import oracle_run, serialize_sst
    src = open(os.path.join(ROOT, "programs", PROGRAM + ".kex"), encoding="utf-8").read()
    blobs = [serialize_sst(s) for s in build_ssts(src)]
    data = workloads.gen_csv(sample_bytes, seed=1234).tobytes()
    t0 = time.perf_counter()
    st, _, _ = oracle_run(blobs, data)
    dt = time.perf_counter() - t0
    assert st == 0
    return dt, len(data)


def cpu_baseline(sample_bytes=1 << 30):
    if ref_binary():
        dt, nb = time_reference(sample_bytes, 1)
        return {"value": nb / GIB / dt, "unit": "GiB/s", "cores": 1, "kind": "reference",
                "sample": "%d MiB synthetic CSV, 1 process of oracle/_ref/csv2json (emitted C + verbatim crt.c, "
                          "cc -O3 -D FLAG_WORDALIGNED, --opt 3 --la=false --act=false), stdin from /dev/shm, "
                          "stdout /dev/null" % (nb >> 20)}
    dt, nb = time_oracle_port(64 << 20)
    return {"value": nb / GIB / dt, "unit": "GiB/s", "cores": 1, "kind": "port",
            "sample": "%d MiB synthetic CSV through oracle/kex_oracle.c (interpreting C restatement)" % (nb >> 20)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = "reference" if ref_binary() else "port"
    # bounded sample: at most ~8 GiB of /dev/shm over all processes, 64-256 MiB each
    per = max(64 << 20, min(256 << 20, (8 << 30) // cores))
    vals = []
    files, total = make_reference_inputs(per, cores) if kind == "reference" else ([], 0)
    try:
        for i in range(args.warmup + args.steps):
            if kind == "reference":
                dt, nb = run_reference_once(files), total
            else:
                dt, nb = time_oracle_port(32 << 20)
            if i >= args.warmup:
                vals.append((dt, nb))
    finally:
        for f in files:
            os.unlink(f)
    tot_t = sum(v[0] for v in vals)
    tot_b = sum(v[1] for v in vals)
    value = tot_b / GIB / tot_t
    used = cores if kind == "reference" else 1
    line = {"impl": "reference", "metric": "input GiB/s on csv2json.kex", "value": value, "unit": "GiB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * tot_t / len(vals), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "csv2json.kex on synthetic CSV (gen_csv.pl distribution); each step = "
                                   "%d MiB per process" % (per >> 20 if kind == "reference" else 32)},
            "cpu_baseline": {"value": value, "unit": "GiB/s", "cores": used, "kind": kind,
                             "sample": "%d concurrent processes of the reference C binary, one per host core, "
                                       "%d MiB each per step" % (used, per >> 20) if kind == "reference" else
                                       "oracle/kex_oracle.c port, single thread, 32 MiB per step"},
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


# ------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--gib", type=float, default=16.0, help="input GiB per GPU")
    ap.add_argument("--e2e-gib", type=float, default=4.0, help="host buffer size of one kex_run_host call in the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    # libraries (NCCL's version banner) may write to stdout: keep fd 1 for the ONE JSON line
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    from kleenexlang_b200 import workloads
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

    # ---- synthetic input: a 64 MiB block of whole rows per rank, tiled in HBM
    block = workloads.gen_csv(64 << 20, seed=100 + rank)
    rows_per_block = int((block == 10).sum())
    reps = max(1, int(args.gib * GIB) // len(block))
    d_block = torch.from_numpy(block).cuda()
    d_in = d_block.repeat(reps)
    n = d_in.numel()
    expect_out = n + 127 * rows_per_block * reps
    d_out = torch.empty(expect_out + (1 << 20), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step_single():
        st, olen, _ = prog.run_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(), stream)
        assert st == 0 and olen == expect_out, (st, olen, expect_out)
        return prog.launch_count()

    q1, nseam = info["nstates"] + 1, prog.seam_bytes()
    init_state = 0
    shard_out = [0]

    def step_sharded():
        # state maps locally, all-gather them, seams locally, all-gather the seam
        # summaries + end-of-input code, emit locally; output stays sharded in rank order
        m = prog.shard_summarize(d_in.data_ptr(), n, stream)
        t = torch.tensor(m, dtype=torch.int32, device="cuda")
        allm = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allm, t)
        starts = stitch_states([x.tolist() for x in allm], init_state)
        end, fail, seam = prog.shard_walk(starts[rank], stream)
        assert fail is None
        acc, code, tail = prog.final_action(end)
        t2 = torch.tensor(list(seam) + [code if acc else 0], dtype=torch.int32, device="cuda")
        allf = [torch.empty_like(t2) for _ in range(world)]
        dist.all_gather(allf, t2)
        fl = [x.tolist() for x in allf]
        lives = stitch_live(prog, [bytes(f[:nseam]) for f in fl], fl[-1][nseam])
        olen = prog.shard_emit(lives[rank], n, d_out.data_ptr(), d_out.numel(), stream)
        # a shard that is not the last one closes its final record with the text that opens the next
        # one, so single shards differ from their stand-alone length; the sum over ranks is checked below
        shard_out[0] = olen
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
    kms = [0.0, 0.0, 0.0, 0.0]
    launches = 0
    ev0.record()
    for _ in range(args.steps):
        launches += step()
        if world == 1:
            k = prog.kernel_ms()
            kms = [a + b for a, b in zip(kms, k)]
    ev1.record()
    torch.cuda.synchronize()
    sampler.mark(end=True)
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.tensor([shard_out[0], expect_out], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        assert int(tot[0]) == int(tot[1]), ("output bytes over all shards", tot.tolist())
    ms_per_step = ms / args.steps
    value = world * n / GIB / (ms_per_step / 1000.0)

    # ---- bit-exactness at full size (outside the timed region): the input is a block of whole
    # records tiled `reps` times and the grammar is record*, so the output must be `reps` copies of
    # the block's output (which tests/ compare with the oracle byte for byte at 4 MiB)
    verified = None
    if world == 1:
        per = expect_out // reps
        o = d_out[:per * reps].view(reps, per)
        verified = bool(per * reps == expect_out and all(
            bool((o[i:i + 32] == o[0]).all()) for i in range(0, reps, 32)))
        assert verified, "output of the tiled input is not the tiled output"

    # ---- end to end through the C ABI with host buffers (H2D + run + D2H inside)
    e2e = None
    wave_reps = max(1, int(args.e2e_gib * GIB) // len(block))
    wave_n = wave_reps * len(block)
    waves = max(1, n // wave_n)
    h_in = torch.from_numpy(np.tile(block, wave_reps)).pin_memory()
    wave_out = wave_n + 127 * rows_per_block * wave_reps
    h_out = torch.empty(wave_out + 4096, dtype=torch.uint8).pin_memory()
    import ctypes
    L = prog._L
    ol, stt, fc = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_size_t()

    def e2e_step():
        for _ in range(waves):
            rc = L.kex_run_host(prog._h, ctypes.c_char_p(h_in.data_ptr()), wave_n, h_out.data_ptr(), h_out.numel(),
                                ctypes.byref(ol), ctypes.byref(stt), ctypes.byref(fc))
            assert rc == 0 and stt.value == 0 and ol.value == wave_out, (rc, stt.value, ol.value)

    e2e_steps = min(args.steps, 2)
    prog.set_timing(False)
    L.kex_run_host(prog._h, ctypes.c_char_p(h_in.data_ptr()), wave_n, h_out.data_ptr(), h_out.numel(),
                   ctypes.byref(ol), ctypes.byref(stt), ctypes.byref(fc))      # warm the staging buffers
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
    e2e = {"value": world * waves * wave_n * e2e_steps / GIB / dt, "unit": "GiB/s",
           "h2d_bytes_per_step": waves * wave_n, "d2h_bytes_per_step": waves * wave_out,
           "note": "kex_run_host over pinned host buffers, %d calls of %.2f GiB per step, %d steps; inside a call "
                   "64 MiB sub-waves are copied in, evaluated and copied out on three streams" % (
               waves, wave_n / GIB, e2e_steps)}

    if rank == 0:
        peak, how = peaks()
        line = {"metric": "input GiB/s on csv2json.kex", "value": value, "unit": "GiB/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "csv2json.kex on %.2f GiB synthetic CSV per GPU (gen_csv.pl distribution, "
                                       "64 MiB seeded block tiled in HBM)" % (n / GIB),
                           "input_bytes_per_gpu": n, "output_bytes_per_gpu": expect_out,
                           "program": "programs/csv2json.kex --opt 3 --la=false --act=false",
                           "sst": {"states": info["nstates"], "classes": info["nclasses"], "registers": info["nregs"]},
                           "l2": "inputs (%.1f GiB) far exceed the 126 MB L2; no explicit flush" % (n / GIB),
                           "verified": "output == %d x the 64 MiB block's output, compared on the device after the timed "
                                       "region" % reps if verified else None,
                           "parallelism": ("1 shard per GPU, 2 all-gathers of seam summaries; rank 0 bound to NUMA node %s"
                                           % numa) if world > 1 else "1 GPU"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches}
        if world == 1:
            emit_ms = kms[2] / args.steps
            ach = ALGO_BYTES_PER_IN_measured(n, expect_out) / (emit_ms / 1000.0) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": (EMIT_TRAFFIC_PER_IN * n) if EMIT_TRAFFIC_PER_IN else None,
                                "kernel": "k3_emit" if info["chunk_bytes"] == 1024 else "k_emit_fast",
                                "peak_source": "%s (MEASURED_PEAKS.json hbm_gbs)" % how,
                                "algorithmic_bytes_per_launch": n + expect_out,
                                "kernel_ms": {"forward_monoid": kms[0] / args.steps, "seams": kms[1] / args.steps,
                                              "emit": emit_ms, "all_device_work": kms[3] / args.steps},
                                "pipeline_frac": (n + expect_out) / (ms_per_step / 1000.0) / 1e9 / peak}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def ALGO_BYTES_PER_IN_measured(n, out):
    # algorithmic bytes of one launch: every input byte read once, every output byte written once
    return float(n + out)


if __name__ == "__main__":
    main()
