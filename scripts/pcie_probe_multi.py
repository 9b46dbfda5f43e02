"""Concurrent host<->device copy probe, one process per GPU (run under torchrun like bench.py):
every rank copies 4 GiB H2D and 7.8 GiB D2H (the byte ratio of the csv2json end-to-end leg) from / to
pinned host memory, alone and with all other ranks at the same time.  The aggregate of the concurrent
both-directions case is the ceiling of bench.py's `e2e` on this host."""
import os, sys, time
import torch, torch.distributed as dist
GIB = 1 << 30
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_in, n_out = 4 * GIB, int(4 * GIB * 1.954)
h_a = torch.empty(n_in, dtype=torch.uint8).pin_memory(); h_a.fill_(1)
h_b = torch.empty(n_out, dtype=torch.uint8).pin_memory(); h_b.fill_(2)
d_a = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(f, concurrent):
    if world > 1 and concurrent:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if world > 1 and concurrent:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
    return dt


def both():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)


both(); torch.cuda.synchronize()
res = {}
for name, f, nb in (("h2d", lambda: d_a.copy_(h_a, non_blocking=True), n_in), ("d2h", lambda: h_b.copy_(d_b, non_blocking=True), n_out),
                    ("both", both, n_in + n_out)):
    res[name] = min(timed(f, True) for _ in range(3)), nb
if rank == 0:
    print("pcie probe, %d GPU(s) concurrently, per rank: %.0f GiB H2D + %.1f GiB D2H, pinned host memory" % (world, n_in / GIB, n_out / GIB))
    for name, (dt, nb) in res.items():
        print("  %-5s %7.1f ms  aggregate %6.1f GB/s (%.1f GB/s per GPU)" % (name, dt * 1e3, world * nb / dt / 1e9, nb / dt / 1e9))
    dt, nb = res["both"]
    print("  ceiling of the end-to-end leg (input bytes over the time both directions need): %.1f GiB/s" % (world * n_in / dt / GIB))
if world > 1:
    dist.destroy_process_group()
