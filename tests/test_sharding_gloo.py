"""The N>1 protocol on CPU: two processes over gloo each own one
byte range of the input, exchange the two seam summaries with all_gather
exactly as `bench.py: step_sharded` does, and evaluate their shard with the
table model of the monoid kernels.  The concatenated shard outputs must equal
the sequential oracle's output.  (The CUDA entry points themselves are covered
by tests/test_gpu_parity.py::test_sharded_entry_points on the GPU box.)"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import program_source, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_model(t, f, data, start_state=None, lam_end=None):
    """Whole-shard statements of kex_shard_summarize / _walk / _emit."""
    C, A = t.C, t.A
    if start_state is None:                      # summarize: state map of the shard
        m = 0
        for b in data:
            m = f.mulF[m * C + t.cls[b]]
        return list(f.elemsF[m])
    s, mb, acts = start_state, 0, []
    for b in data:
        e = f.trans2[s * C + t.cls[b]]
        s = e & 0xFFFF
        assert s != t.Q
        acts.append((e >> 16) & 0xFF)
        mb = f.mulB[mb * f.NG + (e >> 24)]
    if lam_end is None:                          # walk: end state + seam summary
        return s, list(f.belems[mb])
    out = []
    L = lam_end
    for a, b in zip(reversed(acts), reversed(data)):
        e = f.BE[L * A + a]
        L = (e & 0xFFFC) // (4 * A)
        if e & 1:
            out.append(bytes([b]))
        elif e & 2:
            info, hm = f.tplinfo[2 * (e >> 24)], f.tplinfo[2 * (e >> 24) + 1]
            seg = bytearray(f.pool[info & 0xFFFF:(info & 0xFFFF) + (info >> 16)])
            for h in range(min(len(seg), 32)):
                if (hm >> h) & 1:
                    seg[h] = b
            out.append(bytes(seg))
    return b"".join(reversed(out))


def _worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kleenexlang_b200 import workloads, fasttab
    from kleenexlang_b200.frontend.driver import build_ssts
    from kleenexlang_b200.kexprog import build_phase
    from kleenexlang_b200.sharding import stitch_states, stitch_codes
    sst = build_ssts(program_source(name))[0]
    t = build_phase(sst)
    f = fasttab.build_fast(t)
    data = workloads.GENERATORS[name](6000, seed=11).tobytes()
    cut = [0, 2777, len(data)]                    # an arbitrary byte position, not a record boundary
    mine = data[cut[rank]:cut[rank + 1]]
    # 1. state maps
    m = torch.tensor(_shard_model(t, f, mine), dtype=torch.int32)
    allm = [torch.empty_like(m) for _ in range(world)]
    dist.all_gather(allm, m)
    starts = stitch_states([x.tolist() for x in allm], t.init)
    # 2. seam summaries + end-of-input code
    end, seam = _shard_model(t, f, mine, start_state=starts[rank])
    code = f.lam_final[end] if rank == world - 1 else 0
    s = torch.tensor(seam + [code], dtype=torch.int32)
    alls = [torch.empty_like(s) for _ in range(world)]
    dist.all_gather(alls, s)
    rows = [x.tolist() for x in alls]
    assert rows[-1][-1] != 0xFF, "input must be accepted"
    codes = stitch_codes([r[:-1] for r in rows], rows[-1][-1])
    # 3. local emit; output stays sharded in rank order
    out = _shard_model(t, f, mine, start_state=starts[rank], lam_end=codes[rank])
    if rank == world - 1:
        for tgt, kind, ln, off in t.pieces[t.final[end]]:
            out += t.consts[off:off + ln]
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["csv2json", "thousand_sep", "iso_datetime_to_json"])
def test_two_rank_sharding_gloo(name):
    from kleenexlang_b200 import workloads
    from kleenexlang_b200.frontend.driver import build_ssts
    from oracle.sstbin import oracle_run
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    parts = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = workloads.GENERATORS[name](6000, seed=11).tobytes()
    st, exp, _ = oracle_run(build_ssts(program_source(name)), data)
    assert st == 0 and parts[0] + parts[1] == exp
