"""bench.py's host-side pieces that do not need a GPU: the seeded block, the
shard cuts, the periodic fill / compare used for the in-run verification, and
the reference arm's JSON line."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench


@pytest.mark.parametrize("cfg", sorted(bench.CONFIGS))
def test_block_is_whole_records_and_16_aligned(cfg):
    gen = bench.CONFIGS[cfg]["gen"]
    b = bench.gen_block(gen, 1 << 20, seed=3)
    assert len(b) % 16 == 0 and b[-1] == 10 and len(b) > (1 << 20) - 65536
    if gen == "fastq2fasta":
        assert int((b == 10).sum()) % 4 == 0 and b[0] == ord("@")


def test_shard_cuts_are_not_record_aligned():
    block = bench.gen_block("csv2json", 1 << 20, seed=3)
    B = len(block)
    for world in (1, 2, 4, 8):
        total = 64 * B
        offs = bench.shard_cuts(total, world, block, B)
        assert offs[0] == 0 and offs[-1] == total and offs == sorted(offs) and len(offs) == world + 1
        for o in offs[1:-1]:
            assert block[(o - 1) % B] != 10 and o % B != 0


def test_fill_and_compare_periodic_on_cpu(monkeypatch):
    """fill_periodic / equals_periodic with CPU tensors (the bench runs them on the device)."""
    block = torch.arange(0, 1000, dtype=torch.int64).remainder(251).to(torch.uint8)
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{kk: v for kk, v in k.items() if kk != "device"}))
    for start, n in ((0, 3000), (123, 1), (999, 2500), (1234, 877)):
        buf = bench.fill_periodic(torch, block, start, n)
        want = torch.tensor([(int(block[(start + i) % 1000])) for i in range(n)], dtype=torch.uint8)
        assert torch.equal(buf, want)
        assert bench.equals_periodic(torch, buf, n, start, block)
        assert bench.equals_periodic(torch, buf, n, start + 1000 * 70, block)
        if n > 1:
            bad = buf.clone()
            bad[n // 2] ^= 1
            assert not bench.equals_periodic(torch, bad, n, start, block)
            assert not bench.equals_periodic(torch, buf, n, start + 1, block)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "csv2json")), reason="oracle/_ref not built")
def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GiB/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "reference"
    assert line["value"] == max(line["cpu_baseline"]["variants"].values())
    assert len(line["cpu_baseline"]["variants"]) >= 2
