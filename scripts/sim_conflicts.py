"""Shared-memory wavefront model of the G-mode write pass (kex_v4.cuh v4_write_half):
per tile, 32 lanes x 2 halves x 16 steps of byte stores at window offsets given by
the G-mode emission lengths; counts wavefronts (max distinct words per bank) for
candidate window address maps.  CPU only; guides the layout choice."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from kleenexlang_b200 import fasttab, workloads
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200.kexprog import build_phase
name = sys.argv[1] if len(sys.argv) > 1 else "csv2json"
src = open("/root/repo/programs/%s.kex" % name).read()
t = build_phase(build_ssts(src)[0]); f = fasttab.build_fast(t)
G, gtab, _ = fasttab.build_gmode(t, f)
data = workloads.GENERATORS[name](1 << 19, seed=5)
n = len(data) // 1024 * 1024
cls = np.array(t.cls, dtype=np.int64)[data[:n]]
gt = np.array(gtab, dtype=np.int64)
q = t.init; lens = np.zeros(n, np.int64); kind = np.zeros(n, np.int64)
for i in range(n):
    e = gt[q * t.C + cls[i]]; q = e & 0xFFFF
    lens[i] = (e >> 16) & 0xFF; kind[i] = e >> 24
ntiles = n // 1024
L = lens.reshape(ntiles, 1024)
K = kind.reshape(ntiles, 1024)
off = np.cumsum(L, axis=1) - L            # tile-relative output offset of each byte's emission
copy = (K & 0x40) != 0
print(name, "tiles", ntiles, "out/in", lens.sum() / n, "templates/tile", ((K & 0x80) != 0).sum() / ntiles)

def wavefronts(addr, active):
    """addr, active: [ntiles, 32 lanes] for one store instruction -> total wavefronts"""
    word = addr >> 2; bank = word & 31
    tot = 0
    for b in range(32):
        m = active & (bank == b)
        # distinct words per bank
        w = np.where(m, word, -1)
        w.sort(axis=1)
        distinct = ((w[:, 1:] != w[:, :-1]) & (w[:, 1:] >= 0)).sum(axis=1) + (w[:, 0] >= 0)
        tot = np.maximum(tot, distinct)
    return tot.sum()

maps = {
    "identity": lambda a: a,
    "xor16(row)": lambda a: a ^ ((a >> 3) & 0x70),
    "xor4(row)": lambda a: a ^ ((a >> 5) & 0x1C),
    "xor4(row*?)": lambda a: a ^ ((a >> 5) & 0x7C),
    "pad4/128": lambda a: a + ((a >> 7) << 2),
    "pad16/128": lambda a: a + ((a >> 7) << 4),
}
for nm, fn in maps.items():
    tot = 0
    for half in range(2):
        for j in range(16):
            idx = np.arange(32) * 32 + half * 16 + j
            tot += wavefronts(fn(off[:, idx]), copy[:, idx])
    print("%-14s STS.U8 wavefronts/tile %.1f (ideal 32)" % (nm, tot / ntiles))

def chunkmap(fn):
    def m(a):
        row = a >> 7; ch = (a >> 4) & 7
        return (a & ~0x70) | ((fn(row, ch) & 7) << 4)
    return m
more = {
    "ch^row^row>>3": chunkmap(lambda r, c: c ^ r ^ (r >> 3)),
    "ch^row^row>>2": chunkmap(lambda r, c: c ^ r ^ (r >> 2)),
    "ch+row": chunkmap(lambda r, c: c + r),
    "ch+3row": chunkmap(lambda r, c: c + 3 * r),
    "ch^(row>>1)": chunkmap(lambda r, c: c ^ (r >> 1)),
    "ch^row^(row>>1)": chunkmap(lambda r, c: c ^ r ^ (r >> 1)),
    "word^row (4B)": lambda a: a ^ ((a >> 5) & 0x7C),
    "word^row^row>>5": lambda a: a ^ (((a >> 5) ^ (a >> 10)) & 0x7C),
}
for nm, fn in more.items():
    tot = 0
    for half in range(2):
        for j in range(16):
            idx = np.arange(32) * 32 + half * 16 + j
            tot += wavefronts(fn(off[:, idx]), copy[:, idx])
    print("%-16s STS.U8 wavefronts/tile %.1f (ideal 32)" % (nm, tot / ntiles))
# random-bank reference: expected max load of `k` balls in 32 bins
rng = np.random.default_rng(0)
act = copy[:, np.arange(32) * 32]
kk = act.sum(axis=1)
ref = np.mean([np.bincount(rng.integers(0, 32, k), minlength=32).max() for k in kk for _ in range(4)])
print("random banks with the same active-lane counts: %.2f wavefronts per store -> %.1f per tile" % (ref, ref * 32))
