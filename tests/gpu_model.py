"""Executable model of the algorithm the CUDA back end runs, over the same
`PhaseTables` the blob is serialised from.  Test infrastructure only: it lets
the CPU suite check the table construction (fates, piece order, liveness) and
the chunked scan formulation against the sequential SST semantics without a
GPU.  Pure Python, small inputs.

Passes (same names as kleenexlang_b200/csrc/kexcuda.cu):
  chunk_maps   per-chunk state map  Q -> Q
  scan_states  compose maps -> true start state per chunk
  true_walk    per chunk: fate map of the registers, bytes resolved to the
               output, bytes left pending per register, first failure
  scan_fates   backward: which registers at a chunk's end reach the output
  emit         per chunk: backward survival walk, then write surviving bytes
"""
from kleenexlang_b200.kexprog import DEAD, PIECE_SYM


def run_model(t, data: bytes, chunk: int = 7):
    Q, C, R = t.Q, t.C, t.R
    FAIL = Q
    n = len(data)
    nchunks = max(1, (n + chunk - 1) // chunk)
    bounds = [(k * chunk, min(n, (k + 1) * chunk)) for k in range(nchunks)]

    def step(q, b):
        e = t.trans[q * C + t.cls[b]]
        return e & 0xFFFF, e >> 16

    # chunk_maps + scan_states
    maps = []
    for lo, hi in bounds:
        m = []
        for q in range(Q + 1):
            s = q
            for b in data[lo:hi]:
                s = step(s, b)[0]
            m.append(s)
        maps.append(m)
    start = []
    s = t.init
    for m in maps:
        start.append(s)
        s = m[s]
    end_state = s

    # true_walk: find first failure
    fail_pos = None
    for k, (lo, hi) in enumerate(bounds):
        s = start[k]
        for i in range(lo, hi):
            s2, _ = step(s, data[i])
            if s2 == FAIL:
                fail_pos = i
                break
            s = s2
        if fail_pos is not None:
            break
    accepted = fail_pos is None and t.final[end_state] >= 0
    n_eff = n if fail_pos is None else fail_pos
    # on reject, everything up to the failing transition is replayed with an
    # empty live set: only bytes already flushed to the stream count
    nchunks = max(1, (n_eff + chunk - 1) // chunk)
    bounds = [(k * chunk, min(n_eff, (k + 1) * chunk)) for k in range(nchunks)]

    def actions_of(k):
        lo, hi = bounds[k]
        s = start[k]
        acts = []
        for i in range(lo, hi):
            s, a = step(s, data[i])
            acts.append(a)
        return acts, s

    F = []
    for k in range(nchunks):
        acts, _ = actions_of(k)
        f = list(range(R))
        for a in acts:
            f = [x if x in (0, DEAD) else t.fate[a][x] for x in f]
        F.append(f)

    # scan_fates (backward)
    if accepted:
        fa = t.final[end_state]
        live = {r for r in range(1, R) if t.fate[fa][r] == 0}
    else:
        live = set()
    live_end = [None] * nchunks
    for k in range(nchunks - 1, -1, -1):
        live_end[k] = set(live)
        live = {r for r in range(1, R) if F[k][r] == 0 or F[k][r] in live}

    # emit
    out = bytearray()
    for k in range(nchunks):
        acts, _ = actions_of(k)
        lo, _ = bounds[k]
        L = set(live_end[k])
        masks = [None] * len(acts)
        for j in range(len(acts) - 1, -1, -1):
            a = acts[j]
            masks[j] = L | {0}
            L = {r for r in range(1, R) if t.fate[a][r] != DEAD and t.fate[a][r] in masks[j]}
        for j, a in enumerate(acts):
            for tgt, kind, ln, off in t.pieces[a]:
                if tgt in masks[j]:
                    if kind == PIECE_SYM:
                        out.append(data[lo + j])
                    else:
                        out += t.consts[off:off + ln]
    if accepted:
        for tgt, kind, ln, off in t.pieces[t.final[end_state]]:
            out += t.consts[off:off + ln]
        return True, bytes(out), n
    # reject: the C runtime only ever wrote whole 16 KiB flushes
    keep = len(out) // 16384 * 16384
    return False, bytes(out[:keep]), n_eff
