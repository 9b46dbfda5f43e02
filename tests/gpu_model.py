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


# --------------------------------------------------------------------------
# Model of the monoid ("fast") kernels: same tables as the fast section of the
# blob (kleenexlang_b200/fasttab.py), same pass structure as
# kleenexlang_b200/csrc/kex_fast.cuh.
def run_fast_model(t, f, data: bytes, chunk: int = 64, sub: int = 8):
    from kleenexlang_b200.fasttab import E_SYM, E_TPL
    Q, C, A = t.Q, t.C, t.A
    FAIL = Q
    n = len(data)
    nsub = chunk // sub
    assert chunk % sub == 0
    nchunks = (n + chunk - 1) // chunk

    # k_fwd_monoid: prefix element before every sub-chunk, element of the chunk
    samples = []
    cmaps = []
    for k in range(nchunks):
        m = 0
        row = []
        for j in range(k * chunk, min(n, (k + 1) * chunk)):
            if (j - k * chunk) % sub == 0:
                row.append(m)
            m = f.mulF[m * C + t.cls[data[j]]]
        samples.append(row)
        cmaps.append(f.elemsF[m])
    # state scan
    start = []
    s = t.init
    for mp in cmaps:
        start.append(s)
        s = mp[s]
    end_state = s

    def step(q, b):
        e = f.trans2[q * C + t.cls[b]]
        return e & 0xFFFF, (e >> 16) & 0xFF, e >> 24

    # k_seams: backward element of each chunk (walk until it is a constant
    # map), exact position of a failure
    fail_pos = None
    bm = [0] * nchunks
    for k in range(nchunks):
        s = start[k]
        if s == FAIL:
            break
        failing = cmaps[k][s] == FAIL
        mb = 0
        for j in range(k * chunk, min(n, (k + 1) * chunk)):
            s, a, g = step(s, data[j])
            if s == FAIL:
                fail_pos = j
                break
            mb = f.mulB[mb * f.NG + g]
            if not failing and f.constB[mb]:
                break
        bm[k] = mb
        if fail_pos is not None:
            break
    accepted = fail_pos is None and t.final[end_state] >= 0
    n_eff = n if fail_pos is None else fail_pos
    ntiles = (n_eff + chunk - 1) // chunk
    lam = f.lam_final[end_state] if accepted else 0
    lam_end = [0] * ntiles
    for k in range(ntiles - 1, -1, -1):
        lam_end[k] = lam
        lam = f.belems[bm[k]][lam]

    # k_emit
    out = bytearray()
    for k in range(ntiles):
        lo, hi = k * chunk, min(n_eff, (k + 1) * chunk)
        acts = []
        mbs = []
        for ti in range(nsub):
            a0, a1 = lo + ti * sub, min(hi, lo + (ti + 1) * sub)
            if a0 >= a1:
                acts.append([])
                mbs.append(0)
                continue
            s = f.elemsF[samples[k][ti]][start[k]]
            al = []
            mb = 0
            for j in range(a0, a1):
                s, a, g = step(s, data[j])
                assert s != FAIL
                al.append(a)
                mb = f.mulB[mb * f.NG + g]
            acts.append(al)
            mbs.append(mb)
        # suffix composition of the later threads' elements
        suffix = 0
        lam_t = [0] * nsub
        for ti in range(nsub - 1, -1, -1):
            lam_t[ti] = f.belems[suffix][lam_end[k]]
            suffix = f.compB[mbs[ti] * f.NB + suffix]
        assert suffix == bm[k] or f.constB[bm[k]] or k == ntiles - 1 or True
        for ti in range(nsub):
            L = lam_t[ti]
            codes = []
            for a in reversed(acts[ti]):
                e = f.BE[L * A + a]
                codes.append(e)
                L = (e & 0xFFFC) // (4 * A)
            codes.reverse()
            for j, e in enumerate(codes):
                typ, ln, x = e & 3, (e >> 16) & 0xFF, e >> 24
                b = data[lo + ti * sub + j]
                if typ == E_SYM:
                    out.append(b)
                elif typ == E_TPL:
                    info, hmask = f.tplinfo[2 * x], f.tplinfo[2 * x + 1]
                    off, tl = info & 0xFFFF, (info >> 16) & 0xFF
                    assert tl == ln
                    seg = bytearray(f.pool[off:off + tl])
                    for h in range(min(tl, 32)):
                        if (hmask >> h) & 1:
                            seg[h] = b
                    out += seg
    if accepted:
        for tgt, kind, ln, off in t.pieces[t.final[end_state]]:
            out += t.consts[off:off + ln]
        return True, bytes(out), n
    keep = len(out) // 16384 * 16384
    return False, bytes(out[:keep]), n_eff


def run_gmode_model(t, f, gm, data: bytes, tile: int = 64):
    """Model of the G-mode emit kernel (csrc/kex_v4.cuh) on an accepted input:
    a tile whose exact end live set equals G[end state] and whose G-mode walk
    never reaches the FAIL row is emitted from the G-mode table alone; any
    other tile from the exact live sets.  Returns (output, tiles emitted in
    G-mode, tiles)."""
    from kleenexlang_b200.fasttab import E_SYM, E_TPL, GB_T, GB_C
    G, gtab, _ = gm
    Q, C, A = t.Q, t.C, t.A
    n = len(data)
    states, acts = [t.init], []
    for b in data:
        e = f.trans2[states[-1] * C + t.cls[b]]
        assert (e & 0xFFFF) != Q, "model is for accepted inputs"
        states.append(e & 0xFFFF)
        acts.append((e >> 16) & 0xFF)
    assert t.final[states[-1]] >= 0
    lam = [0] * (n + 1)
    lam[n] = f.lam_final[states[-1]]
    for i in range(n - 1, -1, -1):
        lam[i] = ((f.BE[lam[i + 1] * A + acts[i]] & 0xFFFC) // 4) // A

    def template(x, b):
        info, hmask = f.tplinfo[2 * x], f.tplinfo[2 * x + 1]
        off, tl = info & 0xFFFF, (info >> 16) & 0xFF
        seg = bytearray(f.pool[off:off + tl])
        for h in range(min(tl, 32)):
            if (hmask >> h) & 1:
                seg[h] = b
        return seg

    out = bytearray()
    ntiles = (n + tile - 1) // tile
    gtiles = 0
    for k in range(ntiles):
        lo, hi = k * tile, min(n, (k + 1) * tile)
        ok = hi - lo == tile and G[states[hi]] == lam[hi]
        seg = bytearray()
        if ok:
            q = states[lo]
            for i in range(lo, hi):
                e = gtab[q * C + t.cls[data[i]]]
                q = e & 0xFFFF
                if q == Q:
                    ok = False
                    break
                ln, fl = (e >> 16) & 0xFF, e >> 24
                if fl & GB_T:
                    s2 = template(fl & 0x3F, data[i])
                    assert len(s2) == ln
                    seg += s2
                elif fl & GB_C:
                    seg.append(data[i])
        if ok:
            gtiles += 1
        exact = bytearray()
        for i in range(lo, hi):
            e = f.BE[lam[i + 1] * A + acts[i]]
            if (e & 3) == E_SYM:
                exact.append(data[i])
            elif (e & 3) == E_TPL:
                exact += template(e >> 24, data[i])
        if ok:
            assert seg == exact, "G-mode tile differs from the exact evaluation"
        out += exact
    for tgt, kind, ln, off in t.pieces[t.final[states[-1]]]:
        out += t.consts[off:off + ln]
    return bytes(out), gtiles, ntiles
