"""CPU model of the action-interpreter kernels (kleenexlang_b200/csrc/kex_act.cuh):
the same five passes, tile by tile, in plain Python, so the algorithm and its
seam handling can be checked against the oracle without a GPU.

Registers: slot h = builder at stack height h, slot NSLOT-1-r = action register r.
  P0  per tile: height delta, min/max height               -> height at tile starts
  P1  per tile: (fate, added length) summary                -> slot lengths at tile starts (forward scan)
  P2  per tile, exact lengths: length of r at every `write r`; backward summary (fate, tail offset)
  P3  backward scan: end position of every slot's content in the final output at tile ends
  P4  per tile, backwards: every byte is written at end[slot] - 1
"""
ESC = 0xFF
NSLOT = 32
DEAD = 0xFF
NOPOS = 0xFFFFFFFF


def tokens(s, lo, hi):
    """Tokens whose operand (or plain byte) lies in [lo, hi): (pos, kind, arg)."""
    out = []
    for i in range(lo, hi):
        b = s[i]
        if i > 0 and s[i - 1] == ESC and not _is_operand(s, i - 1):
            if b == 0:
                out.append((i, "byte", ESC))
            elif b == 1:
                out.append((i, "push", 0))
            elif b & 1:
                out.append((i, "write", (b - 3) // 2))
            else:
                out.append((i, "pop", (b - 2) // 2))
        elif b != ESC:
            out.append((i, "byte", b))
    return out


def _is_operand(s, i):
    # operands are never ESC, so a byte equal to ESC is always a lead
    return False


def run_model(s: bytes, nregs: int, T: int = 8, G: int = 4):
    n = len(s)
    ntiles = (n + T - 1) // T
    if ntiles == 0:
        return b""
    tl = [tokens(s, t * T, min(n, (t + 1) * T)) for t in range(ntiles)]
    # ---- P0
    delta, mn, mx = [], [], []
    for t in range(ntiles):
        h = lo = hi = 0
        for _, k, _ in tl[t]:
            if k == "push":
                h += 1
            elif k == "pop":
                h -= 1
            lo, hi = min(lo, h), max(hi, h)
        delta.append(h)
        mn.append(lo)
        mx.append(hi)
    h0 = [0] * (ntiles + 1)
    for t in range(ntiles):
        h0[t + 1] = h0[t] + delta[t]
        assert h0[t] + mn[t] >= 0, "pop on the bottom builder"
        assert h0[t] + mx[t] + 1 + nregs <= NSLOT, "stack too deep"

    # ---- P1: forward summaries
    def fwd_summary(t):
        fate = list(range(NSLOT))
        add = [0] * NSLOT
        h = h0[t]
        for _, k, r in tl[t]:
            N = NSLOT - 1 - r
            if k == "byte":
                add[h] += 1
            elif k == "push":
                h += 1
            elif k == "pop":
                for x in range(NSLOT):
                    if fate[x] == N:
                        fate[x] = DEAD
                for x in range(NSLOT):
                    if fate[x] == h:
                        fate[x] = N
                add[N] = add[h]
                add[h] = 0
                h -= 1
            else:
                for x in range(NSLOT):
                    if fate[x] == N:
                        fate[x] = h
                add[h] += add[N]
                add[N] = 0
        return fate, add

    def compose(s1, s2):
        f1, a1 = s1
        f2, a2 = s2
        f = [DEAD if f1[x] == DEAD else f2[f1[x]] for x in range(NSLOT)]
        a = list(a2)
        for m in range(NSLOT):
            if f2[m] != DEAD:
                a[f2[m]] += a1[m]
        return f, a

    def apply(sm, v):
        f, a = sm
        o = list(a)
        for m in range(NSLOT):
            if f[m] != DEAD:
                o[f[m]] += v[m]
        return o

    sums = [fwd_summary(t) for t in range(ntiles)]
    ngroups = (ntiles + G - 1) // G
    gsum = []
    for g in range(ngroups):
        acc = sums[g * G]
        for t in range(g * G + 1, min(ntiles, (g + 1) * G)):
            acc = compose(acc, sums[t])
        gsum.append(acc)
    gstart = [[0] * NSLOT]
    for g in range(ngroups):
        gstart.append(apply(gsum[g], gstart[g]))
    len0 = [None] * ntiles
    for g in range(ngroups):
        v = gstart[g]
        for t in range(g * G, min(ntiles, (g + 1) * G)):
            len0[t] = v
            v = apply(sums[t], v)
    total = gstart[ngroups][0]

    # ---- P2: exact lengths, write lengths, backward summaries
    wlen = {}
    bsum = []
    for t in range(ntiles):
        ln = list(len0[t])
        fate = list(range(NSLOT))
        endoff = list(len0[t])
        h = h0[t]
        for i, k, r in tl[t]:
            N = NSLOT - 1 - r
            if k == "byte":
                ln[h] += 1
            elif k == "push":
                h += 1
            elif k == "pop":
                for x in range(NSLOT):
                    if fate[x] == N:
                        fate[x] = DEAD
                for x in range(NSLOT):
                    if fate[x] == h:
                        fate[x] = N
                ln[N] = ln[h]
                ln[h] = 0
                h -= 1
            else:
                wlen[i >> 1] = ln[N]
                for x in range(NSLOT):
                    if fate[x] == N:
                        fate[x] = h
                        endoff[x] += ln[h]
                ln[h] += ln[N]
                ln[N] = 0
        tail = [0 if fate[x] == DEAD else ln[fate[x]] - endoff[x] for x in range(NSLOT)]
        bsum.append((fate, tail))

    # ---- P3: backward scan of end positions
    def bcompose(s1, s2):        # s1 earlier tile, s2 later tile
        f1, o1 = s1
        f2, o2 = s2
        f, o = [], []
        for x in range(NSLOT):
            if f1[x] == DEAD or f2[f1[x]] == DEAD:
                f.append(DEAD)
                o.append(0)
            else:
                f.append(f2[f1[x]])
                o.append(o1[x] + o2[f1[x]])
        return f, o

    def bapply(sm, e):
        f, o = sm
        return [NOPOS if f[x] == DEAD or e[f[x]] == NOPOS else e[f[x]] - o[x] for x in range(NSLOT)]

    gb = []
    for g in range(ngroups):
        acc = bsum[g * G]
        for t in range(g * G + 1, min(ntiles, (g + 1) * G)):
            acc = bcompose(acc, bsum[t])
        gb.append(acc)
    gend = [None] * ngroups
    e = [NOPOS] * NSLOT
    e[0] = total
    for g in range(ngroups - 1, -1, -1):
        gend[g] = e
        e = bapply(gb[g], e)
    eend = [None] * ntiles
    for g in range(ngroups):
        e = gend[g]
        for t in range(min(ntiles, (g + 1) * G) - 1, g * G - 1, -1):
            eend[t] = e
            e = bapply(bsum[t], e)

    # ---- P4: write
    out = bytearray(total)
    written = 0
    for t in range(ntiles):
        e = list(eend[t])
        h = h0[t + 1]
        for i, k, r in reversed(tl[t]):
            N = NSLOT - 1 - r
            if k == "byte":
                if e[h] != NOPOS:
                    e[h] -= 1
                    out[e[h]] = r
                    written += 1
            elif k == "push":
                h -= 1
            elif k == "pop":
                h += 1
                e[h] = e[N]
                e[N] = NOPOS
            else:
                e[N] = e[h]
                if e[h] != NOPOS:
                    e[h] -= wlen[i >> 1]
    assert written == total
    return bytes(out)
