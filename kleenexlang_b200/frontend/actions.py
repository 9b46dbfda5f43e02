"""Register actions (`reg@t`, `!reg`, `[reg <- ..]`, `[reg += ..]`) as a
two-phase stage: a transducer that writes an *action stream*, then an action
interpreter over that stream.

The reference compiles such programs only in its oracle/action mode
(`compileOracleAction`, src/KMC/Frontend/Commands.hs:204-244): an oracle SST
emits choice codes, an action SST (src/KMC/SymbolicSST/ActionSST.hs:47-129)
replays them and performs the register effects.  Here the split is made one
step later: the transducer keeps its output labels, the action symbols are
*part of its output alphabet* (escaped, below), so path-tree determinization
applies unchanged (no register action ever sits on an input edge), and the
second phase is the action semantics of src/KMC/Kleenex/Actions.hs:14-58 over
that stream -- the same interpretation `interp` gives the symbols in
ActionSST.hs:83-104 (push = next stack register, pop r = move, write r =
append and clear).  On accepted inputs the result is the reference's; like the
reference's oracle/action pair the stage occupies two phases.

Action stream encoding (bytes):
    b            (b != ESC)   output byte b
    ESC 0                     output byte ESC
    ESC 1                     push
    ESC 2+2r                  pop  register r
    ESC 3+2r                  write register r
"""
from .fst import FST

ESC = 0xFF
MAX_ACT_REGS = 120


class ActStage:
    """Second phase of a stage with register actions."""

    def __init__(self, regs):
        self.regs = list(regs)          # register names; index = r of the encoding
        self.nregs = len(self.regs)

    def __repr__(self):
        return "ActStage(%r)" % (self.regs,)


def registers_of(fst):
    """`registers` (ActionSST.hs:70-81): all register names of a transducer."""
    regs = set()

    def scan(ys):
        for y in ys:
            if not isinstance(y, int) and y[0] in ("pop", "write"):
                regs.add(y[1])
    for es in fst.sym.values():
        for _, f, _ in es:
            if f != "copy":
                scan(f[1])
    for es in fst.eps.values():
        for ys, _ in es:
            scan(ys)
    return sorted(regs)


def _encode(ys, rix):
    out = []
    for y in ys:
        if isinstance(y, int):
            out += [ESC, 0] if y == ESC else [y]
        elif y[0] == "push":
            out += [ESC, 1]
        elif y[0] == "pop":
            out += [ESC, 2 + 2 * rix[y[1]]]
        elif y[0] == "write":
            out += [ESC, 3 + 2 * rix[y[1]]]
        else:
            raise AssertionError(y)
    return tuple(out)


def action_stream_fst(fst):
    """-> (FST over plain bytes whose output is the action stream, ActStage)."""
    regs = registers_of(fst)
    if len(regs) > MAX_ACT_REGS:
        raise ValueError("more than %d action registers" % MAX_ACT_REGS)
    rix = {r: i for i, r in enumerate(regs)}
    esc_bit = 1 << ESC
    sym = {}
    for q, es in fst.sym.items():
        out = []
        for p, f, q2 in es:
            if f == "copy":
                if p & esc_bit:
                    # a copied ESC byte must leave escaped; the two halves of the
                    # predicate are disjoint, so the edge order is immaterial
                    if p & ~esc_bit:
                        out.append((p & ~esc_bit, "copy", q2))
                    out.append((esc_bit, ("const", (ESC, 0)), q2))
                else:
                    out.append((p, f, q2))
            else:
                out.append((p, ("const", _encode(f[1], rix)), q2))
        sym[q] = out
    eps = {q: [(_encode(ys, rix), q2) for ys, q2 in es] for q, es in fst.eps.items()}
    return FST(fst.states, sym, eps, fst.initial), ActStage(regs)


def decode_stream(stream: bytes):
    """Action stream -> symbols of `run_actions` (register = its index).  A
    trailing lone ESC (truncated stream) is ignored."""
    out = []
    i, n = 0, len(stream)
    while i < n:
        b = stream[i]
        if b != ESC:
            out.append(b)
            i += 1
            continue
        if i + 1 >= n:
            break
        c = stream[i + 1]
        i += 2
        if c == 0:
            out.append(ESC)
        elif c == 1:
            out.append(("push",))
        elif c & 1:
            out.append(("write", (c - 3) // 2))
        else:
            out.append(("pop", (c - 2) // 2))
    return out


def run_act_stream(stream: bytes) -> bytes:
    """The action interpreter (Actions.hs:14-58) over an action stream; the
    level-0 builder is the result even if the stream stops inside a push."""
    store = {}
    stack = [bytearray()]
    for s in decode_stream(stream):
        if isinstance(s, int):
            stack[-1].append(s)
        elif s[0] == "push":
            stack.append(bytearray())
        elif s[0] == "pop":
            if len(stack) < 2:
                raise ValueError("malformed action stream: pop on the bottom builder")
            store[s[1]] = stack.pop()
        else:
            stack[-1].extend(store.get(s[1], b""))
            store[s[1]] = bytearray()
    return bytes(stack[0])
