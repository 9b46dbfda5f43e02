"""Byte predicates as 256-bit integer masks.

Plays the role of the reference's sorted-interval byte sets
(src/KMC/RangeSet.hs:27-139) and of `coarsestPartition`
(src/KMC/Theories.hs:58-80).  A set is a Python int whose bit b is set iff
byte b is a member; this keeps conjunction/disjunction/complement O(1).
"""

UNIVERSE = (1 << 256) - 1
EMPTY = 0


def singleton(b: int) -> int:
    return 1 << b


def from_ranges(ranges) -> int:
    m = 0
    for lo, hi in ranges:
        if lo <= hi:
            m |= ((1 << (hi - lo + 1)) - 1) << lo
    return m


def complement(m: int) -> int:
    return UNIVERSE & ~m


def member(b: int, m: int) -> bool:
    return (m >> b) & 1 == 1


def size(m: int) -> int:
    return bin(m).count("1")


def to_list(m: int):
    return [b for b in range(256) if (m >> b) & 1]


def to_ranges(m: int):
    """Sorted maximal intervals, as RangeSet.ranges would hold them."""
    out = []
    b = 0
    while b < 256:
        if (m >> b) & 1:
            lo = b
            while b + 1 < 256 and (m >> (b + 1)) & 1:
                b += 1
            out.append((lo, b))
        b += 1
    return out


def is_subset(a: int, b: int) -> bool:
    return a & ~b == 0


def coarsest_partition(preds):
    """Coarsest partition refining every predicate in `preds`
    (Theories.hs:58-80).  Result order is by smallest member, which is a
    deterministic stand-in for the reference's greedy order; the members are
    pairwise disjoint so the order never changes a transducer's behaviour."""
    parts = []
    for p in preds:
        if p == 0:
            continue
        rest = p
        nxt = []
        for q in parts:
            i = q & rest
            if i:
                if q & ~rest:
                    nxt.append(q & ~rest)
                nxt.append(i)
                rest &= ~q
            else:
                nxt.append(q)
        if rest:
            nxt.append(rest)
        parts = nxt
    parts.sort(key=lambda m: (m & -m).bit_length())
    return parts
