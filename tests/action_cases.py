"""Seeded inputs for the register-action programs under programs/actions/."""
import os
import random

from conftest import PROGRAMS

NAMES = ["swap_fields", "reverse_items", "partition", "dropped", "nested", "swap_bytes", "pipeline_actions"]


def source(name):
    return open(os.path.join(PROGRAMS, "actions", name + ".kex"), encoding="utf-8").read()


def _word(rng, lo, hi, alphabet="abcdefghijklmnopqrstuvwxyz"):
    return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))


def gen(name, size, seed=0):
    """An accepted input of roughly `size` bytes."""
    rng = random.Random(seed * 7919 + len(name))
    out = []
    n = 0

    def put(s):
        nonlocal n
        out.append(s)
        n += len(s)
    if name == "swap_fields":
        while n < size:
            put((_word(rng, 0, 12) + "=" + _word(rng, 0, 40, "abcXYZ 0123=;") + "\n").encode())
    elif name == "pipeline_actions":
        while n < size:
            put((_word(rng, 0, 12, "abcdXYZ") + "=" + _word(rng, 0, 30, "abcXYZ=") + "\n").encode())
    elif name == "reverse_items":
        while n < size:
            put((_word(rng, 0, 9, "abc\n xyz") + ";").encode())
    elif name == "partition":
        while n < size:
            put((_word(rng, 0, 60, "aeioubcdfg xyz") + "\n").encode())
    elif name == "dropped":
        while n < size:
            put(((_word(rng, 1, 8) if rng.random() < 0.6 else _word(rng, 1, 5, "0123456789")) + " ").encode())
        put(b"\n")
    elif name == "nested":
        while n < size:
            put((_word(rng, 0, 20) + ",\n").encode())
    elif name == "swap_bytes":
        k = max(2, size // 2 * 2)
        put(bytes(rng.choice([0xFF, 0xFE, 0, 1, 2, 3, 65, 66, 10]) if rng.random() < 0.5 else rng.randrange(256)
                  for _ in range(k)))
    else:
        raise KeyError(name)
    return b"".join(out)


def rejecting(name, data, seed=0):
    """An input the program rejects (where the grammar allows one), else None."""
    if name == "swap_fields":          # (a pipeline's status is its last phase's: crt/crt.c:414-455)
        return data + b"no separator"
    if name == "reverse_items":
        return data + b"unterminated"
    if name == "dropped":
        return data[:-1] + b"?\n"
    if name == "nested":
        return data + b"abc"
    if name == "swap_bytes":
        return data + b"x"
    return None
