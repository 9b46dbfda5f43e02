"""Seeded synthetic inputs for the BASELINE.json configs, following the
distributions of the reference's generators (numpy, vectorised; no reliance
on Perl's rand):

  gen_csv       test/data/csv/gen_csv.pl:30-85   id=int(U*1e7); three names of
                exactly 20 letters; mail = 21 letters @ 21 letters . 3 letters;
                ip = four int(U*255)
  gen_datetime  test/data/datetime/gen_datetime.pl:22-45
  gen_numbers   test/data/numbers/gen_numbers.pl:37-56  (avglen digits per line)
  gen_fastq     4-line records: @id(30-50 bytes), 150 bases, +, 150 qualities

Every generator returns whole records only (the grammars are record*), so
blocks can be concatenated or tiled freely.
"""
import numpy as np

_ALPHA = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ", dtype=np.uint8)


def _digits(v, width):
    """[N, width] left-aligned decimal digits of v (0 = unused cell)."""
    v = v.astype(np.int64)
    nd = np.ones(len(v), dtype=np.int64)
    for k in range(1, width):
        nd += (v >= 10 ** k)
    out = np.zeros((len(v), width), dtype=np.uint8)
    for j in range(width):
        e = nd - 1 - j
        d = (v // np.power(10, np.maximum(e, 0))) % 10
        out[:, j] = np.where(e >= 0, d + 48, 0)
    return out


def _letters(rng, n, k):
    return _ALPHA[rng.integers(0, 52, size=(n, k))]


def _lit(n, s):
    return np.broadcast_to(np.frombuffer(s, dtype=np.uint8), (n, len(s)))


def _pack(cols):
    a = np.concatenate(cols, axis=1)
    flat = a.reshape(-1)
    return flat[flat != 0]


def gen_csv(nbytes, seed=0):
    """>= nbytes of CSV rows (mean row ~133 B), whole rows only."""
    rng = np.random.default_rng(seed)
    n = int(nbytes / 132.0) + 16
    while True:
        cols = [_digits((rng.random(n) * 1e7).astype(np.int64), 7), _lit(n, b","),
                _letters(rng, n, 20), _lit(n, b","), _letters(rng, n, 20), _lit(n, b","),
                _letters(rng, n, 21), _lit(n, b"@"), _letters(rng, n, 21), _lit(n, b"."), _letters(rng, n, 3),
                _lit(n, b","), _letters(rng, n, 20), _lit(n, b",")]
        for k in range(4):
            cols.append(_digits((rng.random(n) * 255).astype(np.int64), 3))
            cols.append(_lit(n, b"." if k < 3 else b"\n"))
        out = _pack(cols)
        if len(out) >= nbytes:
            return _trim_records(out, nbytes)
        n = int(n * 1.1) + 16


def _trim_records(buf, nbytes):
    """Shortest prefix of whole records that is >= nbytes."""
    nl = np.flatnonzero(buf[nbytes - 1:] == 10)
    end = nbytes - 1 + int(nl[0]) + 1
    return np.ascontiguousarray(buf[:end])


def gen_datetime(nbytes, seed=0):
    rng = np.random.default_rng(seed)
    n = int(nbytes / 24.0) + 16

    def two(lo, hi):
        v = rng.integers(lo, hi + 1, size=n)
        return np.stack([v // 10 + 48, v % 10 + 48], axis=1).astype(np.uint8)

    year = _digits(rng.integers(1000, 10000, size=n), 4)
    z = rng.random(n) < 0.3
    sign = np.where(rng.random(n) < 0.5, 43, 45).astype(np.uint8)[:, None]
    tzh, tzm = two(0, 23), two(0, 59)
    tz = np.concatenate([sign, tzh, _lit(n, b":"), tzm], axis=1).copy()
    tz[z] = 0
    tz[z, 0] = 90
    cols = [year, _lit(n, b"-"), two(1, 12), _lit(n, b"-"), two(1, 31), _lit(n, b"T"), two(0, 23), _lit(n, b":"),
            two(0, 59), _lit(n, b":"), two(0, 59), tz, _lit(n, b"\n")]
    out = _pack(cols)
    if len(out) < nbytes:
        return np.concatenate([out, gen_datetime(nbytes - len(out), seed + 1)])
    return _trim_records(out, nbytes)


def gen_numbers(nbytes, seed=0, avglen=1000):
    """gen_numbers.pl: each position is a newline with probability
    1/(avglen+1) unless the previous one was, else a uniform digit."""
    rng = np.random.default_rng(seed)
    out = (rng.integers(0, 10, size=nbytes) + 48).astype(np.uint8)
    nl = rng.random(nbytes) < 1.0 / (avglen + 1)
    nl[1:] &= ~nl[:-1]
    nl[0] = False
    out[nl] = 10
    out[-1] = 10
    if nbytes > 1 and out[-2] == 10:
        out[-2] = 48 + (seed % 10)
    return out


def gen_fastq(nbytes, seed=0):
    rng = np.random.default_rng(seed)
    n = int(nbytes / 340.0) + 8
    idlen = rng.integers(30, 51, size=n)
    ids = (rng.integers(0, 62, size=(n, 50)))
    idc = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", dtype=np.uint8)[ids]
    idc = np.where(np.arange(50)[None, :] < idlen[:, None], idc, 0).astype(np.uint8)
    bases = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, size=(n, 150))]
    qual = rng.integers(33, 74, size=(n, 150)).astype(np.uint8)   # '!'..'I', includes '@' and '+'
    cols = [_lit(n, b"@"), idc, _lit(n, b"\n"), bases, _lit(n, b"\n+\n"), qual, _lit(n, b"\n")]
    out = _pack(cols)
    if len(out) < nbytes:
        return np.concatenate([out, gen_fastq(nbytes - len(out), seed + 1)])
    # trim to whole 4-line records
    starts = np.flatnonzero(out == 10)
    rec_ends = starts[3::4] + 1
    k = int(np.searchsorted(rec_ends, nbytes))
    return np.ascontiguousarray(out[:rec_ends[min(k, len(rec_ends) - 1)]])


GENERATORS = {
    "csv2json": gen_csv,
    "iso_datetime_to_json": gen_datetime,
    "thousand_sep": gen_numbers,
    "add-commas": gen_numbers,
    "fastq2fasta": gen_fastq,
}
