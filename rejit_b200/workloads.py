"""Deterministic synthetic inputs and the pattern sets of BASELINE.json's configs.

  random_ascii ....... restates the reference benchmark's text generator,
                       /root/reference/tools/benchmarks/engines/bench_engine.cc:201-206
                       (`low + r % (high - low)` per byte, default range ['0','z'),
                       tools/benchmarks/run.py:313), with an explicit seed and no
                       NUL bytes.
  fasta_sequence ..... the Benchmarks-Game `fasta` program's three sections
                       (LCG IM=139968 IA=3877 IC=29573, seed 42) with headers and
                       line breaks already stripped, i.e. the text regex-dna
                       searches after its first ReplaceAll(">.*\\n|\\n", "")
                       (/root/reference/sample/regexdna.cc:49).  The generator is
                       not part of the reference (sample/regexdna.cc:32-37 only
                       points to it); it is restated here from its published
                       algorithm.
  fasta_file ......... the same with ">..." headers and 60-column lines.
"""
from __future__ import annotations

import numpy as np

# sample/regexdna.cc:52-62 — the nine variants (BASELINE.json says "8 patterns")
DNA_PATTERNS = [
    "agggtaaa|tttaccct",
    "[cgt]gggtaaa|tttaccc[acg]",
    "a[act]ggtaaa|tttacc[agt]t",
    "ag[act]gtaaa|tttac[agt]ct",
    "agg[act]taaa|ttta[agt]cct",
    "aggg[acg]aaa|ttt[cgt]ccct",
    "agggt[cgt]aa|tt[acg]accct",
    "agggta[cgt]a|t[acg]taccct",
    "agggtaa[cgt]|[acg]ttaccct",
]
# sample/regexdna.cc:69-85 — IUB code substitutions
IUB_SUBSTITUTIONS = [
    ("B", "(c|g|t)"), ("D", "(a|g|t)"), ("H", "(a|c|t)"), ("K", "(g|t)"), ("M", "(a|c)"),
    ("N", "(a|c|g|t)"), ("R", "(a|g)"), ("S", "(c|g)"), ("V", "(a|c|g)"), ("W", "(a|t)"),
    ("Y", "(c|t)"),
]
STRIP_PATTERN = ">.*\n|\n"          # sample/regexdna.cc:49
# tools/benchmarks/run.py:351, README.md:87
COMPLEX_PATTERN = "([complex]|(regexp)){2,7}abcdefgh(at|the|[e-nd]as well)"
COMPLEX_HITS = [b"ccregexpabcdefghthe", b"omabcdefghdas well", b"xcregexpregexpabcdefghat"]
LITERAL_PATTERN = "regexp"
JREP_PATTERN = ";\n}"               # jrep-style literal spanning a line break (sample/jrep.cc:230)


def random_ascii(n: int, seed: int = 1, low: int = ord("0"), high: int = ord("z")) -> np.ndarray:
    rng = np.random.RandomState(seed)
    # rand() yields 31-bit values; low + r % (high - low)
    r = rng.randint(0, 2 ** 31 - 1, size=n, dtype=np.int64)
    return (low + (r % (high - low))).astype(np.uint8)


def plant(text: np.ndarray, needles, every: int, seed: int = 7) -> np.ndarray:
    """Overwrites `text` with one of `needles` roughly every `every` bytes."""
    rng = np.random.RandomState(seed)
    pos = every // 2
    k = 0
    n = len(text)
    while True:
        nd = np.frombuffer(needles[k % len(needles)], dtype=np.uint8)
        if pos + len(nd) >= n:
            break
        text[pos:pos + len(nd)] = nd
        pos += every + int(rng.randint(0, max(1, every // 4)))
        k += 1
    return text


_ALU = (b"GGCCGGGCGCGGTGGCTCACGCCTGTAATCCCAGCACTTTGGGAGGCCGAGGCGGGCGGATCACCTGAGGTCAGGAGTTCGAGA"
        b"CCAGCCTGGCCAACATGGTGAAACCCCGTCTCTACTAAAAATACAAAAATTAGCCGGGCGTGGTGGCGCGCGCCTGTAATCCCAGC"
        b"TACTCGGGAGGCTGAGGCAGGAGAATCGCTTGAACCCGGGAGGCGGAGGTTGCAGTGAGCCGAGATCGCGCCACTGCACTCCAGC"
        b"CTGGGCGACAGAGCGAGACTCCGTCTCAAAAA")
_IUB = [(b"a", 0.27), (b"c", 0.12), (b"g", 0.12), (b"t", 0.27)] + \
       [(c, 0.02) for c in (b"B", b"D", b"H", b"K", b"M", b"N", b"R", b"S", b"V", b"W", b"Y")]
_HOMO = [(b"a", 0.3029549426680), (b"c", 0.1979883004921), (b"g", 0.1975473066391), (b"t", 0.3015094502008)]
_IM, _IA, _IC = 139968, 3877, 29573


def _lcg_cycle(seed: int):
    """One full period of the fasta LCG starting after `seed` (the modulus is
    small, so the stream is periodic and can be tiled)."""
    seq = []
    seen = {}
    x = seed
    while True:
        x = (x * _IA + _IC) % _IM
        if x in seen:
            start = seen[x]
            return np.array(seq[:start], dtype=np.int64), np.array(seq[start:], dtype=np.int64)
        seen[x] = len(seq)
        seq.append(x)


def _random_section(table, count: int, state: dict) -> np.ndarray:
    prefix, cycle = state["prefix"], state["cycle"]
    at = state["at"]
    idx = np.arange(at, at + count, dtype=np.int64)
    vals = np.where(idx < len(prefix), prefix[np.minimum(idx, max(len(prefix) - 1, 0))] if len(prefix) else 0,
                    cycle[(idx - len(prefix)) % len(cycle)])
    state["at"] = at + count
    r = vals.astype(np.float64) / _IM
    cum = np.cumsum([p for _, p in table])
    cum[-1] = 1.0
    chars = np.frombuffer(b"".join(c for c, _ in table), dtype=np.uint8)
    return chars[np.searchsorted(cum, r, side="right").clip(0, len(chars) - 1)]


def fasta_sequence(n: int, seed: int = 42) -> np.ndarray:
    """ONE (2n, ALU repeat) + TWO (3n, IUB) + THREE (5n, homo sapiens), no
    headers, no line breaks: 10n bytes."""
    alu = np.frombuffer(_ALU, dtype=np.uint8)
    one = np.tile(alu, (2 * n) // len(alu) + 1)[:2 * n]
    prefix, cycle = _lcg_cycle(seed)
    state = {"prefix": prefix, "cycle": cycle, "at": 0}
    two = _random_section(_IUB, 3 * n, state)
    three = _random_section(_HOMO, 5 * n, state)
    return np.concatenate([one, two, three])


def fasta_file(n: int, seed: int = 42) -> bytes:
    seq = fasta_sequence(n, seed)
    parts = []
    at = 0
    for header, ln in ((b">ONE Homo sapiens alu\n", 2 * n), (b">TWO IUB ambiguity codes\n", 3 * n),
                       (b">THREE Homo sapiens frequency\n", 5 * n)):
        parts.append(header)
        body = seq[at:at + ln].tobytes()
        at += ln
        parts.append(b"\n".join(body[i:i + 60] for i in range(0, len(body), 60)) + b"\n")
    return b"".join(parts)


def source_blob(n: int, seed: int = 3) -> np.ndarray:
    """A C-like source tree flattened into one text: printable ASCII lines of
    20-100 characters, a fraction of them ending in ";" followed by a "}" line
    (the jrep multi-line literal ';\\n}' then occurs naturally)."""
    rng = np.random.RandomState(seed)
    out = random_ascii(n, seed=seed + 11, low=0x20, high=0x7F)
    pos = 0
    while pos < n:
        ln = int(rng.randint(20, 101))
        pos += ln
        if pos >= n:
            break
        out[pos] = 0x0A
        if rng.randint(0, 16) == 0 and pos + 2 < n and pos >= 1:
            out[pos - 1] = ord(";")
            out[pos + 1] = ord("}")
        pos += 1
    return out
