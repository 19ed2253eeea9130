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


# ---------------------------------------------------------------------------
# Generators for the multi-GB configurations (BASELINE.json configs[3], [4]): any byte range [lo, hi) of the
# text is a pure function of the position, so a rank makes only its own slab, and the whole 5 GB text can be made
# ON THE DEVICE (torch element-wise ops: plumbing for synthetic input, not part of the product) instead of
# through a Python loop over lines.  The same code runs on the CPU for the tests.
# ---------------------------------------------------------------------------
_FASTA_HEADERS = (b">ONE Homo sapiens alu\n", b">TWO IUB ambiguity codes\n", b">THREE Homo sapiens frequency\n")


def _fasta_char_tables(seed: int = 42):
    """The two random sections as lookup tables over the LCG's (prefix, cycle): byte per stream position."""
    prefix, cycle = _lcg_cycle(seed)

    def chars(table, vals):
        if len(vals) == 0:
            return np.zeros(0, dtype=np.uint8)
        cum = np.cumsum([p for _, p in table])
        cum[-1] = 1.0
        alphabet = np.frombuffer(b"".join(c for c, _ in table), dtype=np.uint8)
        return alphabet[np.searchsorted(cum, vals.astype(np.float64) / _IM, side="right").clip(0, len(alphabet) - 1)]
    return len(prefix), len(cycle), [(chars(t, prefix), chars(t, cycle)) for t in (_IUB, _HOMO)]


def fasta_file_size(n: int) -> int:
    return sum(len(h) + ln + (ln + 59) // 60 for h, ln in zip(_FASTA_HEADERS, (2 * n, 3 * n, 5 * n)))


def fasta_file_range(n: int, lo: int, hi: int, device="cpu", seed: int = 42, chunk: int = 1 << 27):
    """Bytes [lo, hi) of fasta_file(n) as a torch uint8 tensor on `device`."""
    import torch
    n_prefix, n_cycle, tabs = _fasta_char_tables(seed)
    dev = torch.device(device)
    alu = torch.from_numpy(np.frombuffer(_ALU, dtype=np.uint8).copy()).to(dev)
    lut = []
    for pre, cyc in tabs:
        lut.append((torch.from_numpy(pre.copy()).to(dev) if len(pre) else None, torch.from_numpy(cyc.copy()).to(dev)))
    heads = [torch.from_numpy(np.frombuffer(h, dtype=np.uint8).copy()).to(dev) for h in _FASTA_HEADERS]
    lens = (2 * n, 3 * n, 5 * n)
    out = torch.empty(hi - lo, dtype=torch.uint8, device=dev)
    starts, at = [], 0
    for h, ln in zip(_FASTA_HEADERS, lens):
        starts.append(at)
        at += len(h) + ln + (ln + 59) // 60
    stream_at = (0, 0, 3 * n)                      # random-stream position of a section's first letter
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        p = torch.arange(c0, c1, dtype=torch.int64, device=dev)
        res = torch.zeros(c1 - c0, dtype=torch.uint8, device=dev)
        for s in range(3):
            body = lens[s] + (lens[s] + 59) // 60
            q = p - starts[s]
            sel = (q >= 0) & (q < len(_FASTA_HEADERS[s]) + body)
            if not bool(sel.any()):
                continue
            q = q[sel]
            r = q - len(_FASTA_HEADERS[s])
            is_head = r < 0
            line, col = torch.div(r.clamp(min=0), 61, rounding_mode="floor"), r.clamp(min=0) % 61
            newline = (col == 60) | (r == body - 1)
            i = (line * 60 + col).clamp(max=lens[s] - 1)
            if s == 0:
                letter = alu[i % len(_ALU)]
            else:
                k = i + stream_at[s]
                pre, cyc = lut[s - 1]
                letter = cyc[(k - n_prefix).clamp(min=0) % n_cycle]
                if pre is not None:
                    letter = torch.where(k < n_prefix, pre[k.clamp(max=n_prefix - 1)], letter)
            val = torch.where(newline, torch.full_like(letter, 10), letter)
            val = torch.where(is_head, heads[s][q.clamp(max=len(_FASTA_HEADERS[s]) - 1)], val)
            res[sel] = val
        out[c0 - lo:c1 - lo] = res
    return out


_MIX1, _MIX2 = -7046029254386353131, -4658895280553007687       # 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9 as int64


def _mix64(x, xp):
    """A 64-bit mixer on int64 arrays / tensors (two's-complement wrap-around; logical shifts by masking)."""
    h = x * _MIX1
    h = h ^ ((h >> 32) & 0xFFFFFFFF)
    h = h * _MIX2
    h = h ^ ((h >> 29) & 0x7FFFFFFFF)
    return h


def source_text_range(lo: int, hi: int, device="cpu", seed: int = 3, chunk: int = 1 << 27):
    """Bytes [lo, hi) of an unbounded C-like source text (BASELINE.json configs[3], the jrep workload): printable
    ASCII, one line break per 64-byte cell at offset 20..59 (lines of 25..103 characters), every sixteenth line ends
    in ';' and is followed by a '}' (the multi-line literal ';\\n}' of sample/jrep.cc then occurs naturally, about
    once per KB).  A pure function of (position, seed)."""
    import torch
    dev = torch.device(device)
    out = torch.empty(hi - lo, dtype=torch.uint8, device=dev)
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        p = torch.arange(c0, c1, dtype=torch.int64, device=dev)
        cell, off = p >> 6, p & 63
        hc = _mix64(cell + seed * 1000003, torch)
        brk = 20 + (hc & 0xFFFF) % 40
        plant = ((hc >> 16) & 15) == 0
        b = 0x20 + (_mix64(p ^ (seed * 7919), torch) & 0xFFFFFF) % 95
        b = torch.where(off == brk, torch.full_like(b, 10), b)
        b = torch.where(plant & (off == brk - 1), torch.full_like(b, ord(";")), b)
        b = torch.where(plant & (off == brk + 1), torch.full_like(b, ord("}")), b)
        out[c0 - lo:c1 - lo] = b.to(torch.uint8)
    return out


def random_ascii_range(lo: int, hi: int, device="cpu", seed: int = 21, low: int = ord("0"), high: int = ord("z"),
                       chunk: int = 1 << 27):
    """Bytes [lo, hi) of an unbounded random text uniform in [low, high) (the reference benchmark's alphabet,
    tools/benchmarks/run.py:313), as a pure function of the position."""
    import torch
    dev = torch.device(device)
    out = torch.empty(hi - lo, dtype=torch.uint8, device=dev)
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        p = torch.arange(c0, c1, dtype=torch.int64, device=dev)
        out[c0 - lo:c1 - lo] = (low + (_mix64(p + seed * 1000003, torch) & 0xFFFFFF) % (high - low)).to(torch.uint8)
    return out
