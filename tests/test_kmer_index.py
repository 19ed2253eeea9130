"""The k-mer index of a fused pattern set (host/automaton.cc BuildKmerIndex) and
the arithmetic k_set_kmer does with it, restated in numpy on the CPU: pack 16
bytes into 2-bit codes with one AND + one multiply per word, look R
consecutive ends up in the bitmap indexed by 7 + R codes, check a hit exactly
against the bytes.  The candidate ends must equal every (overlapping) occurrence Python's
`re` finds; the GPU tier (test_gpu_parity.py) checks the kernel itself against
the oracle.  The patterns are the regex-dna variants of
/root/reference/sample/regexdna.cc:52-62.
"""
import re

import numpy as np
import pytest

import rejit_b200
from rejit_b200 import workloads as W

STREAM = 272          # kDfaStreamBytes
GROUPS = 17


def _pack(words, fm, mult):
    """KmerPack: four little-endian words -> 16 codes (32 bits)."""
    p = [((w & fm) * mult) & 0xFFFFFFFF for w in words]
    return (p[0] >> 24) | ((p[1] >> 24) << 8) | ((p[2] >> 24) << 16) | ((p[3] >> 24) << 24)


def emulate(kset, text: bytes, seed=0):
    info, bitmap, mask16 = kset.kmer_tables()
    shift, canon, canon_ok, K = int(info[0]), int(info[1]), int(info[2]), int(info[3])
    len_le = [int(x) for x in info[4:13]]
    lens = [len(re.sub(r"\[[^\]]*\]", "x", p.split("|")[0])) for p in kset.patterns]
    fm = np.uint64((0x03030303 << shift) & 0xFFFFFFFF)
    mult = np.uint64(0x01041040 >> shift)
    n = len(text)
    rng = np.random.RandomState(seed)
    nstreams = (n + STREAM - 1) // STREAM
    # what a lane sees: the 16 bytes before its sub-stream + the sub-stream; bytes
    # outside the text are arbitrary (stale shared memory in the kernel)
    buf = rng.randint(0, 256, size=16 + nstreams * STREAM + 16, dtype=np.uint8)
    buf[16:16 + n] = np.frombuffer(text, dtype=np.uint8)
    # adversarial garbage: bytes of the alphabet itself
    alpha = np.frombuffer(b"acgtACGTBD", dtype=np.uint8)
    buf[:16] = alpha[rng.randint(0, len(alpha), 16)]
    buf[16 + n:] = alpha[rng.randint(0, len(alpha), len(buf) - 16 - n)]
    idx = (np.arange(nstreams)[:, None] * STREAM + np.arange(16 + STREAM)[None, :])
    lanes = buf[idx]                                            # [nstreams, 288]
    words = lanes.reshape(nstreams, -1, 4).astype(np.uint64)
    words = words[:, :, 0] | (words[:, :, 1] << np.uint64(8)) | (words[:, :, 2] << np.uint64(16)) | (words[:, :, 3] << np.uint64(24))
    codes = [_pack([words[:, 4 * g + i] for i in range(4)], fm, mult) for g in range(GROUPS + 1)]
    found = [[] for _ in range(K)]
    R = int(info[13])
    idx_bits = 2 * (7 + R)
    word_bits = idx_bits - 5
    assert len(bitmap) == 1 << word_bits
    tests = (16 + R - 1) // R
    a = np.arange(nstreams, dtype=np.int64) * STREAM
    limit = np.minimum(a + STREAM, n)
    for g in range(GROUPS):
        P, Q = codes[g], codes[g + 1]
        both = P | (Q << np.uint64(32))
        for t in range(tests):
            x = min(t * R, 16 - R)                     # KmerTestX: first letter of the R ends this lookup answers
            w = (both >> np.uint64(16 + 2 * x)) & np.uint64(0xFFFFFFFF)
            word = bitmap[((w & np.uint64((4 << word_bits) - 4)) >> np.uint64(2)).astype(np.int64)].astype(np.uint64)
            hit = ((word << ((w >> np.uint64(word_bits + 2)) & np.uint64(31))) >> np.uint64(31)) & np.uint64(1)
            for s in np.nonzero(hit)[0]:
                idx = (int(both[s]) >> (18 + 2 * x)) & ((1 << idx_bits) - 1)
                for k in range(t * R - x, R):          # the last lookup overlaps the one before it
                    r = 16 * g + x + 1 + k
                    e = int(a[s]) + r
                    if e > limit[s]:
                        continue
                    x16 = (idx >> (2 * k)) & 0xFFFF
                    v, run = 0, True
                    for i in range(8):                  # byte e-8+i against the byte its code stands for
                        byte = int(lanes[s, 16 + r - 8 + i]) if e >= 8 - i else 0x100
                        code = (x16 >> (2 * i)) & 3
                        ok = ((canon >> (8 * code)) & 0xFF) == byte and (canon_ok >> code) & 1
                        if not ok:
                            v = 0
                        else:
                            v += 1
                    m = int(mask16[x16]) & len_le[v]
                    for j in range(K):
                        if (m >> j) & 1:
                            found[j].append((e - lens[j], e))
    return [sorted(f) for f in found]


def occurrences(pattern: str, text: bytes):
    rx = re.compile(b"(?=(" + pattern.encode("latin-1") + b"))")
    return [(m.start(1), m.end(1)) for m in rx.finditer(text)]


class _Set(rejit_b200.RegejSet):
    def __init__(self, patterns):
        super().__init__(patterns)
        self.patterns = list(patterns)


def test_dna_set_has_a_kmer_index():
    s = _Set(W.DNA_PATTERNS)
    t = s.kmer_tables()
    assert t is not None and "k-mer index" in s.describe()
    info, bitmap, mask16 = t
    assert int(info[0]) == 1                                  # (byte >> 1) & 3 separates a, c, g, t
    assert bytes(int(info[1]).to_bytes(4, "little")) == b"actg"
    # 2 + 8 * 6 distinct 8-mers are accepted by some member
    assert int(np.count_nonzero(mask16)) == 50


@pytest.mark.parametrize("n,seed", [(3000, 1), (20000, 2)])
def test_kmer_arithmetic_finds_every_occurrence(n, seed):
    s = _Set(W.DNA_PATTERNS)
    text = W.fasta_sequence(n).tobytes()
    # make the aliases bite: upper-case copies of real hits, hits cut by the text end and start
    text = b"ggtaaa" + text[:5000] + b"AGGGTAAA" + b"agggtaaB" + text[5000:] + b"agggtaaa" + b"tttaccc"
    got = emulate(s, text, seed)
    for j, p in enumerate(W.DNA_PATTERNS):
        assert got[j] == occurrences(p, text), p


def test_short_members_and_few_live_bytes():
    pats = ["ab", "b[ab]a", "abba|baab"]
    s = _Set(pats)
    assert s.kmer_tables() is not None
    rng = np.random.RandomState(3)
    text = bytes(rng.choice(np.frombuffer(b"abab ABc", dtype=np.uint8), 5000))
    got = emulate(s, text, 4)
    for j, p in enumerate(pats):
        assert got[j] == occurrences(p, text), p


def test_sets_without_a_kmer_index():
    assert _Set(["abcde", "edcba"]).kmer_tables() is None          # five live bytes
    assert _Set(["aaaaaaaaa", "ccccccccc"]).kmer_tables() is None  # nine bytes long
    assert _Set(["a[bB]", "ba"]).kmer_tables() is None             # no two adjacent bits tell a, b, B apart


@pytest.mark.parametrize("pats", [W.DNA_PATTERNS, ["ab", "b[ab]a", "abba|baab"]])
def test_bitmap_built_from_the_accepted_list(pats):
    """k_set_kmer builds the bitmap (and a hash of the member masks) in shared memory from the list of
    accepted 8-mers when that list is short: the same arithmetic on the CPU must give the host's bitmap."""
    info, bitmap, mask16 = _Set(pats).kmer_tables()
    R = int(info[13])
    word_bits = 2 * (7 + R) - 5
    built = np.zeros(1 << word_bits, dtype=np.uint32)
    slots = 512
    hkey = np.zeros(slots, dtype=np.uint32)
    hval = np.zeros(slots, dtype=np.uint32)
    accepted = [int(a) for a in np.nonzero(mask16)[0]]
    for a in accepted:
        # one atomicOr per (a, k, fl): the codes above the 8-mer (fh) only pick bits of the word
        for k in range(R):
            for fl in range(1 << (2 * k)):
                xl = fl | (a << (2 * k))
                sh = 16 + 2 * k - word_bits
                bits = 0
                for fh in range(1 << (2 * (R - 1 - k))):
                    bits |= 1 << (31 - ((xl >> word_bits) | (fh << sh)))
                built[xl & ((1 << word_bits) - 1)] |= np.uint32(bits)
        if len(accepted) <= 128:
            slot = ((a * 0x9E3B) >> 5) & (slots - 1)
            while hkey[slot]:
                slot = (slot + 1) & (slots - 1)
            hkey[slot], hval[slot] = a | 0x10000, mask16[a]
    assert np.array_equal(built, bitmap)
    if len(accepted) <= 128:                                  # kKmerListMax: the hash must hold every entry
        for a in accepted:
            slot = ((a * 0x9E3B) >> 5) & (slots - 1)
            while hkey[slot] != (a | 0x10000):
                assert hkey[slot] != 0
                slot = (slot + 1) & (slots - 1)
            assert hval[slot] == mask16[a]


def test_verify_swar_arithmetic():
    """KmerVerify: the eight bytes before an end as one 64-bit word; codes packed with the AND + multiply of
    KmerPack, the canonical byte of every code fetched with one byte-permute per four bytes."""
    info, _, _ = _Set(W.DNA_PATTERNS).kmer_tables()
    shift, canon, canon_ok = int(info[0]), int(info[1]), int(info[2])
    for cd in range(4):                                       # engine.cu: a code without a live byte gets a byte of another code
        if not (canon_ok >> cd) & 1:
            canon |= ((((cd + 1) & 3) << shift) & 0xFF) << (8 * cd)
    fm, mult = (0x03030303 << shift) & 0xFFFFFFFF, 0x01041040 >> shift
    rng = np.random.RandomState(11)
    alpha = np.frombuffer(b"acgtacgtacgtACGTBn\xff\x00", dtype=np.uint8)

    def prmt(a, sel):
        return sum(((a >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i) for i in range(4))

    for _ in range(3000):
        by = [int(b) for b in alpha[rng.randint(0, len(alpha), 8)]]
        lo = by[0] | by[1] << 8 | by[2] << 16 | by[3] << 24
        hi = by[4] | by[5] << 8 | by[6] << 16 | by[7] << 24
        x16 = ((((lo & fm) * mult) & 0xFFFFFFFF) >> 24) | (((((hi & fm) * mult) & 0xFFFFFFFF) >> 24) << 8)
        diff = 0
        for half, word in ((0, lo), (1, hi)):
            c = (word >> shift) & 0x03030303
            sel = ((c | (c >> 4)) & 0xFF) | (((c >> 8) | (c >> 12)) & 0xFF00)
            diff |= (prmt(canon, sel) ^ word) << (32 * half)
        v = 8 if diff == 0 else (64 - diff.bit_length()) >> 3
        # the definition, byte by byte
        exp_x, run = 0, 0
        for i, b in enumerate(by):
            code = (b >> shift) & 3
            exp_x |= code << (2 * i)
            run = run + 1 if ((canon >> (8 * code)) & 0xFF) == b else 0
        assert (x16, v) == (exp_x, run), by
