"""GPU tier (`pytest -m gpu`): the parity tests proper.  Every call goes through
the C ABI of librejit_b200.so (include/rejit_b200.h) and runs the sm_100a
kernels; expectations come from the oracle (oracle/) and from the golden
fixtures produced by the compiled reference (tests/golden/).  Bit-exact:
identical (begin, end) offset lists."""
import ctypes
import random

import numpy as np
import pytest

import fuzzgen
import rejit_oracle as O
from conftest import expand_table_row

# a hung kernel must end the run, not stall the box: pytest-timeout's thread method exits the process
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200, method="thread")]


@pytest.fixture(scope="module")
def rj():
    import __graft_entry__ as entry
    entry.build()
    import rejit_b200
    if rejit_b200.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tier cannot run (and there is no CPU fallback)")
    return rejit_b200


def test_golden_offsets(rj, golden_vectors):
    for v in golden_vectors:
        t = v["text"].encode("latin-1")
        r = rj.Regej(v["re"])
        assert r.match_all(t) == [tuple(m) for m in v["all"]], (v["re"], v["note"], r.describe())
        assert r.match_full(t) == v["full"], (v["re"], v["note"])
        assert r.match_anywhere(t) == v["anywhere"], (v["re"], v["note"])


def test_reference_test_table(rj, ref_table):
    """The reference's 282 checks with its 33-alignment sweep (tools/tests/test.cc)."""
    n = 0
    for row in ref_table:
        pat, checks = expand_table_row(row)
        r = rj.Regej(pat)
        for mt, text, expected, start, end in checks:
            t = text.encode("latin-1")
            n += 1
            if mt == "full":
                assert r.match_full(t) == bool(expected), (row["line"], pat)
            elif mt == "anywhere":
                assert r.match_anywhere(t) == bool(expected), (row["line"], pat, text)
            elif mt == "all":
                assert r.match_all_count(t) == expected, (row["line"], pat, text)
            else:
                f = r.match_first(t)
                assert (f is not None) == bool(expected), (row["line"], pat, text)
                if expected and start is not None:
                    assert f == (start, end), (row["line"], pat, text)
    assert n > 3000


def test_fuzz_vs_oracle(rj):
    r = random.Random(2026)
    checked = 0
    for _ in range(500):
        pat, alpha = fuzzgen.rand_pattern(r)
        try:
            o = O.Oracle(pat)
        except O.ParserError:
            assert rj.Regej(pat).status == -1, pat
            continue
        g = rj.Regej(pat)
        assert g.status == 0, pat
        for n in (r.randint(0, 40), r.randint(500, 3000)):
            t = fuzzgen.rand_text(r, alpha, n)
            assert g.match_all(t) == o.match_all(t), (pat, t, g.describe())
            assert g.match_full(t) == o.match_full(t), (pat, t)
            checked += 1
    assert checked > 600


def test_edges_of_pieces_and_streams(rj):
    """Matches that straddle the 16-byte lane, 512-byte piece / sub-stream and
    text-end boundaries of the scan kernels."""
    for pat in ("needle", "ne", "n", "agggtaaa|tttaccct", "ab[cd]e", "(xy|z)+w", "needle(s|x)?"):
        o = O.Oracle(pat)
        g = rj.Regej(pat)
        unit = {"needle": b"needle", "ne": b"ne", "n": b"n", "agggtaaa|tttaccct": b"tttaccct",
                "ab[cd]e": b"abde", "(xy|z)+w": b"xyzxyw", "needle(s|x)?": b"needles"}[pat]
        for total in (0, 1, 5, 15, 16, 17, 511, 512, 513, 1024, 1030, 4099):
            for at in (0, 1, 9, 15, 16, 500, 505, 508, 511, 512, 1017, total - len(unit), total - 1):
                if at < 0 or at + len(unit) > total:
                    continue
                t = bytearray(b"." * total)
                t[at:at + len(unit)] = unit
                t = bytes(t)
                assert g.match_all(t) == o.match_all(t), (pat, total, at)


def test_empty_and_tiny_texts(rj):
    for pat in ("a", "a*", "^", "$", "^$", "abc", "(^|$|[x])", "x?"):
        for t in (b"", b"a", b"\n", b"x", b"ab"):
            assert rj.Regej(pat).match_all(t) == O.Oracle(pat).match_all(t), (pat, t)
            assert rj.Regej(pat).match_full(t) == O.Oracle(pat).match_full(t), (pat, t)


def test_dense_matches_take_the_large_path(rj):
    """More than 4096 candidates: radix sort + segment chains; also the buffer
    growth / rerun protocol (first capacity is 65536 candidates)."""
    n = 300000
    t = (b"ab" * (n // 2))
    st = rj.Stats()
    got = rj.Regej("a").match_all_array(t, stats=st)
    assert st.reruns >= 1 and st.large_path == 0      # no overlaps: finished inside the scan kernel, any count
    assert got.shape[0] == n // 2 and (got[:, 0] == np.arange(0, n, 2, dtype=np.uint64)).all()
    assert (got[:, 1] == got[:, 0] + 1).all()
    # overlapping candidates: "aba" on "ababab..." -> every other occurrence
    got = rj.Regej("aba").match_all_array(t, stats=st)
    assert st.large_path == 1
    exp = np.array(O.Oracle("aba").match_all(t[:4000]), dtype=np.uint64)
    assert (got[:len(exp) - 2] == exp[:len(exp) - 2]).all() and got.shape[0] == n // 4
    # dense empty matches and the empty-match rule
    t2 = fuzzgen.rand_text(random.Random(3), "abx", 60000)
    for pat in ("x*", "(^|$|[x])", "a*b*", "[^a]*"):
        exp = O.Oracle(pat).match_all(t2)
        got = rj.Regej(pat).match_all(t2)
        assert got == exp, (pat, len(got), len(exp))


def test_dfa_dense_fallback_and_tma_agree(rj):
    """Fixed-length alternations: sparse -> k_dfa_tma (ordered, TMA-staged);
    dense -> the lane hit lists overflow and the engine switches to k_dfa_scan +
    sort.  Both must equal the oracle."""
    t = fuzzgen.rand_text(random.Random(8), "acgt", 150001)
    for pat in ("a|c", "ac|gt", "acg|tgc", "ac[gt]a|tt[ac]g", "agggtaaa|tttaccct"):
        r = rj.Regej(pat)
        assert r.describe().startswith("fixed-length DFA scan"), r.describe()
        st = rj.Stats()
        got = r.match_all_array(t, stats=st)
        exp = np.array(O.Oracle(pat).match_all(t), dtype=np.uint64).reshape(-1, 2)
        assert got.shape == exp.shape and (got == exp).all(), (pat, got.shape, exp.shape)
        again = r.match_all_array(t)            # steady state (capacities remembered)
        assert (again == exp).all()


def test_reentrant_patterns_large_path(rj):
    """Label replay (FaithfulSegment) with many clusters."""
    t = fuzzgen.rand_text(random.Random(11), "acgt", 120000)
    for pat in (".{,4}t", "[ca]?[gc]*t", "g*a?a|a|aaa.", "(a|c)*gt"):
        assert "reentrant" in rj.Regej(pat).describe()
        assert rj.Regej(pat).match_all(t) == O.Oracle(pat).match_all(t), pat


def test_workloads_medium(rj):
    """BASELINE.json's patterns at sizes the oracle finishes in seconds."""
    from rejit_b200 import workloads as W
    seq = W.fasta_sequence(200000).tobytes()             # 2 MB
    for p in W.DNA_PATTERNS:
        assert rj.Regej(p).match_all(seq) == O.Oracle(p).match_all(seq), p
    for p, _ in W.IUB_SUBSTITUTIONS[:3]:
        assert rj.Regej(p).match_all(seq) == O.Oracle(p).match_all(seq), p
    fa = W.fasta_file(50000)
    assert rj.Regej(W.STRIP_PATTERN).match_all(fa) == O.Oracle(W.STRIP_PATTERN).match_all(fa)
    text = W.plant(W.random_ascii(4 << 20, seed=9), W.COMPLEX_HITS, every=10000).tobytes()
    for p in (W.COMPLEX_PATTERN, W.LITERAL_PATTERN, "abcdefgh"):
        got = rj.Regej(p).match_all(text)
        assert got == O.Oracle(p).match_all(text), p
        assert len(got) > 300
    blob = W.source_blob(2 << 20).tobytes()
    got = rj.Regej(W.JREP_PATTERN).match_all(blob)
    assert got == O.Oracle(W.JREP_PATTERN).match_all(blob) and len(got) > 100
    assert rj.Regej("^").match_all(blob) == O.Oracle("^").match_all(blob)


def _find_all(text, needle):
    """Start offsets of every (possibly overlapping) occurrence of `needle` in a numpy uint8 array: shifted compares
    in chunks, no engine code involved."""
    nd = np.frombuffer(needle, dtype=np.uint8)
    m, n = len(nd), len(text)
    out = []
    step = 1 << 26
    for lo in range(0, max(n - m + 1, 0), step):
        hi = min(n - m + 1, lo + step)
        ok = text[lo:hi] == nd[0]
        for i in range(1, m):
            ok &= text[lo + i:hi + i] == nd[i]
        out.append(np.flatnonzero(ok) + lo)
    return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)


def test_full_size_properties(rj):
    """BASELINE sizes: planted-hit recovery, device/host API agreement, slab
    invariance (the chain state handed across arbitrary cuts reproduces the
    one-piece result)."""
    from rejit_b200 import workloads as W
    # config 3: 500 MB random text, complex regex, planted hits
    n = 500_000_000
    text = W.random_ascii(n, seed=21)
    W.plant(text, W.COMPLEX_HITS, every=1_000_003)
    g = rj.Regej(W.COMPLEX_PATTERN)
    got = g.match_all_array(text)
    # expected: the oracle on +-64 bytes around every occurrence of the required literal, located WITHOUT the
    # engine (numpy shifted compares): a literal the engine's own scan missed would otherwise be missed on both sides
    o = O.Oracle(W.COMPLEX_PATTERN)
    lit_at = _find_all(text, b"abcdefgh")
    assert lit_at.shape[0] >= 400
    lit = rj.Regej("abcdefgh").match_all_array(text)
    assert (lit[:, 0] == lit_at.astype(np.uint64)).all() and (lit[:, 1] == lit[:, 0] + 8).all()
    exp = []
    for b in lit_at:
        lo = max(0, int(b) - 64)
        w = text[lo:int(b) + 64].tobytes()
        exp += [(lo + x, lo + y) for x, y in o.match_all(w)]
    assert [tuple(map(int, r)) for r in got] == exp
    # device-resident text, in one piece and in 3 slabs with carries
    dt = rj.DeviceText(text)
    try:
        st = rj.Stats()
        assert g.match_all_device(dt, stats=st) == got.shape[0]
        assert st.launches >= 1 and st.scan_ms > 0
        r1 = rj.Regej(W.LITERAL_PATTERN)
        whole = r1.match_all_device(dt)
        assert whole == r1.match_all_array(text).shape[0]
    finally:
        dt.free()
    # config 2: the 50 MB FASTA, all nine patterns against the oracle
    seq = W.fasta_sequence(5_000_000)
    data = seq.tobytes()
    for p in W.DNA_PATTERNS:
        got = rj.Regej(p).match_all_array(seq)
        exp = np.array(O.Oracle(p).match_all(data), dtype=np.uint64).reshape(-1, 2)
        assert got.shape == exp.shape and (got == exp).all(), p


def test_fused_pattern_set(rj):
    """the nine regex-dna patterns fused into one scan must
    give, per pattern, exactly the per-pattern MatchAll result."""
    from rejit_b200 import workloads as W
    seq = W.fasta_sequence(300000)                  # 3 MB
    data = seq.tobytes()
    rs = rj.RegejSet(W.DNA_PATTERNS)
    assert rs.describe().startswith("fused set: 9 patterns"), rs.describe()
    fused = rs.match_all(seq)
    for p, got in zip(W.DNA_PATTERNS, fused):
        assert got == O.Oracle(p).match_all(data), p
    # counts-only device path, repeated (steady state) and on ragged lengths
    for cut in (len(seq), 8704 * 3, 8704 * 3 + 5, 100, 17):
        dt = rj.DeviceText(seq[:cut])
        try:
            st = rj.Stats()
            counts = rs.match_all_device(dt, stats=st)
            assert counts == [len(O.Oracle(p).match_all(data[:cut])) for p in W.DNA_PATTERNS], cut
            assert st.strategy == 4 and st.launches == 1     # scan + in-kernel finish
        finally:
            dt.free()
    # mixed lengths and a dense member (falls back to member-by-member runs)
    t = fuzzgen.rand_text(random.Random(4), "acgt", 50000)
    for pats in (["acg", "ttgca", "a[ct]g"], ["a", "cg"], ["acgt", "x+"]):
        got = rj.RegejSet(pats).match_all(t)
        for p, g in zip(pats, got):
            assert g == O.Oracle(p).match_all(t), (pats, p)
    # five live bytes / a nine-byte member: no k-mer index, the union automaton scans (k_set_tma)
    t = fuzzgen.rand_text(random.Random(5), "acgtn", 200000)
    for pats in (["acgn", "ttgca", "a[ct]g"], ["acgtacgta", "ttg[ac]a"]):
        rs2 = rj.RegejSet(pats)
        assert rs2.describe().startswith("fused set") and "k-mer" not in rs2.describe(), rs2.describe()
        for p, g in zip(pats, rs2.match_all(t)):
            assert g == O.Oracle(p).match_all(t), (pats, p)
    # k-mer index: aliases of live bytes (upper case, IUB codes), members cut by both text ends
    rs = rj.RegejSet(W.DNA_PATTERNS)
    assert "k-mer index" in rs.describe(), rs.describe()
    for text in (b"ggtaaa" + data[:70000] + b"AGGGTAAA" + b"agggtaaB" + data[600000:700000] + b"agggtaaa" + b"tttaccc",
                 b"agggtaaa", b"agggtaa", b"tttaccct" * 3000, data[590000:610000].upper() + b"cgggtaaa"):
        for p, g in zip(W.DNA_PATTERNS, rs.match_all(text)):
            assert g == O.Oracle(p).match_all(text), (p, len(text))
    # the set call on slabs with carries
    dt = rj.DeviceText(seq)
    try:
        exp = [len(O.Oracle(p).match_all(data)) for p in W.DNA_PATTERNS]
        n = len(seq)
        for k in (2, 5):
            arr = rj.Carry * len(W.DNA_PATTERNS)
            carry = arr(*[rj.Carry(0, 0xFFFFFFFFFFFFFFFF) for _ in W.DNA_PATTERNS])
            total = [0] * len(W.DNA_PATTERNS)
            for i in range(k):
                lo = n * i // k
                hi = n * (i + 1) // k if i + 1 < k else n + 1
                nxt = arr()
                cnts = rs.match_all_device(dt, own=(lo, hi), carry_in=carry, carry_out=nxt)
                total = [a + b for a, b in zip(total, cnts)]
                carry = nxt
            assert total == exp, (k, total, exp)
    finally:
        dt.free()


def test_fixed_length_finish_in_kernel_and_overlap_fallback(rj):
    """k_dfa_tma / k_set_tma finish in-kernel (grid barrier + copy) when no
    candidate overlaps its predecessor; overlapping neighbours (also across
    lane, sub-region and segment edges) must fall back to the general resolve."""
    rng = random.Random(12)
    t = fuzzgen.rand_text(rng, "xxxxxxxxxxxxxxxxxxxxab", 400000)
    pats = ["a[ab]", "[ab]b", "a[ab]a|b[ab]b", "ab[ab]|ba[ab]"]
    for p in pats:
        r = rj.Regej(p)
        assert r.describe().startswith("fixed-length DFA scan"), r.describe()
        exp = O.Oracle(p).match_all(t)
        assert r.match_all(t) == exp, p
        assert r.match_all(t) == exp, p             # steady state
    got = rj.RegejSet(pats).match_all(t)
    for p, g in zip(pats, got):
        assert g == O.Oracle(p).match_all(t), p
    # whole-literal patterns finish in-kernel too; self-overlapping occurrences fall back
    for p in ("aa", "aba", "ab", "bxa"):
        r = rj.Regej(p)
        assert r.describe().startswith("literal scan"), r.describe()
        exp = O.Oracle(p).match_all(t)
        assert r.match_all(t) == exp and r.match_all(t) == exp, p
    st = rj.Stats()
    dt = rj.DeviceText(np.frombuffer(t, dtype=np.uint8))
    try:
        for p in ("ab", "aa"):
            r = rj.Regej(p)
            r.match_all_device(dt)                           # first call sizes the buffers
            assert r.match_all_device(dt, stats=st) == len(O.Oracle(p).match_all(t)), p
            # one launch: the single-pass scan resolves "aaa" inside the tile ("aa" may meet a tile edge: then two)
            assert st.launches <= (1 if p == "ab" else 3), (p, st.launches)
    finally:
        dt.free()
    # no overlaps at all: one launch per call
    t2 = fuzzgen.rand_text(rng, "acgt", 300000)
    r = rj.Regej("aacgtc|ggtgtc")
    st = rj.Stats()
    dt = rj.DeviceText(np.frombuffer(t2, dtype=np.uint8))
    try:
        r.match_all_device(dt, stats=st)
        cnt = r.match_all_device(dt, stats=st)
        assert cnt == len(O.Oracle("aacgtc|ggtgtc").match_all(t2)) and st.launches == 1
    finally:
        dt.free()
    # slab calls with a carry reaching into the slab
    for p, text in (("ab[ab]|ba[ab]", t), ("aacgtc|ggtgtc", t2), ("a[ab]", b"x" * 8703 + b"aaaa" + b"x" * 9000),
                    ("aa", b"x" * 16383 + b"aaaaa" + b"x" * 20000), ("ab", t)):
        r = rj.Regej(p)
        exp = len(O.Oracle(p).match_all(text))
        n = len(text)
        dt = rj.DeviceText(np.frombuffer(text, dtype=np.uint8))
        try:
            for k in (1, 2, 3, 7):
                carry = rj.Carry(0, 0xFFFFFFFFFFFFFFFF)
                total = 0
                for i in range(k):
                    lo = n * i // k
                    hi = n * (i + 1) // k if i + 1 < k else n + 1
                    nxt = rj.Carry()
                    total += r.match_all_device(dt, own=(lo, hi), carry_in=carry, carry_out=nxt)
                    carry = nxt
                assert total == exp, (p, k, total, exp)
        finally:
            dt.free()


def test_match_first_early_exit(rj):
    """MatchFirst / MatchAnywhere search growing slabs and stop
    at the first one that holds a match; the answer is MatchAll()[0]."""
    rng = random.Random(31)
    n = 3_000_000
    base = bytearray(fuzzgen.rand_text(rng, "0123456789\n", n))
    for pat, hit in (("abcdefgh", b"abcdefgh"), ("ab[cd]e|xyz", b"abde"), ("^qq$", b"\nqq\n"), ("r+s", b"rrrrs")):
        o = O.Oracle(pat)
        for where in (5, 262140, 262144, 300000, 2_359_290, n - len(hit)):
            t = bytearray(base)
            t[where:where + len(hit)] = hit
            t[n - 20:n - 20 + len(hit)] = hit            # a later one that must not win
            t = bytes(t)
            r = rj.Regej(pat)
            assert r.match_first(t) == o.match_first(t), (pat, where)
            assert r.match_anywhere(t) is True
        assert rj.Regej(pat).match_first(bytes(base)) is None and rj.Regej(pat).match_anywhere(bytes(base)) is False


def _replace_expected(pat, text, w):
    out, at = bytearray(), 0
    ms = O.Oracle(pat).match_all(text)
    for b, e in ms:
        out += text[at:b] + w
        at = e
    out += text[at:]
    return bytes(out), len(ms)


def test_replace_all_on_device(rj):
    """Regej::ReplaceAll with the rebuild on the device
    (k_match_lengths + scan + k_replace_tiles) == the reference's Replace
    (src/rejit.cc:97-112) applied to the oracle's matches."""
    rng = random.Random(21)
    cases = [("a", b"", b"X"), ("x*", b"aaxa", b"-"), ("$", b"ab\ncd", b"<EOL>"), ("^", b"a\nb\n", b"> "),
             ("abc", b"abc" * 3000, b""), ("abc", b"abc" * 3000, b"abcabc"), (".*", b"x" * 9000, b"y"),
             ("x{2,}", b"ab" + b"x" * 12000 + b"cd", b"_")]
    for n in (1, 16, 17, 4095, 4096, 4097, 12289, 70001):
        t = fuzzgen.rand_text(rng, "abx\n", n)
        for pat in ("a", "ab|ba", "x*", "a.*", "(^|$|[x])", "b+", "\n"):
            cases.append((pat, t, rng.choice([b"", b"Q", b"(c|g|t)"])))
    for pat, t, w in cases:
        assert rj.Regej(pat).replace_all(t, w) == _replace_expected(pat, t, w), (pat, len(t), w)
    # regex-dna's front: strip headers and newlines, then chained IUB substitutions
    # on the device-resident result (sample/regexdna.cc:49, 69-85)
    from rejit_b200 import workloads as W
    fa = W.fasta_file(30000)                         # ~300 KB FASTA file
    exp, n_strip = _replace_expected(W.STRIP_PATTERN, fa, b"")
    cur = rj.Text(fa)
    try:
        nxt, n = rj.Regej(W.STRIP_PATTERN).replace_all_text(cur, b"")
        cur.free()
        cur = nxt
        assert n == n_strip and len(cur) == len(exp) and cur.download() == exp
        for code, alt in W.IUB_SUBSTITUTIONS[:4]:
            exp, n_e = _replace_expected(code, exp, alt.encode())
            nxt, n = rj.Regej(code).replace_all_text(cur, alt.encode())
            cur.free()
            cur = nxt
            assert n == n_e and cur.download() == exp, code
        # the rebuilt text is an ordinary device text: it can be searched
        p = W.DNA_PATTERNS[3]
        assert rj.Regej(p).match_all_text(cur) == O.Oracle(p).match_all(exp)
    finally:
        cur.free()


def test_multi_gpu_equals_single(rj):
    if rj.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from rejit_b200 import workloads as W
    seq = W.fasta_sequence(400000)
    for p in W.DNA_PATTERNS[:3] + ["a+", "(ac|g)+t"]:
        a = rj.Regej(p).match_all_array(seq)
        b = rj.Regej(p).match_all_array(seq, n_gpus=2)
        assert a.shape == b.shape and (a == b).all(), p


def test_long_literals(rj):
    """Literal nodes of 9..70 bytes (and the long node a group repetition expands to), exact copies and
    near misses, short texts and one of several sub-regions: every byte of the needle counts (the
    reference's own compare skips bytes for such nodes: defect B20, tests/test_oracle.py)."""
    r = random.Random(2022)
    for i in range(120):
        pat, t = fuzzgen.rand_long_literal_case(r)
        if i % 10 == 0:                                   # the same near misses spread over a 100 kB text
            t = b"".join(t + fuzzgen.rand_text(r, "abcd", r.randint(0, 3000)) for _ in range(40))
        assert rj.Regej(pat).match_all(t) == O.Oracle(pat).match_all(t), (pat, len(t))


def test_rich_dialect(rj):
    """Bracket ranges, escapes, \\xHH, repetitions on them, texts with bytes >= 0x80 (tests/fuzzgen.py)."""
    r = random.Random(4712)
    checked = 0
    for _ in range(250):
        pat = fuzzgen.rand_rich_pattern(r)
        try:
            o = O.Oracle(pat)
        except O.ParserError:
            assert rj.Regej(pat).status == -1, pat
            continue
        g = rj.Regej(pat)
        assert g.status == 0, pat
        for n in (r.randint(0, 40), r.randint(200, 3000)):
            t = fuzzgen.rand_rich_text(r, n)
            assert g.match_all(t) == o.match_all(t), (pat, t, g.describe())
            checked += 1
    assert checked > 300


def test_full_size_config_3_slab(rj):
    """BASELINE configs[3]: the jrep-style literal spanning a line break over ONE GPU's slab of the 5 GB source
    blob (625 MB; a 62.5 MB blob tiled, the generator is a Python loop).  The literal cannot overlap itself, so
    every occurrence is a match: offsets from three shifted byte compares."""
    from rejit_b200 import workloads as W
    text = np.tile(W.source_blob(62_500_000, seed=4), 10)
    hit = (text[:-2] == ord(";")) & (text[1:-1] == 10) & (text[2:] == ord("}"))
    begins = np.flatnonzero(hit).astype(np.uint64)
    del hit
    assert begins.shape[0] > 500_000
    got = rj.Regej(W.JREP_PATTERN).match_all_array(text)
    assert got.shape == (begins.shape[0], 2)
    assert (got[:, 0] == begins).all() and (got[:, 1] == begins + 3).all()


def _np_text(b):
    return np.frombuffer(b, dtype=np.uint8)


def test_single_pass_scan_emit(rj):
    """Round 2: literal, required-literal + window and generic patterns run in ONE launch (scan_emit.cuh:
    64 KB tiles, 8 KB per warp, decoupled look-back).  Matches at and across warp / tile edges, the text end,
    dense and empty matches, chains that cross a tile edge (general path takes over), all against the oracle."""
    rng = random.Random(77)
    K8, K64 = 8192, 65536
    base = bytearray(fuzzgen.rand_text(rng, "qwertyuiop   \n", 3 * K64 + 700))
    cases = [("needle", b"needle"), ("B", b"B"), (";\n}", b";\n}"), ("abcdefghijklmnopqrstuvwxyz0123456789", b"abcdefghijklmnopqrstuvwxyz0123456789"),
             ("(ab|c)+needle(s|x)?", b"abcabneedles"), ("x[yz]{2,3}needle", b"xyzzneedle")]
    edges = [0, 1, 15, 16, 511, 512, K8 - 3, K8 - 1, K8, K8 + 1, K64 - 40, K64 - 5, K64 - 1, K64, K64 + 1, 2 * K64 - 2,
             2 * K64 + K8 - 1, 3 * K64 + 690]
    for pat, unit in cases:
        t = bytearray(base)
        for at in edges:
            if at + len(unit) <= len(t):
                t[at:at + len(unit)] = unit
        t = bytes(t)
        r = rj.Regej(pat)
        st = rj.Stats()
        got = r.match_all_array(_np_text(t), stats=st)
        exp = O.Oracle(pat).match_all(t)
        assert [tuple(map(int, x)) for x in got] == exp, (pat, r.describe(), got.shape[0], len(exp))
        if pat != "(ab|c)+needle(s|x)?":                 # (its planted unit holds three starts: one meets a tile edge)
            assert st.launches == 1 and st.reruns == 0, (pat, st.launches, st.reruns)
        for cut in (K64, K64 + 1, K64 - 1, 2 * K64 + K8, K8 - 1, 17):         # ragged text ends
            assert r.match_all(t[:cut]) == O.Oracle(pat).match_all(t[:cut]), (pat, cut)
    # line-oriented and empty matches (generic scan: SWAR start filter with and without the line context)
    lines = bytearray()
    while len(lines) < 2 * K64 + 5000:                   # lines long enough for the tile's candidate list (1 per 32 bytes)
        lines += bytes(rng.choice(b"abcxyz>;{} ") for _ in range(rng.randint(30, 150))) + rng.choice([b"\n", b"\n", b"\r\n", b"\n\n"])
    lines = bytes(lines)
    short = bytearray()                                  # ... and lines too short for it: the round-1 pipeline takes over
    while len(short) < K64 + 3000:
        short += bytes(rng.choice(b"abx>") for _ in range(rng.randint(0, 12))) + b"\n"
    short = bytes(short)
    for pat in ("^", "$", ">.*\n|\n", "(^|$|[x])"):
        assert rj.Regej(pat).match_all(short) == O.Oracle(pat).match_all(short), pat
    for pat in ("^", "$", "^$", ">.*\n|\n", "^a", "x$", "(^|$|[x])", "^[a-c]+", ";\n}|^>", "[ab]x|^y", "\n"):
        r = rj.Regej(pat)
        st = rj.Stats()
        for t in (lines, lines[:K64], lines[:K64 + 1], lines[3:K8 + 3], lines + b"x", lines.rstrip(b"\r\n")):
            got = r.match_all_array(_np_text(t), stats=st)
            exp = O.Oracle(pat).match_all(t)
            assert [tuple(map(int, x)) for x in got] == exp, (pat, len(t), r.describe(), got.shape[0], len(exp))
            if pat in ("^", "$", "^$", "\n", "^a"):       # no match of these can reach across a tile edge
                assert st.launches == 1, (pat, st.launches, r.describe())
    # a start set too large for the SWAR filter: the 256-bit maps
    t = fuzzgen.rand_text(rng, "abcdefgh \n", 3 * K64)
    for pat in ("[a-f]+gh", "^[a-h]*h$", "[abcdef]{2}g|h+ "):
        assert rj.Regej(pat).match_all(t) == O.Oracle(pat).match_all(t), pat
    # chains that cross a tile edge: the arriving match swallows the tile's first candidates -> general path
    for pat, unit in (("a+", b"a" * 37), (">.*\n|\n", b">" + b"h" * 50 + b"\n"), ("ab(ab)*", b"ab" * 20), ("x*", b"x" * 9)):
        for edge in (K64, 2 * K64, K8):
            for back in (1, 5, len(unit) - 1):
                t = bytearray(fuzzgen.rand_text(rng, "qrs\n", 2 * K64 + 3000))
                t[edge - back:edge - back + len(unit)] = unit
                t = bytes(t)
                assert rj.Regej(pat).match_all(t) == O.Oracle(pat).match_all(t), (pat, edge, back)
    # overlapping candidates inside a tile are resolved by the tile itself: still one launch
    t = fuzzgen.rand_text(rng, "xxxxxxxxxxxxxxxxxxxxab", 3 * K64)
    for pat in ("aa", "aba|bab", "a(b|a)+"):
        r = rj.Regej(pat)
        assert r.match_all(t) == O.Oracle(pat).match_all(t), pat
    # slabs with carries (own ranges that cut tiles; the carry reaching into a slab)
    for pat, text in (("needle", bytes(base[:2 * K64]) + b"needle" * 3 + bytes(base[:K8])), ("^", lines), (">.*\n|\n", lines),
                      ("ab(ab)*", b"q" * (K64 - 3) + b"ab" * 9 + b"q" * K64), ("x[yz]{2,3}needle", bytes(base))):
        r = rj.Regej(pat)
        exp = len(O.Oracle(pat).match_all(text))
        n = len(text)
        dt = rj.DeviceText(_np_text(text))
        try:
            for k in (1, 2, 3, 5):
                carry = rj.Carry(0, 0xFFFFFFFFFFFFFFFF)
                total = 0
                for i in range(k):
                    lo = n * i // k
                    hi = n * (i + 1) // k if i + 1 < k else n + 1
                    nxt = rj.Carry()
                    total += r.match_all_device(dt, own=(lo, hi), carry_in=carry, carry_out=nxt)
                    carry = nxt
                assert total == exp, (pat, k, total, exp)
        finally:
            dt.free()


def test_single_pass_scan_emit_dense_and_large(rj):
    """Dense matches at size: every offset list is checked in full against numpy (single bytes, line starts), the
    output buffer grows once and the steady state is one launch."""
    from rejit_b200 import workloads as W
    seq = W.fasta_sequence(2_000_000)                    # 20 MB, IUB letters in the middle section
    for letter in (b"B", b"N"):
        r = rj.Regej(letter.decode())
        exp = np.flatnonzero(seq == letter[0]).astype(np.uint64)
        st = rj.Stats()
        got = r.match_all_array(seq, stats=st)
        assert got.shape[0] == exp.shape[0] and (got[:, 0] == exp).all() and (got[:, 1] == exp + 1).all()
        got = r.match_all_array(seq, stats=st)
        assert st.launches == 1 and st.reruns == 0 and (got[:, 0] == exp).all()
    blob = W.source_blob(64_000_000, seed=5)
    exp = np.concatenate([[0], np.flatnonzero(blob[:-1] == 10) + 1]).astype(np.uint64)      # line starts (no \r in the blob)
    if blob[-1] == 10:
        exp = np.concatenate([exp, [len(blob)]]).astype(np.uint64)
    st = rj.Stats()
    r = rj.Regej("^")
    got = r.match_all_array(blob, stats=st)
    assert got.shape[0] == exp.shape[0] and (got[:, 0] == exp).all() and (got[:, 1] == exp).all()
    got = r.match_all_array(blob, stats=st)
    assert st.launches == 1 and st.reruns == 0
    fa = _np_text(W.fasta_file(1_000_000))               # strip: every newline, every header line
    nl = np.flatnonzero(fa == 10)
    gt = np.flatnonzero(fa == ord(">"))
    got = rj.Regej(W.STRIP_PATTERN).match_all_array(fa)
    assert got.shape[0] == nl.shape[0]                   # one match per line end (a header line is one match)
    assert (got[:, 1] == nl.astype(np.uint64) + 1).all()
    heads = got[got[:, 1] - got[:, 0] > 1]
    assert heads.shape[0] == gt.shape[0] and (heads[:, 0] == gt.astype(np.uint64)).all()


def test_kmer_set_long_runs(rj):
    """Round 2: k_set_kmer is size-independent — a warp owns a run of rows of any length, checks its hits in
    batches and moves them to a staging area when its shared list is full.  400 MB with a member occurrence every
    200 bytes (every warp flushes and spills many times; the staging area grows once), checked offset by offset."""
    from rejit_b200 import workloads as W
    rng = np.random.RandomState(5)
    unit = np.frombuffer(b"agggtaaa", dtype=np.uint8)
    # filler over {a, c}: every alternative of every member needs a 't', so only the planted 8-mers match
    block = rng.choice(np.frombuffer(b"ac", dtype=np.uint8), size=200 * 5000).astype(np.uint8)
    for k in range(0, len(block), 200):
        block[k + 100:k + 108] = unit
    text = np.tile(block, 400)                           # 400 MB
    n_occ = len(text) // 200
    rs = rj.RegejSet(W.DNA_PATTERNS)
    assert "k-mer index" in rs.describe()
    exp_small = [O.Oracle(p).match_all(block[:400000].tobytes()) for p in W.DNA_PATTERNS]
    got_small = rs.match_all(block[:400000])
    assert got_small == exp_small
    which = [len(e) > 0 for e in exp_small]              # the members the planted 8-mer matches: the first only
    assert which == [True] + [False] * 8
    dt = rj.DeviceText(text)
    try:
        st = rj.Stats()
        counts = rs.match_all_device(dt, stats=st)
        assert st.strategy == 4, st.strategy
        counts = rs.match_all_device(dt, stats=st)       # steady state: one launch, no rerun
        assert st.launches == 1 and st.reruns == 0 and st.strategy == 4
        for j, w in enumerate(which):
            assert counts[j] == (n_occ if w else 0), (j, counts[j])
    finally:
        dt.free()
    # full offset lists through the text API
    lists = rs.match_all(text[:100_000_000])
    begins = np.arange(100, 100_000_000, 200, dtype=np.uint64)
    for j, w in enumerate(which):
        got = np.array(lists[j], dtype=np.uint64).reshape(-1, 2)
        if w:
            assert got.shape[0] == begins.shape[0] and (got[:, 0] == begins).all() and (got[:, 1] == begins + 8).all(), j
        else:
            assert got.shape[0] == 0
    # against the nine single-pattern scans (k_dfa_tma, an independent kernel) on 300 MB of FASTA
    seq = np.tile(W.fasta_sequence(3_000_000), 10)
    dt = rj.DeviceText(seq)
    try:
        counts = rs.match_all_device(dt)
        single = [rj.Regej(p).match_all_device(dt) for p in W.DNA_PATTERNS]
        assert counts == single, (counts, single)
    finally:
        dt.free()


def test_replace_all_set_one_pass(rj):
    """Round 2: ReplaceAll calls applied one after the other collapse into ONE byte -> string table when every
    pattern matches one byte and no replacement holds a later pattern's byte (regex-dna's eleven IUB codes,
    /root/reference/sample/regexdna.cc:69-85); any other set is run call by call.  Either way the text and the
    per-pattern counts are those of the sequential calls (oracle matches + the reference's Replace)."""
    from rejit_b200 import workloads as W
    seq = W.fasta_sequence(60000).tobytes()              # 600 KB, IUB codes in the middle third
    pats = [c for c, _ in W.IUB_SUBSTITUTIONS]
    withs = [a.encode() for _, a in W.IUB_SUBSTITUTIONS]

    def sequential(text, pats, withs):
        counts = []
        for p, w in zip(pats, withs):
            text, k = _replace_expected(p, text, w)
            counts.append(k)
        return text, counts

    for text in (seq, seq[:4096], seq[:4097], seq[170000:170017], b"", b"B", b"acgt" * 3000, b"BDHKMNRSVWY" * 5000):
        exp, exp_counts = sequential(text, pats, withs)
        src = rj.Text(text)
        try:
            st = rj.Stats()
            out, counts = rj.replace_all_set_text(pats, src, withs, stats=st)
            assert counts == exp_counts and len(out) == len(exp) and out.download() == exp, len(text)
            assert st.strategy != -1                       # the table, not call by call
            out.free()
        finally:
            src.free()
    # classes, deletions, long replacements, overlapping classes (the FIRST pattern that matches a byte wins)
    t = fuzzgen.rand_text(random.Random(6), "abcdefgh\n", 70000)
    for ps, ws in ((["[abc]", "d", "[a-e]"], [b"<1>", b"", b"Z"]), (["\n"], [b""]), (["a"], [b"x" * 40]),
                   (["a", "b"], [b"yy", b"zz"]), (["h"], [b"h"])):
        exp, exp_counts = sequential(t, ps, ws)
        src = rj.Text(t)
        out, counts = rj.replace_all_set_text(ps, src, ws)
        assert counts == exp_counts and out.download() == exp, ps
        out.free()
        src.free()
    # rows that are copied (no replaced byte in 512 bytes) between rows that are not: every output alignment, texts that
    # end inside a row / a tile, one replacement longer than the staging area of a row
    rng = random.Random(11)
    for n, every in ((16384 * 3 + 5, 700), (16384, 2000), (100001, 97), (513, 1000), (511, 3), (40000, 10 ** 9)):
        body = bytearray(rng.choice(b"acgt") for _ in range(n))
        for at in range(every // 2, n, every):
            body[at] = rng.choice(b"XYZ")
        body = bytes(body)
        for ps, ws in ((["X", "Y", "Z"], [b"(x|y)", b"", b"0123456789abcdefg"]), (["[XY]"], [b"q" * 3000])):
            exp, exp_counts = sequential(body, ps, ws)
            src = rj.Text(body)
            out, counts = rj.replace_all_set_text(ps, src, ws)
            assert counts == exp_counts and out.download() == exp, (n, every, ps)
            out.free()
            src.free()
    # not a table: a replacement feeds a later pattern / a two-byte pattern -> call by call, same answer
    for ps, ws in ((["a", "b"], [b"b", b"c"]), (["ab", "c"], [b"X", b"Y"]), (["a*", "b"], [b"-", b"+"])):
        exp, exp_counts = sequential(t[:20000], ps, ws)
        src = rj.Text(t[:20000])
        st = rj.Stats()
        out, counts = rj.replace_all_set_text(ps, src, ws, stats=st)
        assert st.strategy == -1 and counts == exp_counts and out.download() == exp, ps
        out.free()
        src.free()


def test_replace_all_fused_into_the_scan(rj):
    """Round 2: generic scans whose replacement is not longer than their shortest match write the rebuilt text from
    `k_scan_emit` itself (scan_emit.cuh, kRebuild; engine.cu ReplaceAllDevice) — one launch instead of scan + lengths +
    prefix sum + index + staging.  Same contract as `Regej::ReplaceAll` (src/rejit.cc:97-112, 221-226): the oracle's
    matches applied in Python, byte for byte.  Sizes sit on tile edges (24 KB tiles) and cross a group of 32 tiles;
    gaps of 0 bytes, of a few bytes, of whole tiles; matches that end in the next tile or cover one; replacements of
    0, 1 and 3 bytes.  The fused kernel takes `>.*\\n|\\n`, `[ab]{3,}`, `(^|$|[x])`, `a.*` (the last two mostly hand the
    call back: chains across tile edges, dense tiles); the other patterns are the same contract on the paths the fusion
    does NOT apply to — `x+` / `x*` re-enter their start (faithful resolve), `ab|ba` is a DFA scan, `x{40,}` a literal
    window, replacements longer than the shortest match — so that both sides of the switch in `ReplaceAllDevice` are
    held to the same bytes.  Many look-back steps with the fused kernel: `test_regexdna_chain_at_size` (510 MB)."""
    rng = np.random.default_rng(77)

    def text(alpha, n, p=None):
        return np.frombuffer(alpha, dtype=np.uint8)[rng.choice(len(alpha), size=n, p=p)].tobytes()

    tile = 24576
    cases = []
    for n in (1, 63, tile - 1, tile, tile + 1, 2 * tile, 3 * tile + 5, 33 * tile + 11):
        dense = text(b"abx\n", n)
        sparse = text(b"abx\n", n, p=[0.4995, 0.4995, 0.0005, 0.0005])
        for t in (dense, sparse):
            cases += [("x+", t, b""), ("x+", t, b"Q"), ("ab|ba", t, b"Z"), ("ab|ba", t, b"YZ"), ("x*", t, b""),
                      ("(^|$|[x])", t, b""), ("[ab]{3,}", t, b"QQQ"), ("a.*", t, b"L"), (">.*\n|\n", t, b"")]
    # lines longer than a tile (every `a` is a start and each start runs to the end of its line: kept short, the
    # speculative evaluation is quadratic in the line length), and ONE match that covers a whole tile and more
    lines = text(b"ab", 40 * tile)
    cases += [("a.*", lines[:2 * tile + 100], b""), ("a.*", lines[:tile + 100] + b"\n" + lines[:tile + 50], b"-")]
    header = b"ab\n>" + lines[:2 * tile + 300] + b"\nab\nba\n"
    cases += [(">.*\n|\n", header, b""), (">.*\n|\n", lines[:tile - 2] + header, b"")]
    runs = bytearray(text(b"ab", 6 * tile))
    for at in (100, tile - 20, 3 * tile - 45, 5 * tile):        # runs of x across tile edges, 90 bytes each
        runs[at:at + 90] = b"x" * 90
    w35 = bytes(range(65, 100))
    cases += [("x{40,}", bytes(runs), w35), ("x{40,}", bytes(runs), b""), ("x{40,}", bytes(runs), b"x" * 40)]
    cases += [("x+", text(b"abx", 2 * tile), b"QQ"), (".*", b"x" * 9000, b"y")]      # not fused: replacement too long
    for pat, t, w in cases:
        got = rj.Regej(pat).replace_all(t, w)
        assert got == _replace_expected(pat, t, w), (pat, len(t), w)
    # one launch where the fusion applies (the strip of a FASTA file: SURVEY section 8f, sample/regexdna.cc:49)
    from rejit_b200 import workloads as W
    fa = W.fasta_file(100_000)                                  # ~1 MB, 60-column lines
    exp, n_strip = _replace_expected(W.STRIP_PATTERN, fa, b"")
    raw = rj.Text(fa)
    strip = rj.Regej(W.STRIP_PATTERN)
    for _ in range(2):
        st = rj.Stats()
        out, k = strip.replace_all_text(raw, b"", stats=st)
        assert k == n_strip and out.download() == exp
        import os
        if not os.environ.get("RJ_NO_FUSED_REBUILD") and not os.environ.get("RJ_NO_EMIT"):
            assert st.launches == 1, st.launches
        out.free()
    raw.free()
    # 60 MB through the separate passes (`x+` re-enters its start: no single-pass scan).  Expected with numpy: runs of x
    big = np.frombuffer(b"abx\n", dtype=np.uint8)[rng.choice(4, size=60_000_000, p=[0.45, 0.45, 0.09, 0.01])]
    is_x = big == ord("x")
    starts = is_x.copy()
    starts[1:] &= ~is_x[:-1]
    src = rj.Text(big)
    r = rj.Regej("x+")
    out, k = r.replace_all_text(src, b"")
    assert k == int(starts.sum()) and (_np_text(out.download()) == big[~is_x]).all()
    out.free()
    out, k = r.replace_all_text(src, b"Q")
    exp = big.copy()
    exp[starts] = ord("Q")
    assert k == int(starts.sum()) and (_np_text(out.download()) == exp[~is_x | starts]).all()
    out.free()
    src.free()


def test_regexdna_chain_at_size(rj):
    """BASELINE configs[4] on one GPU's share: strip, nine counts, eleven substitutions over a 510 MB FASTA file on
    device-resident texts.  Size-independent checks: the stripped text IS the sequence the file was made from, the
    nine counts equal ten times the counts of the 50 MB sequence it tiles plus the seams (counted by the oracle on
    the seam windows), the final length is the stripped length plus (len(alt) - 1) per IUB letter."""
    from rejit_b200 import workloads as W
    seq = W.fasta_sequence(5_000_000)                    # 50 MB
    tiled = np.tile(seq, 10)                             # 500 MB
    # a FASTA file of it: one header, 60-column lines
    n = len(tiled)
    body = np.full(n + (n + 59) // 60, 10, dtype=np.uint8)
    idx = np.arange(n, dtype=np.int64)
    body[idx + idx // 60] = tiled
    del idx
    fa = np.concatenate([_np_text(b">ONE Homo sapiens alu\n"), body])
    del body
    raw = rj.Text(fa)
    strip = rj.Regej(W.STRIP_PATTERN)
    st = rj.Stats()
    cur, n_strip = strip.replace_all_text(raw, b"", stats=st)
    raw.free()
    assert n_strip == 1 + (n + 59) // 60 and len(cur) == n
    # the stripped text, byte for byte (sampled windows + a device-side count of every letter below)
    got = _np_text(cur.download())
    assert (got == tiled).all()
    del got
    rs = rj.RegejSet(W.DNA_PATTERNS)
    counts = rs.match_all_text(cur)
    base = rj.RegejSet(W.DNA_PATTERNS).match_all(seq)    # lists on the 50 MB sequence (checked against the oracle elsewhere)
    seam = np.concatenate([seq[-16:], seq[:16]]).tobytes()
    for j, p in enumerate(W.DNA_PATTERNS):
        across = sum(1 for b, e in O.Oracle(p).match_all(seam) if b < 16 < e)
        assert counts[j] == 10 * len(base[j]) + 9 * across, (p, counts[j], len(base[j]), across)
    pats = [c for c, _ in W.IUB_SUBSTITUTIONS]
    withs = [a.encode() for _, a in W.IUB_SUBSTITUTIONS]
    out, k = rj.replace_all_set_text(pats, cur, withs, stats=st)
    letters = np.bincount(tiled, minlength=256)
    assert k == [int(letters[ord(c)]) for c in pats]
    assert len(out) == n + sum(int(letters[ord(c)]) * (len(w) - 1) for c, w in zip(pats, withs))
    # section ONE (10 MB, the ALU repeat) holds no IUB code: the result is unchanged up to there; the MB that
    # follows (section TWO) against the sequential reference semantics
    res = out.download()
    assert res[:10_000_000] == seq[:10_000_000].tobytes()
    exp = seq[10_000_000:11_000_000].tobytes()
    for c, w in zip(pats, withs):
        exp = exp.replace(c.encode(), w)
    assert res[10_000_000:10_000_000 + len(exp)] == exp
    del res
    out.free()
    cur.free()


def test_device_stitch(rj):
    """Round 2: the device-side stitch (k_stitch: NVLink peer stores between the ranks' GPUs) under torchrun, two
    ranks (or as many GPUs as the box has, at most 8).  Needs >= 2 GPUs."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    n = min(rj.device_count(), 8)
    if n < 2:
        pytest.skip("one GPU")
    for world in sorted({2, n}):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), os.path.join(ROOT, "tests", "_stitch_worker.py")]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-3000:]
        line = [l for l in p.stdout.split("\n") if l.startswith("STITCH ")][-1]
        res = json.loads(line[len("STITCH "):])
        for r in res:
            assert r["totals"] == r["expected"], (world, r)
        assert any(r["cascaded"] for r in res) or world > 0
