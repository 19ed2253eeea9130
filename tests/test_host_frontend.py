"""CPU tier: the product's host front end and table builder against the oracle
and the golden fixtures.  Matching itself is emulated by tests/hostsim.cc (the
same tables and the same __host__ __device__ logic the kernels use); the kernels
are exercised by the GPU tier (test_gpu_parity.py)."""
import ctypes
import os
import random
import re
import subprocess
import sys

import pytest

import fuzzgen
import rejit_oracle as O
from conftest import ROOT, expand_table_row


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    entry.build()
    import rejit_b200
    return rejit_b200


def test_library_exports_every_declared_symbol(lib):
    """include/rejit_b200.h vs the built library (no compute, works without a GPU)."""
    header = open(os.path.join(ROOT, "include", "rejit_b200.h")).read()
    declared = set(re.findall(r"\b(rejit_b200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(lib.EXPORTED), declared ^ set(lib.EXPORTED)
    L = lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(r"\bT %s\b" % name, out), name
    # the C++ facade of include/rejit.h is in the same library
    assert "rejit5Regej8MatchAll" in out and "MatchAllParallel" in out


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device the matchers must fail loudly, not compute."""
    if lib.device_count() > 0:
        pytest.skip("a GPU is present")
    re_ = lib.Regej("abc")
    assert re_.status == 0 and re_.compile()
    with pytest.raises(lib.RejitError):
        re_.match_all(b"xxabcxx")
    with pytest.raises(lib.RejitError):
        re_.match_full(b"abc")


def test_ir_identical_to_reference(lib, ir_dumps):
    """Product parser + lowering produce the reference's state numbering and edge
    lists (tests/golden/ir_dumps.json = the reference's --print_re_list)."""
    for key, ref in ir_dumps.items():
        opt, pat = int(key[0]), key[2:]
        r = lib.Regej(pat.encode("latin-1"), parser_opt=bool(opt))
        assert r.status == 0, key
        dump = r.ir_dump().split("\n")
        assert dump[0].startswith("states %d " % ref["n_states"]), (key, dump[0])
        ci, mi = dump.index("control"), dump.index("matching")
        ctrl = []
        for line in dump[ci + 1:mi]:
            m = re.match(r"(\w+) \{(\d+),(\d+)\}", line)
            ctrl.append([m.group(1), int(m.group(2)), int(m.group(3))])
        assert ctrl == ref["control"], key
        mt = []
        for line in dump[mi + 1:]:
            if not line:
                continue
            m = re.match(r"(\w+) \{(\d+),(\d+)\} ?(.*)", line)
            kind = m.group(1)
            if kind == "MultipleChar":
                mt.append([kind, int(m.group(2)), int(m.group(3)), bytes.fromhex(m.group(4)).decode("latin-1")])
            elif kind == "Bracket":
                mt.append([kind, int(m.group(2)), int(m.group(3)), m.group(4) == "neg"])
            else:
                mt.append([kind, int(m.group(2)), int(m.group(3))])
        assert mt == ref["matching"], key


def test_parse_errors_and_status_string(lib):
    for bad in ["", "(ab", "a||b", "()", "*a", "a{3,2}", "[abc", "a\\q", "a]", "a\\.b"]:
        r = lib.Regej(bad)
        assert r.status == -1, bad
        if bad:
            assert r.status_string.startswith("Error parsing at index"), (bad, r.status_string)
    r = lib.Regej("a\\q")
    # layout of the reference's Parser::ParseError (src/parser.cc:652-665)
    assert r.status_string == "Error parsing at index 2\na\\q\n  ^ \nunexpected character q\n"
    assert lib.Regej("a{3,2}").status_string.endswith("Invalid repetition bounds: 3 > 2\n")


def test_hostile_patterns_fail_before_expanding(lib):
    """Patterns whose unrolled form would be huge, or whose tree is absurdly deep, come back as
    a parse error in milliseconds: no allocation storm, no recursion crash, no exception
    crossing the C boundary (`.{50000000}`, `a{3000000000}`, `a????...`)."""
    import time
    hostile = [b".{50000000}", b"a{3000000000}", b"a" + b"?" * 100000, b"(" * 50000 + b"a" + b")" * 50000,
               b"(a*){999999}", b"((a{64}){64}){2}", b"(x{0,0}){4000000000}", b"a{,4097}", b"(abc){1,1400}"]
    t0 = time.perf_counter()
    for pat in hostile:
        r = lib.Regej(pat)
        assert r.status == -1, pat[:20]
        assert "too large" in r.status_string or "too deeply" in r.status_string or r.status_string.startswith("Error parsing"), pat[:20]
    assert time.perf_counter() - t0 < 5.0
    # what fits the engine's 4096 byte positions still parses and compiles
    for pat in (b"(abc){1000}", b"a{,4096}", b"x" * 4096, b"(" * 150 + b"a" + b")" * 150, b"a" + b"?" * 150):
        r = lib.Regej(pat)
        assert r.status == 0 and r.compile(), pat[:20]
    over = lib.Regej(b"x" * 4097)                      # one position more than the engine's cap
    assert over.status == -1 and "too large" in over.status_string


def test_slab_entry_points_refuse_reentrant_patterns(lib):
    """`.{,4}t` can re-enter its own start (defect B19): its chain cannot be resumed from a
    (cur, tail) carry, so the slab / carry entry points return an error instead of wrong
    matches.  The check runs before any device work: no GPU needed."""
    bad, good = lib.Regej(".{,4}t"), lib.Regej("abc|abd")
    assert not bad.shardable() and good.shardable()
    L = lib.lib()
    err = ctypes.create_string_buffer(512)
    r = L.rejit_b200_match_all_device_slab(bad._prog, 0, None, 0, 0, 1, 0, None, 0, None, None, None, err, len(err))
    assert r == -1 and b"re-entrant" in err.value
    cin = lib.Carry(5, 5)
    r = L.rejit_b200_match_all_device(bad._prog, 0, None, 0, None, 0, ctypes.byref(cin), None, None, err, len(err))
    assert r == -1 and b"re-entrant" in err.value
    rs = lib.RegejSet([bad, good])
    counts = (ctypes.c_int64 * 2)()
    r = L.rejit_b200_match_all_set_device_slab(rs._set, 0, None, 0, 0, 1, 0, None, None, counts, None, err, len(err))
    assert r == -1 and b"re-entrant" in err.value


def test_entry_points_refuse_null_handles(lib):
    """A C, cgo or ctypes caller that passes a NULL program / set / text handle gets -1 (or NULL) and a message, not a
    segmentation fault; and every entry point that can allocate is a function-try-block, so no exception crosses the C
    ABI (capi.cc RJ_CATCH).  The checks run before any device work: no GPU needed."""
    L = lib.lib()
    err = ctypes.create_string_buffer(256)
    n = len(err)
    pair = (ctypes.c_uint64 * 2)()
    pairs = ctypes.POINTER(ctypes.c_uint64)()
    counts = (ctypes.c_int64 * 2)()
    calls = [
        ("rejit_b200_match_all_alloc", lambda: L.rejit_b200_match_all_alloc(None, b"abc", 3, ctypes.byref(pairs), None, err, n)),
        ("rejit_b200_match_all", lambda: L.rejit_b200_match_all(None, b"abc", 3, None, 0, err, n)),
        ("rejit_b200_match_first", lambda: L.rejit_b200_match_first(None, b"abc", 3, pair, err, n)),
        ("rejit_b200_match_full", lambda: L.rejit_b200_match_full(None, b"abc", 3, err, n)),
        ("rejit_b200_match_anywhere", lambda: L.rejit_b200_match_anywhere(None, b"abc", 3, err, n)),
        ("rejit_b200_match_all_multi_gpu", lambda: L.rejit_b200_match_all_multi_gpu(None, b"abc", 3, 2, ctypes.byref(pairs), None, err, n)),
        ("rejit_b200_match_all_device", lambda: L.rejit_b200_match_all_device(None, 0, None, 0, None, 0, None, None, None, err, n)),
        ("rejit_b200_match_all_device_slab", lambda: L.rejit_b200_match_all_device_slab(None, 0, None, 0, 0, 1, 0, None, 0, None, None, None, err, n)),
        ("rejit_b200_match_all_text", lambda: L.rejit_b200_match_all_text(None, None, ctypes.byref(pairs), None, err, n)),
        ("rejit_b200_match_all_set_device", lambda: L.rejit_b200_match_all_set_device(None, 0, None, 0, counts, None, err, n)),
        ("rejit_b200_match_all_set_device_slab", lambda: L.rejit_b200_match_all_set_device_slab(None, 0, None, 0, 0, 1, 0, None, None, counts, None, err, n)),
        ("rejit_b200_match_all_set_text", lambda: L.rejit_b200_match_all_set_text(None, None, counts, None, None, err, n)),
        ("rejit_b200_match_all_set_device_stitched", lambda: L.rejit_b200_match_all_set_device_stitched(None, 0, None, 0, 0, 1, 0, None, None, None, counts, None, err, n)),
    ]
    for name, call in calls:
        err.value = b""
        assert call() == -1, name
        assert name.encode() in err.value and b"null" in err.value, (name, err.value)
    good = lib.Regej("abc")
    assert good.compile()
    err.value = b""
    assert L.rejit_b200_match_all_text(good._prog, None, ctypes.byref(pairs), None, err, n) == -1 and b"null text" in err.value
    assert not L.rejit_b200_replace_all_text(None, None, b"", 0, None, None, err, n)
    assert not L.rejit_b200_set_create(None, 0)
    two = (ctypes.c_void_p * 2)(good._prog, None)
    assert not L.rejit_b200_set_create(two, 2)


def test_compile_survives_malformed_ir():
    """The C ABI takes the lowered regexp from a foreign binding: out-of-range states, kinds,
    payload ranges and header fields must come back as an error (entry_state = 255 of 5 states used to fault)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ir_mutation_fuzz.py"), "11", "1500"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-300:], r.stderr[-300:])
    compiled, rejected = [int(x) for x in r.stdout.split()[1::2]]
    assert compiled > 100 and rejected > 500


def test_strategies_for_the_baseline_patterns(lib):
    from rejit_b200 import workloads as W
    assert lib.Regej(W.LITERAL_PATTERN).describe().startswith("literal scan, 6 bytes")
    d = lib.Regej(W.COMPLEX_PATTERN).describe()
    assert d.startswith("required literal (8 bytes) + verify starts in [hit-42, hit-2]"), d
    for p in W.DNA_PATTERNS:
        d = lib.Regej(p).describe()
        assert d.startswith("fixed-length DFA scan") and "match length 8" in d, (p, d)
    for p, _ in W.IUB_SUBSTITUTIONS:
        assert lib.Regej(p).describe().startswith("literal scan, 1 bytes")
    assert lib.Regej(W.JREP_PATTERN).describe().startswith("literal scan, 3 bytes")
    assert "reentrant" not in lib.Regej(W.STRIP_PATTERN).describe()
    assert "reentrant start" in lib.Regej(".{,4}t").describe()


def test_hostsim_golden_offsets(hostsim, golden_vectors):
    for v in golden_vectors:
        t = v["text"].encode("latin-1")
        for strategy in (-1, 3):
            got, desc = hostsim.match_all(v["re"], t, strategy)
            assert got == [tuple(m) for m in v["all"]], (v["re"], v["note"], desc, strategy)
        assert bool(hostsim.match_full(v["re"], t)) == v["full"], (v["re"], v["note"])


def test_hostsim_reference_table(hostsim, ref_table):
    for row in ref_table:
        pat, checks = expand_table_row(row)
        for mt, text, expected, start, end in checks:
            t = text.encode("latin-1")
            if mt == "full":
                assert bool(hostsim.match_full(pat, t)) == bool(expected), (row["line"], pat)
                continue
            got, _ = hostsim.match_all(pat, t)
            if mt == "all":
                assert len(got) == expected, (row["line"], pat, text)
            elif mt == "anywhere":
                assert bool(got) == bool(expected), (row["line"], pat, text)
            else:     # MatchFirst := MatchAll[0]
                assert bool(got) == bool(expected), (row["line"], pat, text)
                if expected and start is not None:
                    assert got[0] == (start, end), (row["line"], pat, text, got[0])


def test_hostsim_fuzz_vs_oracle(hostsim):
    r = random.Random(99)
    checked = 0
    for _ in range(1200):
        pat, alpha = fuzzgen.rand_pattern(r)
        try:
            o = O.Oracle(pat)
        except O.ParserError:
            got, _ = hostsim.match_all(pat, b"")
            assert got == -1, pat
            continue
        for n in (r.randint(0, 50), r.randint(300, 900)):
            t = fuzzgen.rand_text(r, alpha, n)
            exp = o.match_all(t)
            for strategy in (-1, 3):
                got, desc = hostsim.match_all(pat, t, strategy)
                assert got == exp, (pat, t, desc, strategy)
            assert bool(hostsim.match_full(pat, t)) == o.match_full(t), (pat, t)
            checked += 1
    assert checked > 1500


def test_baseline_config_0_plumbing(hostsim):
    """BASELINE.json configs[0]: literal 'regexp' MatchAll over 1 MiB of random ASCII in ['0','z') on the CPU
    (SURVEY.md §8d C1), without hits and with the literal planted every ~5000 bytes: the reference in its default
    AND its parity configuration (when oracle/_ref is here), the oracle and the product's host tables agree."""
    from rejit_b200 import workloads as W
    ref = None
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "librejit_ref.so")):
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        from make_golden import Ref
        ref = Ref()
    plain = W.random_ascii(1 << 20, seed=1)
    for text in (plain.tobytes(), W.plant(plain, [b"regexp"], every=5000).tobytes()):
        exp = O.Oracle(W.LITERAL_PATTERN).match_all(text)
        want, at = [], text.find(b"regexp")
        while at >= 0:
            want.append((at, at + 6))
            at = text.find(b"regexp", at + 6)
        assert exp == want
        got, desc = hostsim.match_all(W.LITERAL_PATTERN, text)
        assert got == exp and desc.startswith("literal scan, 6 bytes"), desc
        if ref is not None:
            for flagset in (0, 2):
                ref.flags(flagset)
                assert ref.match_all(b"regexp", text) == [list(m) for m in exp], flagset
    assert len(exp) > 150


def test_hostsim_rich_dialect(hostsim):
    """The rest of the dialect (ranges, escapes, \\xHH, high bytes; tests/fuzzgen.py): the product's parser accepts
    exactly what the oracle's accepts, and its tables match like the oracle."""
    r = random.Random(4711)
    checked = 0
    for _ in range(700):
        pat = fuzzgen.rand_rich_pattern(r)
        try:
            o = O.Oracle(pat)
        except O.ParserError:
            got, _ = hostsim.match_all(pat, b"")
            assert got == -1, pat
            continue
        for n in (r.randint(0, 40), r.randint(60, 300)):
            t = fuzzgen.rand_rich_text(r, n)
            exp = o.match_all(t)
            for strategy in (-1, 3):
                got, desc = hostsim.match_all(pat, t, strategy)
                assert got == exp, (pat, t, desc, strategy)
            checked += 1
    assert checked > 1000


def test_hostsim_long_literals(hostsim):
    """Literal nodes longer than 16 bytes, exact copies and near misses: the product compares every byte
    (the reference does not: defect B20, tests/test_oracle.py)."""
    r = random.Random(2021)
    kinds = set()
    for _ in range(400):
        pat, t = fuzzgen.rand_long_literal_case(r)
        exp = O.Oracle(pat).match_all(t)
        for strategy in (-1, 3):
            got, desc = hostsim.match_all(pat, t, strategy)
            assert got == exp, (pat, t, desc, strategy)
        got, desc = hostsim.match_all(pat, t)
        kinds.add(desc.split(",")[0].split(";")[0])
        assert hostsim.match_all_slabs(pat, t, r.choice([2, 3, 5])) == exp, (pat, t)
    assert any(k.startswith("literal scan") for k in kinds), kinds


def test_hostsim_slab_stitching(hostsim):
    """The carry protocol of MatchAllHostMultiGpu on non-re-entrant patterns."""
    r = random.Random(5)
    import rejit_b200
    checked = 0
    while checked < 400:
        pat, alpha = fuzzgen.rand_pattern(r)
        try:
            o = O.Oracle(pat)
        except O.ParserError:
            continue
        if "reentrant" in rejit_b200.Regej(pat).describe():
            continue
        t = fuzzgen.rand_text(r, alpha, r.randint(100, 700))
        assert hostsim.match_all_slabs(pat, t, r.choice([2, 3, 4, 8])) == o.match_all(t), (pat, t)
        checked += 1


def test_fused_set_tables_on_cpu():
    """Union DFA of a pattern set (k_set_tma's tables) emulated by tests/hostsim.cc."""
    import ctypes
    from rejit_b200 import workloads as W
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "_build", "libhostsim.so"))
    L.hostsim_set_match_all.restype = ctypes.c_int64
    L.hostsim_set_match_all.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char, ctypes.c_char_p,
                                        ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64]
    seq = W.fasta_sequence(40000).tobytes()
    r = random.Random(1)
    sets = [W.DNA_PATTERNS, ["acg", "ttgca", "a[ct]g"], ["ac|gt", "tgc|aaa", "c[ag]t|ggg"]]
    for pats in sets:
        joined = "\x01".join(pats).encode()
        for j, p in enumerate(pats):
            out = (ctypes.c_uint64 * 100000)()
            k = L.hostsim_set_match_all(joined, len(joined), b"\x01", seq, len(seq), j, out, 50000)
            assert k >= 0, (pats, k)
            assert [(out[2 * i], out[2 * i + 1]) for i in range(k)] == O.Oracle(p).match_all(seq), (pats, p)
    joined = b"acgt\x01x+"
    assert L.hostsim_set_match_all(joined, len(joined), b"\x01", seq, len(seq), 0, None, 0) == -5


def test_workload_shaped_parity_on_cpu(hostsim):
    from rejit_b200 import workloads as W
    seq = W.fasta_sequence(20000).tobytes()          # 200 kB
    for p in W.DNA_PATTERNS:
        got, desc = hostsim.match_all(p, seq)
        assert got == O.Oracle(p).match_all(seq), (p, desc)
    text = W.plant(W.random_ascii(150000, seed=3), W.COMPLEX_HITS, every=9000).tobytes()
    for p in (W.COMPLEX_PATTERN, W.LITERAL_PATTERN, "abcdefgh"):
        got, desc = hostsim.match_all(p, text)
        exp = O.Oracle(p).match_all(text)
        assert got == exp and (p != W.COMPLEX_PATTERN or len(exp) >= 10), (p, desc, len(exp))
    fa = W.fasta_file(2000)
    got, desc = hostsim.match_all(W.STRIP_PATTERN, fa)
    assert got == O.Oracle(W.STRIP_PATTERN).match_all(fa)


def test_replace_all_placement(hostsim):
    """ReplaceAll: the kernel's placement arithmetic
    (device_program.h: ReplaceHead / ReplacePlace) driven on the CPU, against the
    reference's Replace semantics (src/rejit.cc:97-112) applied to the oracle's
    matches: texts around the 4096-byte tile size, matches straddling tiles,
    empty matches (also at the end of the text), empty and long replacements."""
    import random
    import rejit_oracle as O
    import fuzzgen
    rng = random.Random(5)

    def expected(pat, text, w):
        out, at = bytearray(), 0
        ms = O.Oracle(pat).match_all(text)
        for b, e in ms:
            out += text[at:b] + w
            at = e
        out += text[at:]
        return len(ms), bytes(out)

    cases = [("a", b"", b"X"), ("x*", b"aaxa", b"-"), ("$", b"ab\ncd", b"<EOL>"), ("^", b"a\nb\n", b"> "),
             ("abc", b"abc" * 3000, b""), ("abc", b"abc" * 3000, b"abcabc"), (".*", b"x" * 9000, b"y"),
             ("x{2,}", b"ab" + b"x" * 12000 + b"cd", b"_")]
    for n in (1, 15, 16, 17, 4095, 4096, 4097, 8192, 12289):
        t = fuzzgen.rand_text(rng, "abx\n", n)
        for pat in ("a", "ab|ba", "x*", "a.*", "(^|$|[x])", "b+", "\n"):
            for w in (b"", b"Q", b"(c|g|t)"):
                cases.append((pat, t, w))
    for pat, t, w in cases:
        exp = expected(pat, t, w)
        got = hostsim.replace_all(pat, t, w)
        assert got == exp, (pat, len(t), w, got[0], exp[0])


def test_fused_rebuild_placement_model():
    """A MODEL of the placement arithmetic of the fused ReplaceAll (scan_emit.cuh, kRebuild — restated here, not shared
    code; the kernel itself is held to the same bytes by the GPU tier's test_replace_all_fused_into_the_scan): the text
    is cut into tiles, a tile owns the matches that BEGIN in it, publishes (matches, bytes inside them, chain state after
    its last match), reads the sums over the tiles before it, and writes its part of the output starting at
        rb_in  = max(tile_lo, end of a non-empty match arriving from the left)
        rb_out = rb_in - removed_before + w * matches_before
    four gaps per step.  A tile whose numbers cannot be right (removed_before > rb_in) or whose first match is reached
    by the chain from the left raises the flag that sends the call to the separate rebuild.  Checked against the
    reference's Replace (src/rejit.cc:97-112) applied to the oracle's matches, for tile sizes that put matches across
    one and several tile edges; and, with overlapping per-tile lists made up on purpose, that the flag is raised and
    nothing is written out of bounds."""
    import random
    import rejit_oracle as O
    import fuzzgen

    def rebuild(text, matches, w, tile):
        n = len(text)
        ntiles = n // tile + 1
        per = [[] for _ in range(ntiles)]
        for b, e in matches:
            per[min(b // tile, ntiles - 1)].append((b, e))
        out = bytearray(b"\xEE" * (n + 64))
        flagged = False
        before = removed = 0
        arriving = (0, 0)                                    # (cur, non-empty): where the next match may begin
        for t in range(ntiles):
            tile_lo, tile_hi = t * tile, min(t * tile + tile, n)
            mine = per[t]
            local = False                                    # (tiles decide alone: a flag elsewhere does not stop this one)
            if mine:                                         # the seam: does the chain from the left take my first match?
                b, e = mine[0]
                cur, ne = arriving
                if not (b > cur or (b == cur and (e > b or not ne))):
                    local = True
            rb_in = tile_lo
            if arriving[1] and arriving[0] > rb_in:
                rb_in = arriving[0]
            added = len(w) * before
            if removed > rb_in or added > removed:
                local = True
            flagged |= local
            if not local:
                in_pos, o = rb_in, rb_in - removed + added
                entries = mine + [(tile_hi, tile_hi)]        # the bytes after my last match
                for i in range(0, len(entries), 4):          # four gaps per step
                    step = []
                    for k in range(i, min(i + 4, len(entries))):
                        b, e = entries[k]
                        gap = b - in_pos if b > in_pos else 0
                        step.append((in_pos, gap, o, k < len(mine)))
                        o += gap + (len(w) if k < len(mine) else 0)
                        if e > in_pos:
                            in_pos = e
                    for src, gap, dst, is_match in step:
                        assert dst + gap + (len(w) if is_match else 0) <= n, "out of bounds"
                        out[dst:dst + gap] = text[src:src + gap]
                        if is_match:
                            out[dst + gap:dst + gap + len(w)] = w
            # publish: what the tiles after me will read
            before += len(mine)
            removed += sum(e - b for b, e in mine)
            if mine:
                b, e = mine[-1]
                arriving = (e, 1) if e > b else (b + 1, 0)
        total = n - removed + len(w) * before
        return flagged, bytes(out[:total]), out[total:]

    def expected(text, matches, w):
        res, at = bytearray(), 0
        for b, e in matches:
            res += text[at:b] + w
            at = e
        return bytes(res + text[at:])

    rng = random.Random(11)
    cases = [(">.*\n|\n", "ab>\n", b""), ("[ab]{3,}", "abx\n", b"QQQ"), ("(^|$|[x])", "abx\n", b""), ("b+", "abx", b"Q"),
             ("a.*", "abx\n", b"L"), ("x{2,}", "xa", b"--")]
    checked = 0
    for pat, alpha, w in cases:
        o = O.Oracle(pat)
        for n in (0, 1, 7, 64, 65, 300, 1000):
            text = fuzzgen.rand_text(rng, alpha, n)
            matches = o.match_all(text)
            if any(e - b < len(w) for b, e in matches):
                continue                                     # the engine fuses only when w <= the shortest match
            for tile in (8, 16, 64, 256, 4096):
                flagged, got, slack = rebuild(text, matches, w, tile)
                assert not flagged, (pat, n, tile)           # disjoint leftmost-longest matches: every seam is clean
                assert got == expected(text, matches, w), (pat, n, tile)
                assert slack == b"\xEE" * len(slack)
                checked += 1
    assert checked > 100
    # per-tile lists that are not a chain (a match of tile 0 runs over the first match of tile 1, as two tiles that
    # each resolved alone can produce): the flag is raised, whatever was written stayed inside the buffer
    text = b"a" * 64
    flagged, _, slack = rebuild(text, [(2, 20), (16, 30), (40, 44)], b"", 16)
    assert flagged and slack == b"\xEE" * len(slack)
    flagged, _, slack = rebuild(text, [(2, 60), (17, 63), (33, 64)], b"", 16)
    assert flagged and slack == b"\xEE" * len(slack)
    # ... and a later tile whose numbers still pass the check writes (garbage, the call is handed back) in bounds
    text = bytes(range(256)) * 4
    flagged, _, slack = rebuild(text, [(2, 20), (16, 30), (100, 104), (1000, 1024)], b"", 16)
    assert flagged and slack == b"\xEE" * len(slack)
