#!/usr/bin/env python3
"""tests/golden/make_golden.py — regenerates the committed golden fixtures.

Runs ONLY in the build container (needs /root/reference and oracle/_ref, built
by `make -C oracle ref`).  Nothing here is imported by the product.

Outputs (all next to this script):
  ref_test_table.json    the reference's own test table, tools/tests/test.cc:193-534,
                         transcribed as DATA (macro kind, line, regexp, text,
                         expected count/bool, expected MatchFirst [start,end)).
  matchall_offsets.json  MatchAll (begin,end) lists + MatchFirst/Full/Anywhere
                         results produced by the compiled reference with
                         use_fast_forward=0 ("noff") for every
                         table row and for an extra set of workload-shaped and
                         quirk-probing cases.  The reference's tests pin only
                         COUNTS for MatchAll; these pin the offsets.
  ir_dumps.json          the reference's lowered IR (--print_re_list) for every
                         distinct regexp above: state count and edge lists.
"""
import ctypes
import json
import os
import random
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_TEST = "/root/reference/tools/tests/test.cc"
REF_SO = os.path.join(ROOT, "oracle", "_ref", "librejit_ref.so")
REF_IR = os.path.join(ROOT, "oracle", "_ref", "ref_ir")


# ---------------------------------------------------------------- C literals
_ESC = {"n": "\n", "r": "\r", "t": "\t", "\\": "\\", '"': '"', "0": "\0", "'": "'"}


def _c_unescape(s: str) -> str:
    out, i = [], 0
    while i < len(s):
        if s[i] == "\\":
            c = s[i + 1]
            if c == "x":
                j = i + 2
                while j < len(s) and s[j] in "0123456789abcdefABCDEF":
                    j += 1
                out.append(chr(int(s[i + 2:j], 16)))
                i = j
                continue
            out.append(_ESC[c])
            i += 2
        else:
            out.append(s[i])
            i += 1
    return "".join(out)


_TOK = re.compile(r'\s*(?:(x100|x50|x10)\s*\(|("(?:[^"\\]|\\.)*")|(\))|(,)|(-?\d+)|(kMatch\w+))')


def _parse_args(argstr: str):
    """Evaluates the macro argument list: ints, kMatchX, and string expressions
    built from adjacent literals and the x10/x50/x100 repeat macros
    (tools/tests/test.cc:118-120)."""
    args, pos = [], 0
    cur = None          # current string value being concatenated
    stack = []          # (repeat, saved_cur)
    while pos < len(argstr):
        m = _TOK.match(argstr, pos)
        if not m:
            if argstr[pos:].strip() == "":
                break
            raise ValueError("cannot tokenise: " + argstr[pos:])
        pos = m.end()
        rep, lit, rpar, comma, num, ident = m.groups()
        if rep:
            stack.append((int(rep[1:]), cur))
            cur = None
        elif lit:
            cur = (cur or "") + _c_unescape(lit[1:-1])
        elif rpar:
            n, saved = stack.pop()
            cur = (saved or "") + (cur or "") * n
        elif comma:
            if stack:
                raise ValueError("comma inside repeat macro")
            args.append(cur)
            cur = None
        elif num is not None:
            cur = int(num)
        elif ident:
            cur = ident
    args.append(cur)
    return args


def parse_test_table():
    rows = []
    src = open(REF_TEST, encoding="latin-1").read().split("\n")
    for lineno, line in enumerate(src, 1):
        if not (193 <= lineno <= 534):
            continue
        s = line.strip()
        m = re.match(r"(TEST_Full|TEST_Multiple_unbound|TEST_Multiple|TEST)\((.*)\);\s*$", s)
        if not m:
            continue
        kind, a = m.group(1), _parse_args(m.group(2))
        if kind == "TEST_Full":
            rows.append({"kind": "full", "line": lineno, "expected": a[0], "re": a[1], "text": a[2]})
        elif kind in ("TEST_Multiple", "TEST_Multiple_unbound"):
            rows.append({"kind": "multiple_unbound" if kind.endswith("unbound") else "multiple",
                         "line": lineno, "expected": a[0], "re": a[1], "text": a[2],
                         "start": a[3], "end": a[4]})
        else:
            rows.append({"kind": "test", "line": lineno, "match_type": a[0],
                         "expected": a[1], "re": a[2], "text": a[3]})
    return rows


# ------------------------------------------------------------ reference calls
class Ref:
    def __init__(self):
        self.lib = ctypes.CDLL(REF_SO)
        L = self.lib
        L.ref_match_all.restype = ctypes.c_int64
        L.ref_match_all.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t,
                                    ctypes.POINTER(ctypes.c_uint64), ctypes.c_size_t]
        L.ref_match_first.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t,
                                      ctypes.POINTER(ctypes.c_uint64)]
        L.ref_match_full.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        L.ref_match_anywhere.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        L.ref_parse_status.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]

    def flags(self, flagset, parser_opt=1):
        self.lib.ref_set_flagset(flagset)
        self.lib.ref_set_parser_opt(parser_opt)

    def match_all(self, pat: bytes, text: bytes):
        cap = 4096
        while True:
            out = (ctypes.c_uint64 * (2 * cap))()
            n = self.lib.ref_match_all(pat, text, len(text), out, cap)
            if n < 0:
                return None
            if n <= cap:
                return [[out[2 * i], out[2 * i + 1]] for i in range(n)]
            cap = n

    def match_first(self, pat, text):
        out = (ctypes.c_uint64 * 2)()
        r = self.lib.ref_match_first(pat, text, len(text), out)
        return [out[0], out[1]] if r == 1 else None

    def match_full(self, pat, text):
        return self.lib.ref_match_full(pat, text, len(text)) == 1

    def match_anywhere(self, pat, text):
        return self.lib.ref_match_anywhere(pat, text, len(text)) == 1

    def parse_ok(self, pat):
        buf = ctypes.create_string_buffer(256)
        return self.lib.ref_parse_status(pat, buf, 256) == 0


# ------------------------------------------------------------ extra cases
DNA_PATTERNS = [    # sample/regexdna.cc:52-62 (nine variants)
    "agggtaaa|tttaccct", "[cgt]gggtaaa|tttaccc[acg]", "a[act]ggtaaa|tttacc[agt]t",
    "ag[act]gtaaa|tttac[agt]ct", "agg[act]taaa|ttta[agt]cct", "aggg[acg]aaa|ttt[cgt]ccct",
    "agggt[cgt]aa|tt[acg]accct", "agggta[cgt]a|t[acg]taccct", "agggtaa[cgt]|[acg]ttaccct"]
COMPLEX = "([complex]|(regexp)){2,7}abcdefgh(at|the|[e-nd]as well)"   # tools/benchmarks/run.py:351


def extra_cases():
    rnd = random.Random(20260925)
    cases = []

    def add(pat, text, note=""):
        cases.append({"re": pat, "text": text, "note": note})

    # vectors quoted in SURVEY.md §8a (captured from the noff reference)
    add("ab?c", "abbc abc ac", "B5: (ab)?c")
    add("ab{0,1}c", "abbc", "B4")
    add("ab{2,3}c", "abbbbc", "B4 via parser opt")
    add("x*", "aaxa", "empty-match rule")
    add("a*", "baaab")
    add(".*", "ab\ncd")
    add("$", "ab\r\ncd")
    add("^", "ab\r\ncd")
    add("(ab|ab[c][d][e][f]X|de)", "abcdefY", "B11")
    add("(ab|a)(bc|c)?", "abc ab ac abcbc")
    add("(a|b)*c", "ababc c abd bbbc")
    add("(a?){2}a{2}", "aaaa a aa aaa")
    add("[^a]{2}", "abcabbca", "Bracket::DeepCopy drops non_matching")
    add("\\D{3}", "ab1c23d456", "same quirk through \\D")
    add("\\x4A", "@J", "B6 hex letters decode 0-5")
    add(">.*\n|\n", ">ONE Homo\nacgt\nacgt\n>TWO x\nttt\n", "regex-dna strip")
    add("a\nb", "xa\nb a\nb\na\nba\nb", "jrep multi-line literal")
    # workload-shaped
    alpha = "acgt"
    for k, pat in enumerate(DNA_PATTERNS):
        t = "".join(rnd.choice(alpha) for _ in range(6000))
        # plant a few hits of each alternative shape
        for _ in range(12):
            pos = rnd.randrange(0, len(t) - 8)
            alt = rnd.choice(pat.split("|"))
            lit = re.sub(r"\[([a-z]+)\]", lambda m: rnd.choice(m.group(1)), alt)
            t = t[:pos] + lit + t[pos + len(lit):]
        add(pat, t, "regex-dna #%d" % (k + 1))
    for seed in range(3):
        t = "".join(chr(rnd.randrange(0x30, 0x7A)) for _ in range(5000))
        hits = ["ccregexpabcdefghthe", "omabcdefghdas well", "xcregexpregexpabcdefghat",
                "regexpregexpregexpregexpregexpregexpregexpregexpabcdefghthe", "cabcdefghat",
                "oxabcdefghthe"]
        for h in hits:
            pos = rnd.randrange(0, len(t) - len(h))
            t = t[:pos] + h + t[pos + len(h):]
        add(COMPLEX, t, "complex regex, random ['0','z') with planted hits")
        add("regexp", t, "literal over the same text")
        add("abcdefgh", t)
    for b in "BDHKMNRSVWY":                     # IUB singles, sample/regexdna.cc:69-85
        t = "".join(rnd.choice("acgtBDHKMNRSVWY") for _ in range(400))
        add(b, t, "IUB single")
    # overlapping literal chains
    add("aa", "aaaaaaa_aa_aaa")
    add("aba", "abababababa_aba")
    add("abcabc", "abcabcabcabcabc")
    return cases


# ------------------------------------------------------------ IR dumps
def ir_dump(pat: str, parser_opt: int):
    p = subprocess.run([REF_IR, pat, str(parser_opt)], capture_output=True)
    out = p.stdout.decode("latin-1")
    if "PARSE_ERROR" in out or p.returncode != 0:
        return None
    n_states = int(re.search(r"n_states : (\d+)", out).group(1))
    ctrl_txt = out.split("Control regexps list")[1].split("End of control regexp list")[0]
    match_txt = out.split("Matching regexps list")[1].split("End of matching regexp list")[0]
    ctrl = [[m.group(1), int(m.group(2)), int(m.group(3))]
            for m in re.finditer(r"Regexp \((\w+)\) \{(-?\d+), (-?\d+)\}", ctrl_txt)]
    matching = []
    for m in re.finditer(r"(?:MultipleChar \[((?:.|\n)*?)\] \{(-?\d+), (-?\d+)\}|"
                         r"Regexp \((Period)\) \{(-?\d+), (-?\d+)\}|"
                         r"Bracket (\(non_matching\) )?\[ \{(-?\d+), (-?\d+)\})", match_txt):
        if m.group(2) is not None:
            matching.append(["MultipleChar", int(m.group(2)), int(m.group(3)), m.group(1)])
        elif m.group(4):
            matching.append(["Period", int(m.group(5)), int(m.group(6))])
        else:
            matching.append(["Bracket", int(m.group(8)), int(m.group(9)), bool(m.group(7))])
    return {"n_states": n_states, "control": ctrl, "matching": matching}


def main():
    if not (os.path.exists(REF_SO) and os.path.exists(REF_IR) and os.path.exists(REF_TEST)):
        sys.exit("needs /root/reference and `make -C oracle ref`")
    ref = Ref()
    table = parse_test_table()
    json.dump(table, open(os.path.join(HERE, "ref_test_table.json"), "w"), indent=0)
    print("test table rows:", len(table))

    ref.flags(2)      # noff
    vectors = []
    seen = set()

    def record(pat, text, note):
        key = (pat, text)
        if key in seen:
            return
        seen.add(key)
        pb, tb = pat.encode("latin-1"), text.encode("latin-1")
        if not ref.parse_ok(pb):
            vectors.append({"re": pat, "text": text, "note": note, "parse_error": True})
            return
        vectors.append({"re": pat, "text": text, "note": note,
                        "all": ref.match_all(pb, tb), "first": ref.match_first(pb, tb),
                        "full": ref.match_full(pb, tb), "anywhere": ref.match_anywhere(pb, tb)})

    for row in table:
        record(row["re"], row["text"], "test.cc:%d" % row["line"])
        if row["kind"] == "multiple_unbound":   # alignment sweep, test.cc:687-700
            for i in (1, 7, 15, 16, 17, 31, 32):
                record(row["re"], " " * i + row["text"] + " " * (32 - i), "test.cc:%d align %d" % (row["line"], i))
    for c in extra_cases():
        record(c["re"], c["text"], c["note"])
    json.dump(vectors, open(os.path.join(HERE, "matchall_offsets.json"), "w"), indent=0)
    print("offset vectors:", len(vectors))

    dumps = {}
    for pat in sorted({v["re"] for v in vectors}):
        if "\0" in pat:
            continue
        for opt in (1, 0):
            d = ir_dump(pat, opt)
            if d is not None:
                dumps["%d:%s" % (opt, pat)] = d
    json.dump(dumps, open(os.path.join(HERE, "ir_dumps.json"), "w"), indent=0)
    print("ir dumps:", len(dumps))


if __name__ == "__main__":
    main()
