"""tests/fuzzgen.py — seeded random (pattern, text) generators shared by the
oracle-vs-reference and product-vs-oracle differential tests.

Patterns are drawn from the subset of ERE the reference accepts without
crashing (no empty alternatives, no unbalanced '(', no
leading repetition operator, only the escapes of src/parser.cc:53-117)."""
import random

ALPHABETS = {
    "ab": "ab",
    "abc": "abcx",
    "dna": "acgt",
    "nl": "ab\n",
    "crlf": "a\n\rb",
}


def rand_atom(r: random.Random, alpha: str, depth: int) -> str:
    k = r.random()
    if k < 0.45:
        return "".join(r.choice(alpha.replace("\n", "").replace("\r", "") or "a")
                       for _ in range(r.randint(1, 3)))
    if k < 0.55:
        return "."
    if k < 0.70:
        body = "".join(r.sample(alpha.replace("\n", "").replace("\r", "") + "z", r.randint(1, 2)))
        return "[" + ("^" if r.random() < 0.25 else "") + body + "]"
    if k < 0.74 and "\n" in alpha:
        return "\\n"
    if k < 0.78:
        return r.choice(["^", "$"])
    if depth > 0:
        return "(" + rand_alt(r, alpha, depth - 1) + ")"
    return r.choice(alpha.replace("\n", "").replace("\r", "") or "a")


def rand_piece(r: random.Random, alpha: str, depth: int) -> str:
    a = rand_atom(r, alpha, depth)
    k = r.random()
    if a in ("^", "$"):
        return a
    if k < 0.62:
        return a
    if k < 0.72:
        return a + "*"
    if k < 0.80:
        return a + "+"
    if k < 0.86:
        return a + "?"
    lo = r.randint(0, 3)
    form = r.random()
    if form < 0.3:
        return a + "{%d}" % max(lo, 1)
    if form < 0.6:
        return a + "{%d,%d}" % (lo, lo + r.randint(0, 3))
    if form < 0.8:
        return a + "{%d,}" % lo
    return a + "{,%d}" % (lo + 1)


def rand_concat(r: random.Random, alpha: str, depth: int) -> str:
    return "".join(rand_piece(r, alpha, depth) for _ in range(r.randint(1, 4)))


def rand_alt(r: random.Random, alpha: str, depth: int) -> str:
    return "|".join(rand_concat(r, alpha, depth) for _ in range(1 if r.random() < 0.6 else r.randint(2, 3)))


def rand_pattern(r: random.Random, alpha_name: str = None) -> tuple:
    name = alpha_name or r.choice(list(ALPHABETS))
    alpha = ALPHABETS[name]
    return rand_alt(r, alpha, 2), alpha


def rand_text(r: random.Random, alpha: str, n: int) -> bytes:
    return "".join(r.choice(alpha) for _ in range(n)).encode("latin-1")


def rand_long_literal_case(r: random.Random) -> tuple:
    """(pattern, text) around literals longer than the reference's 8-byte quadword compare: pure
    literals of 9..70 bytes (nodes coalesce up to 64), a literal inside a larger pattern, and a group
    repetition that the parser expands into one long node (`(abcab){3,6}` = 15 bytes + a bounded
    repetition).  The text holds exact copies and near misses (1-3 bytes changed), so that a compare
    that skips bytes shows up as a false positive."""
    alpha = "abcd"
    n = r.randint(9, 70)
    lit = "".join(r.choice(alpha) for _ in range(n))
    form = r.randint(0, 3)
    if form == 0:
        pat = lit
    elif form == 1:
        pat = "[ab]" + lit
    elif form == 2:
        pat = lit + "(a|bd)"
    else:
        unit = "".join(r.choice(alpha) for _ in range(r.randint(4, 8)))
        lo = r.randint(3, 6)
        pat = ".(" + unit + "){%d,%d}" % (lo, lo + r.randint(0, 3))
        lit = unit * lo
    parts = []
    for _ in range(r.randint(2, 7)):
        parts.append("".join(r.choice(alpha) for _ in range(r.randint(0, 40))))
        w = list(lit)
        for _ in range(r.choice([0, 0, 1, 1, 2, 3])):
            w[r.randrange(len(w))] = r.choice(alpha)
        parts.append(r.choice(alpha) + "".join(w) + r.choice(["a", "bd", "cc"]))
    return pat, "".join(parts).encode("latin-1")


# ---- the rest of the dialect -------------------------------------------------------------------
# Bracket ranges (also negated, with a leading / trailing '-'), the escapes of src/parser.cc:53-117
# (classes, escaped metacharacters, \\xHH with the reference's letter quirk), repetitions on all of
# them, and texts with bytes >= 0x80 (ranges compare as signed char, x64/codegen-x64.cc:898-908).
RICH_ESCAPES = ["\\d", "\\D", "\\s", "\\S", "\\t", "\\n", "\\(", "\\)", "\\[", "\\]", "\\{", "\\}", "\\|", "\\*",
                "\\+", "\\^", "\\$", "\\\\", "\\x41", "\\x4A", "\\x7e", "\\x09", "\\xaB"]
RICH_TEXT = "ab19AzJ@ \t\n(~Z5" + "\x80\xab\xff"


def _rich_atom(r: random.Random, depth: int) -> str:
    k = r.random()
    if k < 0.30:
        return "".join(r.choice("ab19Az") for _ in range(r.randint(1, 3)))
    if k < 0.45:
        return r.choice(RICH_ESCAPES)
    if k < 0.50:
        return "."
    if k < 0.75:
        items = []
        for _ in range(r.randint(1, 3)):
            if r.random() < 0.5:
                lo, hi = sorted(r.sample("09azAZ15bJ", 2))
                items.append(lo + "-" + hi)
            else:
                items.append(r.choice("ab19AZ@ -"))
        body = "".join(items)
        if r.random() < 0.15:
            body = "-" + body
        if r.random() < 0.15:
            body = body + "-"
        return "[" + ("^" if r.random() < 0.3 else "") + body + "]"
    if k < 0.80:
        return r.choice(["^", "$"])
    if depth > 0:
        return "(" + _rich_alt(r, depth - 1) + ")"
    return r.choice("ab19")


def _rich_piece(r: random.Random, depth: int) -> str:
    a = _rich_atom(r, depth)
    if a in ("^", "$"):
        return a
    k = r.random()
    if k < 0.6:
        return a
    if k < 0.7:
        return a + "*"
    if k < 0.8:
        return a + "+"
    if k < 0.86:
        return a + "?"
    lo = r.randint(0, 2)
    return a + r.choice(["{%d}" % max(lo, 1), "{%d,%d}" % (lo, lo + r.randint(0, 2)), "{%d,}" % lo, "{,%d}" % (lo + 1)])


def _rich_concat(r: random.Random, depth: int) -> str:
    return "".join(_rich_piece(r, depth) for _ in range(r.randint(1, 4)))


def _rich_alt(r: random.Random, depth: int) -> str:
    return "|".join(_rich_concat(r, depth) for _ in range(1 if r.random() < 0.6 else r.randint(2, 3)))


def rand_rich_pattern(r: random.Random) -> str:
    return _rich_alt(r, 2)


def rand_rich_text(r: random.Random, n: int) -> bytes:
    return "".join(r.choice(RICH_TEXT) for _ in range(n)).encode("latin-1")
