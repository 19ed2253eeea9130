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
