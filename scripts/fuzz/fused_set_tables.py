"""scripts/fuzz/fused_set_tables.py — long-running differential run (scratch tooling behind the totals in DESIGN.md section 2;
the committed tests run seeded, bounded versions of the same comparisons).  Usage: python scripts/fuzz/fused_set_tables.py <seed> <seconds>.
Needs the build container (/root/reference, oracle/_ref) where it talks to the compiled reference."""
import os as _os
ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
TMP = _os.environ.get("REJIT_FUZZ_TMP", "/tmp/rejit_fuzz")
_os.makedirs(TMP, exist_ok=True)
import ctypes, sys, os, random, time
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import conftest, fuzzgen, rejit_oracle as O
conftest.hostsim.__wrapped__()
L = ctypes.CDLL(ROOT + "/tests/_build/libhostsim.so")
L.hostsim_set_match_all.restype = ctypes.c_int64
L.hostsim_set_match_all.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64]
r = random.Random(int(sys.argv[1])); budget = float(sys.argv[2]); t0 = time.time(); n = fused = bad = 0
def member(r):
    # fixed-length alternations over acgt with classes, like the regex-dna variants but of any length 1..10
    def word(k): return "".join(r.choice(["a", "c", "g", "t", "[ac]", "[cgt]", "[agt]", "."]) if r.random() < 0.9 else r.choice("acgt") for _ in range(k))
    k = r.randint(1, 10)
    return "|".join(word(k) for _ in range(r.randint(1, 3)))
while time.time() - t0 < budget:
    pats = [member(r) for _ in range(r.randint(1, 9))]
    joined = "\x01".join(pats).encode()
    t = fuzzgen.rand_text(r, r.choice(["acgt", "acgtN", "acgtBD\n"]), r.choice([50, 700, 9000]))
    for j, p in enumerate(pats):
        out = (ctypes.c_uint64 * (2 * (len(t) + 2)))()
        k = L.hostsim_set_match_all(joined, len(joined), b"\x01", t, len(t), j, out, len(t) + 2)
        n += 1
        if k == -5: break            # not fusable
        fused += 1
        got = [(out[2 * i], out[2 * i + 1]) for i in range(k)] if k >= 0 else k
        if got != O.Oracle(p).match_all(t):
            bad += 1
            if bad < 5: print("DIFF", pats, j, t[:80], k, flush=True)
print("member runs", n, "fused", fused, "bad", bad)
