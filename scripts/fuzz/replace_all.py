"""scripts/fuzz/replace_all.py — long-running differential run (scratch tooling behind the totals in DESIGN.md section 2;
the committed tests run seeded, bounded versions of the same comparisons).  Usage: python scripts/fuzz/replace_all.py <seed> <seconds>.
Needs the build container (/root/reference, oracle/_ref) where it talks to the compiled reference."""
import os as _os
ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
TMP = _os.environ.get("REJIT_FUZZ_TMP", "/tmp/rejit_fuzz")
_os.makedirs(TMP, exist_ok=True)
import sys, random, time
sys.path[:0] = [ROOT, ROOT + "/tests", ROOT + "/oracle"]
import conftest, fuzzgen, rejit_oracle as O
hostsim = conftest.hostsim.__wrapped__()
r = random.Random(int(sys.argv[1])); budget = float(sys.argv[2]); t0 = time.time(); n = bad = 0
while time.time() - t0 < budget:
    pat, alpha = fuzzgen.rand_pattern(r)
    try: o = O.Oracle(pat)
    except O.ParserError: continue
    for ln in (r.randint(0, 60), r.choice([4000, 4096, 4097, 9000, 13000])):
        t = fuzzgen.rand_text(r, alpha, ln)
        w = r.choice([b"", b"Q", b"(c|g|t)", b"xy" * 20])
        ms = o.match_all(t)
        out, at = bytearray(), 0
        for b, e in ms:
            out += t[at:b] + w; at = e
        out += t[at:]
        got = hostsim.replace_all(pat, t, w)
        n += 1
        if got != (len(ms), bytes(out)):
            bad += 1
            if bad < 5: print("DIFF", repr(pat), len(t), w, got[0], len(ms), flush=True)
print("cases", n, "bad", bad)
