"""scripts/fuzz/hostsim_vs_oracle.py — long-running differential run (scratch tooling behind the totals in DESIGN.md section 2;
the committed tests run seeded, bounded versions of the same comparisons).  Usage: python scripts/fuzz/hostsim_vs_oracle.py <seed> <seconds>.
Needs the build container (/root/reference, oracle/_ref) where it talks to the compiled reference."""
import os as _os
ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
TMP = _os.environ.get("REJIT_FUZZ_TMP", "/tmp/rejit_fuzz")
_os.makedirs(TMP, exist_ok=True)
import sys, os, random, time
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import conftest, fuzzgen
import rejit_oracle as O
import rejit_b200
hostsim = conftest.hostsim.__wrapped__() if hasattr(conftest.hostsim, "__wrapped__") else None
seed0 = int(sys.argv[1]); budget = float(sys.argv[2])
t0 = time.time(); checked = 0; fails = 0
seed = seed0
while time.time() - t0 < budget:
    r = random.Random(seed); seed += 1
    for _ in range(200):
        pat, alpha = fuzzgen.rand_pattern(r)
        try:
            o = O.Oracle(pat)
        except O.ParserError:
            got, _ = hostsim.match_all(pat, b"")
            if got != -1:
                print("PARSE-DIFF", repr(pat), flush=True); fails += 1
            continue
        reent = "reentrant" in rejit_b200.Regej(pat).describe()
        for n in (r.randint(0, 40), r.randint(100, 2500)):
            t = fuzzgen.rand_text(r, alpha, n)
            exp = o.match_all(t)
            for strategy in (-1, 3):
                got, desc = hostsim.match_all(pat, t, strategy)
                if got != exp:
                    print("DIFF", repr(pat), repr(t[:200]), len(t), desc, strategy, flush=True); fails += 1
            if bool(hostsim.match_full(pat, t)) != o.match_full(t):
                print("FULLDIFF", repr(pat), repr(t[:200]), flush=True); fails += 1
            if not reent:
                k = r.choice([2, 3, 4, 8])
                if hostsim.match_all_slabs(pat, t, k) != exp:
                    print("SLABDIFF", repr(pat), repr(t[:200]), len(t), k, flush=True); fails += 1
            checked += 1
print("done seeds", seed0, "..", seed, "checked", checked, "fails", fails, flush=True)
