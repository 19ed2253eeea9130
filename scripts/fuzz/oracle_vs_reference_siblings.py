"""scripts/fuzz/oracle_vs_reference_siblings.py — long-running differential run (scratch tooling behind the totals in DESIGN.md section 2;
the committed tests run seeded, bounded versions of the same comparisons).  Usage: python scripts/fuzz/oracle_vs_reference_siblings.py <seed> <seconds>.
Needs the build container (/root/reference, oracle/_ref) where it talks to the compiled reference."""
import os as _os
ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
TMP = _os.environ.get("REJIT_FUZZ_TMP", "/tmp/rejit_fuzz")
_os.makedirs(TMP, exist_ok=True)
import sys, os, random, time, subprocess
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests", ROOT + "/tests/golden"]
import fuzzgen, rejit_oracle as O
from make_golden import Ref
import test_oracle as T
ref = Ref(); ref.flags(2)
seed = int(sys.argv[1]); budget = float(sys.argv[2]); t0 = time.time(); checked = fails = 0
while time.time() - t0 < budget:
    r = random.Random(seed); seed += 1
    for _ in range(100):
        pat, alpha = fuzzgen.rand_pattern(r)
        try: o = O.Oracle(pat, long_literal_defect=True)
        except O.ParserError: continue
        pb = pat.encode("latin-1")
        if T._has_reference_ub(pat) or not ref.parse_ok(pb): continue
        for _ in range(3):
            t = fuzzgen.rand_text(r, alpha, r.choice([r.randint(0, 48), r.randint(100, 400)]))
            checked += 1
            full = bool(ref.match_full(pb, t)); anyw = bool(ref.match_anywhere(pb, t))
            if o.match_full(t) != full:
                fails += 1; print("FULL", repr(pat), t.hex(), full, flush=True)
            if o.match_anywhere(t) != anyw:
                fails += 1; print("ANYWHERE", repr(pat), t.hex(), anyw, flush=True)
            # MatchFirst of the product is defined as MatchAll[0] (reference quirk B11): compare existence only
            f = ref.match_first(pb, t)
            if bool(f) != bool(o.match_all(t)):
                fails += 1; print("FIRST-EXISTS", repr(pat), t.hex(), f, flush=True)
print("checked", checked, "fails", fails, flush=True)
