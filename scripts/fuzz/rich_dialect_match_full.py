"""scripts/fuzz/rich_dialect_match_full.py — long-running differential run (scratch tooling behind the totals in DESIGN.md section 2;
the committed tests run seeded, bounded versions of the same comparisons).  Usage: python scripts/fuzz/rich_dialect_match_full.py <seed> <seconds>.
Needs the build container (/root/reference, oracle/_ref) where it talks to the compiled reference."""
import os as _os
ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
TMP = _os.environ.get("REJIT_FUZZ_TMP", "/tmp/rejit_fuzz")
_os.makedirs(TMP, exist_ok=True)
import sys, random, time
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests", ROOT + "/tests/golden"]
import fuzzgen, rejit_oracle as O, conftest, test_oracle as T
from make_golden import Ref
ref = Ref(); ref.flags(2); hostsim = conftest.hostsim.__wrapped__()
r = random.Random(int(sys.argv[1])); budget = float(sys.argv[2]); t0 = time.time(); n = bad = 0
while time.time() - t0 < budget:
    pat = fuzzgen.rand_rich_pattern(r); pb = pat.encode("latin-1")
    try: o = O.Oracle(pat)
    except O.ParserError: continue
    use_ref = ref.parse_ok(pb) and not T._has_reference_ub(pat) and o.longest_literal <= 16
    for _ in range(4):
        t = fuzzgen.rand_rich_text(r, r.randint(0, 12))
        e = o.match_full(t); n += 1
        if bool(hostsim.match_full(pat, t)) != e: bad += 1; print("PRODUCT", repr(pat), t)
        if use_ref and bool(ref.match_full(pb, t)) != e: bad += 1; print("REF", repr(pat), t)
print("cases", n, "bad", bad)
