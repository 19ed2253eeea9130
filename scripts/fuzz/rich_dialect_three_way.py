"""scripts/fuzz/rich_dialect_three_way.py — long-running differential run (scratch tooling behind the totals in DESIGN.md section 2;
the committed tests run seeded, bounded versions of the same comparisons).  Usage: python scripts/fuzz/rich_dialect_three_way.py <seed> <seconds>.
Needs the build container (/root/reference, oracle/_ref) where it talks to the compiled reference."""
import os as _os
ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
TMP = _os.environ.get("REJIT_FUZZ_TMP", "/tmp/rejit_fuzz")
_os.makedirs(TMP, exist_ok=True)
import sys, os, random, time, subprocess
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests", ROOT + "/tests/golden", _os.path.dirname(_os.path.abspath(__file__))]
import fuzzgen as gen2, rejit_oracle as O, conftest
from make_golden import Ref
import test_oracle as T
ref = Ref(); ref.flags(2)
hostsim = conftest.hostsim.__wrapped__()
seed = int(sys.argv[1]); budget = float(sys.argv[2]); t0 = time.time(); checked = 0; fails = {}
def note(kind, pat, t):
    fails[kind] = fails.get(kind, 0) + 1
    if fails[kind] <= 6: print(kind, repr(pat), t.hex(), flush=True)
while time.time() - t0 < budget:
    r = random.Random(seed); seed += 1
    for _ in range(100):
        pat = gen2.rand_rich_pattern(r); pb = pat.encode("latin-1")
        try: o = O.Oracle(pat, long_literal_defect=True); oe = O.Oracle(pat)
        except O.ParserError:
            if ref.parse_ok(pb): note("PARSE-oracle-rejects-ref-accepts", pat, b"")
            got, _ = hostsim.match_all(pat, b"")
            if got != -1: note("PARSE-product-accepts-oracle-rejects", pat, b"")
            continue
        got, _ = hostsim.match_all(pat, b"")
        if got == -1: note("PARSE-product-rejects-oracle-accepts", pat, b""); continue
        ref_ok = ref.parse_ok(pb) and not T._has_reference_ub(pat)
        if not ref.parse_ok(pb): note("PARSE-oracle-accepts-ref-rejects", pat, b"")
        for _ in range(3):
            t = gen2.rand_rich_text(r, r.choice([r.randint(0, 40), r.randint(60, 300)]))
            checked += 1
            exp = oe.match_all(t)
            for strategy in (-1, 3):
                g, desc = hostsim.match_all(pat, t, strategy)
                if g != exp: note("PRODUCT-vs-oracle", pat, t)
            if ref_ok and b"\x00" not in t:
                if [list(m) for m in o.match_all(t)] != ref.match_all(pb, t):
                    fresh = subprocess.run([sys.executable, "-c", T._FRESH, T.REF_SO, pb.hex(), t.hex()], capture_output=True, text=True).stdout.strip()
                    if str([list(m) for m in o.match_all(t)]) != fresh: note("ORACLE-vs-reference", pat, t)
print("checked", checked, "fails", fails, flush=True)
