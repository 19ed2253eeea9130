"""scripts/fuzz/kmer_index.py — long-running differential run (scratch tooling behind the totals in DESIGN.md section 2;
the committed tests run seeded, bounded versions of the same comparisons).  Usage: python scripts/fuzz/kmer_index.py <seed> <seconds>.
Needs the build container (/root/reference, oracle/_ref) where it talks to the compiled reference."""
import os as _os
ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
TMP = _os.environ.get("REJIT_FUZZ_TMP", "/tmp/rejit_fuzz")
_os.makedirs(TMP, exist_ok=True)
import sys, random, time
sys.path[:0] = [ROOT, ROOT + "/tests", ROOT + "/oracle"]
import test_kmer_index as K, fuzzgen
r = random.Random(int(sys.argv[1])); budget = float(sys.argv[2]); t0 = time.time(); sets = idx = bad = 0
while time.time() - t0 < budget:
    live = r.choice(["acgt", "acgt", "ab", "xyz", "qrst", "aceg", "ACGT", "0123"])
    def word(k):
        out = []
        for _ in range(k):
            if r.random() < 0.75: out.append(r.choice(live))
            else:
                cls = "".join(sorted(set(r.sample(live, r.randint(1, len(live))))))
                out.append("[" + cls + "]")
        return "".join(out)
    pats = []
    for _ in range(r.randint(1, 9)):
        k = r.randint(1, 8)
        pats.append("|".join(word(k) for _ in range(r.randint(1, 2))))
    try:
        s = K._Set(pats)
    except Exception as e:
        continue
    sets += 1
    if s.kmer_tables() is None: continue
    idx += 1
    others = "".join(c for c in "acgtNBxyzqrsACGT0123\n" if c not in live)[:6]
    t = fuzzgen.rand_text(r, live * 6 + others, r.choice([40, 600, 3000]))
    got = K.emulate(s, t, seed=r.randint(0, 99))
    for j, p in enumerate(pats):
        if got[j] != K.occurrences(p, t):
            bad += 1
            if bad < 5: print("DIFF", pats, j, t[:100], flush=True)
print("sets", sets, "with k-mer index", idx, "bad", bad)
