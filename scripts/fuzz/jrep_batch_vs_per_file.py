"""scripts/fuzz/jrep_batch_vs_per_file.py — long-running differential run (scratch tooling behind the totals in DESIGN.md section 2;
the committed tests run seeded, bounded versions of the same comparisons).  Usage: python scripts/fuzz/jrep_batch_vs_per_file.py <seed> <seconds>.
Needs the build container (/root/reference, oracle/_ref) where it talks to the compiled reference."""
import os as _os
ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
TMP = _os.environ.get("REJIT_FUZZ_TMP", "/tmp/rejit_fuzz")
_os.makedirs(TMP, exist_ok=True)
import os, sys, random, subprocess, shutil, time
sys.path[:0] = [ROOT, ROOT + "/tests", ROOT + "/oracle"]
import fuzzgen, rejit_oracle as O, test_oracle as T
OURS = TMP + "/jrep_on_ref"
subprocess.run(["g++", "-std=c++11", "-O2", "-I/root/reference/include", ROOT + "/samples/jrep.cc", "-L" + ROOT + "/oracle/_ref", "-lrejit_ref", "-Wl,-rpath," + ROOT + "/oracle/_ref", "-o", OURS], check=True)
r = random.Random(int(sys.argv[1])); budget = float(sys.argv[2]); t0 = time.time()
root = TMP + "/jf%s" % sys.argv[1]
n = bad = 0
while time.time() - t0 < budget:
    shutil.rmtree(root, ignore_errors=True); os.makedirs(root)
    name = r.choice(["nl", "crlf", "ab"])
    alpha = fuzzgen.ALPHABETS[name] + ("\n" if name == "ab" else "")
    files = []
    for k in range(r.randint(2, 9)):
        body = fuzzgen.rand_text(r, alpha, r.choice([1, 2, 5, 30, 200]))
        p = os.path.join(root, "f%02d" % k); open(p, "wb").write(body); files.append(p)
    for _ in range(6):
        pat, _a = fuzzgen.rand_pattern(r, name if name != "ab" else "nl")
        try: O.Oracle(pat)
        except O.ParserError: continue
        if T._has_reference_ub(pat) or "(" in pat and pat.count("(") != pat.count(")"): continue
        opts = r.choice([["-n"], ["-H", "-n", "-A1"], ["-H", "-B2"], []])
        a = subprocess.run([OURS] + opts + ["--batch-bytes=0", pat] + files, capture_output=True)
        b = subprocess.run([OURS] + opts + [r.choice(["-j0", "-j2", "-j5"]), r.choice(["--gpus=1", "--gpus=2", "--gpus=3"]), "--batch-bytes=" + r.choice(["67108864", "300", "40"])] + [pat] + files, capture_output=True)
        n += 1
        if a.returncode < 0 or b.returncode < 0: continue      # the reference aborts on some patterns
        if (a.returncode, a.stdout) != (b.returncode, b.stdout):
            bad += 1; print("DIFF", repr(pat), opts, [open(f, "rb").read() for f in files], flush=True)
print("cases", n, "bad", bad)
