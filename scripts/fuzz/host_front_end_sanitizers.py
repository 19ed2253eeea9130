"""scripts/fuzz/host_front_end_sanitizers.py — long-running differential run (scratch tooling behind the totals in DESIGN.md section 2;
the committed tests run seeded, bounded versions of the same comparisons).  Usage: python scripts/fuzz/host_front_end_sanitizers.py <seed> <seconds>.
Needs the build container (/root/reference, oracle/_ref) where it talks to the compiled reference."""
import os as _os
ROOT = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
TMP = _os.environ.get("REJIT_FUZZ_TMP", "/tmp/rejit_fuzz")
_os.makedirs(TMP, exist_ok=True)
import ctypes, sys, os, random, time
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import fuzzgen
L = ctypes.CDLL(TMP + "/asan/libhostsim.so")
L.hostsim_match_all.restype = ctypes.c_int64
L.hostsim_match_all.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64, ctypes.c_char_p, ctypes.c_size_t]
L.hostsim_match_all_slabs.restype = ctypes.c_int64
L.hostsim_match_all_slabs.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64]
L.hostsim_match_full.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint64]
L.hostsim_set_match_all.restype = ctypes.c_int64
L.hostsim_set_match_all.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64]
r = random.Random(int(sys.argv[1])); budget = float(sys.argv[2]); t0 = time.time(); n = 0
META = "()[]{}|*+?^$\\.,-0123456789abcxdn\n"
while time.time() - t0 < budget:
    if r.random() < 0.5:
        pat, alpha = fuzzgen.rand_pattern(r)
    else:
        pat = "".join(r.choice(META) for _ in range(r.randint(0, 14))); alpha = "abcx\n"
    pb = pat.encode("latin-1")
    for ln in (0, r.randint(1, 40), r.randint(100, 1500)):
        t = fuzzgen.rand_text(r, alpha, ln)
        cap = len(t) + 2
        out = (ctypes.c_uint64 * (2 * cap))(); d = ctypes.create_string_buffer(512)
        for strategy in (-1, 3):
            L.hostsim_match_all(pb, len(pb), r.randint(0, 1), t, len(t), strategy, out, cap, d, 512)
        L.hostsim_match_full(pb, len(pb), t, len(t))
        L.hostsim_match_all_slabs(pb, len(pb), t, len(t), r.choice([2, 3, 8]), out, cap)
    # sets
    pats = [fuzzgen.rand_pattern(r, "dna")[0] for _ in range(r.randint(1, 5))]
    joined = "\x01".join(pats).encode("latin-1"); t = fuzzgen.rand_text(r, "acgt", 800)
    out = (ctypes.c_uint64 * 4000)()
    for j in range(len(pats)):
        L.hostsim_set_match_all(joined, len(joined), b"\x01", t, len(t), j, out, 2000)
    n += 1
print("patterns", n, "ok")
