"""The C++ sample written against include/rejit.h (source compatibility with the
reference's public header): it must compile and link against
librejit_b200.so, refuse to run without a GPU (no CPU fallback), and on a GPU
print what the oracle says."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import __graft_entry__ as entry
    entry.build()
    exe = str(tmp_path / "regexdna")
    libdir = os.path.join(ROOT, "rejit_b200")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "samples", "regexdna.cc"), "-L" + libdir, "-lrejit_b200",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def test_sample_compiles_against_rejit_h_and_needs_a_gpu(tmp_path):
    exe = _build(tmp_path)
    import rejit_b200
    if rejit_b200.device_count() > 0:
        pytest.skip("a GPU is present: covered by the gpu tier")
    r = subprocess.run([exe], input=b">ONE x\nacgt\n", capture_output=True)
    assert r.returncode != 0 and b"no CUDA device" in r.stderr        # loud, not a silent CPU path


@pytest.mark.gpu
def test_sample_regexdna_output(tmp_path):
    import rejit_oracle as O
    from rejit_b200 import workloads as W
    exe = _build(tmp_path)
    fa = W.fasta_file(20000)
    r = subprocess.run([exe], input=fa, capture_output=True, check=True)

    def replace(pat, text, w):
        out, at = bytearray(), 0
        for b, e in O.Oracle(pat).match_all(text):
            out += text[at:b] + w
            at = e
        return bytes(out + text[at:])

    seq = replace(W.STRIP_PATTERN, fa, b"")
    lines = ["%s %d" % (p, len(O.Oracle(p).match_all(seq))) for p in W.DNA_PATTERNS]
    cur = seq
    for code, alt in W.IUB_SUBSTITUTIONS:
        cur = replace(code, cur, alt.encode())
    expected = "\n".join(lines) + "\n\n%d\n%d\n%d\n" % (len(fa), len(seq), len(cur))
    assert r.stdout.decode() == expected
