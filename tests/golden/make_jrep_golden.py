#!/usr/bin/env python3
"""tests/golden/make_jrep_golden.py — regenerates tests/golden/jrep_cases.json.

Runs ONLY in the build container: it executes the reference's own jrep
(oracle/_ref/jrep_ref, built by `make -C oracle ref` from
/root/reference/sample/jrep.cc) on the tree of tests/jrep_tree.py, file names
given explicitly in sorted order (a directory walk's order depends on the file
system), and records its stdout per case.  Before recording, every pattern is
matched over every file with the compiled reference (default flags, what jrep_ref
runs) AND with the parity oracle (oracle/rejit_oracle.py): a case where the two
disagree would pin a fast-forward defect of the reference
instead of the parity semantics, and is refused."""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), HERE]
import jrep_tree                       # noqa: E402
import rejit_oracle as O               # noqa: E402
from make_golden import Ref            # noqa: E402

JREP_REF = os.path.join(ROOT, "oracle", "_ref", "jrep_ref")


def main():
    ref = Ref()
    ref.flags(0)
    out = []
    with tempfile.TemporaryDirectory() as root:
        paths = jrep_tree.make_tree(root)
        bodies = [open(os.path.join(root, p), "rb").read() for p in paths]
        all_paths, all_bodies = paths, bodies
        for case in jrep_tree.CASES:
            pat, opts = case[0], case[1]
            only = case[2] if len(case) > 2 else ""
            paths = [p for p in all_paths if p.startswith(only)]
            bodies = [b for p, b in zip(all_paths, all_bodies) if p.startswith(only)]
            o = O.Oracle(pat)
            for p, body in zip(paths, bodies):
                if body:
                    assert ref.match_all(pat.encode("latin-1"), body) == [list(m) for m in o.match_all(body)], (pat, p)
            r = subprocess.run([JREP_REF] + opts + [pat] + paths, cwd=root, capture_output=True, check=True)
            expected = r.stdout
            n_matches = None
            if "-c" in opts:
                # The reference's Linux branch drops the colour (sample/jrep.cc:329-341, only its macOS
                # branch emits the escape codes): the recorded text is the uncoloured one, and the test
                # strips "\x1B[31m" / "\x1B[0m" from the sample's output and counts them.
                assert b"\x1b" not in expected
                n_matches = sum(len(o.match_all(body)) for body in bodies if body)
            out.append({"re": pat, "options": opts, "only": only, "stdout": expected.decode("latin-1"), "bytes": len(expected),
                        "matches": n_matches})
            print(repr(pat), opts, "->", "%d bytes" % len(expected))
    json.dump({"tree": "tests/jrep_tree.py make_tree()", "files": len(all_paths), "cases": out},
              open(os.path.join(HERE, "jrep_cases.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
