"""tests/jrep_tree.py — the deterministic source tree and the case list shared by
tests/golden/make_jrep_golden.py (which runs the reference's jrep on it) and the
jrep tests in tests/test_samples.py (which run samples/jrep.cc on it).

The files are cut from rejit_b200.workloads.source_blob (printable ASCII lines,
';\\n}' occurring naturally) and dressed so that the ends of files meet in the
ways that matter to a batching front end: no trailing newline, a trailing blank
line, a file ending in ';' followed by a file beginning with '}', empty files,
one-byte files."""
import os
import random

from rejit_b200 import workloads as W

N_BYTES = 120000

# (pattern, options[, only the files under this prefix]).  Patterns are ones for which the reference's default
# configuration equals its fast-forward-free one (literals, classes, anchors;
# checked by make_jrep_golden.py against the oracle), and whose output stays small.
CASES = [
    (";\n}", ["-H", "-n"]),
    (";\n}", ["-n", "-A1", "-B1"]),
    ("Qu", []),
    ("ab", ["-H", "-n", "-C2"]),
    ("^}", ["-H", "-n"]),
    (";$", ["-n", "-A2"]),
    ("[0-9][0-9][0-9]x", ["-H", "-n", "-B2"]),
    ("\n}[A-Z]", ["-H", "-n"]),
    (";\n*}", ["-H", "-n"]),
    ("}\n*[a-z]", ["-n"]),
    ("zz$", ["-H", "-c"]),
    # matches the empty string everywhere; over a batch, ";\n" at the end of g000 / g003 / g006 swallows the
    # separator and ENDS at the first byte of the next file, whose own empty match there must not be lost
    ("(;\n)*", ["-n"], "d9/"),
]
BATCHES = ["268435456", "0", "5000", "30000"]


def make_tree(root: str):
    """Writes the tree under `root`; returns the file paths, relative to root, sorted."""
    rng = random.Random(7)
    blob = W.source_blob(N_BYTES, seed=5).tobytes()
    at, k, paths = 0, 0, []
    while at < len(blob):
        d = os.path.join("d%d" % rng.randint(0, 3), "s%d" % rng.randint(0, 1))
        os.makedirs(os.path.join(root, d), exist_ok=True)
        n = rng.choice([0, 1, 7, 300, 3000, 9000])
        body = blob[at:at + n]
        at += n
        mode = rng.randint(0, 5)
        if mode == 0 and body and not body.endswith(b"\n"):
            body += b"\n"
        elif mode == 1:
            body = body.rstrip(b"\n") + b";"            # the next file may begin with '}'
        elif mode == 2:
            body = b"}" + body
        elif mode == 3:
            body += b"zz\n\n"
        rel = os.path.join(d, "f%03d.c" % k)
        with open(os.path.join(root, rel), "wb") as f:
            f.write(body)
        paths.append(rel)
        k += 1
    # Neighbours (in sorted order) whose ends meet: a match of ';\\n}' or ';\\n*}' over the batch
    # swallows the separator between g000|g001, and the two around g004 ('\\n' alone) at once.
    meet = [b"int a;\nint b;", b"}\nint c;\n", b"}", b"x = 1;", b"\n", b"} // end\nab;\n}\n", b";", b"};"]
    os.makedirs(os.path.join(root, "d9"), exist_ok=True)
    for i, body in enumerate(meet):
        rel = os.path.join("d9", "g%03d.c" % i)
        with open(os.path.join(root, rel), "wb") as f:
            f.write(body)
        paths.append(rel)
    return sorted(paths)
