"""world_size-2 (and 3) gloo test of the N>1 path on CPU: every rank resolves
its slab, one all-gather stitches the chain at the slab edges, the summed match
count must equal the one-piece oracle result."""
import json
import os
import random
import subprocess
import sys

import pytest

import fuzzgen
import rejit_oracle as O
from conftest import ROOT


def _run(world, cases, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_gloo_worker.py"), json.dumps(cases)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.split("\n") if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


@pytest.mark.parametrize("world,port", [(2, 29611), (3, 29612)])
def test_slab_stitching_over_gloo(hostsim, world, port):
    import rejit_b200
    r = random.Random(77 + world)
    cases, expect = [], []
    # adversarial: matches straddling / abutting the cut, runs of overlapping candidates
    fixed = [("aa", b"a" * 101), ("aba", b"ab" * 60), ("ab+", b"abbbbbbbbbbbbbbbbbbbbbbbbbbbbbbbb" * 3),
             ("x|$", b"ax\n" * 40), ("abc", b"abc" * 41)]
    for pat, text in fixed:
        cases.append([pat, text.hex()])
        expect.append(len(O.Oracle(pat).match_all(text)))
    while len(cases) < 40:
        pat, alpha = fuzzgen.rand_pattern(r)
        try:
            o = O.Oracle(pat)
        except O.ParserError:
            continue
        if "reentrant" in rejit_b200.Regej(pat).describe():
            continue            # not sharded (label replay cannot be cut)
        text = fuzzgen.rand_text(r, alpha, r.randint(50, 400))
        cases.append([pat, text.hex()])
        expect.append(len(o.match_all(text)))
    got = _run(world, cases, port)
    assert [g[0] for g in got] == expect
    assert max(g[1] for g in got) <= world       # collective rounds stay bounded
