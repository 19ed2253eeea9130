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


def _run(world, cases, port, set_cases=None, stitch="nccl"):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_gloo_worker.py"), json.dumps(cases)]
    if set_cases is not None:
        cmd.append(json.dumps(set_cases))
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RJ_STITCH=stitch))
    assert p.returncode == 0, p.stderr[-2000:]
    tag = "SETRESULT " if set_cases is not None else "RESULT "
    line = [l for l in p.stdout.split("\n") if l.startswith(tag)][-1]
    return json.loads(line[len(tag):])


@pytest.mark.parametrize("world,port,stitch", [(2, 29613, "nccl"), (3, 29614, "nccl"), (2, 29615, "shm"), (4, 29616, "shm"), (8, 29617, "shm"),
                                               (2, 29618, "p2p"), (3, 29619, "p2p"), (8, 29620, "p2p")])
def test_set_stitching_over_gloo(hostsim, world, port, stitch):
    """The fused-set stitch of bench.py --gpus N: k members' records in one all-gather per round."""
    from rejit_b200 import workloads as W
    seq = W.fasta_sequence(300).tobytes()
    third = len(seq) // world
    # matches of several members straddling and abutting every cut
    planted = bytearray(seq)
    for cut in range(third, len(seq) - 8, third):
        planted[cut - 3:cut + 5] = b"agggtaaa"
        planted[cut + 5:cut + 13] = b"tttaccct"
    set_cases = [[W.DNA_PATTERNS, bytes(planted).hex()],
                 [["aa", "aba", "ab"], (b"ab" * 40 + b"a" * 41).hex()],
                 [["abc", "bca", "cab"], (b"abc" * 50).hex()]]
    expect = [[len(O.Oracle(p).match_all(bytes.fromhex(t))) for p in pats] for pats, t in set_cases]
    fixed = [["aa", (b"a" * 101).hex()], ["x|$", (b"ax\n" * 40).hex()], ["abc", (b"abc" * 41).hex()]]
    got = _run(world, fixed * 5, port, set_cases * 5, stitch)
    assert [g[0] for g in got] == expect * 5
    assert max(g[1] for g in got) <= world + 1


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
