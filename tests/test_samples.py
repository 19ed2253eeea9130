"""The C++ samples written against include/rejit.h (source compatibility with the
reference's public header): they must compile and link against
librejit_b200.so, refuse to run without a GPU (no CPU fallback), and on a GPU
print what the oracle / the reference's own program says.

samples/jrep.cc also compiles against the REFERENCE's
header and library; that build is the CPU-tier checker of its front end (file
batching, line index, context printing) against the reference's own jrep."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INCLUDE = "/root/reference/include"
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
RUN_TIMEOUT = 600          # seconds for one run of a sample: a hang must fail the test, not stall the tier


def _build(tmp_path, name="regexdna"):
    import __graft_entry__ as entry
    entry.build()
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "rejit_b200")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "samples", name + ".cc"), "-L" + libdir, "-lrejit_b200",
                    "-Wl,-rpath," + libdir, "-lpthread", "-o", exe], check=True)
    return exe


@pytest.mark.parametrize("name", ["regexdna", "regexdna_device"])
def test_sample_compiles_against_rejit_h_and_needs_a_gpu(tmp_path, name):
    exe = _build(tmp_path, name)
    import rejit_b200
    if rejit_b200.device_count() > 0:
        pytest.skip("a GPU is present: covered by the gpu tier")
    r = subprocess.run([exe], input=b">ONE x\nacgt\n", capture_output=True, timeout=RUN_TIMEOUT)
    assert r.returncode != 0 and b"no CUDA device" in r.stderr        # loud, not a silent CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["regexdna", "regexdna_device"])
def test_sample_regexdna_output(tmp_path, name):
    """regexdna: written against the reference's surface only (twelve host round trips);
    regexdna_device: one upload, rejit::Text chain + the fused set count.  Same output."""
    import rejit_oracle as O
    from rejit_b200 import workloads as W
    exe = _build(tmp_path, name)
    fa = W.fasta_file(20000)
    r = subprocess.run([exe], input=fa, capture_output=True, check=True, timeout=RUN_TIMEOUT)

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


# ---- jrep --------------------------------------------------------------------------
def _jrep_cases():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "jrep_cases.json")))["cases"]


def _check_jrep_against_golden(exe, root, paths, quick=False):
    """quick: every batching mode for the first three cases and the last one, two modes for the rest (each run
    on a GPU box pays for a CUDA context)."""
    import jrep_tree
    cases = _jrep_cases()
    for k, case in enumerate(cases):
        expected = case["stdout"].encode("latin-1")
        files = [p for p in paths if p.startswith(case.get("only", ""))]
        for batch in (jrep_tree.BATCHES if not quick or k < 3 or k == len(cases) - 1 else ["268435456", "5000"]):
            # staging: the build's default (threads on librejit_b200, none on the reference), none, three threads
            jobs = {"5000": ["-j0"], "30000": ["-j3"]}.get(batch, [])
            r = subprocess.run([exe] + case["options"] + jobs + ["--batch-bytes=" + batch, case["re"]] + files, cwd=root,
                               capture_output=True, timeout=RUN_TIMEOUT)
            assert r.returncode == 0, (case["re"], case["options"], batch, r.stderr[-300:])
            got = r.stdout
            if "-c" in case["options"]:
                assert got.count(b"\x1B[31m") == got.count(b"\x1B[0m") == case["matches"]
                got = got.replace(b"\x1B[31m", b"").replace(b"\x1B[0m", b"")
            assert got == expected, (case["re"], case["options"], batch, len(got), len(expected))


def test_jrep_compiles_against_rejit_h_and_needs_a_gpu(tmp_path):
    exe = _build(tmp_path, "jrep")
    r = subprocess.run([exe], capture_output=True, timeout=RUN_TIMEOUT)
    assert r.returncode == 64 and b"Usage" in r.stderr
    import rejit_b200
    if rejit_b200.device_count() > 0:
        pytest.skip("a GPU is present: covered by the gpu tier")
    (tmp_path / "a.c").write_bytes(b"int x;\n}\n")
    r = subprocess.run([exe, "x", str(tmp_path / "a.c")], capture_output=True, timeout=RUN_TIMEOUT)
    assert r.returncode != 0 and b"no CUDA device" in r.stderr and r.stdout == b""


@pytest.mark.skipif(not (os.path.exists(os.path.join(REF_INCLUDE, "rejit.h")) and
                         os.path.exists(os.path.join(REF_DIR, "jrep_ref"))),
                    reason="needs the reference's header and oracle/_ref (build container only)")
def test_jrep_front_end_on_the_reference_library(tmp_path):
    """samples/jrep.cc built UNCHANGED against the reference's header and library: same bytes as the
    committed golden output (= the reference's jrep on explicit file lists), in every batching mode,
    and the same bytes and exit code as the reference's jrep run live on a recursive walk."""
    import jrep_tree
    exe = str(tmp_path / "jrep_on_ref")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I" + REF_INCLUDE, os.path.join(ROOT, "samples", "jrep.cc"),
                    "-L" + REF_DIR, "-lrejit_ref", "-Wl,-rpath," + REF_DIR, "-lpthread", "-o", exe], check=True)
    root = str(tmp_path / "tree")
    os.makedirs(root)
    paths = jrep_tree.make_tree(root)
    _check_jrep_against_golden(exe, root, paths)
    ref = os.path.join(REF_DIR, "jrep_ref")
    noff = dict(os.environ, REJIT_REF_FLAGSET="2")        # both programs in the parity configuration (oracle/ref_shim.cc)
    for pat in (";\n}", "x*", "\n", "a.*b", "(;|\n)+}", "$"):      # incl. empty matches, matches that swallow separators
        for opts in (["-n"], ["-H", "-n", "-A2", "-B1"], ["-H", "-C1"]):
            a = subprocess.run([ref] + opts + ["-r", pat, "."], cwd=root, capture_output=True, env=noff, timeout=RUN_TIMEOUT)
            assert a.returncode == 0 and a.stdout
            for batch in jrep_tree.BATCHES:
                b = subprocess.run([exe] + opts + ["-r", "--batch-bytes=" + batch, pat, "."], cwd=root, capture_output=True,
                                   env=noff, timeout=RUN_TIMEOUT)
                assert (a.returncode, a.stdout) == (b.returncode, b.stdout), (pat, opts, batch)
    # found by fuzzing batch against per-file mode: "aa\\n" swallows the separator after 'bbaaa' and ends at the
    # first byte of the next file, whose empty match at that offset used to be lost
    meet = str(tmp_path / "meet")
    os.makedirs(meet)
    for i, body in enumerate([b"ba\n\n\n", b"bba\na", b"bbaaa", b"\n", b"b\n\naaab\n", b"a\n\nb\nbbaa"]):
        with open(os.path.join(meet, "f%d" % i), "wb") as f:
            f.write(body)
    names = sorted(os.listdir(meet))
    for pat in ("aa\\n*", "(aa\n)*", "a*\n*"):
        a = subprocess.run([ref, "-n", pat] + names, cwd=meet, capture_output=True, env=noff, timeout=RUN_TIMEOUT)
        for batch in jrep_tree.BATCHES:
            b = subprocess.run([exe, "-n", "--batch-bytes=" + batch, pat] + names, cwd=meet, capture_output=True, env=noff, timeout=RUN_TIMEOUT)
            assert a.returncode == 0 and (a.returncode, a.stdout) == (b.returncode, b.stdout), (pat, batch)
    # --gpus N on this build = N matcher threads calling the shared compiled Regej (as the reference's jrep does)
    for extra in (["--gpus=3", "--batch-bytes=7000", "-j2"], ["--gpus=2", "-j4", "--batch-bytes=0"]):
        a = subprocess.run([ref, "-H", "-n", "-B1", "-r", "ab", "."], cwd=root, capture_output=True, env=noff, timeout=RUN_TIMEOUT)
        b = subprocess.run([exe, "-H", "-n", "-B1", "-r", *extra, "ab", "."], cwd=root, capture_output=True, env=noff, timeout=RUN_TIMEOUT)
        assert a.stdout and (a.returncode, a.stdout) == (b.returncode, b.stdout), extra
    # -j N: N threads stage the batch; same bytes whatever N and the batch size
    a = subprocess.run([ref, "-H", "-n", "-A1", "-r", ";\n}", "."], cwd=root, capture_output=True, env=noff, timeout=RUN_TIMEOUT)
    for jobs in ("-j1", "-j4", "-j16"):
        for batch in jrep_tree.BATCHES:
            b = subprocess.run([exe, jobs, "-H", "-n", "-A1", "-r", "--batch-bytes=" + batch, ";\n}", "."], cwd=root,
                               capture_output=True, env=noff, timeout=RUN_TIMEOUT)
            assert a.stdout and (a.returncode, a.stdout) == (b.returncode, b.stdout), (jobs, batch)
    # a file that cannot be opened ends the run with its errno; what came before it is printed, what follows is not
    # (sample/jrep.cc:269-274, 540-541).  Root opens anything, except a write-only sysfs attribute.
    locked = []
    for d, _, fs in os.walk("/sys/class") if os.path.isdir("/sys/class") else []:
        for f in fs:
            try:
                if not locked and (os.stat(os.path.join(d, f)).st_mode & 0o777) == 0o200:
                    locked.append(os.path.join(d, f))
            except OSError:
                pass
    if locked:
        names3 = [names[0], locked[0], names[1]]
        a = subprocess.run([ref, "-H", "a", *names3], cwd=meet, capture_output=True, env=noff, timeout=RUN_TIMEOUT)
        assert a.returncode == 13 and a.stdout
        for extra in ([], ["-j3"], ["-j3", "--batch-bytes=0"], ["--batch-bytes=0"]):
            b = subprocess.run([exe, "-H", *extra, "a", *names3], cwd=meet, capture_output=True, env=noff, timeout=RUN_TIMEOUT)
            assert (a.returncode, a.stdout) == (b.returncode, b.stdout), extra
    for args in (["x", "missing.c"], ["x", "."], ["x", "d0"]):       # stat failure (exit 255), directory without -r
        a = subprocess.run([ref] + args, cwd=root, capture_output=True, timeout=RUN_TIMEOUT)
        b = subprocess.run([exe] + args, cwd=root, capture_output=True, timeout=RUN_TIMEOUT)
        assert (a.returncode, a.stdout, a.stderr) == (b.returncode, b.stdout, b.stderr), args


@pytest.mark.gpu
def test_sample_jrep_output(tmp_path):
    """The same program on librejit_b200.so: one pinned blob and one upload per batch, matches that
    swallow a separator re-run per file; byte-identical to what the reference's jrep printed."""
    import jrep_tree
    exe = _build(tmp_path, "jrep")
    root = str(tmp_path / "tree")
    os.makedirs(root)
    paths = jrep_tree.make_tree(root)
    _check_jrep_against_golden(exe, root, paths, quick=True)


# ---- bench_engine --------------------------------------------------------------------
def _table(stdout: bytes):
    """Parses an engine's table the way tools/benchmarks/run.py:192-211 does."""
    lines = stdout.decode().split("\n")
    labels = lines[0].split()
    assert "text_size" in labels
    rows = {}
    for raw in lines[1:]:
        cells = raw.split()
        if cells:
            rows[int(cells[0])] = {labels[i]: float(v) for i, v in enumerate(cells[1:], start=1)}
    return labels, rows


@pytest.mark.skipif(not (os.path.exists(os.path.join(REF_INCLUDE, "rejit.h")) and
                         os.path.exists(os.path.join(REF_DIR, "bench_ref"))),
                    reason="needs the reference's header and oracle/_ref (build container only)")
def test_bench_engine_speaks_the_reference_harness_format(tmp_path):
    """samples/bench_engine.cc on the reference library against the reference's own engine binary:
    same labels line (byte for byte), same sizes, same column layout, for both table shapes."""
    exe = str(tmp_path / "bench_on_ref")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I" + REF_INCLUDE, os.path.join(ROOT, "samples", "bench_engine.cc"),
                    "-L" + REF_DIR, "-lrejit_ref", "-Wl,-rpath," + REF_DIR, "-o", exe], check=True)
    ref = os.path.join(REF_DIR, "bench_ref")
    for extra in ([], ["--run_worst_case=0"]):
        args = ["regexp", "--iterations=3", "--low_char=0", "--high_char=z", "--size=8,4096,1048576"] + extra
        a = subprocess.run([ref] + args, capture_output=True, check=True, timeout=RUN_TIMEOUT).stdout
        b = subprocess.run([exe] + args, capture_output=True, check=True, timeout=RUN_TIMEOUT).stdout
        assert a.split(b"\n")[0] == b.split(b"\n")[0]
        (la, ra), (lb, rb) = _table(a), _table(b)
        assert la == lb and sorted(ra) == sorted(rb) == [8, 4096, 1048576]
        assert [len(x) for x in a.split(b"\n")] == [len(x) for x in b.split(b"\n")]
        assert all(v > 0 for row in rb.values() for v in row.values())
    r = subprocess.run([exe, ""], capture_output=True, timeout=RUN_TIMEOUT)
    assert r.returncode == 1 and b"Cannot test an empty regular expression." in r.stdout


@pytest.mark.gpu
def test_sample_bench_engine_runs(tmp_path):
    exe = _build(tmp_path, "bench_engine")
    for extra in ([], ["--resident=1"]):
        r = subprocess.run([exe, "regexp", "--iterations=5", "--low_char=0", "--high_char=z", "--size=4096,4194304"] + extra,
                           capture_output=True, check=True, timeout=RUN_TIMEOUT)
        labels, rows = _table(r.stdout)
        assert labels == ["text_size", "worse", "amortised", "best"] and sorted(rows) == [4096, 4194304]
        assert all(v > 0 for row in rows.values() for v in row.values())


# ---- concurrent callers ----------------------------------------------------------------
@pytest.mark.gpu
def test_concurrent_match_all_on_shared_programs(tmp_path):
    """Eight threads, six compiled patterns shared by all of them (the reference's jrep calls MatchAll that
    way, sample/jrep.cc:461-493): every result equals the one computed by a single thread."""
    exe = _build(tmp_path, "threads")
    r = subprocess.run([exe, "8", "12"], capture_output=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith(b"ok "), (r.returncode, r.stdout[-300:], r.stderr[-300:])


def test_jrep_patterns_through_the_host_tables(hostsim, tmp_path):
    """What the GPU tier's jrep test will ask of the engine, asked of the product's own tables on the CPU
    (tests/hostsim.cc): every golden pattern and the line index over every file of the tree and over the
    batch blob (files joined by the separator), against the oracle."""
    import jrep_tree
    import rejit_oracle as O
    root = str(tmp_path / "tree")
    os.makedirs(root)
    paths = jrep_tree.make_tree(root)
    bodies = [open(os.path.join(root, p), "rb").read() for p in paths]
    blob = b"\n".join(b for b in bodies if b)
    for pat in sorted({c["re"] for c in _jrep_cases()} | {"^"}):
        o = O.Oracle(pat)
        for text in [blob] + [b for b in bodies if b]:
            got, desc = hostsim.match_all(pat, text)
            assert got == o.match_all(text), (pat, desc, len(text))


# ---- the samples end to end on the product's own front end, no GPU ---------------------------------
def _build_on_double(tmp_path, libdir, name):
    exe = str(tmp_path / (name + "_on_double"))
    subprocess.run(["g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "samples", name + ".cc"),
                    "-L" + libdir, "-lhostsim_rejit", "-Wl,-rpath," + libdir, "-lpthread", "-o", exe], check=True)
    return exe


def test_samples_end_to_end_on_the_host_tables(rejit_double, tmp_path):
    """The samples built as for librejit_b200.so (REJIT_B200 defined: pinned staging, rejit::Text, MatchAllParallel,
    device-side chains) but linked against tests/hostsim_rejit.cc, which answers through the product's parser,
    lowering and tables with the kernels emulated on the CPU: the library-specific branches of the samples and
    the front end are exercised together where no GPU exists.  jrep: every golden case in every batching mode,
    and sharded over "two devices"; regex-dna (both programs): the oracle's counts and lengths; threads."""
    import jrep_tree
    import rejit_oracle as O
    from rejit_b200 import workloads as W
    root = str(tmp_path / "tree")
    os.makedirs(root)
    paths = jrep_tree.make_tree(root)
    jrep = _build_on_double(tmp_path, rejit_double, "jrep")
    _check_jrep_against_golden(jrep, root, paths)
    # several devices: batches in rotation (one matcher thread per device, output in batch order), or every
    # batch cut into slabs (--shard)
    for case in _jrep_cases()[:4]:
        files = [p for p in paths if p.startswith(case.get("only", ""))]
        for extra in (["--gpus=2"], ["--gpus=3", "--batch-bytes=7000"], ["--gpus=2", "--shard"], ["--gpus=4", "--batch-bytes=0"]):
            r = subprocess.run([jrep] + case["options"] + extra + [case["re"]] + files, cwd=root, capture_output=True, timeout=RUN_TIMEOUT)
            assert r.returncode == 0 and r.stdout == case["stdout"].encode("latin-1"), (case["re"], extra, r.stderr[-200:])

    fa = W.fasta_file(3000)

    def replace(pat, text, w):
        out, at = bytearray(), 0
        for b, e in O.Oracle(pat).match_all(text):
            out += text[at:b] + w
            at = e
        return bytes(out + text[at:])

    seq = replace(W.STRIP_PATTERN, fa, b"")
    cur = seq
    for code, alt in W.IUB_SUBSTITUTIONS:
        cur = replace(code, cur, alt.encode())
    expected = "\n".join("%s %d" % (p, len(O.Oracle(p).match_all(seq))) for p in W.DNA_PATTERNS) + \
        "\n\n%d\n%d\n%d\n" % (len(fa), len(seq), len(cur))
    for name in ("regexdna", "regexdna_device"):
        r = subprocess.run([_build_on_double(tmp_path, rejit_double, name)], input=fa, capture_output=True, check=True, timeout=RUN_TIMEOUT)
        assert r.stdout.decode() == expected, name

    r = subprocess.run([_build_on_double(tmp_path, rejit_double, "threads"), "4", "2"], capture_output=True, timeout=RUN_TIMEOUT)
    assert r.returncode == 0 and r.stdout.startswith(b"ok "), r.stdout[-200:]
