#!/usr/bin/env python3
"""jrep on a GPU next to the reference's jrep on the host cores (reference: README.md:27-45,
sample/jrep.cc): a synthetic source tree of F files (rejit_b200.workloads.source_text_range cut into files), the
pattern ';\\n}' with -H -n -r.  Prints ONE JSON line: wall seconds and GB/s of both programs (page cache warm: the
tree was just written), the phases our jrep reports (JREP_TRACE=1) and whether the two outputs are the same bytes.
   usage: jrep_bench.py [--bytes N] [--files F] [--gpus G] [--keep]
The reference binary (oracle/_ref/jrep_ref) is the reference's own sample compiled by oracle/Makefile; it is run with
-j<nproc> for its time and with -j0 for the comparison (its threaded output order is not deterministic)."""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_jrep():
    exe = os.path.join(ROOT, "samples", "_build", "jrep")
    src = os.path.join(ROOT, "samples", "jrep.cc")
    lib = os.path.join(ROOT, "rejit_b200", "librejit_b200.so")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++11", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                               "-L", os.path.join(ROOT, "rejit_b200"), "-lrejit_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "rejit_b200")])
    return exe


def make_tree(root, total, files):
    from rejit_b200 import workloads as W
    per = total // files
    chunk_files = max(1, (1 << 27) // per)
    k = 0
    while k < files:
        m = min(chunk_files, files - k)
        blob = W.source_text_range(k * per, (k + m) * per).numpy().tobytes()
        for i in range(m):
            d = os.path.join(root, "d%02d" % ((k + i) % 64), "s%d" % ((k + i) // 64 % 4))
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "f%05d.c" % (k + i)), "wb") as f:
                f.write(blob[i * per:(i + 1) * per])
        k += m
    return per * files


def run(cmd, out_path, env=None, reps=2):
    best, err = 1e18, b""
    for _ in range(reps):
        with open(out_path, "wb") as out:
            t0 = time.perf_counter()
            p = subprocess.run(cmd, stdout=out, stderr=subprocess.PIPE, env=env)
            dt = time.perf_counter() - t0
        if p.returncode not in (0, 1):
            raise RuntimeError("%s exited with %d: %s" % (cmd[0], p.returncode, p.stderr[-300:]))
        if dt < best:
            best, err = dt, p.stderr
    return best, err.decode("latin-1")


def sha_sorted_lines(path):
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    lines.sort()
    return hashlib.sha256(b"\n".join(lines)).hexdigest(), len(lines)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--files", type=int, default=2048)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--keep", action="store_true")
    args = ap.parse_args()
    exe = build_jrep()
    ref = os.path.join(ROOT, "oracle", "_ref", "jrep_ref")
    base = tempfile.mkdtemp(prefix="jrep_bench_", dir=os.environ.get("RJ_TMP", "/dev/shm" if os.path.isdir("/dev/shm") else None))
    try:
        tree = os.path.join(base, "tree")
        os.makedirs(tree)
        total = make_tree(tree, args.bytes, args.files)
        pattern = ";\n}"
        stagers = max(1, min(16, (os.cpu_count() or 2) - 2))                # threads that read files into the pinned blob
        ours_cmd = [exe, "-H", "-n", "-r", "-j%d" % stagers, pattern, tree] + (["--gpus=%d" % args.gpus] if args.gpus > 1 else [])
        env = dict(os.environ, JREP_TRACE="1")
        run(ours_cmd, os.path.join(base, "warm.out"), env=env, reps=1)                   # context, kernels, page cache
        t_ours, trace = run(ours_cmd, os.path.join(base, "ours.out"), env=env)
        line = {"config": "jrep -H -n -r ';\\n}' over a source tree of %d files, %d MB (SURVEY §8f rank 3)" % (args.files, total // 1000000),
                "bytes": total, "files": args.files, "n_gpus": args.gpus,
                "ours": {"wall_s": round(t_ours, 4), "gbs": round(total / t_ours / 1e9, 3),
                         "trace": [ln for ln in trace.strip().split("\n") if ln][-6:]}}
        h_ours, n_lines = sha_sorted_lines(os.path.join(base, "ours.out"))
        line["output_lines"] = n_lines
        if os.path.exists(ref):
            ncpu = os.cpu_count() or 1
            t_ref, _ = run([ref, "-H", "-n", "-r", "-j%d" % ncpu, pattern, tree], os.path.join(base, "ref.out"))
            t_ref1, _ = run([ref, "-H", "-n", "-r", "-j0", pattern, tree], os.path.join(base, "ref0.out"), reps=1)
            h_ref, _ = sha_sorted_lines(os.path.join(base, "ref0.out"))
            with open(os.path.join(base, "ours.out"), "rb") as a, open(os.path.join(base, "ref0.out"), "rb") as b:
                same_order = a.read() == b.read()
            line["reference"] = {"wall_s": round(t_ref, 4), "gbs": round(total / t_ref / 1e9, 3), "threads": ncpu,
                                 "single_thread_wall_s": round(t_ref1, 4)}
            line["output_identical"] = same_order
            line["output_identical_as_sorted_lines"] = h_ours == h_ref
            line["speedup_vs_reference"] = round(t_ref / t_ours, 2)
        else:
            line["reference"] = None
        print(json.dumps(line), flush=True)
    finally:
        if not args.keep:
            shutil.rmtree(base, ignore_errors=True)


if __name__ == "__main__":
    main()
