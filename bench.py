#!/usr/bin/env python3
"""bench.py — GB/s of text scanned by MatchAll (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (config.workload): BASELINE.json configs[1] — the regex-dna
alternation set (the reference sample has NINE variants, sample/regexdna.cc:52-62)
counted over a 50 MB synthetic FASTA sequence.  A "step" = the nine patterns over
the text once.

  value ......... PHYSICAL rate: text bytes / step time (round 2; round 1 counted
                  the text once per pattern, k*N/time — kept as `value_per_pattern`).
                  Text resident in HBM, ONE fused set call per step (k_set_kmer);
                  step time = the call's device pipeline time (CUDA events on the
                  engine's stream, first launch to the counts being on the host).
                  The K timed steps run back to back between barrier + synchronize
                  brackets and rotate over six device copies of the text (300 MB > the
                  126 MB L2), so every step reads HBM-cold bytes without a flush kernel
                  in between; `ms_per_step_l2_flushed` is the round-1 protocol (one
                  buffer, L2 flushed before every call) measured in the same run.
  e2e ........... the text starts in pinned host memory: uploaded once per step
                  (H2D inside the timed region), one fused set call through the C ABI,
                  every match list copied back (D2H inside); host wall clock.
  e2e_dropin .... a C++ program (samples/e2e_dropin.cc) that calls the UNMODIFIED
                  signature Regej::MatchAll(const char*, size_t, ...) nine times on a
                  std::string (pageable memory), timed inside the program.
  roofline ...... the dominant kernel (k_set_kmer): algorithmic bytes (N + 16*M per
                  launch) / its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline .. the reference's own JIT (oracle/_ref, flag set "noreduce": its
                  fastest configuration that is correct on these patterns), 1 core.
  configs ....... one row per BASELINE.json configuration at its stated size:
                  C1 literal over 1 MiB (reference JIT on the CPU, and the engine),
                  C3 complex regex over 500 MB random text (without / with hits),
                  C4 jrep literal over a 5 GB source text (one text; 16 MiB batches),
                  C5 regex-dna chain (strip, nine counts, eleven substitutions) over a
                  5 GB FASTA file — each with GB/s (physical), the scan kernel's
                  roofline fraction, launches and an independent count check.
                  At N > 1 the 5 GB texts are strong-scaled: rank r owns slab r.
                  Last row (N = 1): samples/jrep.cc over a 1 GiB source tree of 2048
                  files next to the reference's own jrep (oracle/_ref/jrep_ref, all host
                  cores), outputs compared (scripts/jrep_bench.py; RJ_BENCH_JREP=0 skips).
                  RJ_BENCH_CONFIGS=0 skips the rows (the headline only).
  N > 1 ......... weak scaling of the headline: every rank owns one 50 MB slab (plus a
                  right halo) of an N*50 MB text.  The chain is stitched on the
                  DEVICES, inside the scan kernel: the CTA of k_set_kmer that reports the
                  result sends the chain state leaving its slab into the right
                  neighbour's HBM with a peer store over NVLink (inbox mapped through
                  CUDA IPC) and waits for the one arriving from the left
                  (config.stitch = "nvlink"; rejit_b200_match_all_set_device_stitched).
                  Reported next to it: the same exchange as a second launch (k_stitch,
                  `stitch_separate_launch`) and through an NCCL all-gather
                  (torch.distributed, `nccl_stitch`).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FASTA_N = 5_000_000            # -> 50,000,000 bytes per slab
HALO = 64
ROTATE = 6                     # device copies of the slab the timed steps rotate over (300 MB > L2)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ref_lib():
    so = os.path.join(ROOT, "oracle", "_ref", "librejit_ref.so")
    if not os.path.exists(so):
        return None
    L = ctypes.CDLL(so)
    L.ref_compile.restype = ctypes.c_void_p
    L.ref_compile.argtypes = [ctypes.c_char_p]
    L.ref_free.argtypes = [ctypes.c_void_p]
    L.ref_run_match_all.restype = ctypes.c_int64
    L.ref_run_match_all.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.ref_run_match_all_mt.restype = ctypes.c_int64
    L.ref_run_match_all_mt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t]
    return L


def cpu_reference_run(seq, patterns, threads, steps, sample_bytes, mode="proc"):
    """Times the reference JIT (or, if it is not built, the oracle port) over the
    first `sample_bytes` of the text; returns (GB/s, kind, cores, counts, sample).

    threads > 1: the compiled matcher is single-threaded per call and — measured
    on this image — calls made from several THREADS of one process do not run
    concurrently at all (8 threads, each with its own compiled Regej, take as long
    as 1; 8 processes scale linearly), so the all-cores number uses worker
    PROCESSES: the text is cut into contiguous slabs (+7 bytes of overlap, the
    longest match minus one), every worker compiles its own nine matchers, and a
    step is timed from "go" to the last worker's "done" (spin flags in shared
    memory).  Matches that begin in the overlap are not counted twice."""
    n = min(len(seq), sample_bytes)
    L = ref_lib()
    if L is None:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import rejit_oracle
        n = min(n, 2_000_000)
        data = seq[:n].tobytes()
        t0 = time.perf_counter()
        counts = [rejit_oracle.Oracle(p).match_all_count(data) for p in patterns]
        dt = time.perf_counter() - t0
        return len(patterns) * n / dt / 1e9, "port", 1, counts, "first %d bytes x %d patterns, one pass of the C oracle" % (n, len(patterns))
    L.ref_set_flagset(1)               # "noreduce": FF on, ff_reduce off
    if threads <= 1:
        handles = [L.ref_compile(p.encode()) for p in patterns]
        ptr = ctypes.c_void_p(seq.ctypes.data)
        best, counts = None, []
        for _ in range(max(1, steps)):
            t0 = time.perf_counter()
            counts = [int(L.ref_run_match_all(h, ptr, n)) for h in handles]
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        for h in handles:
            L.ref_free(h)
        return len(patterns) * n / best / 1e9, "reference", 1, counts, \
            "first %d bytes of the 50 MB text x %d patterns, 1 thread, best of %d passes, flags noreduce" % (n, len(patterns), max(1, steps))

    if mode == "omp":
        # threads of one process (OpenMP team, one privately compiled matcher per thread)
        handles = [L.ref_compile(p.encode()) for p in patterns]
        ptr = ctypes.c_void_p(seq.ctypes.data)
        best, counts = None, []
        for rep in range(max(1, steps) + 1):
            t0 = time.perf_counter()
            counts = [int(L.ref_run_match_all_mt(h, ptr, n, threads, 7)) for h in handles]
            dt = time.perf_counter() - t0
            if rep > 0:
                best = dt if best is None else min(best, dt)
        for h in handles:
            L.ref_free(h)
        return len(patterns) * n / best / 1e9, "reference", threads, counts, \
            "first %d bytes of the 50 MB text x %d patterns, %d OpenMP threads (4 slabs per thread), best of %d passes, flags noreduce" % (n, len(patterns), threads, max(1, steps))

    import multiprocessing as mp
    ctx = mp.get_context("fork")
    workers = threads
    go = ctx.RawValue("i", 0)
    done = ctx.RawArray("i", workers)
    cnt = ctx.RawArray("q", workers * len(patterns))
    base = seq.ctypes.data
    slab = (n + workers - 1) // workers

    def worker(w):
        lo, hi = min(n, w * slab), min(n, (w + 1) * slab)
        ext = min(n, hi + 7)
        hs = [L.ref_compile(p.encode()) for p in patterns]
        out = (ctypes.c_uint64 * 2048)()
        L.ref_match_all_handle.restype = ctypes.c_int64
        L.ref_match_all_handle.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
        step = 0
        while True:
            step += 1
            while go.value < step and go.value >= 0:
                pass
            if go.value < 0:
                os._exit(0)
            for k, h in enumerate(hs):
                cnt[w * len(patterns) + k] = L.ref_match_all_handle(h, ctypes.c_void_p(base + lo), ext - lo, hi - lo) if hi > lo else 0
            done[w] = step

    procs = [ctx.Process(target=worker, args=(w,), daemon=True) for w in range(workers)]
    for p in procs:
        p.start()
    best = None
    for step in range(1, max(1, steps) + 2):          # first step = warm-up (compile, page faults)
        t0 = time.perf_counter()
        go.value = step
        while any(done[w] < step for w in range(workers)):
            pass
        dt = time.perf_counter() - t0
        if step > 1:
            best = dt if best is None else min(best, dt)
    go.value = -1
    for p in procs:
        p.join(timeout=5)
    counts = [sum(cnt[w * len(patterns) + k] for w in range(workers)) for k in range(len(patterns))]
    return len(patterns) * n / best / 1e9, "reference", workers, counts, \
        "first %d bytes of the 50 MB text x %d patterns, %d worker processes (one slab each), best of %d steps, flags noreduce" % (n, len(patterns), workers, max(1, steps))


def traffic_of(kernel):
    """Measured DRAM bytes per launch of `kernel` (ncu --set full), or None."""
    try:
        here = os.path.dirname(os.path.abspath(__file__))
        return int(json.load(open(os.path.join(here, "profiles", "traffic.json")))[kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


# ---------------------------------------------------------------------------
# rows for the other BASELINE.json configurations
# ---------------------------------------------------------------------------
def _timed_calls(rj, regej, dt, reps, device, flush=True, **kw):
    """Device-resident MatchAll `reps` times (L2 flushed before each); returns (count, pipeline_ms, scan_ms, launches)."""
    st = rj.Stats()
    cnt = 0
    for _ in range(2):
        cnt = regej.match_all_device(dt, stats=st, **kw)
    tot = scan = 0.0
    for _ in range(reps):
        if flush:
            rj.lib().rejit_b200_flush_l2(device)
        cnt = regej.match_all_device(dt, stats=st, **kw)
        tot += st.total_ms
        scan += st.scan_ms
    return cnt, tot / reps, scan / reps, st.launches


def _row(name, n_bytes, matches, pipeline_ms, scan_ms, launches, peak, kernel, **extra):
    alg = n_bytes + 16.0 * matches
    row = {"config": name, "bytes": int(n_bytes), "matches": int(matches), "gbs": round(n_bytes / pipeline_ms / 1e6, 1),
           "pipeline_ms": round(pipeline_ms, 4), "launches": int(launches),
           "roofline": {"kernel": kernel, "bound": "hbm", "achieved": round(alg / scan_ms / 1e6, 1), "peak": peak,
                        "unit": "GB/s", "frac": round(alg / scan_ms / 1e6 / peak, 4), "avg_launch_ms": round(scan_ms, 4)}}
    row.update(extra)
    return row


def _device_text(rj, tensor, n, device):
    """A torch uint8 tensor (n bytes + 64 bytes of padding) as a device text the engine borrows."""
    return rj.DeviceText(nbytes=n, device=device, borrowed_ptr=tensor.data_ptr())


def _count_literal(torch, t, n, needle):
    """Occurrences of `needle` in t[:n] by shifted compares in torch (no engine code): (count, positions or None)."""
    total = 0
    step = 1 << 28
    m = len(needle)
    for lo in range(0, max(n - m + 1, 0), step):
        hi = min(n - m + 1, lo + step)
        ok = t[lo:hi] == needle[0]
        for i in range(1, m):
            ok &= t[lo + i:hi + i] == needle[i]
        total += int(ok.sum().item())
    return total


def config_rows(rj, W, torch, peak, device, rank, world, dist, tdev, reps):
    """One row per BASELINE.json configuration (see the module docstring).  Every row is independent: a failure
    becomes {"config": ..., "error": ...} instead of taking the bench line down."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    rows = []
    cuda = torch.device("cuda", device)

    def guarded(name, fn):
        try:
            out = fn()
            if out is not None:
                rows.extend(out if isinstance(out, list) else [out])
        except Exception as exc:                        # noqa: BLE001 — a row must not take the line down
            rows.append({"config": name, "error": "%s: %s" % (type(exc).__name__, str(exc)[:200])})
        torch.cuda.empty_cache()

    # ---- C1: literal over 1 MiB random ASCII — the reference JIT on the CPU (plumbing) and the engine --------
    def c1():
        if rank != 0:
            return None
        text = W.random_ascii(1 << 20, seed=1)
        L = ref_lib()
        ref = None
        if L is not None:
            L.ref_set_flagset(0)
            h = L.ref_compile(W.LITERAL_PATTERN.encode())
            best = None
            for _ in range(20):
                t0 = time.perf_counter()
                cnt_ref = int(L.ref_run_match_all(h, ctypes.c_void_p(text.ctypes.data), len(text)))
                dt_ = time.perf_counter() - t0
                best = dt_ if best is None else min(best, dt_)
            L.ref_free(h)
            ref = {"gbs": round(len(text) / best / 1e9, 3), "matches": cnt_ref, "cores": 1, "flags": "default"}
        r = rj.Regej(W.LITERAL_PATTERN)
        dt = rj.DeviceText(text, device=device)
        cnt, pms, sms, la = _timed_calls(rj, r, dt, reps, device, flush=False)
        dt.free()
        return _row("C1 literal 'regexp', 1 MiB random ASCII (L2 resident; the reference runs it on the CPU)", len(text), cnt,
                    pms, sms, la, peak, "k_scan_emit<literal>", reference_cpu=ref,
                    oracle_counts_equal=(ref is None or ref["matches"] == cnt))
    guarded("C1", c1)

    # ---- C3: complex regex over 500 MB random text, without and with hits ---------------------------------------
    def c3():
        n = 500_000_000
        if world > 1 and rank != 0:
            return None                                 # a single-GPU configuration
        import rejit_oracle
        t = torch.empty(n + 64, dtype=torch.uint8, device=cuda)
        t[:n] = W.random_ascii_range(0, n, device=cuda, seed=21)
        t[n:] = 0
        r = rj.Regej(W.COMPLEX_PATTERN)
        dt = _device_text(rj, t, n, device)
        out = []
        lit = torch.tensor(list(b"abcdefgh"), dtype=torch.uint8, device=cuda)
        natural = _count_literal(torch, t, n, lit)
        cnt, pms, sms, la = _timed_calls(rj, r, dt, reps, device)
        out.append(_row("C3 complex regex, 500 MB random text, no hits", n, cnt, pms, sms, la, peak, "k_scan_emit<window>",
                        oracle_counts_equal=(natural == 0 and cnt == 0), oracle="torch: the required literal does not occur"))
        # hits every ~10 kB: the three needles in turn
        pos = torch.arange(5003, n - 64, 10007, dtype=torch.int64, device=cuda)
        for j, nd in enumerate(W.COMPLEX_HITS):
            pj = pos[j::len(W.COMPLEX_HITS)]
            for k2, byte in enumerate(nd):
                t[pj + k2] = byte
        torch.cuda.synchronize(device)
        cnt, pms, sms, la = _timed_calls(rj, r, dt, reps, device)
        # the oracle on +-64 bytes around every planted needle (bounded: the first 3000 plants), and the plant count
        o = rejit_oracle.Oracle(W.COMPLEX_PATTERN)
        sample = pos[:3000].cpu().numpy()
        exp = 0
        for pp in sample:
            w = bytes(t[int(pp) - 64:int(pp) + 64].cpu().numpy())
            exp += len(o.match_all(w))
        hi_off = int(sample[-1]) + 64
        got_sample = r.match_all_device(_device_text(rj, t, hi_off, device))
        out.append(_row("C3 complex regex, 500 MB random text, a hit every ~10 kB", n, cnt, pms, sms, la, peak, "k_scan_emit<window>",
                        oracle_counts_equal=(exp == got_sample and cnt == int(pos.numel())),
                        oracle="oracle windows around the first 3000 plants (%d matches) + one match per plant (%d)" % (exp, int(pos.numel()))))
        del t
        return out
    guarded("C3", c3)

    # ---- C4: jrep literal over a 5 GB source text, sharded by slab ---------------------------------------------------
    def c4():
        total = 5_000_000_000
        slab = total // world
        lo, hi = rank * slab, (total if rank + 1 == world else (rank + 1) * slab)
        halo = 16 if rank + 1 < world else 0
        n = hi - lo + halo
        t = torch.empty(n + 64, dtype=torch.uint8, device=cuda)
        t[:n] = W.source_text_range(lo, lo + n, device=cuda)
        t[n:] = 0
        torch.cuda.synchronize(device)
        r = rj.Regej(W.JREP_PATTERN)
        dt = _device_text(rj, t, n, device)
        own = (0, (hi - lo) if rank + 1 < world else (1 << 62))
        if dist is not None:
            dist.barrier()
        cnt, pms, sms, la = _timed_calls(rj, r, dt, max(2, reps // 2), device, own=own, base_offset=lo)
        needle = torch.tensor(list(W.JREP_PATTERN.encode()), dtype=torch.uint8, device=cuda)
        exp = _count_literal(torch, t, (hi - lo) + (2 if rank + 1 < world else 0), needle)
        vals = torch.tensor([float(pms), float(cnt), float(exp), float(sms)], dtype=torch.float64, device=cuda)
        if dist is not None:
            allv = [torch.zeros_like(vals) for _ in range(world)]
            dist.all_gather(allv, vals)
            pms_max = max(float(v[0]) for v in allv)
            sms_max = max(float(v[3]) for v in allv)
            cnt_all, exp_all = int(sum(float(v[1]) for v in allv)), int(sum(float(v[2]) for v in allv))
        else:
            pms_max, sms_max, cnt_all, exp_all = pms, sms, cnt, exp
        out = []
        if rank == 0:
            out.append(_row("C4 jrep literal ';\\n}' over a 5 GB source text, one text, %d slab(s) of %d MB" % (world, slab // 1_000_000),
                            total, cnt_all, pms_max, sms_max, la, peak, "k_scan_emit<literal>", n_gpus=world,
                            oracle_counts_equal=(cnt_all == exp_all), oracle="torch shifted compares on every slab",
                            stitch="none needed: the literal cannot overlap itself, a match belongs to the slab it begins in"))
        # the same bytes as independent 16 MiB batches (jrep matches file by file: nothing spans two batches)
        if world == 1:
            piece = 16 << 20
            st = rj.Stats()
            tot_ms, got, exp_b = 0.0, 0, 0
            pieces = list(range(0, n, piece))
            for p0 in pieces[:64]:                      # bounded: 64 batches = 1 GiB
                ln = min(piece, n - p0)
                v = rj.DeviceText(nbytes=ln, device=device, borrowed_ptr=t.data_ptr() + p0)
                r.match_all_device(v, stats=st)
                got += r.match_all_device(v, stats=st)
                tot_ms += st.total_ms
                exp_b += _count_literal(torch, t[p0:p0 + ln], ln, needle)
            nb = min(len(pieces), 64)
            out.append({"config": "C4 per-file flavour: the first %d batches of 16 MiB, one MatchAll each" % nb, "bytes": nb * piece,
                        "matches": got, "gbs": round(nb * piece / tot_ms / 1e6, 1), "pipeline_ms": round(tot_ms, 3),
                        "launches": nb, "oracle_counts_equal": got == exp_b})
        del t
        return out
    guarded("C4", c4)

    # ---- C5: the regex-dna chain over a 5 GB FASTA file, sharded by slab ----------------------------------------------
    class _Cai:                                         # a device pointer as a torch tensor (zero copy)
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    def c5():
        n_lines = 500_000_000                           # fasta_file(500 M): 5.08 GB, 5.0 G letters
        size = W.fasta_file_size(n_lines)
        per = -(-size // world)
        lo, hi = rank * per, min(size, (rank + 1) * per)
        margin = 128

        def first_line_start(at):                       # smallest p >= at where a line begins
            if at == 0:
                return 0
            if at >= size:
                return size
            w = W.fasta_file_range(n_lines, at - 1, min(size, at - 1 + margin), device=cuda)
            return at + int((w == 10).nonzero()[0].item())
        begin, end = first_line_start(lo), first_line_start(hi)        # this rank owns the lines that begin in [lo, hi)
        n = end - begin
        t = torch.empty(n + 64, dtype=torch.uint8, device=cuda)
        t[:n] = W.fasta_file_range(n_lines, begin, end, device=cuda)
        t[n:] = 0
        torch.cuda.synchronize(device)
        raw = rj.Text(device=device, device_ptr=t.data_ptr(), nbytes=n)
        del t
        torch.cuda.empty_cache()
        strip = rj.Regej(W.STRIP_PATTERN)
        rs = rj.RegejSet(W.DNA_PATTERNS)
        pats = [c for c, _ in W.IUB_SUBSTITUTIONS]
        withs = [a.encode() for _, a in W.IUB_SUBSTITUTIONS]
        regs5 = [rj.Regej(c) for c in pats]
        best, letters = None, None
        for rep in range(2):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            st = rj.Stats()
            seq5, n_strip = strip.replace_all_text(raw, b"", stats=st)
            strip_ms = st.total_ms
            t1 = time.perf_counter()
            n_seq = len(seq5)
            if world > 1:
                # the "tiny all-gather": every rank's first 16 letters, so that a variant spanning a cut of the
                # sequence is counted by the rank it begins in (halo written into the text's padding)
                view = torch.as_tensor(_Cai(seq5.device_ptr(), n_seq + 16), device=cuda)
                head = view[:16].clone() if n_seq >= 16 else torch.zeros(16, dtype=torch.uint8, device=cuda)
                heads = [torch.zeros_like(head) for _ in range(world)]
                dist.all_gather(heads, head)
                torch.cuda.synchronize(device)
                if rank + 1 < world:
                    view[n_seq:n_seq + 16] = heads[rank + 1]
                    torch.cuda.synchronize(device)
                dt5 = rj.DeviceText(nbytes=n_seq + (16 if rank + 1 < world else 0), device=device, borrowed_ptr=seq5.device_ptr())
                counts = rs.match_all_device(dt5, stats=st, own=(0, n_seq if rank + 1 < world else (1 << 62)))
            else:
                counts = rs.match_all_text(seq5, stats=st)
            count_ms = st.total_ms
            t2 = time.perf_counter()
            outp, ks = rj.replace_all_set_text(regs5, seq5, withs, stats=st)
            iub_ms = st.total_ms
            t3 = time.perf_counter()
            res = {"wall_ms": (t3 - t0) * 1e3, "strip_ms": (t1 - t0) * 1e3, "count9_ms": (t2 - t1) * 1e3, "iub11_ms": (t3 - t2) * 1e3,
                   "device_ms": strip_ms + count_ms + iub_ms, "stripped": n_seq, "final": len(outp), "counts": counts, "iub": ks}
            if rep == 1:
                # independent of the engine: letter counts of the stripped text by torch
                view = torch.as_tensor(_Cai(seq5.device_ptr(), n_seq + (16 if rank + 1 < world else 0)), device=cuda)
                letters = torch.zeros(256, dtype=torch.int64, device=cuda)
                for c0 in range(0, n_seq, 1 << 28):
                    letters += torch.bincount(view[c0:min(n_seq, c0 + (1 << 28))].to(torch.int64), minlength=256)
                lit_a = _count_literal(torch, view, n_seq + (7 if rank + 1 < world else 0), torch.tensor(list(b"agggtaaa"), dtype=torch.uint8, device=cuda))
                lit_b = _count_literal(torch, view, n_seq + (7 if rank + 1 < world else 0), torch.tensor(list(b"tttaccct"), dtype=torch.uint8, device=cuda))
                res["torch_first_variant"] = lit_a + lit_b
                res["torch_letters"] = sum(int(letters[c]) for c in range(256))
                res["torch_final"] = n_seq + sum(int(letters[ord(c)]) * (len(w) - 1) for c, w in zip(pats, withs))
                res["torch_iub"] = [int(letters[ord(c)]) for c in pats]
            seq5.free()
            outp.free()
            if best is None or res["wall_ms"] < best["wall_ms"]:
                keep = {k: res[k] for k in res if k.startswith("torch_")} if rep == 1 else {}
                best = dict(res, **keep)
            if rep == 1:
                for k in ("torch_first_variant", "torch_final", "torch_iub", "torch_letters"):
                    best[k] = res[k]
        raw.free()
        ok_local = (best["final"] == best["torch_final"] and best["iub"] == best["torch_iub"] and
                    best["counts"][0] == best["torch_first_variant"])
        vals = torch.tensor([best["wall_ms"], best["device_ms"], float(n), float(best["stripped"]), float(best["final"]),
                             1.0 if ok_local else 0.0] + [float(c) for c in best["counts"]], dtype=torch.float64, device=cuda)
        if dist is not None:
            allv = [torch.zeros_like(vals) for _ in range(world)]
            dist.all_gather(allv, vals)
        else:
            allv = [vals]
        if rank != 0:
            return None
        wall = max(float(v[0]) for v in allv)
        dev_ms = max(float(v[1]) for v in allv)
        tot_in = int(sum(float(v[2]) for v in allv))
        stripped = int(sum(float(v[3]) for v in allv))
        final = int(sum(float(v[4]) for v in allv))
        counts_all = [int(sum(float(v[6 + j]) for v in allv)) for j in range(len(W.DNA_PATTERNS))]
        return {"config": "C5 regex-dna chain (strip, nine counts, eleven IUB substitutions) over a 5.08 GB FASTA file, %d slab(s)" % world,
                "bytes": size, "n_gpus": world, "gbs": round(size / wall / 1e6, 2), "wall_ms": round(wall, 3),
                "device_ms": round(dev_ms, 3), "gbs_device": round(size / dev_ms / 1e6, 2),
                "phases_ms_rank0": {k: round(best[k], 3) for k in ("strip_ms", "count9_ms", "iub11_ms")},
                "stripped": stripped, "final": final, "counts": counts_all,
                "oracle_counts_equal": (tot_in == size and stripped == 10 * n_lines and all(float(v[5]) == 1.0 for v in allv)),
                "oracle": "file arithmetic (every letter kept, every header / newline removed); torch on the stripped text: letter "
                          "histogram -> substitution counts and final length, shifted compares -> the first variant's count",
                "stitch": "slabs are cut at line ends; the nine counts use a 16-byte halo exchanged by ONE NCCL all-gather per step"
                          if world > 1 else "one slab",
                "launches": "strip 1 (the scan writes the stripped text itself; RJ_NO_FUSED_REBUILD=1: 1 scan + 5 rebuild), counts 1, substitutions 5"}
    guarded("C5", c5)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from rejit_b200 import workloads as W
    patterns = W.DNA_PATTERNS
    K = len(patterns)
    config = {"workload": "regex-dna alternation set (9 patterns, sample/regexdna.cc:52-62; BASELINE says 8) "
                          "over 50 MB synthetic FASTA sequence per GPU (BASELINE.json configs[1])",
              "text_bytes_per_gpu": FASTA_N * 10, "patterns": K,
              "l2": "the timed steps rotate over %d device copies of the slab (%d MB > the 126 MB L2): every step reads "
                    "HBM-cold text, no flush kernel between steps (ms_per_step_l2_flushed: the round-1 protocol)" % (ROTATE, ROTATE * 50),
              "parallelism": "slab%d" % args.gpus,
              "value_definition": "physical: text bytes / step time (the nine patterns share ONE pass); value_per_pattern = 9x"}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        seq = W.fasta_sequence(FASTA_N)
        ncpu = os.cpu_count() or 1
        sample = 50_000_000
        t0 = time.perf_counter()
        tried = []
        best = None
        widths = sorted({w for w in (8, 16, 32, 64, ncpu // 2, ncpu - 2, ncpu) if 1 < w <= ncpu})
        for mode in ("omp", "proc"):
            for w in widths:
                try:
                    r = cpu_reference_run(seq, patterns, w, max(1, args.steps), sample, mode=mode)
                except Exception as exc:          # a configuration that cannot run is skipped, not fatal
                    tried.append({"mode": mode, "width": w, "error": str(exc)[:80]})
                    continue
                tried.append({"mode": mode, "width": w, "gbs_per_pattern": round(r[0], 3)})
                if best is None or r[0] > best[0]:
                    best = r
        single = cpu_reference_run(seq, patterns, 1, 1, sample)
        tried.append({"mode": "single", "width": 1, "gbs_per_pattern": round(single[0], 3)})
        if best is None or single[0] > best[0]:
            best = single
        gbs_k, kind, cores, counts, what = best
        gbs = gbs_k / K                              # physical: the reference reads the text once per pattern
        wall = time.perf_counter() - t0
        line = {"impl": "reference", "metric": "GB/s text scanned (MatchAll)", "value": round(gbs, 4), "unit": "GB/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(min(len(seq), sample) / gbs / 1e6, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config, "gpu_launches": 0, "value_per_pattern": round(gbs_k, 4),
                "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": cores, "kind": kind, "sample": what},
                "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "match_counts": counts, "wall_s": round(wall, 2), "configurations_tried": tried}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm
    import numpy as np
    import torch
    import rejit_b200 as rj
    from rejit_b200 import sharding
    dist = None
    tdev = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        import datetime
        tdev = torch.device("cuda", local_rank)
        dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=240), device_id=tdev)
    if rj.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (rejit_b200 has no CPU fallback)")
    peak, peak_src = load_peaks()

    # every rank generates its own slab (+ the first bytes of its right neighbour as halo)
    seq = W.fasta_sequence(FASTA_N, seed=42 + rank)
    n_own = len(seq)
    if world > 1 and rank + 1 < world:
        nxt = W.fasta_sequence(HALO // 10 + 10, seed=42 + rank + 1)[:HALO]
        buf = np.concatenate([seq, nxt])
    else:
        buf = seq
    slab_lo = rank * n_own
    regs = [rj.Regej(p) for p in patterns]
    for r in regs:
        r.compile()
    dtext = rj.DeviceText(buf, device=local_rank)
    # the timed steps rotate over ROTATE device copies of the text (ROTATE * 50 MB > the 126 MB L2): every step reads
    # HBM-cold bytes without a flush kernel (and, at N > 1, without a barrier) between the steps
    rotation = [dtext] + [rj.DeviceText(buf, device=local_rank) for _ in range(ROTATE - 1)]
    step_no = [0]
    total_text = n_own * world
    rset = rj.RegejSet(regs)

    # ---- the stitch at N > 1: device-side neighbour exchange over NVLink; the NCCL all-gather is timed next to it
    stitch = None
    nccl_ex = None
    if world > 1:
        nccl_ex = sharding.NcclExchange(dist, world, tdev)
        if os.environ.get("RJ_STITCH", "nvlink") == "nvlink":
            try:
                stitch = sharding.DeviceStitch(dist, rank, world, local_rank, tdev)
            except Exception as exc:
                print("bench.py: device-side stitch unavailable (%s)" % exc, file=sys.stderr)
                stitch = None
            ok = torch.tensor([1 if stitch is not None else 0], dtype=torch.int64, device=tdev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                stitch = None
    config["stitch"] = ("nvlink" if stitch is not None else "nccl") if world > 1 else "none (one slab)"
    if world > 1:
        config["timing"] = ("step = device pipeline time of the rank's own call (CUDA events on the engine's stream, as at N=1); "
                            "the scan kernel's reporting CTA sends the chain states to the right neighbour and waits for the "
                            "left neighbour's INSIDE the kernel, so the events include the exchange and any rank skew; K steps "
                            "back to back between barrier + synchronize brackets, no barrier inside; max over ranks")

    def slab_run(stats, acc, dtext=dtext):
        def run(carries):
            cin = (rj.Carry * K)(*[rj.Carry(max(c - slab_lo, 0), t - slab_lo if t != sharding.NO_TAIL and t >= slab_lo else sharding.NO_TAIL)
                                   for c, t in carries])
            cout = (rj.Carry * K)()
            own_end = n_own if rank + 1 < world else (1 << 62)
            cnts = rset.match_all_device(dtext, stats=stats, own=(0, own_end), base_offset=slab_lo, carry_in=cin, carry_out=cout)
            acc[0] += stats.total_ms
            acc[1] += stats.scan_ms
            acc[2] += stats.launches
            outs = [(cout[j].cur + slab_lo, cout[j].tail + slab_lo if cout[j].tail != sharding.NO_TAIL else sharding.NO_TAIL)
                    for j in range(K)]
            return cnts, outs
        return run

    cascades = [0]
    fused_stitch = [os.environ.get("RJ_STITCH_FUSED", "1") != "0"]

    own_end = n_own if rank + 1 < world else (1 << 62)

    def one_step_fused(how=None):
        """One step: (step ms, scan ms, launches, this rank's counts)."""
        st = rj.Stats()
        if how == "flush":                                   # the round-1 protocol: one buffer, L2 flushed before the call
            rj.lib().rejit_b200_flush_l2(local_rank)
            cnts = rset.match_all_device(dtext, stats=st)
            return st.total_ms, st.scan_ms, st.launches, cnts
        dt_now = rotation[step_no[0] % ROTATE]
        step_no[0] += 1
        if world == 1:
            cnts = rset.match_all_device(dt_now, stats=st)
            return st.total_ms, st.scan_ms, st.launches, cnts
        acc = [0.0, 0.0, 0]
        run = slab_run(st, acc, dt_now)
        if how is None and stitch is not None and fused_stitch[0]:
            # scan + stitch in ONE kernel: the reporting CTA of k_set_kmer sends the chain states to the right
            # neighbour's HBM and waits for the left neighbour's; the device time of the call includes that wait
            cnts, couts, arrived, redo = rset.match_all_device_stitched(dt_now, (0, own_end), slab_lo, stats=st)
            acc = [st.total_ms, st.scan_ms, st.launches]
            if redo:
                want = [(max(arrived[j][0], slab_lo), arrived[j][1] if arrived[j][1] == slab_lo else sharding.NO_TAIL)
                        if (redo >> j) & 1 else (slab_lo, sharding.NO_TAIL) for j in range(K)]
                cnts, couts2 = run(want)
                acc[0] += st.total_ms
                acc[1] += st.scan_ms
                acc[2] += st.launches
                couts_g = [(c + slab_lo, t + slab_lo if t != sharding.NO_TAIL else sharding.NO_TAIL) for c, t in couts]
                if couts2 != couts_g:
                    cascades[0] += 1
            return acc[0], acc[1], acc[2], cnts
        t0 = time.perf_counter()
        if how == "nccl" or stitch is None:
            torch.cuda.synchronize(local_rank)
            dist.barrier()                                   # all ranks start the step together (untimed)
            t0 = time.perf_counter()
            t_run = [0.0]

            def timed_run(c):
                a = time.perf_counter()
                out = run(c)
                t_run[0] += time.perf_counter() - a
                return out
            totals, _ = sharding.stitched_counts_set(dist, rank, world, slab_lo, K, timed_run, device=tdev, exchange=nccl_ex)
            ex_s = (time.perf_counter() - t0) - t_run[0]
            return acc[0] + ex_s * 1e3, acc[1], acc[2], totals
        cnts, couts = run([(slab_lo, sharding.NO_TAIL)] * K)
        a = time.perf_counter()
        arrived, redo = stitch.exchange(couts, slab_lo)
        ex_s = time.perf_counter() - a
        if redo:
            want = [(max(arrived[j][0], slab_lo), arrived[j][1] if arrived[j][1] == slab_lo else sharding.NO_TAIL)
                    if (redo >> j) & 1 else (slab_lo, sharding.NO_TAIL) for j in range(K)]
            cnts, couts2 = run(want)
            if couts2 != couts:
                cascades[0] += 1
        return acc[0] + ex_s * 1e3, acc[1], acc[2], cnts

    def one_step_calls():
        """The same workload as nine separate MatchAll calls (N = 1 only)."""
        step_ms = scan_ms = 0.0
        launches, counts = 0, []
        for r in regs:
            rj.lib().rejit_b200_flush_l2(local_rank)
            st = rj.Stats()
            counts.append(r.match_all_device(dtext, stats=st))
            step_ms += st.total_ms
            scan_ms += st.scan_ms
            launches += st.launches
        return step_ms, scan_ms, launches, counts

    sampler = ClockSampler(local_rank) if rank == 0 else None     # covers warm-up + timed region
    warm = max(3, args.warmup)
    for _ in range(warm):
        one_step_fused()
    torch.cuda.synchronize(local_rank)
    if dist is not None:
        dist.barrier()
    # ---- headline: the fused set path: EXACTLY K steps between barrier + synchronize brackets -------------
    t_wall = time.perf_counter()
    f_ms = f_scan = 0.0
    f_launches = 0
    f_counts = []
    for _ in range(args.steps):
        ms, sc, la, f_counts = one_step_fused()
        f_ms += ms
        f_scan += sc
        f_launches += la
    torch.cuda.synchronize(local_rank)
    host_ms_per_step = (time.perf_counter() - t_wall) * 1e3 / args.steps      # host clock incl. the Python call overhead
    if dist is not None:
        t = torch.tensor([f_ms], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        f_ms = float(t.item())
        # the per-rank counts of the last step -> totals; did any stitch cascade
        t = torch.tensor(list(f_counts) + [cascades[0]], dtype=torch.int64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if stitch is not None:
            f_counts = [int(x) for x in t[:K].tolist()]
        n_cascades = int(t[K].item())
        dist.barrier()
    else:
        n_cascades = 0
    # the same step with the stitch as a second launch (k_stitch) after the scan (N > 1 only)
    sep_ms = None
    if world > 1 and stitch is not None and fused_stitch[0]:
        fused_stitch[0] = False
        for _ in range(3):
            one_step_fused()
        sep_ms = 0.0
        for _ in range(args.steps):
            ms, _, _, _ = one_step_fused()
            sep_ms += ms
        fused_stitch[0] = True
        t = torch.tensor([sep_ms], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sep_ms = float(t.item()) / args.steps
        dist.barrier()
    # the round-1 protocol at N = 1: one buffer, the L2 flushed before every call (cross-check of the rotation)
    flush_ms = None
    if world == 1:
        for _ in range(2):
            one_step_fused("flush")
        flush_ms = sum(one_step_fused("flush")[0] for _ in range(args.steps)) / args.steps
    # the same fused step with the stitch records through an NCCL all-gather (N > 1 only)
    nccl_ms = None
    nccl_counts = None
    if world > 1 and stitch is not None:
        for _ in range(3):
            one_step_fused("nccl")
        nccl_ms = 0.0
        for _ in range(args.steps):
            ms, _, _, nccl_counts = one_step_fused("nccl")
            nccl_ms += ms
        t = torch.tensor([nccl_ms], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nccl_ms = float(t.item()) / args.steps
        dist.barrier()
    wall = time.perf_counter() - t_wall
    # ---- nine separate calls (N = 1) ---------------------------------------------------
    calls = None
    if world == 1:
        for _ in range(2):
            one_step_calls()
        c_ms = c_scan = 0.0
        c_launches = 0
        for _ in range(args.steps):
            ms, sc, la, c_counts = one_step_calls()
            c_ms += ms
            c_scan += sc
            c_launches += la
        ms_calls = c_ms / args.steps
        alg_call = len(buf) + 16.0 * sum(c_counts) / K
        scan_call = c_scan / (args.steps * K)
        calls = {"value": round(total_text / ms_calls / 1e6, 3), "unit": "GB/s (physical: the text is read nine times per step)",
                 "ms_per_step": round(ms_calls, 4), "gpu_launches": c_launches, "counts_equal_fused": c_counts == f_counts,
                 "roofline": {"kernel": "k_dfa_tma", "achieved": round(alg_call / scan_call / 1e6, 2),
                              "frac": round(alg_call / scan_call / 1e6 / peak, 4), "avg_launch_ms": round(scan_call, 5)}}
    if sampler and wall < 0.5:
        # the timed region is only milliseconds long: keep the same load running until nvidia-smi (100 ms period) has
        # seen it a few times (rank 0 only, nothing here enters a collective: the local fused call, no stitch)
        t_end = time.perf_counter() + 0.6
        st_keep = rj.Stats()
        while time.perf_counter() < t_end:
            rset.match_all_device(dtext, stats=st_keep)
    clocks = sampler.stop() if sampler else None
    f_ms_per_step = f_ms / args.steps
    value = total_text / (f_ms_per_step / 1e3) / 1e9

    # ---- e2e: host buffers in, match lists out -------------------------------------
    L = rj.lib()
    pinned = L.rejit_b200_pinned_alloc(n_own)
    ctypes.memmove(pinned, seq.ctypes.data, n_own)
    err = ctypes.create_string_buffer(256)
    e2e_d2h = [0]

    def e2e_step_fused():
        handle = L.rejit_b200_text_upload(local_rank, pinned, n_own, err, 256)
        if not handle:
            raise SystemExit(err.value.decode())
        cnts = (ctypes.c_int64 * K)()
        prs = (ctypes.POINTER(ctypes.c_uint64) * K)()
        if L.rejit_b200_match_all_set_text(rset._set, handle, cnts, prs, None, err, 256) != 0:
            raise SystemExit(err.value.decode())
        e2e_d2h[0] = 16 * sum(cnts) + 64 * K
        for j in range(K):
            L.rejit_b200_free(prs[j])
        L.rejit_b200_text_free(handle)

    def e2e_step_percall():
        for r in regs:
            pairs = ctypes.POINTER(ctypes.c_uint64)()
            k = L.rejit_b200_match_all_alloc(r._prog, pinned, n_own, ctypes.byref(pairs), None, err, 256)
            if k < 0:
                raise SystemExit(err.value.decode())
            L.rejit_b200_free(pairs)

    def time_e2e(fn, reps):
        for _ in range(2):
            fn()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dt_ = (time.perf_counter() - t0) / reps
        if dist is not None:
            t = torch.tensor([dt_], dtype=torch.float64, device=tdev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_ = float(t.item())
        return total_text / dt_ / 1e9

    e2e_reps = max(3, min(args.steps, 10))
    e2e_fused = time_e2e(e2e_step_fused, e2e_reps)
    e2e_percall = time_e2e(e2e_step_percall, 3) if world == 1 else None
    L.rejit_b200_pinned_free(pinned)

    # ---- e2e through the unmodified C++ signature, pageable memory (N = 1) -------------------------------
    dropin = None
    if world == 1:
        try:
            exe = os.path.join(ROOT, "samples", "_build", "e2e_dropin")
            src = os.path.join(ROOT, "samples", "e2e_dropin.cc")
            if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
                os.makedirs(os.path.dirname(exe), exist_ok=True)
                subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                                       "-L", os.path.join(ROOT, "rejit_b200"), "-lrejit_b200",
                                       "-Wl,-rpath," + os.path.join(ROOT, "rejit_b200")])
            fa_path = os.path.join(ROOT, "samples", "_build", "seq50.bin")
            seq.tofile(fa_path)
            outp = subprocess.run([exe, fa_path, str(e2e_reps)], capture_output=True, text=True, timeout=300)
            os.remove(fa_path)
            dropin = json.loads(outp.stdout.strip().split("\n")[-1])
        except Exception as exc:                                   # noqa: BLE001
            dropin = {"error": str(exc)[:200]}

    # ---- the other configurations ----------------------------------------------------------------------
    rows = None
    if os.environ.get("RJ_BENCH_CONFIGS", "1") != "0":
        rows = config_rows(rj, W, torch, peak, local_rank, rank, world, dist, tdev, max(3, min(args.steps, 5)))

    # ---- jrep on the GPU next to the reference's jrep on the host cores (N = 1) ---------
    if rows is not None and world == 1 and os.environ.get("RJ_BENCH_JREP", "1") != "0":
        try:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "jrep_bench.py"),
                                  "--bytes", os.environ.get("RJ_BENCH_JREP_BYTES", str(1 << 30)), "--files", "2048"],
                                 capture_output=True, text=True, timeout=600)
            rows.append(json.loads(out.stdout.strip().split("\n")[-1]))
        except Exception as exc:                                   # noqa: BLE001
            rows.append({"config": "jrep", "error": str(exc)[:200]})

    # ---- MatchAllParallel on rank 0 over all N devices, against the compiled reference (N > 1) -----------
    parallel = None
    if world > 1:
        dist.barrier()
        if rank == 0:
            try:
                whole = np.concatenate([W.fasta_sequence(FASTA_N, seed=42 + r2) for r2 in range(world)])
                Lr = ref_lib()
                parallel = {"n_gpus": world, "patterns": []}
                for p in (patterns[0], patterns[4]):
                    t0 = time.perf_counter()
                    got = rj.Regej(p).match_all_array(whole, n_gpus=world)
                    dt_ = time.perf_counter() - t0
                    exp = None
                    if Lr is not None:
                        Lr.ref_set_flagset(1)
                        h = Lr.ref_compile(p.encode())
                        exp = int(Lr.ref_run_match_all_mt(h, ctypes.c_void_p(whole.ctypes.data), len(whole), min(16, os.cpu_count() or 1), 7))
                        Lr.ref_free(h)
                    parallel["patterns"].append({"pattern": p, "matches": int(got.shape[0]), "reference_matches": exp,
                                                 "equal": exp is None or exp == int(got.shape[0]), "wall_ms": round(dt_ * 1e3, 2)})
                parallel["fused_totals_equal_reference"] = None
                if Lr is not None:
                    Lr.ref_set_flagset(1)
                    tot = []
                    for p in patterns:
                        h = Lr.ref_compile(p.encode())
                        tot.append(int(Lr.ref_run_match_all_mt(h, ctypes.c_void_p(whole.ctypes.data), len(whole), min(16, os.cpu_count() or 1), 7)))
                        Lr.ref_free(h)
                    parallel["fused_totals_equal_reference"] = (tot == f_counts)
                    parallel["reference_totals"] = tot
            except Exception as exc:                               # noqa: BLE001
                parallel = {"error": str(exc)[:200]}
        dist.barrier()

    if stitch is not None:
        stitch.close()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    m_total = sum(f_counts) if world == 1 else None
    # roofline of the set kernel on THIS rank's launches
    own_counts = f_counts if world == 1 else None
    f_scan_avg = f_scan / max(1, f_launches) if world > 1 else f_scan / args.steps
    m_rank = sum(f_counts) / world if world > 1 else sum(f_counts)
    alg_bytes = len(buf) + 16.0 * m_rank
    achieved = alg_bytes / (f_scan_avg / 1e3) / 1e9
    cpu = None
    if world == 1:
        gbs, kind, cores, ccounts, what = cpu_reference_run(seq, patterns, 1, 2, 50_000_000)
        cpu = {"value": round(gbs / K, 4), "unit": "GB/s", "cores": cores, "kind": kind, "sample": what,
               "value_per_pattern": round(gbs, 4), "match_counts_equal": ccounts == f_counts}
    set_kernel = "k_set_kmer" if "k-mer index" in rset.describe() and not os.environ.get("RJ_NO_KMER") else "k_set_tma"
    config["how"] = ("the nine patterns are fused into one scan (rejit_b200_match_all_set_device): the text is read ONCE per step")
    line = {"metric": "GB/s text scanned (MatchAll)", "value": round(value, 3), "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": round(f_ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "value_per_pattern": round(K * value, 3),
            "per_pattern_calls": calls,
            "e2e": {"value": round(e2e_fused, 3), "unit": "GB/s", "h2d_bytes_per_step": n_own,
                    "d2h_bytes_per_step": int(e2e_d2h[0]), "api": "rejit_b200_text_upload + rejit_b200_match_all_set_text (C ABI)",
                    "how": "text uploaded once per step from pinned host memory (H2D inside the timed region), one fused "
                           "set call, every match list copied back (D2H inside)"},
            "e2e_per_call_upload": None if e2e_percall is None else
            {"value": round(e2e_percall, 3), "unit": "GB/s", "h2d_bytes_per_step": K * n_own,
             "api": "rejit_b200_match_all_alloc x 9 (pinned source)"},
            "e2e_dropin": dropin,
            "nccl_stitch": None if nccl_ms is None else
            {"value": round(total_text / (nccl_ms / 1e3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(nccl_ms, 4),
             "counts_equal": nccl_counts == f_counts,
             "how": "the same fused step with the stitch records exchanged by an NCCL all-gather (torch.distributed) instead "
                    "of the device-side neighbour exchange"},
            "stitch_separate_launch": None if sep_ms is None else
            {"value": round(total_text / (sep_ms / 1e3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(sep_ms, 4),
             "how": "scan kernel, then k_stitch as a second launch (host time inside the exchange added to the device time)"},
            "ms_per_step_l2_flushed": None if flush_ms is None else round(flush_ms, 4),
            "ms_per_step_host_clock": round(host_ms_per_step, 4),
            "stitch_cascades": n_cascades,
            "parallel_parity": parallel,
            "gpu_launches": f_launches,
            "roofline": {"bound": "hbm", "kernel": set_kernel, "achieved": round(achieved, 2), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic_of(set_kernel),
                         "traffic_source": "profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one "
                                           "ncu --set full capture of this workload)",
                         "note": "one launch = the scan of the text for all nine patterns plus the in-kernel finish "
                                 "(exact check of the hits, grid-wide exchange of the counts, matches written at their "
                                 "final place, report to the host); a 50 MB text is 7.7 us of HBM time: the fixed costs "
                                 "(first byte after ~5 us, ordered finish ~10 us) dominate at this size, see the 625 MB row "
                                 "of `configs` and DESIGN.md",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                         "avg_launch_ms": round(f_scan_avg, 5)},
            "cpu_baseline": cpu, "clocks": clocks, "match_counts": f_counts, "configs": rows, "wall_s": round(wall, 2)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
