"""Test configuration.

Tiers:
  * default (`-m "not gpu"`): runs anywhere — oracle vs golden vectors, the host
    front end and table builder vs the oracle (through tests/hostsim.cc, a CPU
    emulation of what the kernels do with the tables), C-ABI symbol checks,
    world_size-2 gloo test of the slab-stitching protocol, the samples end to end
    (on the reference's library, and on tests/hostsim_rejit.cc = the product's own
    front end behind the rejit.h surface with the kernels emulated).
  * `-m gpu`: the parity tests proper — every call goes through the C ABI of
    librejit_b200.so and runs the sm_100a kernels; results are compared with the
    oracle (oracle/) and the committed golden fixtures (tests/golden/).
"""
import ctypes
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_vectors():
    return json.load(open(os.path.join(GOLDEN, "matchall_offsets.json")))


@pytest.fixture(scope="session")
def ref_table():
    return json.load(open(os.path.join(GOLDEN, "ref_test_table.json")))


@pytest.fixture(scope="session")
def ir_dumps():
    return json.load(open(os.path.join(GOLDEN, "ir_dumps.json")))


@pytest.fixture(scope="session")
def hostsim():
    """tests/hostsim.cc built with g++ (CPU emulation of the kernels' table use)."""
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostsim.so")
    srcs = [os.path.join(ROOT, "tests", "hostsim.cc")] + [
        os.path.join(ROOT, "rejit_b200", "csrc", "host", f) for f in ("parser.cc", "lower.cc", "automaton.cc")]
    deps = srcs + [os.path.join(ROOT, "rejit_b200", "csrc", "cuda", "device_program.h"),
                   os.path.join(ROOT, "rejit_b200", "csrc", "host", "automaton.h"),
                   os.path.join(ROOT, "rejit_b200", "csrc", "host", "ir.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-o", so] + srcs)
    L = ctypes.CDLL(so)
    L.hostsim_match_all.restype = ctypes.c_int64
    L.hostsim_match_all.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p,
                                    ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64),
                                    ctypes.c_uint64, ctypes.c_char_p, ctypes.c_size_t]
    L.hostsim_match_all_slabs.restype = ctypes.c_int64
    L.hostsim_match_all_slabs.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint64,
                                          ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64]
    L.hostsim_match_full.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint64]

    L.hostsim_replace_all.restype = ctypes.c_int64
    L.hostsim_replace_all.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint64,
                                      ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint64,
                                      ctypes.POINTER(ctypes.c_uint64)]

    class Sim:
        def replace_all(self, pat, text, with_):
            pb = pat.encode("latin-1") if isinstance(pat, str) else pat
            cap = len(text) + (len(text) + 2) * len(with_) + 64
            out = ctypes.create_string_buffer(cap)
            n_out = ctypes.c_uint64()
            r = L.hostsim_replace_all(pb, len(pb), text, len(text), with_, len(with_), out, cap, ctypes.byref(n_out))
            if r < 0:
                return int(r), b""
            return int(r), out.raw[:n_out.value]

        def match_all(self, pat, text, strategy=-1, parser_opt=1):
            pb = pat.encode("latin-1") if isinstance(pat, str) else pat
            cap = len(text) + 2
            out = (ctypes.c_uint64 * (2 * cap))()
            d = ctypes.create_string_buffer(512)
            r = L.hostsim_match_all(pb, len(pb), parser_opt, text, len(text), strategy, out, cap, d, 512)
            desc = d.value.decode("latin-1")
            if r < 0:
                return int(r), desc
            return [(out[2 * i], out[2 * i + 1]) for i in range(r)], desc

        def match_all_slabs(self, pat, text, slabs):
            pb = pat.encode("latin-1") if isinstance(pat, str) else pat
            cap = len(text) + 2
            out = (ctypes.c_uint64 * (2 * cap))()
            r = L.hostsim_match_all_slabs(pb, len(pb), text, len(text), slabs, out, cap)
            return [(out[2 * i], out[2 * i + 1]) for i in range(max(r, 0))]

        def match_full(self, pat, text):
            pb = pat.encode("latin-1") if isinstance(pat, str) else pat
            return L.hostsim_match_full(pb, len(pb), text, len(text))

    return Sim()


@pytest.fixture(scope="session")
def rejit_double(hostsim):
    """tests/hostsim_rejit.cc: the include/rejit.h surface on top of tests/hostsim.cc (the product's own front end
    and tables, kernels emulated on the CPU), built as a shared library the samples can link instead of
    librejit_b200.so.  Test infrastructure only.  Returns the directory that holds libhostsim_rejit.so."""
    out_dir = os.path.join(ROOT, "tests", "_build")
    so = os.path.join(out_dir, "libhostsim_rejit.so")
    srcs = [os.path.join(ROOT, "tests", f) for f in ("hostsim_rejit.cc", "hostsim.cc")] + [
        os.path.join(ROOT, "rejit_b200", "csrc", "host", f) for f in ("parser.cc", "lower.cc", "automaton.cc")]
    deps = srcs + [os.path.join(ROOT, "include", "rejit.h"), os.path.join(ROOT, "rejit_b200", "csrc", "cuda", "device_program.h"),
                   os.path.join(ROOT, "rejit_b200", "csrc", "host", "automaton.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so] + srcs)
    return out_dir


def expand_table_row(row):
    """Expands one macro of the reference's test table the way its harness does
    (tools/tests/test.cc:665-715) into (match_type, text, expected, start, end)."""
    out = []
    pat, text = row["re"], row["text"]
    if row["kind"] == "full":
        out.append(("full", text, row["expected"], None, None))
        if row["expected"]:
            out.append(("first", text, 1, None, None))
            out.append(("all", text, 1, None, None))       # harness passes `expected` (==1) as the count
    elif row["kind"] in ("multiple", "multiple_unbound"):
        limit = 32 if row["kind"] == "multiple_unbound" else 0
        for i in range(limit + 1):
            t = (" " * i + text + " " * (limit - i)) if limit else text
            out.append(("first", t, row["expected"], row["start"] + i, row["end"] + i))
            out.append(("anywhere", t, row["expected"], None, None))
            out.append(("all", t, row["expected"], None, None))
    else:
        mt = {"kMatchAll": "all", "kMatchFirst": "first", "kMatchFull": "full",
              "kMatchAnywhere": "anywhere"}[row["match_type"]]
        out.append((mt, text, row["expected"], None, None))
    return pat, out
