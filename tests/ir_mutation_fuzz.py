"""tests/ir_mutation_fuzz.py <seed> <count> — run as a subprocess by tests/test_host_frontend.py (a crash must not take
pytest down): lowers random patterns, mutates the flat IR (state numbers, kinds, payload ranges, list order, header
fields) and hands it to rejit_b200_compile, which has to answer with a program or an error, never with a fault."""
import ctypes
import os
import random
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fuzzgen
L = ctypes.CDLL(os.path.join(ROOT, "rejit_b200", "librejit_b200.so"))
class Edge(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("kind", "entry_state", "exit_state", "payload_offset", "payload_length", "flags")]
class IR(ctypes.Structure):
    _fields_ = [("n_states", ctypes.c_int32), ("entry_state", ctypes.c_int32), ("exit_state", ctypes.c_int32),
                ("n_matching", ctypes.c_int32), ("n_control", ctypes.c_int32), ("edges", ctypes.POINTER(Edge)),
                ("payload", ctypes.POINTER(ctypes.c_uint8)), ("payload_length", ctypes.c_size_t)]
L.rejit_b200_parse.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.POINTER(IR)), ctypes.c_char_p, ctypes.c_size_t]
L.rejit_b200_compile.argtypes = [ctypes.POINTER(IR), ctypes.c_char_p, ctypes.c_size_t]
L.rejit_b200_compile.restype = ctypes.c_void_p
L.rejit_b200_program_free.argtypes = [ctypes.c_void_p]
L.rejit_b200_program_describe.argtypes = [ctypes.c_void_p]; L.rejit_b200_program_describe.restype = ctypes.c_char_p
r = random.Random(int(sys.argv[1])); N = int(sys.argv[2]); ok = bad = 0
err = ctypes.create_string_buffer(256)
for it in range(N):
    pat, _ = fuzzgen.rand_pattern(r)
    pb = pat.encode("latin-1")
    pir = ctypes.POINTER(IR)()
    if L.rejit_b200_parse(pb, len(pb), 1, ctypes.byref(pir), err, 256) != 0:
        continue
    src = pir.contents
    ne = src.n_matching + src.n_control
    edges = (Edge * max(ne, 1))()
    for i in range(ne):
        ctypes.memmove(ctypes.byref(edges[i]), ctypes.byref(src.edges[i]), ctypes.sizeof(Edge))
    pl = (ctypes.c_uint8 * max(src.payload_length, 1))()
    ctypes.memmove(pl, src.payload, src.payload_length)
    m = IR(src.n_states, src.entry_state, src.exit_state, src.n_matching, src.n_control, edges, pl, src.payload_length)
    for _ in range(r.randint(1, 3)):
        k = r.randint(0, 9)
        v = r.choice([-1, 0, 1, 2, 3, 5, 64, 255, 256, 4096, 70000, 2**31 - 1, -2**31])
        if k == 0: m.n_states = v if abs(v) < 100000 else m.n_states
        elif k == 1: m.entry_state = v
        elif k == 2: m.exit_state = v
        elif k == 3 and ne: edges[r.randrange(ne)].kind = v
        elif k == 4 and ne: edges[r.randrange(ne)].entry_state = v
        elif k == 5 and ne: edges[r.randrange(ne)].exit_state = v
        elif k == 6 and ne: edges[r.randrange(ne)].payload_offset = v
        elif k == 7 and ne: edges[r.randrange(ne)].payload_length = v
        elif k == 8 and src.payload_length: pl[r.randrange(src.payload_length)] = r.randrange(256)
        elif k == 9 and ne > 1:
            i, j = r.randrange(ne), r.randrange(ne)
            tmp = Edge(); ctypes.memmove(ctypes.byref(tmp), ctypes.byref(edges[i]), ctypes.sizeof(Edge))
            ctypes.memmove(ctypes.byref(edges[i]), ctypes.byref(edges[j]), ctypes.sizeof(Edge))
            ctypes.memmove(ctypes.byref(edges[j]), ctypes.byref(tmp), ctypes.sizeof(Edge))
    p = L.rejit_b200_compile(ctypes.byref(m), err, 256)
    if p:
        L.rejit_b200_program_describe(p); L.rejit_b200_program_free(p); ok += 1
    else:
        bad += 1
print("compiled", ok, "rejected", bad)
