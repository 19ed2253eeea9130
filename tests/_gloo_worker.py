"""Worker for tests/test_sharding_gloo.py: one rank of a world_size-N gloo job
running the slab-stitching protocol of rejit_b200/sharding.py (the one bench.py
uses under torchrun) with the per-slab resolve emulated on the CPU by
tests/hostsim.cc.  Prints one JSON line per case on rank 0."""
import ctypes
import json
import os
import random
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzzgen  # noqa: E402
from rejit_b200 import sharding  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    L = ctypes.CDLL(os.path.join(ROOT, "tests", "_build", "libhostsim.so"))
    L.hostsim_slab_run.restype = ctypes.c_int64
    L.hostsim_slab_run.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_uint64,
                                   ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64,
                                   ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    cases = json.loads(sys.argv[1])
    out = []
    exchange = None                      # default: the torch.distributed all-gather (gloo here, nccl under bench.py)
    if os.environ.get("RJ_STITCH") == "shm":
        exchange = sharding.ShmExchange(rank, world, 3 * 33, os.environ.get("MASTER_PORT", "0"))
        dist.barrier()
        exchange.attach()
    for pat, text_hex in cases:
        text = bytes.fromhex(text_hex)
        pb = pat.encode("latin-1")
        lo, hi = sharding.slab_bounds(len(text), world, rank)

        def run(cur, tail):
            oc, ot = ctypes.c_uint64(), ctypes.c_uint64()
            c = L.hostsim_slab_run(pb, len(pb), text, len(text), lo, hi, 1 if rank + 1 == world else 0,
                                   cur, tail, ctypes.byref(oc), ctypes.byref(ot))
            return int(c), int(oc.value), int(ot.value)
        total, rounds = sharding.stitched_count(dist, rank, world, lo, run, exchange=exchange)
        out.append([total, rounds])
    # pattern sets (bench.py's fused step): one all-gather carries every member's record
    set_out = []
    for pats, text_hex in (json.loads(sys.argv[2]) if len(sys.argv) > 2 else []):
        text = bytes.fromhex(text_hex)
        lo, hi = sharding.slab_bounds(len(text), world, rank)

        def run_set(carries):
            counts, couts = [], []
            for pat, (cur, tail) in zip(pats, carries):
                pb = pat.encode("latin-1")
                oc, ot = ctypes.c_uint64(), ctypes.c_uint64()
                c = L.hostsim_slab_run(pb, len(pb), text, len(text), lo, hi, 1 if rank + 1 == world else 0,
                                       cur, tail, ctypes.byref(oc), ctypes.byref(ot))
                counts.append(int(c))
                couts.append((int(oc.value), int(ot.value)))
            return counts, couts
        if os.environ.get("RJ_STITCH") == "p2p":
            # the neighbour protocol of the device-side stitch (one send right, one receive left), then ONE gather of
            # the per-rank counts and "did my answer change what I had sent"; a cascade repeats the step the old way
            import torch
            stitch = sharding.GlooNeighbourStitch(dist, rank, world)
            counts, cascaded = sharding.stitched_set_neighbour(stitch, lo, len(pats), run_set)
            mine = torch.tensor(counts + [1 if cascaded else 0], dtype=torch.int64)
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            if any(int(e[-1]) for e in every):
                totals, rounds = sharding.stitched_counts_set(dist, rank, world, lo, len(pats), run_set)
                rounds += 1
            else:
                totals, rounds = [sum(int(e[j]) for e in every) for j in range(len(pats))], 1
        else:
            totals, rounds = sharding.stitched_counts_set(dist, rank, world, lo, len(pats), run_set, exchange=exchange)
        set_out.append([totals, rounds])
    if rank == 0:
        print("RESULT " + json.dumps(out), flush=True)
        print("SETRESULT " + json.dumps(set_out), flush=True)
    if exchange is not None:
        dist.barrier()
        exchange.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
