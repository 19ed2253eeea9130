"""Worker for test_device_stitch (GPU tier, >= 2 GPUs): one rank per GPU under torchrun (backend nccl).  Every rank
owns a slab of one text resident on ITS device; the chain is stitched with the device-side neighbour exchange
(rejit_b200_stitch_*: a peer store over NVLink into the neighbour's HBM, CUDA IPC), and the summed counts must equal
the oracle's one-piece result — also when a match straddles or abuts a cut (redo) and when resolving again changes
what a rank had already sent (cascade: the step is repeated with the iterating all-gather protocol)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

import rejit_b200 as rj  # noqa: E402
from rejit_b200 import sharding  # noqa: E402
from rejit_b200 import workloads as W  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    tdev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=120), device_id=tdev)
    stitch = sharding.DeviceStitch(dist, rank, world, local, tdev)
    nccl = sharding.NcclExchange(dist, world, tdev)
    seq = bytearray(W.fasta_sequence(60000).tobytes())                  # 600 KB
    cut = len(seq) // world
    for c in range(cut, len(seq) - 16, cut):                            # members straddling and abutting every cut
        seq[c - 3:c + 5] = b"agggtaaa"
        seq[c + 5:c + 13] = b"tttaccct"
    cases = [(W.DNA_PATTERNS, bytes(seq)), (["aa", "aba", "ab"], b"ab" * 40000 + b"a" * 40001),
             (["abc", "bca", "cab"], b"abc" * 50000), (["needle"], (b"x" * 997 + b"needle") * 300),
             ([";\n}", "\n"], W.source_text_range(0, 300000).numpy().tobytes())]
    out = []
    for pats, text in cases:
        k = len(pats)
        lo, hi = sharding.slab_bounds(len(text), world, rank)
        halo = 64
        piece = np.frombuffer(text[lo:min(len(text), hi + halo)], dtype=np.uint8)
        dt = rj.DeviceText(piece, device=local)
        rs = rj.RegejSet(pats)

        def run(carries):
            cin = (rj.Carry * k)(*[rj.Carry(max(c - lo, 0), t - lo if t != sharding.NO_TAIL and t >= lo else sharding.NO_TAIL) for c, t in carries])
            cout = (rj.Carry * k)()
            own_end = (hi - lo) if rank + 1 < world else (1 << 62)
            cnts = rs.match_all_device(dt, own=(0, own_end), base_offset=lo, carry_in=cin, carry_out=cout)
            return cnts, [(cout[j].cur + lo, cout[j].tail + lo if cout[j].tail != sharding.NO_TAIL else sharding.NO_TAIL) for j in range(k)]
        counts, cascaded = sharding.stitched_set_neighbour(stitch, lo, k, run)
        mine = torch.tensor(list(counts) + [1 if cascaded else 0], dtype=torch.int64, device=tdev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        any_cascade = any(int(e[-1]) for e in every)
        if any_cascade:
            totals, _ = sharding.stitched_counts_set(dist, rank, world, lo, k, run, device=tdev, exchange=nccl)
        else:
            totals = [sum(int(e[j]) for e in every) for j in range(k)]
        out.append({"patterns": pats, "totals": totals, "cascaded": any_cascade})
        dt.free()
    stitch.close()
    if rank == 0:
        import rejit_oracle as O
        for (pats, text), o in zip(cases, out):
            o["expected"] = [len(O.Oracle(p).match_all(text)) for p in pats]
        print("STITCH " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
