"""Slab sharding of one text over the ranks of a torch.distributed job, and the
chain stitch at slab boundaries.

Every rank scans the starts that fall inside its own contiguous slab (its text
buffer carries a right halo) and resolves its matches as if nothing reached in
from the left.  One all-gather of a 4-word record per rank — (carry.cur,
carry.tail, match count, slab begin) — then tells every rank whether the chain
arriving from its left neighbour differs from that assumption; only in that
case (a match straddling or abutting the boundary) does the rank resolve again
with the real carry, and only then is a further round needed.
"""
from __future__ import annotations

from typing import Callable, Tuple

NO_TAIL = (1 << 64) - 1


def slab_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    lo = (total // world) * rank
    hi = total if rank + 1 == world else (total // world) * (rank + 1)
    return lo, hi


def stitched_count(dist, rank: int, world: int, slab_lo: int,
                   run: Callable[[int, int], Tuple[int, int, int]], device=None) -> Tuple[int, int]:
    """run(carry_cur, carry_tail) -> (count, carry_out_cur, carry_out_tail), all
    offsets global.  Returns (global match count, collective rounds used)."""
    import torch
    used = (slab_lo, NO_TAIL)
    count, ccur, ctail = run(*used)
    rounds = 0
    while True:
        rec = torch.tensor([ccur, ctail if ctail != NO_TAIL else -1, count, slab_lo], dtype=torch.int64,
                           device=device)
        if world > 1:
            gathered = [torch.zeros_like(rec) for _ in range(world)]
            dist.all_gather(gathered, rec)
            rows = [g.tolist() for g in gathered]
        else:
            rows = [rec.tolist()]
        rounds += 1
        # every rank evaluates the same predicate for every rank
        redo = []
        for r in range(1, world):
            left_cur, left_tail = rows[r - 1][0], rows[r - 1][1]
            lo_r = rows[r][3]
            want = (max(left_cur, lo_r), left_tail if left_tail == lo_r else -1)
            redo.append(want)
        changed = False
        if rank > 0:
            want = redo[rank - 1]
            want_carry = (want[0], want[1] if want[1] != -1 else NO_TAIL)
            if want_carry != used:
                used = want_carry
                count, ccur, ctail = run(*used)
                changed = True
        flag = torch.tensor([1 if changed else 0], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()) == 0:
            return sum(int(r[2]) for r in rows), rounds


def stitched_counts_set(dist, rank: int, world: int, slab_lo: int, k: int, run, device=None):
    """Set version: run(carries) -> (counts[k], carries_out[k]) with carries as
    lists of (cur, tail) in global offsets.  One all-gather of a [k, 3] block per
    rank; a rank resolves again only if some member's arriving chain differs.
    Returns (global counts[k], collective rounds)."""
    import torch
    used = [(slab_lo, NO_TAIL)] * k
    counts, couts = run(used)
    rounds = 0
    while True:
        rec = torch.tensor([[couts[j][0], couts[j][1] if couts[j][1] != NO_TAIL else -1, counts[j]] for j in range(k)] +
                           [[slab_lo, 0, 0]], dtype=torch.int64, device=device)
        if world > 1:
            gathered = [torch.zeros_like(rec) for _ in range(world)]
            dist.all_gather(gathered, rec)
            rows = [g.tolist() for g in gathered]
        else:
            rows = [rec.tolist()]
        rounds += 1
        changed = False
        if rank > 0:
            lo_r = rows[rank][k][0]
            want = []
            for j in range(k):
                left_cur, left_tail = rows[rank - 1][j][0], rows[rank - 1][j][1]
                want.append((max(left_cur, lo_r), left_tail if left_tail == lo_r else NO_TAIL))
            if want != used:
                used = want
                counts, couts = run(used)
                changed = True
        flag = torch.tensor([1 if changed else 0], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()) == 0:
            return [sum(int(r[j][2]) for r in rows) for j in range(k)], rounds
