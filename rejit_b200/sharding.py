"""Slab sharding of one text over the ranks of a torch.distributed job, and the
chain stitch at slab boundaries.

Every rank scans the starts that fall inside its own contiguous slab (its text
buffer carries a right halo) and resolves its matches as if nothing reached in
from the left.  One all-gather of a 4-word record per rank — (carry.cur,
carry.tail, match count, slab begin) — then tells every rank whether the chain
arriving from its left neighbour differs from that assumption; only in that
case (a match straddling or abutting the boundary) does the rank resolve again
with the real carry, and only then is a further round needed.  There is no
second collective: all ranks evaluate the same predicate on the same rows.
"""
from __future__ import annotations

from typing import Callable, Tuple

NO_TAIL = (1 << 64) - 1


def slab_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    lo = (total // world) * rank
    hi = total if rank + 1 == world else (total // world) * (rank + 1)
    return lo, hi


class NcclExchange:
    """The all-gather as a torch.distributed collective (backend nccl: the record
    goes host -> device -> NVLink -> device -> host; backend gloo in the CPU tests)."""
    name = "nccl"

    def __init__(self, dist, world: int, device=None):
        self.dist, self.world, self.device = dist, world, device

    def __call__(self, rec):
        import torch
        t = torch.tensor(rec, dtype=torch.int64, device=self.device)
        if self.world == 1:
            return [t.tolist()]
        gathered = [torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(gathered, t)
        return [g.tolist() for g in gathered]


class ShmExchange:
    """The same all-gather through a POSIX shared-memory mailbox, for ranks on ONE
    box (the only layout this engine shards over, SURVEY.md \u00a78e).  The records are
    produced in host memory (the engine reports through mapped pinned memory), so a
    host-to-host exchange skips two PCIe copies, the collective's launch and a
    stream synchronisation: ~2 us instead of ~100 us per round, which matters when
    the whole step is 40 us.

    Layout: [world][2][1 + words] int64; a rank writes its record into the buffer
    (seq & 1), then the sequence number; readers spin until every rank shows the
    current sequence number.  Two buffers suffice: nobody can finish exchange s+1
    before everybody has written it, i.e. before everybody has read exchange s."""
    name = "shm"

    def __init__(self, rank: int, world: int, words: int, key: str, timeout_s: float = 60.0):
        import platform
        import time
        import numpy as np
        from multiprocessing import shared_memory
        # record, then sequence number, with plain numpy stores: that order is only kept by x86 (total store order);
        # on a weakly ordered host (Grace) a reader could see the new number with a stale record
        if platform.machine().lower() not in ("x86_64", "amd64", "i686", "i386"):
            raise RuntimeError("ShmExchange relies on x86 store ordering: use the device-side stitch or NcclExchange on %s"
                               % platform.machine())
        self.rank, self.world, self.words = rank, world, words
        size = world * 2 * (1 + words) * 8
        name = "rejit_b200_" + key
        if rank == 0:
            try:
                stale = shared_memory.SharedMemory(name=name)
                stale.close()
                stale.unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=size)
            self.shm.buf[:size] = bytes(size)
        else:
            self.shm = None
        self._name, self._size, self._timeout = name, size, timeout_s
        self._np, self._time = np, time
        self.a = None
        self.seq = 0

    def attach(self):
        """Call after a barrier that follows rank 0's constructor."""
        from multiprocessing import shared_memory
        if self.shm is None:
            self.shm = shared_memory.SharedMemory(name=self._name)
        self.a = self._np.ndarray((self.world, 2, 1 + self.words), dtype=self._np.int64, buffer=self.shm.buf)

    def __call__(self, rec):
        flat = self._np.asarray(rec, dtype=self._np.int64).reshape(-1)
        assert flat.size <= self.words
        self.seq += 1
        b = self.seq & 1
        mine = self.a[self.rank, b]
        mine[1:1 + flat.size] = flat
        mine[0] = self.seq                                   # published (x86 keeps the store order)
        deadline = self._time.perf_counter() + self._timeout
        col = self.a[:, b, 0]
        while not (col == self.seq).all():
            if self._time.perf_counter() > deadline:
                raise TimeoutError("ShmExchange: a rank did not arrive (seq %d, seen %s)" % (self.seq, col.tolist()))
        shape = self._np.asarray(rec).shape
        return [self.a[r, b, 1:1 + flat.size].reshape(shape).tolist() for r in range(self.world)]

    def close(self):
        try:
            self.a = None
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass


def stitched_count(dist, rank: int, world: int, slab_lo: int,
                   run: Callable[[int, int], Tuple[int, int, int]], device=None, exchange=None) -> Tuple[int, int]:
    """run(carry_cur, carry_tail) -> (count, carry_out_cur, carry_out_tail), all
    offsets global.  Returns (global match count, collective rounds used).

    ONE collective per round: every rank sees every rank's record and therefore
    knows, without asking, which ranks have to resolve again (the carry a rank
    assumed is a function of the records of the round before)."""
    exchange = exchange or NcclExchange(dist, world, device)
    count, ccur, ctail = run(slab_lo, NO_TAIL)
    used = None                       # the carry every rank resolved with; known after the first gather
    rounds = 0
    while True:
        rows = exchange([ccur, ctail if ctail != NO_TAIL else -1, count, slab_lo])
        rounds += 1
        if used is None:
            used = [(rows[r][3], NO_TAIL) for r in range(world)]
        changed = False
        for r in range(1, world):
            left_cur, left_tail = rows[r - 1][0], rows[r - 1][1]
            lo_r = rows[r][3]
            want = (max(left_cur, lo_r), left_tail if left_tail == lo_r else NO_TAIL)
            if want != used[r]:
                used[r] = want
                changed = True
                if r == rank:
                    count, ccur, ctail = run(*want)
        if not changed:
            return sum(int(r[2]) for r in rows), rounds


def stitched_counts_set(dist, rank: int, world: int, slab_lo: int, k: int, run, device=None, exchange=None):
    """Set version: run(carries) -> (counts[k], carries_out[k]) with carries as
    lists of (cur, tail) in global offsets.  One all-gather of a [k + 1, 3] block
    per rank and round; a rank resolves again only if some member's arriving
    chain differs.  Returns (global counts[k], collective rounds)."""
    exchange = exchange or NcclExchange(dist, world, device)
    counts, couts = run([(slab_lo, NO_TAIL)] * k)
    used = None
    rounds = 0
    while True:
        rows = exchange([[couts[j][0], couts[j][1] if couts[j][1] != NO_TAIL else -1, counts[j]] for j in range(k)] +
                        [[slab_lo, 0, 0]])
        rounds += 1
        if used is None:
            used = [[(rows[r][k][0], NO_TAIL)] * k for r in range(world)]
        changed = False
        for r in range(1, world):
            lo_r = rows[r][k][0]
            want = []
            for j in range(k):
                left_cur, left_tail = rows[r - 1][j][0], rows[r - 1][j][1]
                want.append((max(left_cur, lo_r), left_tail if left_tail == lo_r else NO_TAIL))
            if want != used[r]:
                used[r] = want
                changed = True
                if r == rank:
                    counts, couts = run(want)
        if not changed:
            return [sum(int(r[j][2]) for r in rows) for j in range(k)], rounds


# ---------------------------------------------------------------------------
# Neighbour stitch (round 2): the chain state only ever matters to the RIGHT neighbour, so the all-gather above is
# replaced by one send to the right / one receive from the left.  On GPUs the send is a peer-to-peer store over
# NVLink into the neighbour's device memory, issued by a one-warp kernel on the engine's stream
# (rejit_b200_stitch_exchange; the inboxes are mapped once through CUDA IPC); the CPU tests run the same protocol
# over gloo point-to-point messages.
# ---------------------------------------------------------------------------
def _reaches(arrived, slab_lo):
    """The predicate of k_stitch: does the chain that arrives from the left reach into a slab that begins at slab_lo."""
    cur, tail = arrived
    return cur > slab_lo or (tail != NO_TAIL and tail == cur and cur == slab_lo)


class GlooNeighbourStitch:
    """exchange(leaving, slab_lo) -> (arrived[k], redo_mask) over torch.distributed point-to-point messages."""
    name = "p2p"

    def __init__(self, dist, rank: int, world: int):
        self.dist, self.rank, self.world = dist, rank, world

    def exchange(self, leaving, slab_lo):
        import torch
        k = len(leaving)
        out = torch.tensor([[c if (c > slab_lo or t != NO_TAIL) else 0, -1 if t == NO_TAIL else t] for c, t in leaving], dtype=torch.int64)
        req = self.dist.isend(out, self.rank + 1) if self.rank + 1 < self.world else None
        arrived = [(0, NO_TAIL)] * k
        if self.rank > 0:
            got = torch.zeros((k, 2), dtype=torch.int64)
            self.dist.recv(got, self.rank - 1)
            arrived = [(int(c), NO_TAIL if int(t) < 0 else int(t)) for c, t in got.tolist()]
        if req is not None:
            req.wait()
        redo = 0
        for j, a in enumerate(arrived):
            if a[0] and _reaches(a, slab_lo):
                redo |= 1 << j
        return arrived, redo

    def close(self):
        pass


class DeviceStitch:
    """The same over NVLink peer stores (include/rejit_b200.h: rejit_b200_stitch_*).  `dist` is only used once, to
    hand the 64-byte CUDA IPC handles of the inboxes to the neighbours."""
    name = "nvlink"

    def __init__(self, dist, rank: int, world: int, device: int, tdev=None):
        import ctypes
        import torch
        import rejit_b200 as rj
        self.rj, self.device, self.rank, self.world = rj, device, rank, world
        L = rj.lib()
        handle = (ctypes.c_ubyte * 64)()
        err = ctypes.create_string_buffer(256)
        if L.rejit_b200_stitch_open(device, rank, world, handle, err, len(err)) != 0:
            raise rj.RejitError(err.value.decode("latin-1"))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=tdev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        hs = [bytes(t.cpu().tolist()) for t in every]
        left = (ctypes.c_ubyte * 64)(*hs[rank - 1]) if rank > 0 else None
        right = (ctypes.c_ubyte * 64)(*hs[rank + 1]) if rank + 1 < world else None
        if L.rejit_b200_stitch_connect(device, left, right, err, len(err)) != 0:
            raise rj.RejitError(err.value.decode("latin-1"))
        dist.barrier()

    def exchange(self, leaving, slab_lo):
        import ctypes
        rj = self.rj
        k = len(leaving)
        out = (rj.Carry * k)(*[rj.Carry(c, t) for c, t in leaving])
        arr = (rj.Carry * k)()
        redo = ctypes.c_uint32()
        err = ctypes.create_string_buffer(256)
        if rj.lib().rejit_b200_stitch_exchange(self.device, k, out, slab_lo, arr, ctypes.byref(redo), err, len(err)) != 0:
            raise rj.RejitError(err.value.decode("latin-1"))
        return [(int(a.cur), int(a.tail)) for a in arr], int(redo.value)

    def close(self):
        self.rj.lib().rejit_b200_stitch_close(self.device)


def stitched_set_neighbour(stitch, slab_lo: int, k: int, run):
    """One step of the neighbour protocol.  run(carries) -> (counts[k], carries_out[k]) (global offsets, as for
    stitched_counts_set).  Returns (this rank's counts[k], cascaded): `cascaded` says that resolving again with the
    arriving chain changed the state this rank had already sent to its right neighbour (a slab without a
    re-synchronisation point: the caller then repeats the step with stitched_counts_set, which iterates)."""
    counts, couts = run([(slab_lo, NO_TAIL)] * k)
    arrived, redo = stitch.exchange(couts, slab_lo)
    cascaded = False
    if redo:
        want = []
        for j in range(k):
            if (redo >> j) & 1:
                cur, tail = arrived[j]
                want.append((max(cur, slab_lo), tail if tail == slab_lo else NO_TAIL))
            else:
                want.append((slab_lo, NO_TAIL))
        counts, couts2 = run(want)
        cascaded = couts2 != couts
    return counts, cascaded
