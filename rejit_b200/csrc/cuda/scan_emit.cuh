// rejit_b200 — single-pass scan + ordered emit (round 2).
//
// One kernel reads the text ONCE and writes the matches ONCE, at their final
// place, in text order — no slot ranges, no grid barrier, no second kernel:
//
//   tiles     the text is cut into tiles of 32 KB; a WARP takes tiles in order from
//             a ticket counter and works on its own: no block-wide barrier anywhere.
//             It streams the tile as 64 rows of 512 bytes (a lane holds 16 bytes of
//             a row: one coalesced 16-byte load, four rows in flight).
//   filter    literal: the first <= 4 needle bytes at 16 alignments (funnel shifts;
//             the window that begins in the three bytes before the lane's 16 comes
//             from the left neighbour, so nothing waits for a row still in flight);
//             generic: "can a match begin at this byte" as SWAR byte compares
//             (<= 4 start bytes and "right after a line break") or a 256-bit map.
//             Lanes with a survivor leave one word in shared memory (a ballot and
//             a store; nothing else happens inside the streaming loop).
//   evaluate  when its list of survivors fills up, and after the last row, the warp
//             turns survivors into candidates (begin, E(begin)): the rest of the
//             needle / one NFA run per start (device_program.h NfaRun) / for
//             "required literal + window" patterns one NFA run per start in front
//             of every needle hit.  One lane per candidate, results compacted in
//             order.  The rows in flight stay in flight meanwhile.
//   select    candidates of a tile are sorted by construction.  When every candidate
//             begins after its predecessor ended (the rule of ChainTake) the
//             candidates ARE the matches; otherwise one lane walks the list with
//             ChainTake (leftmost-longest, /root/reference/src/x64/codegen-x64.cc:
//             401-522, src/codegen.cc:36-86).
//   place     the warp publishes {count, chain state} of its tile and looks back
//             over the tiles before it (decoupled look-back: records tagged with
//             the call's sequence number, 32 predecessors per step) for the number
//             of matches before it and the chain state arriving from the left; a
//             state that reaches into the tile (a match straddling the tile edge)
//             cannot be repaired locally — counts are already published — so it
//             raises kFinOverlap and the host runs the general path instead.
//   report    the warp that finishes the last tile writes the FinRecord to mapped
//             host memory.
//
// Algorithmic traffic: N bytes read + 16 bytes written per match; the look-back
// records (32 bytes per 32 KB tile) stay in L2.
#ifndef REJIT_B200_CUDA_SCAN_EMIT_CUH_
#define REJIT_B200_CUDA_SCAN_EMIT_CUH_

#include "kernels.cuh"

namespace rejit_b200 {

constexpr uint32_t kEmWarps = 8;
constexpr uint32_t kEmThreads = kEmWarps * 32;
#ifndef RJ_EM_ROWS
#define RJ_EM_ROWS 64
#endif
constexpr uint32_t kEmRows = RJ_EM_ROWS;                      // rows of 512 bytes per tile
constexpr uint32_t kEmTileBytes = kEmRows * 512;              // 32 KB, one warp
constexpr uint32_t kEmEntCap = 512;                           // survivors (16-byte groups) a warp collects before it evaluates them
constexpr uint32_t kEmEntFlush = kEmEntCap - 128;             // ... checked every four rows (at most 128 more)
constexpr uint32_t kEmCandCap = 1024;                         // candidates per tile
constexpr uint32_t kEmWinCandCap = 512;                       // ... in window mode (the other half holds the needle hits)
constexpr uint32_t kEmBias = 8192;                            // offsets in a tile are relative to tile_lo - kEmBias (16 bits)
constexpr uint32_t kEmDropped = 0xFFFFu;                      // length field of a candidate the chain did not take
constexpr uint32_t kEmPending = 0xFFFEu;                      // length field of a start that has not been evaluated yet
constexpr unsigned int kFinLastEmpty = 8u;                    // FinRecord.flags: the last match is empty
constexpr unsigned int kFinStuck = 16u;                       // a look-back gave up waiting (never expected)
// (the pattern's NFA tables stay in global memory: a copy in shared memory made the window / generic kernels slower,
// with and without hits — the rewritten table pointers cost registers in the streaming loop; A/B in profiles/r2j_*)
constexpr size_t kEmSmemBytes = kEmWarps * (kEmEntCap * 4 + kEmCandCap * 4);

// a candidate in shared memory: begin - tile_base in the low half, length in the high half
__device__ __forceinline__ uint32_t EmCand(uint32_t rel, uint32_t len) { return rel | (len << 16); }
__device__ __forceinline__ uint32_t EmRel(uint32_t c) { return c & 0xFFFFu; }
__device__ __forceinline__ uint32_t EmLen(uint32_t c) { return c >> 16; }

enum : int { kEmLiteral = 0, kEmWindow = 1, kEmGeneric = 2 };

struct EmLit {
  const uint8_t* needle;
  uint32_t m, p4, pmask;
  uint32_t win_lo, win_hi;          // window mode: starts in [hit - win_hi, hit - win_lo]
};

// "can a match (or the empty match) begin at a byte c whose predecessor was / was not a line break":
// cand = eq(c, b0[..]) | (after_break & (all1 ? any : eq(c, b1[..])))       (swar)
// or two 256-bit maps indexed by the byte                                     (!swar)
struct EmFilter {
  uint32_t t[2][8];                 // [after a line break][byte >> 5] bit (byte & 31)
  uint32_t swar;                    // the SWAR form is exact for this pattern
  uint32_t n0, b0;                  // start bytes that need no context (packed, n0 <= 4)
  uint32_t n1, b1;                  // further start bytes right after a line break (n1 <= 4)
  uint32_t all1;                    // ... or every byte, right after a line break
  uint32_t use_sol;                 // the context matters at all
};

struct EmitArgs {
  uint4* records;                   // [ntiles][2]: what every tile found
  uint4* group_records;             // [ceil(ntiles / 32)][2]: groups of 32 tiles (own numbers, then inclusive)
  unsigned int* sync;               // [0..1] ticket (64 bit), [2] flags, [3] tiles done; zeroed by the host before the launch
  unsigned long long* final_state;  // [3]: total matches, chain state (cur, non-empty); written by the last tile
  uint64_t tile0, ntiles;           // tiles [tile0, tile0 + ntiles) hold every owned start (and needle hit)
  // ReplaceAll fused into the scan (the kRebuild instantiation; generic scans, replacement not longer than the
  // shortest match, so the output is never longer than the text): instead of the match pairs the kernel writes the
  // rebuilt text — every tile copies the bytes between its matches and the replacement strings to their final place
  // (tile and group records also carry the bytes inside their matches, summed by the same look-back) — and the
  // lengths pass, the prefix sum, the index and the staged rebuild of the separate path (replace.cuh) are not run.
  uint8_t* rep_out;                 // the rebuilt text, >= n + 64 bytes (only read by the kRebuild kernel)
  const uint8_t* rep_with;          // the replacement string, device memory
  uint32_t rep_w;                   // its length
  uint32_t static_stride;           // warps of the grid when the tiles are dealt round robin, 0: ticket counter
  uint32_t rows;                    // rows of 512 bytes per tile (<= kEmRows; the host picks: dense candidates want smaller tiles)
  uint64_t* out_pairs;
  uint64_t out_cap, base_offset;
  FinRecord* host_records;
  unsigned int seq;
  Carry carry_in;
};

// ---- chain state across candidates: where the next match may begin, and whether the last match was non-empty
// (then the end of that match is also the chain's `tail`, ChainTake in device_program.h) ------------------------
struct EmState {
  uint64_t cur;
  uint32_t ne;
};
__device__ __forceinline__ bool EmTakes(const EmState& s, uint64_t b, uint32_t len) {
  return b > s.cur || (b == s.cur && (len > 0 || !s.ne));
}
__device__ __forceinline__ EmState EmAfter(uint64_t b, uint32_t len) {
  EmState s;
  s.cur = len ? b + len : b + 1;
  s.ne = len ? 1u : 0u;
  return s;
}

// ---- byte compares on a 16-byte group, "transposed" result: byte k of word j -> bit 8 k + j -------------------
__device__ __forceinline__ uint32_t EmZeroFlags(uint32_t x) {       // bit 7 of every zero byte (exact)
  return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
__device__ __forceinline__ uint32_t EmEqT(const uint4& v, uint32_t byte) {
  const uint32_t s = byte * 0x01010101u;
  return (EmZeroFlags(v.x ^ s) >> 7) | (EmZeroFlags(v.y ^ s) >> 6) | (EmZeroFlags(v.z ^ s) >> 5) | (EmZeroFlags(v.w ^ s) >> 4);
}
// the flags of the positions one byte further (position 15 falls out, `carry` enters at position 0)
__device__ __forceinline__ uint32_t EmShiftT(uint32_t t, uint32_t carry) {
  return ((t << 8) & 0x0F0F0F00u) | ((t >> 23) & 0x0000000Eu) | carry;
}
__device__ __forceinline__ uint32_t EmLastT(uint32_t t) { return (t >> 27) & 1u; }      // position 15
// transposed bit q = 8 k + j  <->  position 4 j + k; in an entry the flags are packed to 16 bits, bit 4 k + j
__device__ __forceinline__ uint32_t EmTOfPos(uint32_t p) { return ((p & 3u) << 3) | (p >> 2); }
__device__ __forceinline__ uint32_t EmPosOfT16(int q) { return (uint32_t)(((q & 3) << 2) | (q >> 2)); }
// positions >= limit (0..16) removed
__device__ __forceinline__ uint32_t EmKeepBelowT(uint32_t t, uint32_t limit) {
  if (limit >= 16) return t;
  uint32_t keep = 0;
  for (uint32_t p = 0; p < limit; ++p) keep |= 1u << EmTOfPos(p);
  return t & keep;
}

// the streaming load of a row: read once, must not push the pattern's tables (and the bytes the NFA runs re-read) out of L1
__device__ __forceinline__ uint4 EmLoadStream(const uint4* p) {
#if defined(RJ_EM_LDG)
  return __ldg(p);
#else
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#endif
}
__device__ __forceinline__ uint4 EmLoadRow(const uint8_t* __restrict__ text, uint64_t n16, uint64_t at) {
  return at < n16 ? __ldg(reinterpret_cast<const uint4*>(text + at)) : make_uint4(0, 0, 0, 0);
}

// ===========================================================================
// filters: one row (the lane's 16 bytes) -> 16 flag bits
// ===========================================================================
// literal: bit j = the first min(m, 4) needle bytes match at the window that BEGINS three bytes before the lane's
// 16 plus j, i.e. at offset (lane's offset) - 3 + j.  `pw` = the four bytes before the lane's 16.
template <bool kFull4>
__device__ __forceinline__ uint32_t EmLitFlags(const uint4& v, uint32_t pw, uint32_t p4, uint32_t pmask) {
  const uint32_t w[5] = {pw, v.x, v.y, v.z, v.w};
  uint32_t flags = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int o = j + 1;                                      // byte offset in the 20-byte buffer
    const uint32_t x = (o & 3) ? __funnelshift_r(w[o >> 2], w[(o >> 2) + 1], 8 * (o & 3)) : w[o >> 2];
    if (kFull4 ? (x == p4) : (((x ^ p4) & pmask) == 0)) flags |= 1u << j;
  }
  return flags;
}
template <bool kFull4>
__device__ __forceinline__ bool EmLitAny(const uint4& v, uint32_t pw, uint32_t p4, uint32_t pmask) {
  const uint32_t w[5] = {pw, v.x, v.y, v.z, v.w};
  // four independent chains of compares (one chain of sixteen dependent predicate updates stalled every row)
  bool any[4] = {false, false, false, false};
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int o = j + 1;
    const uint32_t x = (o & 3) ? __funnelshift_r(w[o >> 2], w[(o >> 2) + 1], 8 * (o & 3)) : w[o >> 2];
    any[j & 3] |= kFull4 ? (x == p4) : (((x ^ p4) & pmask) == 0);
  }
  return (any[0] | any[1]) | (any[2] | any[3]);
}

// generic: flag bits transposed (EmEqT), 28 bits
__device__ __forceinline__ uint32_t EmFilterGroup(const uint4& v, uint32_t prev_is_break, const EmFilter& f) {
  if (f.swar) {
    uint32_t m0 = 0;
    for (uint32_t i = 0; i < f.n0; ++i) m0 |= EmEqT(v, (f.b0 >> (8 * i)) & 0xFFu);
    if (!f.use_sol) return m0;
    const uint32_t brk = EmEqT(v, 0x0Au) | EmEqT(v, 0x0Du);
    uint32_t m1 = 0x0F0F0F0Fu;
    if (!f.all1) {
      m1 = 0;
      for (uint32_t i = 0; i < f.n1; ++i) m1 |= EmEqT(v, (f.b1 >> (8 * i)) & 0xFFu);
    }
    return m0 | (EmShiftT(brk, prev_is_break) & m1);
  }
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t cand = 0;
  uint32_t sol = prev_is_break;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const uint32_t c = (w[p >> 2] >> (8 * (p & 3))) & 0xFFu;
    const uint32_t bit = (f.t[sol][c >> 5] >> (c & 31)) & 1u;
    cand |= bit << (((p & 3) << 3) | (p >> 2));
    sol = (c == 0x0Au || c == 0x0Du) ? 1u : 0u;
  }
  return cand;
}

// ===========================================================================
// evaluation: survivors -> candidates (begin - tile_base | length << 16), appended to the tile's list in position
// order.  Called by the whole warp; return the new number of candidates (above `cap`: the list overflowed).
// ===========================================================================
// literal occurrences (flag j of a group = the window that begins at group offset - 3 + j): bounds, ownership, the
// rest of the needle.  kHits: the results are needle hits (u32 offsets) for the window stage instead of candidates.
template <bool kHits>
__device__ __forceinline__ uint32_t EmEvalLiteral(const uint8_t* __restrict__ text, uint64_t n, const EmLit& lit,
                                                   uint64_t own_lo, uint64_t own_hi, uint64_t tile_base, uint64_t tile_lo,
                                                   const uint32_t* my_ent, uint32_t n_ent, uint32_t* out, uint32_t k,
                                                   uint32_t cap) {
  const int lane = threadIdx.x & 31;
  for (uint32_t base = 0; base < n_ent; base += 32) {
    const uint32_t ent = base + lane < n_ent ? my_ent[base + lane] : 0u;
    uint32_t valid = ent & 0xFFFFu;
    const uint64_t my = tile_lo + (uint64_t)(ent >> 16) * 16;          // the group's offset; flag j: my - 3 + j
    for (uint32_t hh = valid; hh; hh &= hh - 1) {
      const int j = __ffs(hh) - 1;
      const uint64_t pos = my + j - 3;                                 // (wraps below zero for the text's first bytes)
      bool ok = my + j >= 3 && pos >= own_lo && pos < own_hi && pos + lit.m <= n;
      for (uint32_t i = 4; i < lit.m && ok; ++i) ok = (__ldg(text + pos + i) == __ldg(lit.needle + i));
      if (!ok) valid &= ~(1u << j);
    }
    const uint32_t c = __popc(valid);
    const uint32_t incl = WarpInclusiveScan(c);
    uint32_t idx = k + incl - c;
    for (; valid; valid &= valid - 1) {
      const uint32_t rel = (uint32_t)(my + (__ffs(valid) - 1) - 3 - tile_base);
      if (idx < cap) out[idx] = kHits ? rel : EmCand(rel, lit.m);
      ++idx;
    }
    k += __shfl_sync(kFullMask, incl, 31);
  }
  __syncwarp();
  return k;
}

// one NFA run per start in front of every needle hit (the windows of neighbouring hits of this tile are clipped
// against each other so that every start is tried once, in order).  Two phases: the starts of ALL the tile's hits
// are first put to the one-byte test "can a match begin with this byte here" (nine in ten fail it on the complex
// benchmark pattern) and the survivors collected, in order, in `my_list`; then the NFA runs, 32 survivors at a time.
// (Running the windows hit by hit — two batches of mostly idle lanes per hit, each a chain of dependent table
// lookups — was 60 % of the kernel on a text with a hit every 10 KB.)
// the NFA runs for the starts my_list[0 .. n_list) (offsets from tile_base, ascending), 32 at a time; candidates appended in order
__device__ __forceinline__ uint32_t EmRunList(const uint8_t* __restrict__ text, uint64_t n, const NfaTables& nfa, uint64_t tile_base,
                                               const uint32_t* my_list, uint32_t n_list, uint32_t* my_cand, uint32_t k, uint32_t cap,
                                               unsigned int* flags) {
  const int lane = threadIdx.x & 31;
  for (uint32_t base = 0; base < n_list; base += 32) {
    const uint32_t i = base + lane;
    uint64_t s = 0, e = kNoMatch;
    if (i < n_list) {
      s = tile_base + my_list[i];
      e = NfaRunAny(nfa, text, n, s);
    }
    __syncwarp();
    const bool has = e != kNoMatch;
    if (has && e - s >= kEmPending) { *flags |= kFinDense; e = s; }      // length does not fit: the general path
    const uint32_t bal = __ballot_sync(kFullMask, has);
    if (bal) {
      const uint32_t idx = k + __popc(bal & ((1u << lane) - 1u));
      if (has && idx < cap) my_cand[idx] = EmCand((uint32_t)(s - tile_base), (uint32_t)(e - s));
      k += __popc(bal);
    }
  }
  __syncwarp();
  return k;
}

// (not inlined: its NFA runs would share the streaming loop's 64 registers and push the rows in flight into local memory;
// at the end of a tile, where it is called, no row is in flight and the call saves nothing)
__device__ __noinline__ uint32_t EmEvalWindow(const uint8_t* __restrict__ text, uint64_t n, const NfaTables& nfa,
                                                  const EmLit& lit, const ScanRange& range, uint64_t tile_base,
                                                  const uint32_t* my_hits, uint32_t n_hits, uint64_t* prev_hit,
                                                  uint32_t* my_cand, uint32_t k, uint32_t cap, unsigned int* flags,
                                                  uint32_t* my_list, uint32_t list_cap) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t n_list = 0;
  for (uint32_t q = 0; q < n_hits; ++q) {
    const uint64_t h = tile_base + my_hits[q];
    const uint64_t prev = *prev_hit;                           // the hit before it in this tile (kNoMatch: none)
    *prev_hit = h;
    if (h < lit.win_lo) continue;
    const uint64_t s_max = h - lit.win_lo;                     // inclusive
    uint64_t s_min = h >= lit.win_hi ? h - lit.win_hi : 0;
    if (prev != kNoMatch && prev >= lit.win_lo && prev - lit.win_lo + 1 > s_min) s_min = prev - lit.win_lo + 1;
    for (uint64_t base = s_min; base <= s_max; base += 32) {
      const uint64_t s = base + lane;
      bool ok = false;
      if (s <= s_max && s >= range.own_begin && s < range.own_end && s < n) {
        const int ctx = nfa.has_anchor ? ContextAt(text, n, s) : 0;
        ok = nfa.start_ok[ctx * 256 + text[s]] != 0;
      }
      const uint32_t bal = __ballot_sync(kFullMask, ok);
      if (bal) {
        if (ok) my_list[n_list + __popc(bal & lt_mask)] = (uint32_t)(s - tile_base);
        n_list += __popc(bal);
        __syncwarp();
        if (n_list + 32 > list_cap) { k = EmRunList(text, n, nfa, tile_base, my_list, n_list, my_cand, k, cap, flags); n_list = 0; }
      }
    }
  }
  if (n_list) k = EmRunList(text, n, nfa, tile_base, my_list, n_list, my_cand, k, cap, flags);
  __syncwarp();
  return k;
}

// generic: survivors (transposed flags, flag of position p of a group = a start at group offset + p) -> starts,
// appended to the candidate list as pending entries, then one NFA run per start, compacted in place
__device__ __forceinline__ uint32_t EmEvalGeneric(const uint8_t* __restrict__ text, uint64_t n, const NfaTables& nfa,
                                                   const ScanRange& range, uint64_t tile_base, uint64_t tile_lo,
                                                   const uint32_t* my_ent, uint32_t n_ent, bool add_end, uint32_t* my_cand,
                                                   uint32_t k, uint32_t cap, unsigned int* flags) {
  const int lane = threadIdx.x & 31;
  uint32_t n_start = k;
  for (uint32_t base = 0; base < n_ent; base += 32) {
    const uint32_t ent = base + lane < n_ent ? my_ent[base + lane] : 0u;
    const uint32_t f16 = ent & 0xFFFFu;
    const uint64_t my = tile_lo + (uint64_t)(ent >> 16) * 16;
    const uint32_t c = __popc(f16);
    const uint32_t incl = WarpInclusiveScan(c);
    uint32_t idx = n_start + incl - c;
    if (f16) {
#pragma unroll 1
      for (uint32_t p = 0; p < 16; ++p) {
        if (!((f16 >> EmPosOfT16((int)p)) & 1u)) continue;          // (the 4 x 4 transpose is its own inverse)
        if (idx < cap) my_cand[idx] = EmCand((uint32_t)(my + p - tile_base), kEmPending);
        ++idx;
      }
    }
    n_start += __shfl_sync(kFullMask, incl, 31);
  }
  // the offset n itself: only the empty match can begin there (the run decides)
  if (add_end) {
    if (lane == 0 && n_start < cap) my_cand[n_start] = EmCand((uint32_t)(n - tile_base), kEmPending);
    ++n_start;
  }
  __syncwarp();
  if (n_start > cap) return n_start;
  for (uint32_t base = k; base < n_start; base += 32) {
    const uint32_t i = base + lane;
    uint32_t rel = 0;
    uint64_t s = 0, e = kNoMatch;
    if (i < n_start) {
      rel = EmRel(my_cand[i]);
      s = tile_base + rel;
      if (s >= range.own_begin && s < range.own_end && s <= n) e = NfaRunAny(nfa, text, n, s);
    }
    __syncwarp();
    const bool has = e != kNoMatch;
    if (has && e - s >= kEmPending) { *flags |= kFinDense; e = s; }
    const uint32_t bal = __ballot_sync(kFullMask, has);
    if (has) my_cand[k + __popc(bal & ((1u << lane) - 1u))] = EmCand(rel, (uint32_t)(e - s));
    k += __popc(bal);
    __syncwarp();
  }
  return k;
}

// ===========================================================================
// look-back records: two 16-byte halves per tile, each written with one store and carrying
//   tag = seq << 2 | state   (1: the tile's own numbers, 2: everything up to and including the tile)
// half 0: {count lo, count hi, tag, 0}    half 1: {cur lo, cur hi | ne << 31 | has << 30, tag, 0}
// ===========================================================================
__device__ __forceinline__ void EmStore16(uint4* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 EmLoad16(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// (`removed`: bytes inside the counted matches — only the fused ReplaceAll uses it; low word in half 0, high word in half 1)
__device__ __forceinline__ void EmPublish(uint4* rec, uint32_t tag, uint64_t count, bool has, const EmState& st, uint64_t removed = 0) {
  EmStore16(rec, (uint32_t)count, (uint32_t)(count >> 32), tag, (uint32_t)removed);
  EmStore16(rec + 1, (uint32_t)st.cur, (uint32_t)(st.cur >> 32) | (st.ne << 31) | (has ? 1u << 30 : 0u), tag, (uint32_t)(removed >> 32));
}

// Called by a whole warp.  Returns (in every lane) the number of matches in the tiles before `t` and the chain
// state that arrives at the tile (the state after the last match before it, or the call's carry).
// (A step that looks at 128 records, four per lane with their loads in flight together, was tried against the idea
// that the inclusive records advance by one step's worth of tiles per L2 round trip: it halved the throughput of
// every text with matches in all tiles — four times the polling reads on the lines the publishers are writing.)
//
// Two levels (round 2c).  With one level, the inclusive records advance along the text by at most 32 tiles per L2
// round trip whatever the number of warps: ~1.5-2 TB/s with 32 KB tiles, and the bound of every long text whose
// tiles all hold matches (C4: 2.0 TB/s over 500 MB, 1.3 TB/s over 5 GB).  Now a tile record only ever says what the
// tile itself found; the LAST tile of every group of 32 adds its group's records up and publishes a GROUP record
// (own numbers first, then — after a look-back over the group records before it, 32 groups = 1024 tiles per step —
// everything up to and including the group).  A tile's prefix = the tiles before it in its group + the groups before.
template <bool kRebuild>
__device__ __forceinline__ void EmLookBack(const EmitArgs& em, const uint4* records, uint64_t t, uint64_t* before, EmState* arriving,
                                           bool* found, uint64_t* removed_before) {
  const int lane = threadIdx.x & 31;
  uint64_t excl = 0, excl_rem = 0;
  EmState st;
  st.cur = 0; st.ne = 0;
  bool have = false;
  for (int64_t base = (int64_t)t;; base -= 32) {
    const int64_t idx = base - 1 - lane;
    uint64_t cnt = 0, cur = 0, rem = 0;
    uint32_t state = 0, ne = 0, has = 0;
    if (idx >= 0) {
      const uint4* rec = records + 2 * (uint64_t)idx;
      for (uint32_t polls = 0;; ++polls) {
        if (polls == (1u << 20)) {                 // a predecessor that never reports would hang the device: give up,
          atomicOr(&em.sync[2], kFinOverlap | kFinStuck);      // the host runs the general path and says so
          state = 2;
          break;
        }
        const uint4 a = EmLoad16(rec), b = EmLoad16(rec + 1);
        if (a.z == b.z && (a.z >> 2) == (em.seq & 0x3FFFFFFFu) && (a.z & 3u) != 0) {
          state = a.z & 3u;
          cnt = (uint64_t)a.y << 32 | a.x;
          cur = (uint64_t)(b.y & 0x3FFFFFFFu) << 32 | b.x;
          ne = b.y >> 31;
          has = (b.y >> 30) & 1u;
          if (kRebuild) rem = (uint64_t)b.w << 32 | a.w;
          break;
        }
        // (a short clock-counting spin: __nanosleep(64) here made the chain of look-backs 20-60 % slower on texts
        // with matches, measured — the wake-up granularity is far above the L2 round trip the poll waits for)
        const long long t0 = clock64();
        while (clock64() - t0 < 64) {}
      }
    } else if (idx == -1) {                        // before the first tile: nothing counted, the call's carry
      state = 2;
      cur = em.carry_in.cur;
      ne = (em.carry_in.tail == em.carry_in.cur) ? 1u : 0u;
      has = 1;
    } else {
      state = 2;                                   // further back: nothing
    }
    const uint32_t incl_mask = __ballot_sync(kFullMask, state == 2);
    const int stop = incl_mask ? __ffs(incl_mask) - 1 : 31;
    const bool use = lane <= stop;
    uint64_t sum = use ? cnt : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(kFullMask, sum, d);
    excl += sum;
    if (kRebuild) {
      uint64_t sr = use ? rem : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) sr += __shfl_xor_sync(kFullMask, sr, d);
      excl_rem += sr;
    }
    const uint32_t st_mask = __ballot_sync(kFullMask, use && has);
    if (!have && st_mask) {
      const int src = __ffs(st_mask) - 1;
      st.cur = __shfl_sync(kFullMask, cur, src);
      st.ne = __shfl_sync(kFullMask, ne, src);
      have = true;
    }
    if (incl_mask) break;
  }
  *before = excl;
  *arriving = st;
  *found = have;
  *removed_before = excl_rem;
}

// The tiles of tile t's group that come before it (t & 31 of them, one per lane): their matches, and the chain state
// after the last match among them (*found: there is one).
template <bool kRebuild>
__device__ __forceinline__ void EmGroupBefore(const EmitArgs& em, uint64_t t, uint32_t* before, EmState* after, bool* found,
                                              uint64_t* removed) {
  const int lane = threadIdx.x & 31;
  const uint32_t j = (uint32_t)t & 31u;
  uint32_t cnt = 0, ne = 0, has = 0;
  uint64_t cur = 0, rem = 0;
  if ((uint32_t)lane < j) {
    const uint4* rec = em.records + 2 * (t - j + lane);
    for (uint32_t polls = 0;; ++polls) {
      if (polls == (1u << 20)) { atomicOr(&em.sync[2], kFinOverlap | kFinStuck); break; }
      const uint4 a = EmLoad16(rec), b = EmLoad16(rec + 1);
      if (a.z == b.z && (a.z >> 2) == (em.seq & 0x3FFFFFFFu) && (a.z & 3u) != 0) {
        cnt = a.x;                                 // (a tile holds at most kEmCandCap matches)
        cur = (uint64_t)(b.y & 0x3FFFFFFFu) << 32 | b.x;
        ne = b.y >> 31;
        has = (b.y >> 30) & 1u;
        if (kRebuild) rem = (uint64_t)b.w << 32 | a.w;
        break;
      }
      const long long t0 = clock64();
      while (clock64() - t0 < 64) {}
    }
  }
  *before = __reduce_add_sync(kFullMask, cnt);
  *removed = 0;
  if (kRebuild) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rem += __shfl_xor_sync(kFullMask, rem, d);
    *removed = rem;
  }
  const uint32_t st_mask = __ballot_sync(kFullMask, has != 0);
  const int src = st_mask ? 31 - __clz(st_mask) : 0;        // the nearest tile before me with a match
  after->cur = __shfl_sync(kFullMask, cur, src);
  after->ne = __shfl_sync(kFullMask, ne, src);
  *found = st_mask != 0;
}

// ===========================================================================
// the kernel: every warp is on its own
// ===========================================================================
// kDepth rows in flight per warp.  A warp asks for a new row only when it has looked at one, so its time per row is
// the memory latency / kDepth PLUS its own instructions: with four rows the literal kernels ran at 4.0 TB/s where the
// same tiling with an empty loop body reads 6.9 TB/s (scripts/probe/read_bw.cu); eight rows cost sixteen registers
// (three CTAs per SM instead of four).  The window / generic kernels are at their register limit with four.
template <int kMode, bool kFull4, int kDepth, bool kRebuild = false>
__global__ void __launch_bounds__(kEmThreads, kDepth > 4 ? 3 : 4)
k_scan_emit(const uint8_t* __restrict__ text, uint64_t n, EmLit lit, NfaTables nfa, EmFilter flt, ScanRange range,
            EmitArgs em) {
  extern __shared__ __align__(16) uint8_t em_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t* my_ent = reinterpret_cast<uint32_t*>(em_smem) + warp * kEmEntCap;
  uint32_t* my_cand = reinterpret_cast<uint32_t*>(em_smem + kEmWarps * kEmEntCap * 4) + warp * kEmCandCap;
  uint32_t* my_hits = my_cand + kEmWinCandCap;                                          // window mode: the upper half
  const uint64_t n16 = (n + 15) & ~15ull;
  const uint32_t cap = kMode == kEmWindow ? kEmWinCandCap : kEmCandCap;
  const uint32_t lt_mask = (1u << lane) - 1u;
  // needle hits may sit up to win_hi bytes after an owned start
  const uint64_t hit_hi = kMode == kEmWindow ? range.own_end + lit.win_hi + 1 : range.own_end;
  // a literal window begins up to three bytes before a lane's 16: rows are scanned up to the one that holds n + 2
  const uint64_t scan_end = kMode == kEmGeneric ? n : n + 3;

  // the survivors are evaluated when their list reaches this length (and after the tile's last row).  Generic scans
  // re-read the text around every start (line context, NFA run): evaluated early, those bytes are still in L2 — at the
  // end of a 24 KB tile, with 4736 tiles in flight, half of them had been evicted (DRAM traffic 1.8 x the text).
#ifndef RJ_EM_GEN_STOP
#define RJ_EM_GEN_STOP kEmEntCap
#endif
  constexpr uint32_t kEntStop = kMode == kEmGeneric ? (uint32_t)(RJ_EM_GEN_STOP) : kEmEntCap;
  const uint32_t tile_rows = em.rows, tile_bytes = em.rows * 512u;
  // Tiles are dealt round robin (em.static_stride = warps of the grid; the host launches no more CTAs than are resident
  // together, so the warp that owns a tile someone looks back at is running) or, with static_stride == 0, taken from a
  // ticket counter.  One counter for 4736 warps was the largest single cost of the no-hit scan: 3.9 -> 4.6 TB/s at
  // 500 MB, 4.4 -> 5.1 TB/s at 5 GB, C4 over 5 GB 1.4 -> 2.0 TB/s (profiles/r2_ab_scan_emit.txt).
  const unsigned long long stride = em.static_stride;
  unsigned int done_mine = 0;                               // (lane 0) tiles this warp has finished and not yet reported
  unsigned long long next_ticket = 0;
  if (stride) next_ticket = (unsigned long long)blockIdx.x * kEmWarps + warp;
  else if (lane == 0) next_ticket = atomicAdd(reinterpret_cast<unsigned long long*>(em.sync), 1ull);
  for (;;) {
    const uint64_t t = __shfl_sync(kFullMask, next_ticket, 0);
    if (t >= em.ntiles) {
      if (lane == 0 && done_mine) atomicAdd(&em.sync[3], done_mine);
      break;
    }
    const uint64_t tile_lo = (em.tile0 + t) * tile_bytes;
    const uint64_t tile_base = tile_lo >= kEmBias ? tile_lo - kEmBias : 0;
    const uint64_t mine = tile_lo + (uint64_t)lane * 16;
    unsigned int flags = 0;
    uint32_t cnt = 0;                                       // candidates of the tile so far
    uint64_t prev_hit = kNoMatch;
    // The next ticket is asked for a few rows before this tile's last row, so that its round trip hides behind them (a
    // tile that has an owner but has not been started holds up the look-back of every tile after it: not earlier).
    // (An L2 bulk prefetch a few rows ahead of the loads was tried too: 8 / 16 / 32 rows ahead all cost 10-20 %.)
    bool asked = false;
    const bool live = tile_lo <= n && tile_lo < hit_hi + 16 && tile_lo + tile_bytes + 16 > range.own_begin;
    if (live) {
      // ---- stream the rows; evaluate the survivors whenever their list fills up, and at the end --------------
      uint32_t n_ent = 0;
      // rows [0, rows_ld) have my 16 bytes inside the (padded) text: one compare per load instead of a 64-bit bound
      const uint32_t rows_ld = mine < n16 ? (uint32_t)(((n16 - mine + 511) >> 9) < tile_rows ? ((n16 - mine + 511) >> 9) : tile_rows) : 0u;
      const uint4* src = reinterpret_cast<const uint4*>(text + mine);
      const uint4 zero4 = make_uint4(0, 0, 0, 0);
      uint4 v[kDepth];
#pragma unroll
      for (int u = 0; u < kDepth; ++u) v[u] = (uint32_t)u < rows_ld ? EmLoadStream(src + u * 32) : zero4;
      // what precedes the tile: its last word (literal windows) / whether its last byte is a line break
      uint32_t tail = 0;
      if (kMode == kEmGeneric) {
        tail = 1u;                                          // the text start counts as "after a line break"
        if (tile_lo > 0) { const uint8_t pb = __ldg(text + tile_lo - 1); tail = (pb == 0x0Au || pb == 0x0Du) ? 1u : 0u; }
      } else if (tile_lo >= 4) {
        tail = __ldg(reinterpret_cast<const uint32_t*>(text + tile_lo - 4));
      }
      // rows that hold text (or, for literals, the three bytes after it)
      uint32_t rows = tile_rows;
      if (tile_lo + tile_bytes > scan_end) rows = scan_end > tile_lo ? (uint32_t)((scan_end - tile_lo + 511) >> 9) : 0u;
      auto row = [&](uint4& v, uint32_t r) {
        const uint4 cur = v;
        v = r + kDepth < rows_ld ? EmLoadStream(src + (r + kDepth) * 32) : zero4;
        uint32_t f16 = 0;
        if (kMode == kEmGeneric) {
          uint32_t prev = 0;
          if (flt.use_sol || !flt.swar) {
            const uint32_t w3 = cur.w >> 24;
            const uint32_t my_last = (w3 == 0x0Au || w3 == 0x0Du) ? 1u : 0u;
            prev = __shfl_up_sync(kFullMask, my_last, 1);
            if (lane == 0) prev = tail;
            tail = __shfl_sync(kFullMask, my_last, 31);
          }
          uint32_t cand = EmFilterGroup(cur, prev, flt);
          const uint64_t at = mine + (uint64_t)r * 512;
          if (at + 16 > n) cand = at < n ? EmKeepBelowT(cand, (uint32_t)(n - at)) : 0u;      // bytes beyond the text
          f16 = __byte_perm(cand | (cand >> 4), 0u, 0x4420);                               // bit 8 k + j -> bit 4 k + j
        } else {
          uint32_t pw = __shfl_up_sync(kFullMask, cur.w, 1);
          if (lane == 0) pw = tail;
          tail = __shfl_sync(kFullMask, cur.w, 31);
#ifdef RJ_EM_READONLY      // tuning builds: the memory side of this tiling alone (the filter replaced by four XORs)
          if ((cur.x ^ cur.y ^ cur.z ^ cur.w ^ pw) == 0x12345679u) f16 = 1;
#else
          if (EmLitAny<kFull4>(cur, pw, lit.p4, lit.pmask)) f16 = EmLitFlags<kFull4>(cur, pw, lit.p4, lit.pmask);
#endif
        }
        const uint32_t bal = __ballot_sync(kFullMask, f16 != 0);
        if (bal) {
          if (f16) my_ent[n_ent + __popc(bal & lt_mask)] = ((r * 32u + lane) << 16) | f16;
          n_ent += __popc(bal);
        }
      };
      uint32_t r = 0;
#pragma unroll 1
      for (;;) {
#pragma unroll 1
        for (; r + kDepth <= rows && n_ent + 32 * kDepth <= kEntStop; r += kDepth) {
          if (!asked && r + kDepth + 8 >= rows) {
            asked = true;
            if (!stride && lane == 0) next_ticket = atomicAdd(reinterpret_cast<unsigned long long*>(em.sync), 1ull);
          }
#pragma unroll
          for (int u = 0; u < kDepth; ++u) row(v[u], r + u);
        }
        if (r + kDepth > rows) {
#pragma unroll
          for (int u = 0; u < kDepth - 1; ++u)
            if (r + u < rows) row(v[u], r + u);
          r = rows;
        } else if (kMode == kEmWindow) {
          // more needle hits in one tile than the list holds (one per 64 bytes): the general path.  The windows are
          // evaluated once, after the tile's last row — a call in the middle of the tile, with rows in flight, kept
          // those rows in local memory for the whole streaming loop.
          flags |= kFinDense;
          n_ent = 0;
          r = rows;
        }
        __syncwarp();
        if (kMode == kEmLiteral) {
          if (n_ent) cnt = EmEvalLiteral<false>(text, n, lit, range.own_begin, range.own_end, tile_base, tile_lo, my_ent, n_ent, my_cand, cnt, cap);
        } else if (kMode == kEmWindow) {
          const uint32_t n_hits = n_ent ? EmEvalLiteral<true>(text, n, lit, range.own_begin, hit_hi, tile_base, tile_lo, my_ent, n_ent, my_hits, 0u, kEmWinCandCap) : 0u;
          if (n_hits > kEmWinCandCap) flags |= kFinDense;
          else if (n_hits) cnt = EmEvalWindow(text, n, nfa, lit, range, tile_base, my_hits, n_hits, &prev_hit, my_cand, cnt, cap, &flags,
                                              my_ent, kEmEntCap);             // (the survivors' list is free: they have become hits)
        } else {
          const bool add_end = r >= rows && n >= tile_lo && n < tile_lo + tile_bytes;
          cnt = EmEvalGeneric(text, n, nfa, range, tile_base, tile_lo, my_ent, n_ent, add_end, my_cand, cnt, cap, &flags);
        }
        n_ent = 0;
        if (cnt > cap) { flags |= kFinDense; cnt = 0; }
        if (r >= rows) break;
      }
    }
    if (stride) next_ticket = t + stride;
    else if (!asked && lane == 0) next_ticket = atomicAdd(reinterpret_cast<unsigned long long*>(em.sync), 1ull);
    flags = __reduce_or_sync(kFullMask, flags);
    if (flags) cnt = 0;
    // ---- does every candidate begin after its predecessor ended?  Else one lane walks the chain ----------------
    bool ok = true;
    for (uint32_t base = 0; base < cnt; base += 32) {
      const uint32_t i = base + lane;
      if (i > 0 && i < cnt) {
        const uint32_t p = my_cand[i - 1], c = my_cand[i];
        ok &= EmTakes(EmAfter(EmRel(p), EmLen(p)), EmRel(c), EmLen(c));
        if (EmRel(c) <= EmRel(p)) flags |= kFinOverlap;        // not even sorted (windows of two tiles interleave)
      }
    }
    ok = __all_sync(kFullMask, ok);
    flags = __reduce_or_sync(kFullMask, flags);
    if (flags) cnt = 0;
    if (!ok && cnt) {
      // leftmost-longest over the tile's candidates, as if nothing reached in from the left
      __syncwarp();
      if (lane == 0) {
        ChainState cs;
        cs.cur = 0; cs.tail = kNoMatch;
        for (uint32_t i = 0; i < cnt; ++i) {
          const uint32_t c = my_cand[i];
          if (!ChainTake(&cs, EmRel(c), (uint64_t)EmRel(c) + EmLen(c))) my_cand[i] = EmCand(EmRel(c), kEmDropped);
        }
      }
      __syncwarp();
      uint32_t k = 0;
      for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t c = i < cnt ? my_cand[i] : EmCand(0, kEmDropped);
        __syncwarp();
        const bool keep = EmLen(c) != kEmDropped;
        const uint32_t bal = __ballot_sync(kFullMask, keep);
        if (keep) my_cand[k + __popc(bal & lt_mask)] = c;
        k += __popc(bal);
        __syncwarp();
      }
      cnt = k;
    }
    // ---- publish, look back, check the seam ---------------------------------------------------------------------
    const uint32_t first = cnt ? my_cand[0] : 0u, last = cnt ? my_cand[cnt - 1] : 0u;
    const EmState mine_out = EmAfter(tile_base + EmRel(last), EmLen(last));
    uint4* rec = em.records + 2 * t;
    const uint32_t tag = (em.seq & 0x3FFFFFFFu) << 2;
    // (fused ReplaceAll: the bytes inside this tile's matches)
    uint64_t tile_removed = 0;
    if (kRebuild) {
      uint32_t part = 0;
      for (uint32_t i = lane; i < cnt; i += 32) part += EmLen(my_cand[i]);
      tile_removed = __reduce_add_sync(kFullMask, part);
    }
    if (lane == 0) EmPublish(rec, tag | 1u, cnt, cnt != 0, mine_out, tile_removed);
    // A tile without matches needs nothing from its predecessors — no place for matches, no seam — and its own record
    // already says all there is to say about it: it does not look back at all, unless it closes its group (a look-back
    // blocks the warp with no loads in flight: on a text without hits it was a quarter of the kernel's time).
    uint64_t before = 0, removed_before = 0;
    EmState arriving;
    arriving.cur = 0; arriving.ne = 0;
    const bool closes_group = (t & 31u) == 31u || t + 1 == em.ntiles;        // the last tile of its group
    const bool look = cnt != 0 || closes_group || kRebuild;     // (every tile of a rebuild has bytes to place)
    if (look) {
      uint32_t in_group;
      EmState st_group, st_before;
      bool have_group, have_before;
      uint64_t groups_before, rem_group, rem_groups_before;
      EmGroupBefore<kRebuild>(em, t, &in_group, &st_group, &have_group, &rem_group);
      uint4* grec = em.group_records + 2 * (t >> 5);
      const uint64_t group_count = (uint64_t)in_group + cnt;
      const bool group_has = cnt != 0 || have_group;
      const EmState group_out = cnt ? mine_out : st_group;
      if (closes_group && lane == 0) EmPublish(grec, tag | 1u, group_count, group_has, group_out, rem_group + tile_removed);
      EmLookBack<kRebuild>(em, em.group_records, t >> 5, &groups_before, &st_before, &have_before, &rem_groups_before);
      before = groups_before + in_group;
      removed_before = rem_groups_before + rem_group;
      arriving = have_group ? st_group : st_before;
      if (closes_group && lane == 0)
        EmPublish(grec, tag | 2u, groups_before + group_count, true, group_has ? group_out : st_before,
                  rem_groups_before + rem_group + tile_removed);
      if (cnt && !EmTakes(arriving, tile_base + EmRel(first), EmLen(first))) flags |= kFinOverlap;   // the chain from the left reaches in
    }
    // (fused ReplaceAll) the first text byte that is mine to copy, and its place in the rebuilt text: an output byte
    // sits at (its text offset) - (bytes removed before it) + rep_w x (matches before it).  Every match counted by the
    // look-back ends at or before rb_in, so the sums must fit under it; numbers that do not (a tile before me has
    // raised a flag and published nothing, or its last match runs into my first) write nothing and send the call to
    // the separate rebuild.
    uint64_t rb_in = 0, rb_out = 0;
    if (kRebuild) {
      rb_in = tile_lo;
      if (arriving.ne && arriving.cur > rb_in) rb_in = arriving.cur;           // a match from the left ends here
      const uint64_t added = (uint64_t)em.rep_w * before;
      if (removed_before > rb_in || added > removed_before) flags |= kFinOverlap;
      rb_out = rb_in - removed_before + added;
    }
    if (lane == 0) {
      if (flags) atomicOr(&em.sync[2], flags);
      if (t + 1 == em.ntiles) {
        const EmState fin = cnt ? mine_out : arriving;
        atomicExch(&em.final_state[0], before + cnt);
        atomicExch(&em.final_state[1], fin.cur);
        atomicExch(&em.final_state[2], (unsigned long long)fin.ne);
        if (kRebuild) atomicExch(&em.final_state[3], removed_before + tile_removed);
      }
    }
    if (kRebuild) {
      // ---- fused ReplaceAll: my part of the rebuilt text.  The bytes between my matches are copied (lanes take
      // consecutive bytes: a gap is a few coalesced sectors), each match leaves the replacement string.  Four gaps per
      // step, their loads issued together: one gap per step was one load latency per match (a FASTA tile has 400).
      if (!flags) {
        const uint64_t tile_hi = tile_lo + tile_bytes < n ? tile_lo + tile_bytes : n;
        uint8_t* __restrict__ outb = em.rep_out;
        uint64_t in_pos = rb_in, o = rb_out;
        const uint32_t w = em.rep_w;
        const uint8_t wb = (uint32_t)lane < w ? __ldg(em.rep_with + lane) : (uint8_t)0;
#pragma unroll 1
        for (uint32_t i = 0; i <= cnt; i += 4) {
          uint64_t src[4], dst[4];
          uint32_t gap[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t k = i + u;
            uint64_t b = in_pos, e = in_pos;                  // (entries past the end: nothing)
            if (k < cnt) { const uint32_t c = my_cand[k]; b = tile_base + EmRel(c); e = b + EmLen(c); }
            else if (k == cnt) { b = tile_hi; e = tile_hi; }  // the bytes after my last match
            src[u] = in_pos;
            gap[u] = b > in_pos ? (uint32_t)(b - in_pos) : 0u;
            dst[u] = o;
            o += gap[u] + (k < cnt ? w : 0u);
            if (e > in_pos) in_pos = e;
          }
          uint8_t x[4], y[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            x[u] = 0; y[u] = 0;
            if ((uint32_t)lane < gap[u]) x[u] = text[src[u] + lane];
            if ((uint32_t)lane + 32u < gap[u]) y[u] = text[src[u] + 32 + lane];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if ((uint32_t)lane < gap[u]) outb[dst[u] + lane] = x[u];
            if ((uint32_t)lane + 32u < gap[u]) outb[dst[u] + 32 + lane] = y[u];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            // a long gap: eight bytes per lane in flight
            uint32_t k = 64u + lane;
            for (; k + 224u < gap[u]; k += 256u) {
              uint8_t z[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) z[q] = text[src[u] + k + 32 * q];
#pragma unroll
              for (int q = 0; q < 8; ++q) outb[dst[u] + k + 32 * q] = z[q];
            }
            for (; k < gap[u]; k += 32u) outb[dst[u] + k] = text[src[u] + k];
            if (i + u < cnt && w) {
              const uint64_t at = dst[u] + gap[u];
              if ((uint32_t)lane < w) outb[at + lane] = wb;
              for (uint32_t q = 32u + lane; q < w; q += 32u) outb[at + q] = __ldg(em.rep_with + q);
            }
          }
        }
      }
    } else {
      // ---- my matches, at their final place ----------------------------------------------------------------------
      ulonglong2* outp = reinterpret_cast<ulonglong2*>(em.out_pairs);
      for (uint32_t i = lane; i < cnt; i += 32) {
        const uint32_t c = my_cand[i];
        const uint64_t b = tile_base + EmRel(c) + em.base_offset;
        if (before + i < em.out_cap) outp[before + i] = make_ulonglong2(b, b + EmLen(c));
      }
    }
    __syncwarp();
    // ---- the warp that finishes the last tile reports --------------------------------------------------------------
    // (no fence per tile: a gpu-scope fence invalidates the SM's L1 under the other warps' streaming loads; only the
    // last tile's totals must be visible to the reporter, and the host reads the pairs after the kernel has ended)
    // Every tile counts itself done with a reduction nobody waits for; the warp of the LAST tile waits until all
    // have (it is the last to start, so not for long) and reports.  A tile that raised a flag makes it visible first.
    // (a warp adds up its finished tiles and tells the counter once, when it leaves or when it holds the last tile:
    // one reduction per tile on one address was 150 k of them for a 5 GB text)
    if (lane == 0) {
      if (flags) __threadfence();
      if (t + 1 != em.ntiles) {
        ++done_mine;
      } else {
        if (done_mine) atomicAdd(&em.sync[3], done_mine);
        done_mine = 0;
        __threadfence();
        const volatile unsigned int* done = em.sync + 3;
        for (uint32_t polls = 0; (uint64_t)*done + 1 < em.ntiles; ++polls) {
          if (polls > (1u << 24)) { atomicOr(&em.sync[2], kFinOverlap | kFinStuck); break; }
          __nanosleep(100);
        }
        __threadfence();
        const unsigned int fl = atomicOr(&em.sync[2], 0u);
        const unsigned long long total = __ldcg(&em.final_state[0]), cur = __ldcg(&em.final_state[1]);
        const unsigned int ne = (unsigned int)__ldcg(&em.final_state[2]);
        // FinRecord: last_end / last_nonempty as StatusFromRecord expects them (kFinLastEmpty: cur = last_end + 1)
        const unsigned long long last_end = ne ? cur : (cur ? cur - 1 : 0), last_ne = ne ? cur : 0;
        const unsigned int flg = fl | ((total && !ne) ? kFinLastEmpty : 0u);
        volatile uint4* dst = reinterpret_cast<volatile uint4*>(em.host_records);
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"((unsigned int)total),
                     "r"((unsigned int)(total >> 32)), "r"(flg), "r"(em.seq) : "memory");
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 1), "r"((unsigned int)last_end),
                     "r"((unsigned int)(last_end >> 32)), "r"(0u), "r"(em.seq) : "memory");
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 2), "r"((unsigned int)last_ne),
                     "r"((unsigned int)(last_ne >> 32)), "r"(0u), "r"(em.seq) : "memory");
        if (kRebuild) {
          // a second FinRecord: n_matches = bytes removed (the host computes the length of the rebuilt text from it)
          const unsigned long long removed = __ldcg(&em.final_state[3]);
          asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 3), "r"((unsigned int)removed),
                       "r"((unsigned int)(removed >> 32)), "r"(0u), "r"(em.seq) : "memory");
          asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 4), "r"(0u), "r"(0u), "r"(0u), "r"(em.seq) : "memory");
          asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 5), "r"(0u), "r"(0u), "r"(0u), "r"(em.seq) : "memory");
        }
      }
    }
  }
}

}  // namespace rejit_b200

#endif  // REJIT_B200_CUDA_SCAN_EMIT_CUH_
