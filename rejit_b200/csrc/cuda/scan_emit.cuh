// rejit_b200 — single-pass scan + ordered emit (round 2).
//
// One kernel reads the text ONCE and writes the matches ONCE, at their final
// place, in text order — no slot ranges, no grid barrier, no second kernel:
//
//   tiles     the text is cut into tiles of 64 KB; a CTA (8 warps) takes tiles in
//             order from a ticket counter; warp w owns bytes [8 KB w, 8 KB (w+1))
//             of the tile and streams them as 16 rows of 512 bytes (a lane holds 16
//             bytes of a row: one coalesced 16-byte load, four rows in flight).
//   filter    literal: the first <= 4 needle bytes at the lane's 16 alignments
//             (funnel shifts, the straddling word from the next lane);
//             generic: "can a match begin at this byte" as SWAR byte compares
//             (<= 4 start bytes and "right after a line break") or a 256-bit map.
//             Lanes with a survivor leave one word in shared memory (a ballot and
//             a store; nothing else happens inside the streaming loop).
//   evaluate  after its rows the warp turns survivors into candidates
//             (begin, E(begin)): the rest of the needle / one NFA run per start
//             (device_program.h NfaRun) / for "required literal + window" patterns
//             one NFA run per start in front of every needle hit.  One lane per
//             candidate, results compacted in order.
//   select    candidates of one warp are sorted by construction.  When every
//             candidate begins after its predecessor ended (the rule of ChainTake)
//             the candidates ARE the matches; otherwise one thread walks the tile's
//             lists with ChainTake (leftmost-longest, /root/reference/src/
//             x64/codegen-x64.cc:401-522, src/codegen.cc:36-86).
//   place     the CTA publishes {count, chain state} of its tile and looks back
//             over the tiles before it (decoupled look-back: records tagged with
//             the call's sequence number, 32 predecessors per step) for the number
//             of matches before it and the chain state arriving from the left; a
//             state that reaches into the tile (a match straddling the tile edge)
//             cannot be repaired locally — counts are already published — so it
//             raises kFinOverlap and the host runs the general path instead.
//   report    the CTA that finishes last writes the FinRecord to mapped host memory.
//
// Algorithmic traffic: N bytes read + 16 bytes written per match; the look-back
// records (32 bytes per 64 KB tile) stay in L2.
#ifndef REJIT_B200_CUDA_SCAN_EMIT_CUH_
#define REJIT_B200_CUDA_SCAN_EMIT_CUH_

#include "kernels.cuh"

namespace rejit_b200 {

constexpr uint32_t kEmWarps = 8;
constexpr uint32_t kEmThreads = kEmWarps * 32;
constexpr uint32_t kEmRows = 16;                              // rows of 512 bytes per warp and tile
constexpr uint32_t kEmWarpBytes = kEmRows * 512;              // 8 KB
constexpr uint32_t kEmTileBytes = kEmWarps * kEmWarpBytes;    // 64 KB
constexpr uint32_t kEmEntCap = kEmWarpBytes / 16;             // every 16-byte group may hold a survivor
constexpr uint32_t kEmCandCap = 512;                          // candidates per warp and tile
constexpr uint32_t kEmWinCandCap = 256;                       // ... in window mode (the other half holds the needle hits)
constexpr uint32_t kEmBias = 8192;                            // offsets in a tile are relative to tile_lo - kEmBias
constexpr uint32_t kEmDropped = 0xFFFFFFFFu;                  // length of a candidate the chain did not take
constexpr uint32_t kEmPending = 0xFFFFFFFEu;                  // length of a start that has not been evaluated yet
constexpr unsigned int kFinLastEmpty = 8u;                    // FinRecord.flags: the last match is empty
constexpr unsigned int kFinStuck = 16u;                       // a look-back gave up waiting (never expected)
constexpr size_t kEmSmemBytes = kEmWarps * (kEmEntCap * 4 + kEmCandCap * 8);

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
  uint4* records;                   // [ntiles][2]: look-back records
  unsigned int* sync;               // [0..1] ticket (64 bit), [2] flags, [3] tiles done; zeroed by the host before the launch
  unsigned long long* final_state;  // [3]: total matches, chain state (cur, non-empty); written by the last tile
  uint64_t tile0, ntiles;           // tiles [tile0, tile0 + ntiles) hold every owned start (and needle hit)
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

__device__ __forceinline__ uint4 EmLoadRow(const uint8_t* __restrict__ text, uint64_t n16, uint64_t at) {
  return at < n16 ? __ldg(reinterpret_cast<const uint4*>(text + at)) : make_uint4(0, 0, 0, 0);
}

// ===========================================================================
// filters: the warp's 16 rows -> entries {group << 16 | 16 flag bits} in position order
// ===========================================================================
// literal: flag bits in natural order (bit j = the first min(m, 4) needle bytes match at byte j)
template <bool kFull4>
__device__ __forceinline__ uint32_t EmScanLiteral(const uint8_t* __restrict__ text, uint64_t n, uint64_t n16, uint64_t warp_lo,
                                                   uint32_t p4, uint32_t pmask, uint32_t* my_ent) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint64_t mine = warp_lo + (uint64_t)lane * 16;
  uint32_t n_ent = 0;
  uint4 nxt[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) nxt[u] = EmLoadRow(text, n16, mine + (uint64_t)u * 512);
#pragma unroll 1
  for (uint32_t r0 = 0; r0 < kEmRows; r0 += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = nxt[u];
    if (r0 + 4 < kEmRows) {
#pragma unroll
      for (int u = 0; u < 4; ++u) nxt[u] = EmLoadRow(text, n16, mine + (uint64_t)(r0 + 4 + u) * 512);
    } else {
      // the word that follows the warp's bytes
      const uint64_t after = warp_lo + kEmWarpBytes;
      nxt[0].x = (lane == 0 && after < n16) ? __ldg(reinterpret_cast<const uint32_t*>(text + after)) : 0u;
    }
    if (warp_lo + (uint64_t)r0 * 512 >= n) continue;          // rows beyond the text (uniform)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t up = __shfl_down_sync(kFullMask, v[u].x, 1);
      const uint32_t wrap = __shfl_sync(kFullMask, (u < 3) ? v[u + 1].x : nxt[0].x, 0);
      const uint32_t nx = (lane == 31) ? wrap : up;
      const bool any = LitAny<kFull4>(v[u], nx, p4, pmask);
      const uint32_t bal = __ballot_sync(kFullMask, any);
      if (bal) {
        if (any) my_ent[n_ent + __popc(bal & lt_mask)] = (((r0 + u) * 32u + lane) << 16) | LitMask(v[u], nx, p4, pmask);
        n_ent += __popc(bal);
      }
    }
  }
  return n_ent;
}

// generic: flag bits transposed (EmEqT)
__device__ __forceinline__ uint32_t EmFilterGroup(const uint4& v, uint32_t prev_is_break, const EmFilter& f, uint32_t* brk_out) {
  if (f.swar) {
    uint32_t m0 = 0;
    for (uint32_t i = 0; i < f.n0; ++i) m0 |= EmEqT(v, (f.b0 >> (8 * i)) & 0xFFu);
    if (!f.use_sol) { *brk_out = 0; return m0; }
    const uint32_t brk = EmEqT(v, 0x0Au) | EmEqT(v, 0x0Du);
    *brk_out = brk;
    uint32_t m1 = 0x0F0F0F0Fu;
    if (!f.all1) {
      m1 = 0;
      for (uint32_t i = 0; i < f.n1; ++i) m1 |= EmEqT(v, (f.b1 >> (8 * i)) & 0xFFu);
    }
    return m0 | (EmShiftT(brk, prev_is_break) & m1);
  }
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t cand = 0, brk = 0;
  uint32_t sol = prev_is_break;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const uint32_t c = (w[p >> 2] >> (8 * (p & 3))) & 0xFFu;
    const uint32_t bit = (f.t[sol][c >> 5] >> (c & 31)) & 1u;
    cand |= bit << (((p & 3) << 3) | (p >> 2));
    sol = (c == 0x0Au || c == 0x0Du) ? 1u : 0u;
    brk |= sol << (((p & 3) << 3) | (p >> 2));
  }
  *brk_out = brk;
  return cand;
}

__device__ __forceinline__ uint32_t EmScanGeneric(const uint8_t* __restrict__ text, uint64_t n, uint64_t n16, uint64_t warp_lo,
                                                   const EmFilter& f, uint32_t* my_ent) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint64_t mine = warp_lo + (uint64_t)lane * 16;
  uint32_t n_ent = 0;
  // is the byte before the warp's first byte a line break (offset 0: the text start counts as one)
  uint32_t carry = 1u;
  if (warp_lo > 0) { const uint8_t pb = (warp_lo - 1 < n) ? __ldg(text + warp_lo - 1) : 0; carry = (pb == 0x0Au || pb == 0x0Du) ? 1u : 0u; }
  uint4 nxt[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) nxt[u] = EmLoadRow(text, n16, mine + (uint64_t)u * 512);
#pragma unroll 1
  for (uint32_t r0 = 0; r0 < kEmRows; r0 += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = nxt[u];
    if (r0 + 4 < kEmRows) {
#pragma unroll
      for (int u = 0; u < 4; ++u) nxt[u] = EmLoadRow(text, n16, mine + (uint64_t)(r0 + 4 + u) * 512);
    }
    if (warp_lo + (uint64_t)r0 * 512 >= n) continue;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t at = mine + (uint64_t)(r0 + u) * 512;
      // the line-break flag of the byte before my 16: my left neighbour's last byte (lane 0: the row before)
      uint32_t brk = 0, cand;
      if (f.use_sol || !f.swar) {
        const uint32_t w3 = v[u].w >> 24;
        const uint32_t my_last = (w3 == 0x0Au || w3 == 0x0Du) ? 1u : 0u;
        uint32_t prev = __shfl_up_sync(kFullMask, my_last, 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(kFullMask, my_last, 31);
        cand = EmFilterGroup(v[u], prev, f, &brk);
      } else {
        cand = EmFilterGroup(v[u], 0u, f, &brk);
      }
      if (at + 16 > n) cand = at < n ? EmKeepBelowT(cand, (uint32_t)(n - at)) : 0u;      // bytes beyond the text
      const uint32_t bal = __ballot_sync(kFullMask, cand != 0);
      if (bal) {
        // 28-bit transposed flags (bit 8 k + j) -> 16 bits (bit 4 k + j)
        if (cand) my_ent[n_ent + __popc(bal & lt_mask)] = (((r0 + u) * 32u + lane) << 16) | __byte_perm(cand | (cand >> 4), 0u, 0x4420);
        n_ent += __popc(bal);
      }
    }
  }
  return n_ent;
}


// ===========================================================================
// evaluation: entries -> candidates {begin - tile_base, length}, in position order.
// Every function is called by the whole warp and returns the number of candidates
// (a value above `cap` means the list overflowed: the tile is too dense).
// ===========================================================================
// literal occurrences (natural flag order): bounds, ownership, the rest of the needle.  kHits: the results are
// needle hits (u32 offsets) for the window stage instead of candidates.
template <bool kHits>
__device__ __forceinline__ uint32_t EmEvalLiteral(const uint8_t* __restrict__ text, uint64_t n, const EmLit& lit,
                                                   uint64_t own_lo, uint64_t own_hi, uint64_t tile_base, uint64_t warp_lo,
                                                   const uint32_t* my_ent, uint32_t n_ent, void* out, uint32_t cap) {
  const int lane = threadIdx.x & 31;
  uint32_t k = 0;
  for (uint32_t base = 0; base < n_ent; base += 32) {
    const uint32_t ent = base + lane < n_ent ? my_ent[base + lane] : 0u;
    uint32_t valid = ent & 0xFFFFu;
    const uint64_t my = warp_lo + (uint64_t)(ent >> 16) * 16;
    for (uint32_t hh = valid; hh; hh &= hh - 1) {
      const int j = __ffs(hh) - 1;
      const uint64_t pos = my + j;
      bool ok = pos >= own_lo && pos < own_hi && pos + lit.m <= n;
      for (uint32_t i = 4; i < lit.m && ok; ++i) ok = (__ldg(text + pos + i) == __ldg(lit.needle + i));
      if (!ok) valid &= ~(1u << j);
    }
    const uint32_t c = __popc(valid);
    const uint32_t incl = WarpInclusiveScan(c);
    uint32_t idx = k + incl - c;
    for (; valid; valid &= valid - 1) {
      const uint32_t rel = (uint32_t)(my + (__ffs(valid) - 1) - tile_base);
      if (idx < cap) {
        if (kHits) static_cast<uint32_t*>(out)[idx] = rel;
        else static_cast<uint2*>(out)[idx] = make_uint2(rel, lit.m);
      }
      ++idx;
    }
    k += __shfl_sync(kFullMask, incl, 31);
  }
  __syncwarp();
  return k;
}

// one NFA run per start in front of every needle hit (the windows of neighbouring hits of this warp are clipped
// against each other so that every start is tried once, in order)
__device__ __forceinline__ uint32_t EmEvalWindow(const uint8_t* __restrict__ text, uint64_t n, const NfaTables& nfa,
                                                  const EmLit& lit, const ScanRange& range, uint64_t tile_base,
                                                  const uint32_t* my_hits, uint32_t n_hits, uint2* my_cand, uint32_t cap,
                                                  unsigned int* flags) {
  const int lane = threadIdx.x & 31;
  uint32_t k = 0;
  for (uint32_t q = 0; q < n_hits; ++q) {
    const uint64_t h = tile_base + my_hits[q];
    if (h < lit.win_lo) continue;
    const uint64_t s_max = h - lit.win_lo;                     // inclusive
    uint64_t s_min = h >= lit.win_hi ? h - lit.win_hi : 0;
    if (q > 0) {
      const uint64_t prev = tile_base + my_hits[q - 1];
      if (prev >= lit.win_lo && prev - lit.win_lo + 1 > s_min) s_min = prev - lit.win_lo + 1;
    }
    for (uint64_t base = s_min; base <= s_max; base += 32) {
      const uint64_t s = base + lane;
      uint64_t e = kNoMatch;
      if (s <= s_max && s >= range.own_begin && s < range.own_end && s < n) {
        const int ctx = nfa.has_anchor ? ContextAt(text, n, s) : 0;
        if (nfa.start_ok[ctx * 256 + text[s]]) e = NfaRunAny(nfa, text, n, s);
      }
      __syncwarp();
      const bool has = e != kNoMatch;
      if (has && e - s >= kEmPending) { *flags |= kFinDense; e = s; }      // length does not fit: the general path
      const uint32_t bal = __ballot_sync(kFullMask, has);
      if (bal) {
        const uint32_t idx = k + __popc(bal & ((1u << lane) - 1u));
        if (has && idx < cap) my_cand[idx] = make_uint2((uint32_t)(s - tile_base), (uint32_t)(e - s));
        k += __popc(bal);
      }
    }
  }
  __syncwarp();
  return k;
}

// generic: entries (transposed flags) -> starts, in place in the candidate list, then one NFA run per start
__device__ __forceinline__ uint32_t EmEvalGeneric(const uint8_t* __restrict__ text, uint64_t n, const NfaTables& nfa,
                                                   const ScanRange& range, uint64_t tile_base, uint64_t warp_lo,
                                                   const uint32_t* my_ent, uint32_t n_ent, uint2* my_cand, uint32_t cap,
                                                   unsigned int* flags) {
  const int lane = threadIdx.x & 31;
  uint32_t n_start = 0;
  for (uint32_t base = 0; base < n_ent; base += 32) {
    const uint32_t ent = base + lane < n_ent ? my_ent[base + lane] : 0u;
    const uint32_t f16 = ent & 0xFFFFu;
    const uint64_t my = warp_lo + (uint64_t)(ent >> 16) * 16;
    const uint32_t c = __popc(f16);
    const uint32_t incl = WarpInclusiveScan(c);
    uint32_t idx = n_start + incl - c;
    if (f16) {
#pragma unroll 1
      for (uint32_t p = 0; p < 16; ++p) {
        if (!((f16 >> (((p & 3u) << 2) | (p >> 2))) & 1u)) continue;
        if (idx < cap) my_cand[idx] = make_uint2((uint32_t)(my + p - tile_base), kEmPending);
        ++idx;
      }
    }
    n_start += __shfl_sync(kFullMask, incl, 31);
  }
  // the offset n itself: only the empty match can begin there (the run decides)
  if (n >= warp_lo && n < warp_lo + kEmWarpBytes) {
    if (lane == 0 && n_start < cap) my_cand[n_start] = make_uint2((uint32_t)(n - tile_base), kEmPending);
    ++n_start;
  }
  __syncwarp();
  if (n_start > cap) return n_start;
  uint32_t k = 0;
  for (uint32_t base = 0; base < n_start; base += 32) {
    const uint32_t i = base + lane;
    uint32_t rel = 0;
    uint64_t s = 0, e = kNoMatch;
    if (i < n_start) {
      rel = my_cand[i].x;
      s = tile_base + rel;
      if (s >= range.own_begin && s < range.own_end && s <= n) e = NfaRunAny(nfa, text, n, s);
    }
    __syncwarp();
    const bool has = e != kNoMatch;
    if (has && e - s >= kEmPending) { *flags |= kFinDense; e = s; }
    const uint32_t bal = __ballot_sync(kFullMask, has);
    if (has) my_cand[k + __popc(bal & ((1u << lane) - 1u))] = make_uint2(rel, (uint32_t)(e - s));
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
__device__ __forceinline__ void EmPublish(uint4* rec, uint32_t tag, uint64_t count, bool has, const EmState& st) {
  EmStore16(rec, (uint32_t)count, (uint32_t)(count >> 32), tag, 0u);
  EmStore16(rec + 1, (uint32_t)st.cur, (uint32_t)(st.cur >> 32) | (st.ne << 31) | (has ? 1u << 30 : 0u), tag, 0u);
}

// Called by warp 0.  Returns (in every lane) the number of matches in the tiles before `t` and the chain state
// that arrives at the tile (the state after the last match before it, or the call's carry).
__device__ __forceinline__ void EmLookBack(const EmitArgs& em, uint64_t t, uint64_t* before, EmState* arriving) {
  const int lane = threadIdx.x & 31;
  uint64_t excl = 0;
  EmState st;
  st.cur = 0; st.ne = 0;
  bool have = false;
  for (int64_t base = (int64_t)t;; base -= 32) {
    const int64_t idx = base - 1 - lane;
    uint64_t cnt = 0, cur = 0;
    uint32_t state = 0, ne = 0, has = 0;
    if (idx >= 0) {
      const uint4* rec = em.records + 2 * (uint64_t)idx;
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
          break;
        }
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
}

// ===========================================================================
// the kernel
// ===========================================================================
constexpr uint32_t kEmSeqMax = 4096;        // candidates of a tile up to which one thread resolves an overlap

template <int kMode, bool kFull4>
__global__ void __launch_bounds__(kEmThreads, 4)
k_scan_emit(const uint8_t* __restrict__ text, uint64_t n, EmLit lit, NfaTables nfa, EmFilter flt, ScanRange range,
            EmitArgs em) {
  extern __shared__ __align__(16) uint8_t em_smem[];
  __shared__ unsigned long long s_ticket, s_before;
  __shared__ uint32_t s_cnt[kEmWarps], s_off[kEmWarps], s_ok[kEmWarps];
  __shared__ uint2 s_first[kEmWarps], s_last[kEmWarps];
  __shared__ unsigned int s_flags, s_mode;         // s_mode: 0 write, 1 resolve + compact first, 2 nothing to write
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t* my_ent = reinterpret_cast<uint32_t*>(em_smem) + warp * kEmEntCap;
  uint2* cand_all = reinterpret_cast<uint2*>(em_smem + kEmWarps * kEmEntCap * 4);
  uint2* my_cand = cand_all + warp * kEmCandCap;
  const uint64_t n16 = (n + 15) & ~15ull;
  const uint32_t cap = kMode == kEmWindow ? kEmWinCandCap : kEmCandCap;
  // needle hits may sit up to win_hi bytes after an owned start
  const uint64_t hit_hi = kMode == kEmWindow ? range.own_end + lit.win_hi + 1 : range.own_end;

  for (;;) {
    if (threadIdx.x == 0) { s_ticket = atomicAdd(reinterpret_cast<unsigned long long*>(em.sync), 1ull); s_flags = 0; }
    __syncthreads();
    const uint64_t t = s_ticket;
    if (t >= em.ntiles) break;
    const uint64_t tile_lo = (em.tile0 + t) * kEmTileBytes;
    const uint64_t tile_base = tile_lo >= kEmBias ? tile_lo - kEmBias : 0;
    const uint64_t warp_lo = tile_lo + (uint64_t)warp * kEmWarpBytes;
    unsigned int flags = 0;
    uint32_t cnt = 0;
    // ---- filter + evaluate: my 8 KB -----------------------------------------------------------
    if (warp_lo <= n && warp_lo < hit_hi + 16 && warp_lo + kEmWarpBytes + 16 > range.own_begin) {
      uint32_t n_ent;
      if (kMode == kEmGeneric) n_ent = EmScanGeneric(text, n, n16, warp_lo, flt, my_ent);
      else n_ent = EmScanLiteral<kFull4>(text, n, n16, warp_lo, lit.p4, lit.pmask, my_ent);
      __syncwarp();
      if (kMode == kEmLiteral) {
        cnt = n_ent ? EmEvalLiteral<false>(text, n, lit, range.own_begin, range.own_end, tile_base, warp_lo, my_ent, n_ent, my_cand, cap) : 0u;
      } else if (kMode == kEmWindow) {
        uint32_t* my_hits = reinterpret_cast<uint32_t*>(my_cand + kEmWinCandCap);       // the upper half of my list
        const uint32_t n_hits = n_ent ? EmEvalLiteral<true>(text, n, lit, range.own_begin, hit_hi, tile_base, warp_lo, my_ent, n_ent, my_hits, 2 * kEmWinCandCap) : 0u;
        if (n_hits > 2 * kEmWinCandCap) { flags |= kFinDense; }
        else if (n_hits) cnt = EmEvalWindow(text, n, nfa, lit, range, tile_base, my_hits, n_hits, my_cand, cap, &flags);
      } else {
        cnt = EmEvalGeneric(text, n, nfa, range, tile_base, warp_lo, my_ent, n_ent, my_cand, cap, &flags);
      }
      if (cnt > cap) { flags |= kFinDense; cnt = 0; }
    }
    // ---- does every candidate of my list begin after its predecessor ended? ----------------------
    bool ok = true;
    for (uint32_t base = 0; base < cnt; base += 32) {
      const uint32_t i = base + lane;
      if (i > 0 && i < cnt) {
        const uint2 p = my_cand[i - 1], c = my_cand[i];
        ok &= EmTakes(EmAfter(p.x, p.y), c.x, c.y);
      }
    }
    ok = __all_sync(kFullMask, ok);
    flags = __reduce_or_sync(kFullMask, flags);
    if (lane == 0) {
      s_cnt[warp] = cnt;
      s_ok[warp] = ok ? 1u : 0u;
      if (cnt) { s_first[warp] = my_cand[0]; s_last[warp] = my_cand[cnt - 1]; }
      if (flags) atomicOr(&s_flags, flags);
    }
    __syncthreads();
    // ---- the tile: list boundaries, one-thread resolve when candidates overlap -----------------------
    if (threadIdx.x == 0) {
      bool all_ok = true, ordered = true;
      uint32_t total = 0;
      bool seen = false;
      uint2 last = make_uint2(0, 0);
      for (uint32_t w = 0; w < kEmWarps; ++w) {
        if (!s_cnt[w]) continue;
        all_ok &= s_ok[w] != 0;
        if (seen) {
          all_ok &= EmTakes(EmAfter(last.x, last.y), s_first[w].x, s_first[w].y);
          ordered &= s_first[w].x > last.x;
        }
        seen = true;
        last = s_last[w];
        total += s_cnt[w];
      }
      unsigned int mode = 0;
      if (s_flags) mode = 2;
      else if (!all_ok) {
        if (!ordered || total > kEmSeqMax) { s_flags = kFinOverlap; mode = 2; }
        else {
          // leftmost-longest over the tile's candidates, as if nothing reached in from the left
          ChainState cs;
          cs.cur = 0; cs.tail = kNoMatch;
          for (uint32_t w = 0; w < kEmWarps; ++w) {
            uint2* list = cand_all + w * kEmCandCap;
            for (uint32_t i = 0; i < s_cnt[w]; ++i)
              if (!ChainTake(&cs, list[i].x, (uint64_t)list[i].x + list[i].y)) list[i].y = kEmDropped;
          }
          mode = 1;
        }
      }
      s_mode = mode;
    }
    __syncthreads();
    if (s_mode == 1) {
      // my list without the candidates the chain dropped
      uint32_t k = 0;
      for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t i = base + lane;
        const uint2 c = i < cnt ? my_cand[i] : make_uint2(0, kEmDropped);
        __syncwarp();
        const bool keep = c.y != kEmDropped;
        const uint32_t bal = __ballot_sync(kFullMask, keep);
        if (keep) my_cand[k + __popc(bal & ((1u << lane) - 1u))] = c;
        k += __popc(bal);
        __syncwarp();
      }
      cnt = k;
      if (lane == 0) {
        s_cnt[warp] = cnt;
        if (cnt) { s_first[warp] = my_cand[0]; s_last[warp] = my_cand[cnt - 1]; }
      }
      __syncthreads();
    } else if (s_mode == 2) {
      cnt = 0;
    }
    // ---- publish, look back, check the seam (warp 0) ---------------------------------------------------
    if (warp == 0) {
      uint32_t total = 0;
      bool seen = false;
      uint2 first = make_uint2(0, 0), last = make_uint2(0, 0);
      if (s_mode != 2)
        for (uint32_t w = 0; w < kEmWarps; ++w) {
          if (!s_cnt[w]) continue;
          if (!seen) first = s_first[w];
          seen = true;
          last = s_last[w];
          if (lane == 0) s_off[w] = total;
          total += s_cnt[w];
        }
      EmState mine = EmAfter(tile_base + last.x, last.y);
      uint4* rec = em.records + 2 * t;
      const uint32_t tag = (em.seq & 0x3FFFFFFFu) << 2;
      if (lane == 0) EmPublish(rec, tag | 1u, total, seen, mine);
      uint64_t before;
      EmState arriving;
      EmLookBack(em, t, &before, &arriving);
      if (lane == 0) {
        unsigned int fl = s_flags;
        if (seen && !EmTakes(arriving, tile_base + first.x, first.y)) fl |= kFinOverlap;   // the chain from the left reaches in
        EmPublish(rec, tag | 2u, before + total, true, seen ? mine : arriving);
        if (fl) atomicOr(&em.sync[2], fl);
        s_before = before;
        if (t + 1 == em.ntiles) {
          const EmState fin = seen ? mine : arriving;
          em.final_state[0] = before + total;
          em.final_state[1] = fin.cur;
          em.final_state[2] = fin.ne;
        }
      }
    }
    __syncthreads();
    // ---- my matches, at their final place --------------------------------------------------------------
    if (cnt) {
      const uint64_t at0 = s_before + s_off[warp];
      ulonglong2* outp = reinterpret_cast<ulonglong2*>(em.out_pairs);
      for (uint32_t i = lane; i < cnt; i += 32) {
        const uint2 c = my_cand[i];
        const uint64_t b = tile_base + c.x + em.base_offset;
        if (at0 + i < em.out_cap) outp[at0 + i] = make_ulonglong2(b, b + c.y);
      }
    }
    __syncthreads();
    // ---- the CTA that finishes the last tile reports ------------------------------------------------------
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int done = atomicAdd(&em.sync[3], 1u);
      if ((uint64_t)done + 1 == em.ntiles) {
        __threadfence();
        const unsigned int fl = __ldcg(&em.sync[2]);
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
      }
    }
  }
}

}  // namespace rejit_b200

#endif  // REJIT_B200_CUDA_SCAN_EMIT_CUH_
