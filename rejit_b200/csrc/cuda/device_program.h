// Device-side view of a compiled pattern and the per-start NFA run.
//
// NfaRun restates, for ONE start offset s, what the reference's forward
// matching loop computes for the thread that started at s
// (/root/reference/src/x64/codegen-x64.cc:535-677 with the control edges of
// :366-398): the largest end offset e such that the NFA, entered at s, has its
// exit state live at e — the "leftmost-longest" end for that start
// (SURVEY.md §8a, E(s)).  The reference's start-pointer ring and its
// older-thread-wins merge are devices for doing this for all starts in one
// sequential pass; here every start is an independent lane.
//
// Functions are __host__ __device__ so that the core logic can be exercised by
// the CPU-only unit tests through a separate test library; the product library
// never calls them on the host.
#ifndef REJIT_B200_CUDA_DEVICE_PROGRAM_H_
#define REJIT_B200_CUDA_DEVICE_PROGRAM_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define RJ_HD __host__ __device__ __forceinline__
#else
#define RJ_HD inline
#endif

namespace rejit_b200 {

constexpr int kMaxWords = 128;          // 4096 positions
constexpr uint64_t kNoMatch = ~0ull;

// All pointers are device pointers (or host pointers in the CPU unit tests).
struct NfaTables {
  int n_pos;
  int words;                            // W: 32-bit words per position set
  int has_anchor;                       // 0: context is always 0
  const uint32_t* byte_mask;            // [256][W]
  const uint32_t* first;                // [4][W]
  const uint32_t* follow;               // [4][n_pos][W]
  const uint32_t* accept;               // [4][W]
  const uint32_t* chain;                // [W]  follow == {k+1} in every context
  uint8_t accept_empty[4];
  const uint8_t* start_ok;              // [4][256]
};

RJ_HD bool IsLineBreak(uint8_t c) { return c == '\n' || c == '\r'; }

// Context of offset p (two bits: sol | eol<<1).
RJ_HD int ContextAt(const uint8_t* text, uint64_t n, uint64_t p) {
  int sol = (p == 0) || IsLineBreak(text[p - 1]);
  int eol = (p == n) || IsLineBreak(text[p]);
  return sol | (eol << 1);
}

// One run from start offset s.  Returns E(s) or kNoMatch.  If `full_only`,
// returns n when the exit state is live exactly at offset n, else kNoMatch
// (MatchFull semantics, codegen-x64.cc:252-256, 586-590).
template <int W>
RJ_HD uint64_t NfaRun(const NfaTables& t, const uint8_t* text, uint64_t n, uint64_t s,
                      bool full_only = false) {
  uint32_t cur[W];
  uint64_t best = kNoMatch;
  int ctx = t.has_anchor ? ContextAt(text, n, s) : 0;
  if (t.accept_empty[ctx] && (!full_only || s == n)) best = s;
  if (s >= n) return best;
  {
    const uint32_t* bm = t.byte_mask + (uint64_t)text[s] * W;
    const uint32_t* f = t.first + ctx * W;
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) { cur[i] = f[i] & bm[i]; any |= cur[i]; }
    if (!any) return best;
  }
  uint64_t p = s + 1;
  for (;;) {
    ctx = t.has_anchor ? ContextAt(text, n, p) : 0;
    {
      const uint32_t* acc = t.accept + ctx * W;
      uint32_t hit = 0;
#pragma unroll
      for (int i = 0; i < W; ++i) hit |= cur[i] & acc[i];
      if (hit && (!full_only || p == n)) best = p;
    }
    if (p >= n) break;
    uint32_t nxt[W];
    // chained positions advance by a one-bit shift
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      uint32_t c = cur[i] & t.chain[i];
      nxt[i] = (c << 1) | carry;
      carry = c >> 31;
    }
    const uint32_t* fol = t.follow + (uint64_t)ctx * t.n_pos * W;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      uint32_t rest = cur[i] & ~t.chain[i];
      while (rest) {
#if defined(__CUDA_ARCH__)
        int b = __ffs(rest) - 1;
#else
        int b = __builtin_ctz(rest);
#endif
        rest &= rest - 1;
        const uint32_t* row = fol + (uint64_t)(i * 32 + b) * W;
#pragma unroll
        for (int j = 0; j < W; ++j) nxt[j] |= row[j];
      }
    }
    const uint32_t* bm = t.byte_mask + (uint64_t)text[p] * W;
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) { cur[i] = nxt[i] & bm[i]; any |= cur[i]; }
    if (!any) break;
    ++p;
  }
  return best;
}

// Arbitrary-width variant (position sets live in local memory).
RJ_HD uint64_t NfaRunWide(const NfaTables& t, const uint8_t* text, uint64_t n, uint64_t s,
                          bool full_only = false) {
  const int W = t.words;
  uint32_t cur[kMaxWords], nxt[kMaxWords];
  uint64_t best = kNoMatch;
  int ctx = t.has_anchor ? ContextAt(text, n, s) : 0;
  if (t.accept_empty[ctx] && (!full_only || s == n)) best = s;
  if (s >= n) return best;
  {
    const uint32_t* bm = t.byte_mask + (uint64_t)text[s] * W;
    const uint32_t* f = t.first + ctx * W;
    uint32_t any = 0;
    for (int i = 0; i < W; ++i) { cur[i] = f[i] & bm[i]; any |= cur[i]; }
    if (!any) return best;
  }
  uint64_t p = s + 1;
  for (;;) {
    ctx = t.has_anchor ? ContextAt(text, n, p) : 0;
    {
      const uint32_t* acc = t.accept + ctx * W;
      uint32_t hit = 0;
      for (int i = 0; i < W; ++i) hit |= cur[i] & acc[i];
      if (hit && (!full_only || p == n)) best = p;
    }
    if (p >= n) break;
    uint32_t carry = 0;
    for (int i = 0; i < W; ++i) {
      uint32_t c = cur[i] & t.chain[i];
      nxt[i] = (c << 1) | carry;
      carry = c >> 31;
    }
    const uint32_t* fol = t.follow + (uint64_t)ctx * t.n_pos * W;
    for (int i = 0; i < W; ++i) {
      uint32_t rest = cur[i] & ~t.chain[i];
      while (rest) {
#if defined(__CUDA_ARCH__)
        int b = __ffs(rest) - 1;
#else
        int b = __builtin_ctz(rest);
#endif
        rest &= rest - 1;
        const uint32_t* row = fol + (uint64_t)(i * 32 + b) * W;
        for (int j = 0; j < W; ++j) nxt[j] |= row[j];
      }
    }
    const uint32_t* bm = t.byte_mask + (uint64_t)text[p] * W;
    uint32_t any = 0;
    for (int i = 0; i < W; ++i) { cur[i] = nxt[i] & bm[i]; any |= cur[i]; }
    if (!any) break;
    ++p;
  }
  return best;
}

RJ_HD uint64_t NfaRunAny(const NfaTables& t, const uint8_t* text, uint64_t n, uint64_t s,
                         bool full_only = false) {
  switch (t.words) {
    case 1: return NfaRun<1>(t, text, n, s, full_only);
    case 2: return NfaRun<2>(t, text, n, s, full_only);
    case 3: return NfaRun<3>(t, text, n, s, full_only);
    case 4: return NfaRun<4>(t, text, n, s, full_only);
    default: return NfaRunWide(t, text, n, s, full_only);
  }
}

// ---------------------------------------------------------------------------
// Match selection: leftmost-longest, non-overlapping, with the reference's
// empty-match rule (/root/reference/src/codegen.cc:36-76 and
// codegen-x64.cc:401-466, 469-522; SURVEY.md §8a-7/8 "step 2").
//
// Input: candidates (begin_i, end_i) sorted by strictly increasing begin,
// end_i = E(begin_i).  State threaded through the list:
//   cur ......... smallest offset at which the next match may begin
//   tail ........ end offset of the last selected NON-EMPTY match, or kNoMatch
// A candidate is taken iff begin >= cur and not (empty and begin == tail).
struct ChainState {
  uint64_t cur;
  uint64_t tail;
};

RJ_HD bool ChainTake(ChainState* st, uint64_t b, uint64_t e) {
  if (b < st->cur) return false;
  if (b == e) {
    if (st->tail == b) return false;
    st->cur = b + 1;
    return true;
  }
  st->cur = e;
  st->tail = e;
  return true;
}


// ---------------------------------------------------------------------------
// Exact selection for patterns whose start can be re-entered by a running
// thread (first[ctx] & follow[ctx][k] != 0 for some k: "x*y", ".{,4}t", ...).
//
// For those patterns the reference's single-pass matcher is NOT equivalent to
// "longest end per start + greedy chain": when a match [s,e) is recorded at
// offset e, the thread born at e has already lost, during the control-edge pass
// at e, every state that was occupied by a thread born inside (s,e) — and those
// occupants are then wiped by ClearStates (codegen-x64.cc:401-466, 951-987,
// 1075-1097), leaving the state empty.  Example (verified against the
// reference): ".{,4}t" on "agccttgaact" -> [0,5) [6,11); the match [5,6) is
// lost.  To stay bit-exact we replay the reference's label semantics over each
// cluster of candidates, in position form:
//     lab[k]  = start offset of the oldest thread that has just consumed a byte
//               through position k (older start wins on collision)
//     at offset p:  blocked = U follow[ctx][k] over live k   (before any wipe)
//                   exit label = min lab[k] over live k in accept[ctx]
//                                (else p itself when the empty match is possible)
//                   record [label, p), wipe every lab in (label, p)
//                   step: survivors propagate through follow & byte_mask;
//                         the newborn thread p enters first[ctx] & ~blocked
// A cluster starts at a candidate that no earlier candidate reaches
// (max earlier end < begin, strictly) and runs to the largest end in it; no
// match can begin outside candidate starts, because every recorded match is a
// genuine NFA path.
struct FaithfulScratch {
  uint64_t* lab;       // [n_pos]
  uint64_t* nlab;      // [n_pos]
  uint32_t* act;       // [W]
  uint32_t* nact;      // [W]
  uint32_t* blocked;   // [W]
};

RJ_HD int BitScan(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __ffs(x) - 1;
#else
  return __builtin_ctz(x);
#endif
}

// Candidates [i0, i1) (sorted by begin, equal begins adjacent) form one cluster.
// Writes take[j] (0/1) and fin_end[j] for j in [i0, i1).
RJ_HD void FaithfulSegment(const NfaTables& t, const uint8_t* text, uint64_t n, const uint64_t* b,
                           const uint64_t* e, uint64_t i0, uint64_t i1, FaithfulScratch sc,
                           uint32_t* take, uint64_t* fin_end) {
  const int W = t.words, P = t.n_pos;
  uint64_t* lab = sc.lab;
  uint64_t* nlab = sc.nlab;
  uint32_t* act = sc.act;
  uint32_t* nact = sc.nact;
  for (int i = 0; i < W; ++i) act[i] = 0;
  uint64_t p1 = 0;
  for (uint64_t j = i0; j < i1; ++j) { take[j] = 0; fin_end[j] = e[j]; if (e[j] > p1) p1 = e[j]; }
  int64_t last = -1;                       // index of the most recent recorded match
  for (uint64_t p = b[i0];; ++p) {
    const int ctx = t.has_anchor ? ContextAt(text, n, p) : 0;
    const uint32_t* fol = t.follow + (uint64_t)ctx * P * W;
    const uint32_t* acc = t.accept + ctx * W;
    const uint32_t* fst = t.first + ctx * W;
    uint64_t exitlab = kNoMatch;
    for (int i = 0; i < W; ++i) sc.blocked[i] = 0;
    for (int i = 0; i < W; ++i) {
      uint32_t m = act[i];
      while (m) {
        int k = i * 32 + BitScan(m);
        m &= m - 1;
        const uint32_t* row = fol + (uint64_t)k * W;
        for (int q = 0; q < W; ++q) sc.blocked[q] |= row[q];
        if ((acc[i] >> (k & 31)) & 1u) { if (lab[k] < exitlab) exitlab = lab[k]; }
      }
    }
    if (exitlab == kNoMatch && t.accept_empty[ctx]) exitlab = p;
    if (exitlab != kNoMatch) {
      // MatchAllAppendFilter: drop recorded matches that begin at or after the new one
      while (last >= (int64_t)i0 && b[last] >= exitlab) {
        take[last] = 0;
        int64_t q = last - 1;
        while (q >= (int64_t)i0 && !take[q]) --q;
        last = q;
      }
      bool drop = (exitlab == p) && last >= (int64_t)i0 && fin_end[last] == p;
      if (!drop) {
        // locate the candidate whose begin is the label
        uint64_t lo = i0, hi = i1;
        while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (b[mid] < exitlab) lo = mid + 1; else hi = mid; }
        if (lo < i1 && b[lo] == exitlab) { take[lo] = 1; fin_end[lo] = p; last = (int64_t)lo; }
      }
      for (int i = 0; i < W; ++i) {
        uint32_t m = act[i];
        while (m) {
          int bit = BitScan(m);
          m &= m - 1;
          int k = i * 32 + bit;
          if (lab[k] > exitlab && lab[k] < p) act[i] &= ~(1u << bit);
        }
      }
    }
    if (p >= p1 || p >= n) break;
    const uint32_t* bm = t.byte_mask + (uint64_t)text[p] * W;
    for (int i = 0; i < W; ++i) nact[i] = 0;
    for (int i = 0; i < W; ++i) {
      uint32_t m = act[i];
      while (m) {
        int k = i * 32 + BitScan(m);
        m &= m - 1;
        const uint32_t* row = fol + (uint64_t)k * W;
        const uint64_t lk = lab[k];
        for (int q = 0; q < W; ++q) {
          uint32_t mm = row[q] & bm[q];
          while (mm) {
            int bit = BitScan(mm);
            mm &= mm - 1;
            int j = q * 32 + bit;
            if (!((nact[q] >> bit) & 1u)) { nact[q] |= 1u << bit; nlab[j] = lk; }
            else if (lk < nlab[j]) nlab[j] = lk;
          }
        }
      }
    }
    for (int q = 0; q < W; ++q) {
      uint32_t mm = fst[q] & ~sc.blocked[q] & bm[q];
      while (mm) {
        int bit = BitScan(mm);
        mm &= mm - 1;
        int j = q * 32 + bit;
        if (!((nact[q] >> bit) & 1u)) { nact[q] |= 1u << bit; nlab[j] = p; }
      }
    }
    uint64_t* tl = lab; lab = nlab; nlab = tl;
    uint32_t* ta = act; act = nact; nact = ta;
  }
}

// ---------------------------------------------------------------------------
// ReplaceAll placement (k_replace_tiles; also driven on the CPU by
// tests/hostsim.cc).  `pairs` = the (begin,end) matches, `removed[i]` = total
// length of the matches before match i.
// ---------------------------------------------------------------------------
constexpr uint32_t kReplaceTile = 4096;

struct ReplaceTileHead {
  uint64_t m0, m1;          // the tile's own matches: begin in [tile_lo, tile_hi) (+ end of text for the last tile)
  uint64_t removed_before;  // bytes removed before tile_lo
  uint64_t skip_end;        // end of the match straddling tile_lo (>= tile_lo)
  uint64_t r0;              // removed[m0]
};

RJ_HD uint64_t ReplaceLowerBound(const uint64_t* pairs, uint64_t m, uint64_t key) {
  uint64_t lo = 0, hi = m;                 // first i with begin[i] >= key
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (pairs[2 * mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

RJ_HD void ReplaceHead(const uint64_t* pairs, const uint64_t* removed, uint64_t m, uint64_t tile_lo,
                              ReplaceTileHead* h) {
  const uint64_t m0 = h->m0;
  h->removed_before = 0;
  h->skip_end = tile_lo;
  h->r0 = (m0 < m) ? removed[m0] : 0;
  if (m0 > 0) {
    const uint64_t pb = pairs[2 * (m0 - 1)], pe = pairs[2 * (m0 - 1) + 1];
    h->removed_before = removed[m0 - 1] + ((pe < tile_lo ? pe : tile_lo) - pb);
    if (pe > h->skip_end) h->skip_end = pe;
  }
}

// One thread places input bytes [tile_lo + 16*thread, +16) of the tile and the
// replacements of the matches beginning there.  s_b / s_e: begin / clipped end of
// the own matches relative to tile_lo; s_r: removed[m0+i] - removed[m0].
RJ_HD void ReplacePlace(uint32_t thread, const uint8_t* text, uint64_t tile_lo, uint64_t tile_hi, bool last,
                               const ReplaceTileHead& h, uint32_t cnt, const uint16_t* s_b, const uint16_t* s_e,
                               const uint16_t* s_r, const uint8_t* with, uint32_t w, uint8_t* out) {
  const uint32_t span = (uint32_t)(tile_hi - tile_lo);
  const uint32_t p = thread * 16;
  if (p > span) return;
  const uint64_t out_base = tile_lo - h.removed_before + (uint64_t)w * h.m0;
  const uint32_t head_skip = (uint32_t)((h.skip_end < tile_hi ? h.skip_end : tile_hi) - tile_lo);
  uint32_t lo = 0, hi = cnt;                // i = own matches beginning before p
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (s_b[mid] < p) lo = mid + 1; else hi = mid; }
  uint32_t i = lo;
  uint32_t rem = head_skip < p ? head_skip : p;        // removed inside the tile before p
  uint32_t cur_skip = head_skip;
  if (i > 0) {
    const uint32_t pe = s_e[i - 1];
    rem += s_r[i - 1] + ((pe < p ? pe : p) - s_b[i - 1]);
    if (pe > cur_skip) cur_skip = pe;
  }
  uint64_t q = out_base + p - rem + (uint64_t)w * i;
  // the thread whose range holds the end of the text also places the matches that begin there
  const uint32_t stop = (p + 16 < span) ? p + 16 : span + ((last && p + 16 > span) ? 1u : 0u);
  for (uint32_t pos = p; pos < stop; ++pos) {
    if (i < cnt && s_b[i] == pos) {
      for (uint32_t k = 0; k < w; ++k) out[q + k] = with[k];
      q += w;
      if (s_e[i] > cur_skip) cur_skip = s_e[i];
      ++i;
    }
    if (pos < span && pos >= cur_skip) out[q++] = text[tile_lo + pos];
  }
}

}  // namespace rejit_b200

#endif  // REJIT_B200_CUDA_DEVICE_PROGRAM_H_
