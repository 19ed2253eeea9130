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

}  // namespace rejit_b200

#endif  // REJIT_B200_CUDA_DEVICE_PROGRAM_H_
