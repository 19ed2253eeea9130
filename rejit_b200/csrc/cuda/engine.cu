// rejit_b200 — sm_100a matching engine: kernels and the device pipeline.
//
// What this file replaces in the reference:
//   loop A, the fast-forward scan  /root/reference/src/x64/codegen-x64.cc:1102-1403
//        -> k_lit_scan (literal / required-literal scan, 16-byte vector loads,
//           warp shuffles for the bytes straddling two lanes)
//        -> k_dfa_scan (exact table-driven scan for fixed-length alternations,
//           tables staged in shared memory)
//   loop B, the NFA active-state advance  codegen-x64.cc:535-677
//        -> NfaRun (device_program.h) driven by k_window_verify / k_generic_scan,
//           one lane per start offset, position sets as per-lane bit masks
//   match selection  codegen-x64.cc:401-522 + /root/reference/src/codegen.cc:36-86
//        -> k_resolve_small (one CTA: bitonic sort + greedy chain) or the
//           large path (radix sort + prefix-max + segment-parallel chain)
//
// Results are (begin,end) byte offsets, identical to the reference built with
// fast-forward disabled (the parity configuration).
#include "engine.h"

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "device_program.h"

namespace rejit_b200 {

// ===========================================================================
// device-side structures
// ===========================================================================
struct CandBuf {                       // unordered (begin, end) append buffer
  uint64_t* begin;
  uint64_t* end;
  unsigned long long* count;           // may run past cap: overflow marker
  uint64_t cap;
};

struct ScanRange {                     // which start offsets this launch owns
  uint64_t own_begin;                  // inclusive
  uint64_t own_end;                    // exclusive (n+1 to own the offset n)
};

struct PipelineStatus {                // one per call, read back by the host
  unsigned long long n_candidates;
  unsigned long long n_hits;
  unsigned long long n_matches;
  unsigned long long carry_cur;
  unsigned long long carry_tail;
  unsigned int overflow;               // a buffer was too small: grow and rerun
  unsigned int need_large;             // too many candidates for the one-CTA resolve
  unsigned int full_result;            // MatchFull answer
  unsigned int pad;
};

struct DfaTables {
  const uint16_t* next;                // [n_states * n_classes], entries pre-multiplied by n_classes
  const uint8_t* byte_class;           // [256]
  int n_states, n_classes;
  int first_accept_scaled;             // first accepting state * n_classes
  uint32_t match_len;
};

constexpr unsigned kFullMask = 0xFFFFFFFFu;
constexpr int kSmallResolveMax = 4096;

// ===========================================================================
// small device helpers
// ===========================================================================
__device__ __forceinline__ void AppendAggregated(const CandBuf& buf, uint64_t b, uint64_t e) {
  // warp-aggregated atomic append (one atomic per converged group of lanes)
  unsigned m = __activemask();
  int lane = threadIdx.x & 31;
  int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(buf.count, (unsigned long long)__popc(m));
  base = __shfl_sync(m, base, leader);
  unsigned long long idx = base + __popc(m & ((1u << lane) - 1u));
  if (idx < buf.cap) {
    buf.begin[idx] = b;
    buf.end[idx] = e;
  }
}

__device__ __forceinline__ uint4 LoadText16(const uint8_t* __restrict__ text, uint64_t n, uint64_t at) {
  // 16-byte vector load when the whole vector is inside the text, else a
  // zero-padded byte gather (only ever at the tail)
  if (at + 16 <= n) {
    return __ldg(reinterpret_cast<const uint4*>(text + at));
  }
  uint32_t w[4] = {0, 0, 0, 0};
  for (int i = 0; i < 16; ++i)
    if (at + i < n) w[i >> 2] |= (uint32_t)text[at + i] << (8 * (i & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint32_t LoadText4(const uint8_t* __restrict__ text, uint64_t n, uint64_t at) {
  if (at + 4 <= n) return __ldg(reinterpret_cast<const uint32_t*>(text + at));
  uint32_t w = 0;
  for (int i = 0; i < 4; ++i)
    if (at + i < n) w |= (uint32_t)text[at + i] << (8 * i);
  return w;
}

// ===========================================================================
// K1: literal scan.  One warp owns 512-byte pieces of the text (lane l holds
// bytes [16l, 16l+16) of the piece in four registers); the first min(m,4)
// needle bytes are compared at all 16 alignments with funnel shifts, the word
// straddling into the next lane comes from a shuffle.  Survivors (rare) compare
// the rest of the needle from global memory.
// Algorithmic traffic: N bytes read + 16 bytes written per occurrence.
// ===========================================================================
template <int kUnroll>
__global__ void __launch_bounds__(256)
k_lit_scan(const uint8_t* __restrict__ text, uint64_t n, const uint8_t* __restrict__ needle,
           uint32_t m, uint32_t p4, uint32_t pmask, ScanRange range, CandBuf out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t first_piece = range.own_begin / 512;
  const uint64_t last_byte = (range.own_end < n ? range.own_end : n);   // starts must be < n
  if (last_byte == 0) return;
  const uint64_t npieces = (last_byte + 511) / 512;
  for (uint64_t piece0 = first_piece + warp * kUnroll; piece0 < npieces; piece0 += nwarps * kUnroll) {
    uint4 v[kUnroll];
    uint32_t nx[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      uint64_t my = (piece0 + u) * 512 + (uint64_t)lane * 16;
      v[u] = (my < n) ? LoadText16(text, n, my) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      uint64_t my = (piece0 + u) * 512 + (uint64_t)lane * 16;
      nx[u] = __shfl_down_sync(kFullMask, v[u].x, 1);
      if (lane == 31) nx[u] = (my + 16 < n) ? LoadText4(text, n, my + 16) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint64_t my = (piece0 + u) * 512 + (uint64_t)lane * 16;
      const uint32_t w[5] = {v[u].x, v[u].y, v[u].z, v[u].w, nx[u]};
      uint32_t hits = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint32_t x = __funnelshift_r(w[j >> 2], w[(j >> 2) + 1], 8 * (j & 3));
        if (((x ^ p4) & pmask) == 0) hits |= 1u << j;
      }
      while (hits) {
        int j = __ffs(hits) - 1;
        hits &= hits - 1;
        uint64_t pos = my + j;
        if (pos < range.own_begin || pos >= range.own_end || pos + m > n) continue;
        bool ok = true;
        for (uint32_t i = 4; i < m && ok; ++i) ok = (text[pos + i] == needle[i]);
        if (ok) AppendAggregated(out, pos, pos + m);
      }
    }
  }
}

// ===========================================================================
// K2: exact DFA scan for fixed-length, anchor-free patterns.  Each lane walks
// its own contiguous sub-stream of `stream_bytes` bytes (16-byte loads), after
// warming the automaton up on the preceding round16(L-1) bytes so that its state
// at the sub-stream start equals the state of one sequential pass.  The
// transition table (pre-multiplied uint16 rows) and the byte-class map live in
// shared memory.  A match of length L ending at e is reported as (e-L, e).
// Algorithmic traffic: N bytes read + 16 bytes per match.
// ===========================================================================
__global__ void __launch_bounds__(256)
k_dfa_scan(const uint8_t* __restrict__ text, uint64_t n, DfaTables dfa, uint32_t stream_bytes,
           ScanRange range, CandBuf out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint16_t* s_next = reinterpret_cast<uint16_t*>(smem_raw);
  const int table_entries = dfa.n_states * dfa.n_classes;
  uint8_t* s_class = smem_raw + ((table_entries * 2 + 15) & ~15);
  for (int i = threadIdx.x; i < table_entries; i += blockDim.x) s_next[i] = dfa.next[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_class[i] = dfa.byte_class[i];
  __syncthreads();

  const uint32_t L = dfa.match_len;
  const uint32_t warm = (L - 1 + 15) & ~15u;
  const uint64_t n_streams = (n + stream_bytes - 1) / stream_bytes;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  const int acc = dfa.first_accept_scaled;
  for (uint64_t sidx = tid; sidx < n_streams; sidx += nthreads) {
    const uint64_t a = sidx * stream_bytes;
    const uint64_t b = (a + stream_bytes < n) ? a + stream_bytes : n;
    // skip sub-streams that cannot contain an owned match end
    if (b + 0 <= range.own_begin || a >= range.own_end + L) continue;
    uint64_t p = (a >= warm) ? a - warm : 0;
    uint32_t state = 0;
    for (; p < b; p += 16) {
      uint4 v = LoadText16(text, n, p);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t peak = 0;
      uint32_t s0 = state;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        uint32_t st = s_next[state + s_class[c]];
        // bytes at or beyond b belong to the next lane: freeze the state there
        state = (p + i < b) ? st : state;
        peak = max(peak, state);
      }
      if (peak >= (uint32_t)acc) {
        // rare: replay the 16 bytes to find the exact end offsets
        uint32_t st = s0;
        for (int i = 0; i < 16 && p + i < b; ++i) {
          uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
          st = s_next[st + s_class[c]];
          if (st >= (uint32_t)acc) {
            uint64_t e = p + i + 1;
            if (e > a && e >= L) {
              uint64_t s = e - L;
              if (s >= range.own_begin && s < range.own_end) AppendAggregated(out, s, e);
            }
          }
        }
      }
    }
  }
}

// ===========================================================================
// K3: generic scan — one lane per start offset: start filter on the first
// byte, then the per-start NFA run.
// ===========================================================================
__global__ void __launch_bounds__(256)
k_generic_scan(const uint8_t* __restrict__ text, uint64_t n, NfaTables nfa, ScanRange range,
               CandBuf out) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t s = range.own_begin + tid; s < range.own_end && s <= n; s += nthreads) {
    int ctx = nfa.has_anchor ? ContextAt(text, n, s) : 0;
    bool ok = nfa.accept_empty[ctx] || (s < n && nfa.start_ok[ctx * 256 + text[s]]);
    if (!ok) continue;
    uint64_t e = NfaRunAny(nfa, text, n, s);
    if (e != kNoMatch) AppendAggregated(out, s, e);
  }
}

// ===========================================================================
// K4: verify the window of possible starts in front of every needle hit.
// ===========================================================================
__global__ void __launch_bounds__(256)
k_window_verify(const uint8_t* __restrict__ text, uint64_t n, NfaTables nfa, CandBuf hits,
                uint32_t lo, uint32_t hi, ScanRange range, CandBuf out) {
  unsigned long long nh = *hits.count;
  if (nh > hits.cap) nh = hits.cap;
  const uint32_t wsize = hi - lo + 1;
  const uint64_t total = (uint64_t)nh * wsize;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t idx = tid; idx < total; idx += nthreads) {
    uint64_t h = hits.begin[idx / wsize];
    uint32_t j = (uint32_t)(idx % wsize);        // start = h - hi + j
    if (h + j < hi) continue;
    uint64_t s = h + j - hi;
    if (s < range.own_begin || s >= range.own_end) continue;
    int ctx = nfa.has_anchor ? ContextAt(text, n, s) : 0;
    if (!(s < n && nfa.start_ok[ctx * 256 + text[s]])) continue;
    uint64_t e = NfaRunAny(nfa, text, n, s);
    if (e != kNoMatch) AppendAggregated(out, s, e);
  }
}

// ===========================================================================
// MatchFull: one sequential run from offset 0 (a single lane; MatchFull is a
// latency-bound sibling of the hot path).
// ===========================================================================
__global__ void k_match_full(const uint8_t* __restrict__ text, uint64_t n, NfaTables nfa,
                             PipelineStatus* status) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint64_t e = NfaRunAny(nfa, text, n, 0, /*full_only=*/true);
    status->full_result = (e == n) ? 1u : 0u;
  }
}

// ===========================================================================
// Resolve, small path: one CTA sorts the candidates by begin (bitonic, shared
// memory) and walks the chain.
// ===========================================================================
struct FaithfulArgs {               // only used for re-entrant patterns
  int enabled;
  NfaTables nfa;
  const uint8_t* text;
  uint64_t n;
  uint8_t* scratch;                  // per-walker label scratch
  uint64_t scratch_stride;           // bytes per walker
  uint32_t* take;                    // [candidates]
  uint64_t* fin_end;                 // [candidates]
};

__device__ __forceinline__ FaithfulScratch WalkerScratch(const FaithfulArgs& fa, uint64_t walker) {
  uint8_t* base = fa.scratch + walker * fa.scratch_stride;
  const uint64_t P = fa.nfa.n_pos > 0 ? fa.nfa.n_pos : 1;
  const uint64_t W = fa.nfa.words;
  FaithfulScratch sc;
  sc.lab = reinterpret_cast<uint64_t*>(base);
  sc.nlab = sc.lab + P;
  sc.act = reinterpret_cast<uint32_t*>(sc.nlab + P);
  sc.nact = sc.act + W;
  sc.blocked = sc.nact + W;
  return sc;
}

__global__ void __launch_bounds__(1024)
k_resolve_small(CandBuf cand, Carry carry_in, uint64_t base_offset, uint64_t* __restrict__ out_pairs,
                uint64_t out_cap, FaithfulArgs fa, PipelineStatus* status) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* kb = reinterpret_cast<uint64_t*>(smem_raw);
  unsigned long long m = *cand.count;
  if (threadIdx.x == 0) status->n_candidates = m;
  if (m > cand.cap) {
    if (threadIdx.x == 0) status->overflow = 1;
    return;
  }
  if (m > (unsigned long long)kSmallResolveMax) {
    if (threadIdx.x == 0) status->need_large = 1;
    return;
  }
  int count = (int)m;
  int padded = 1;
  while (padded < count) padded <<= 1;
  uint64_t* ke = kb + padded;
  for (int i = threadIdx.x; i < padded; i += blockDim.x) {
    kb[i] = (i < count) ? cand.begin[i] : ~0ull;
    ke[i] = (i < count) ? cand.end[i] : ~0ull;
  }
  __syncthreads();
  for (int k = 2; k <= padded; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < padded; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          bool up = ((i & k) == 0);
          uint64_t a = kb[i], b = kb[ixj];
          if ((a > b) == up) {
            kb[i] = b; kb[ixj] = a;
            uint64_t t = ke[i]; ke[i] = ke[ixj]; ke[ixj] = t;
          }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0 && fa.enabled) {
    // re-entrant pattern: replay the reference's thread labels cluster by cluster
    FaithfulScratch sc = WalkerScratch(fa, 0);
    int i = 0;
    while (i < count) {
      uint64_t reach = ke[i];
      int j = i + 1;
      while (j < count && !(reach < kb[j])) { reach = reach > ke[j] ? reach : ke[j]; ++j; }
      FaithfulSegment(fa.nfa, fa.text, fa.n, kb, ke, (uint64_t)i, (uint64_t)j, sc, fa.take, fa.fin_end);
      i = j;
    }
    unsigned long long taken = 0;
    uint64_t last_end = carry_in.cur;
    for (int q = 0; q < count; ++q) {
      if (!fa.take[q]) continue;
      if (taken < out_cap) {
        out_pairs[2 * taken] = kb[q] + base_offset;
        out_pairs[2 * taken + 1] = fa.fin_end[q] + base_offset;
      }
      last_end = fa.fin_end[q] > kb[q] ? fa.fin_end[q] : kb[q] + 1;
      ++taken;
    }
    status->n_matches = taken;
    status->carry_cur = last_end;
    status->carry_tail = kNoMatch;
  } else if (threadIdx.x == 0) {
    ChainState st{carry_in.cur, carry_in.tail};
    unsigned long long taken = 0;
    uint64_t prev_b = ~0ull;
    for (int i = 0; i < count; ++i) {
      uint64_t b = kb[i], e = ke[i];
      if (b == prev_b) continue;           // duplicate start (overlapping windows)
      prev_b = b;
      if (ChainTake(&st, b, e)) {
        if (taken < out_cap) {
          out_pairs[2 * taken] = b + base_offset;
          out_pairs[2 * taken + 1] = e + base_offset;
        }
        ++taken;
      }
    }
    status->n_matches = taken;
    status->carry_cur = st.cur;
    status->carry_tail = st.tail;
  }
}

// ===========================================================================
// Resolve, large path (after a radix sort by begin):
//   reach[i] = max(carry.cur, max_{j<i} end[j])           (exclusive max scan)
//   i is a restart point when no earlier candidate can influence it:
//     reach[i] < begin[i], or reach[i] == begin[i] and the candidate is non-empty
//   every restart point walks its segment sequentially (ChainTake);
//   an exclusive sum over the take flags places the matches.
// ===========================================================================
struct MaxOp {
  __host__ __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a > b ? a : b; }
};

__device__ __forceinline__ bool IsRestart(const uint64_t* b, const uint64_t* e, const uint64_t* reach, uint64_t i) {
  return reach[i] < b[i] || (reach[i] == b[i] && e[i] > b[i]);
}

__global__ void k_segment_chain(const uint64_t* __restrict__ b, const uint64_t* __restrict__ e,
                                const uint64_t* __restrict__ reach, uint64_t m, Carry carry_in,
                                uint32_t* __restrict__ take) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = tid; i < m; i += nthreads) {
    bool head = (i == 0) || (b[i] != b[i - 1] && IsRestart(b, e, reach, i));
    if (!head) continue;
    ChainState st;
    if (i == 0) { st.cur = carry_in.cur; st.tail = carry_in.tail; }
    else { st.cur = 0; st.tail = kNoMatch; }
    uint64_t prev_b = kNoMatch;
    for (uint64_t j = i; j < m; ++j) {
      if (j > i && b[j] != b[j - 1] && IsRestart(b, e, reach, j)) break;
      if (b[j] == prev_b) { take[j] = 0; continue; }
      prev_b = b[j];
      take[j] = ChainTake(&st, b[j], e[j]) ? 1u : 0u;
    }
  }
}

// Large path for re-entrant patterns: one walker per cluster head (no earlier
// candidate reaches the head's begin, strictly).
__global__ void k_segment_faithful(const uint64_t* __restrict__ b, const uint64_t* __restrict__ e,
                                   const uint64_t* __restrict__ reach, uint64_t m, FaithfulArgs fa) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  FaithfulScratch sc = WalkerScratch(fa, tid);
  for (uint64_t i = tid; i < m; i += nthreads) {
    bool head = (i == 0) || (reach[i] < b[i]);
    if (!head) continue;
    uint64_t j = i + 1;
    while (j < m && !(reach[j] < b[j])) ++j;
    FaithfulSegment(fa.nfa, fa.text, fa.n, b, e, i, j, sc, fa.take, fa.fin_end);
  }
}

__global__ void k_scatter_matches(const uint64_t* __restrict__ b, const uint64_t* __restrict__ e,
                                  const uint32_t* __restrict__ take, const uint64_t* __restrict__ slot,
                                  uint64_t m, uint64_t base_offset, uint64_t* __restrict__ out_pairs,
                                  uint64_t out_cap, unsigned long long* last_any,
                                  unsigned long long* last_nonempty) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = tid; i < m; i += nthreads) {
    if (!take[i]) continue;
    uint64_t at = slot[i];
    if (at < out_cap) {
      out_pairs[2 * at] = b[i] + base_offset;
      out_pairs[2 * at + 1] = e[i] + base_offset;
    }
    atomicMax(last_any, (unsigned long long)(i + 1));
    if (e[i] > b[i]) atomicMax(last_nonempty, (unsigned long long)(i + 1));
  }
}

__global__ void k_finish_large(const uint64_t* __restrict__ b, const uint64_t* __restrict__ e,
                               const uint32_t* __restrict__ take, const uint64_t* __restrict__ slot,
                               uint64_t m, Carry carry_in, const unsigned long long* last_any,
                               const unsigned long long* last_nonempty, PipelineStatus* status) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  status->n_matches = m ? slot[m - 1] + take[m - 1] : 0;
  uint64_t cur = carry_in.cur, tail = carry_in.tail;
  if (*last_any) {
    uint64_t i = *last_any - 1;
    cur = (e[i] > b[i]) ? e[i] : b[i] + 1;
  }
  if (*last_nonempty) tail = e[*last_nonempty - 1];
  status->carry_cur = cur;
  status->carry_tail = tail;
}

__global__ void k_widen_flags(const uint32_t* __restrict__ take, uint64_t* __restrict__ wide, uint64_t m) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = tid; i < m; i += nthreads) wide[i] = take[i];
}

__global__ void k_fill_u32(uint32_t* p, uint64_t count, uint32_t v) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = tid; i < count; i += nthreads) p[i] = v;
}

// ===========================================================================
// host side: device contexts
// ===========================================================================
namespace {

bool Check(cudaError_t e, const char* what, std::string* error) {
  if (e == cudaSuccess) return true;
  if (error) *error = std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e);
  return false;
}
#define RJ_TRY(call)                                 \
  do {                                               \
    if (!Check((call), #call, error)) return false;  \
  } while (0)

struct Buffer {
  void* p = nullptr;
  size_t bytes = 0;
  bool Reserve(size_t want, std::string* error) {
    if (want <= bytes) return true;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    size_t grow = want + want / 4 + 256;
    if (!Check(cudaMalloc(&p, grow), "cudaMalloc", error)) return false;
    bytes = grow;
    return true;
  }
  void Release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

class DeviceContext {
 public:
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::mutex mu;                      // one MatchAll at a time per device context
  Buffer text;                        // staging for host-text calls
  Buffer cand_b, cand_e, hit_b, hit_e, out_pairs;
  Buffer counters;                    // [0]=cand count [1]=hit count [2]=last_any [3]=last_nonempty
  Buffer status;
  Buffer sorted_b, sorted_e, reach, take, wide, slot, cub_tmp;
  Buffer fscratch, fin_end;           // label scratch / final ends for re-entrant patterns
  Buffer flush;
  uint64_t cand_cap = 0, hit_cap = 0;
  PipelineStatus* h_status = nullptr; // pinned

  bool Init(int dev, std::string* error) {
    device = dev;
    RJ_TRY(cudaSetDevice(dev));
    cudaDeviceProp prop;
    RJ_TRY(cudaGetDeviceProperties(&prop, dev));
    sm_count = prop.multiProcessorCount;
    RJ_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (auto& e : ev) RJ_TRY(cudaEventCreate(&e));
    RJ_TRY(cudaMallocHost(&h_status, sizeof(PipelineStatus)));
    if (!counters.Reserve(64, error)) return false;
    if (!status.Reserve(sizeof(PipelineStatus), error)) return false;
    return true;
  }
  bool ReserveCandidates(uint64_t cap, std::string* error) {
    if (cap <= cand_cap) return true;
    if (!cand_b.Reserve(cap * 8, error) || !cand_e.Reserve(cap * 8, error) ||
        !out_pairs.Reserve(cap * 16, error)) return false;
    cand_cap = cap;
    return true;
  }
  bool ReserveHits(uint64_t cap, std::string* error) {
    if (cap <= hit_cap) return true;
    if (!hit_b.Reserve(cap * 8, error) || !hit_e.Reserve(cap * 8, error)) return false;
    hit_cap = cap;
    return true;
  }
};

class DeviceProgram {
 public:
  NfaTables nfa{};
  DfaTables dfa{};
  const uint8_t* needle = nullptr;
  uint32_t needle_len = 0, p4 = 0, pmask = 0;
  size_t dfa_smem = 0;
  std::vector<void*> allocs;
  ~DeviceProgram() { for (void* p : allocs) cudaFree(p); }

  template <class T>
  bool Upload(const std::vector<T>& host, const T** dev, std::string* error) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(host.size() * sizeof(T), 16);
    RJ_TRY(cudaMalloc(&p, bytes));
    allocs.push_back(p);
    if (!host.empty()) RJ_TRY(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<const T*>(p);
    return true;
  }

  bool Build(const CompiledAutomaton& ca, std::string* error) {
    FlatTables ft;
    FlattenTables(ca, &ft);
    nfa.n_pos = ft.n_pos;
    nfa.words = ft.words;
    nfa.has_anchor = ca.nfa.has_anchor ? 1 : 0;
    for (int c = 0; c < 4; ++c) nfa.accept_empty[c] = ft.accept_empty[c];
    if (!Upload(ft.byte_mask, &nfa.byte_mask, error) || !Upload(ft.first, &nfa.first, error) ||
        !Upload(ft.follow, &nfa.follow, error) || !Upload(ft.accept, &nfa.accept, error) ||
        !Upload(ft.chain, &nfa.chain, error) || !Upload(ft.start_ok, &nfa.start_ok, error)) return false;

    if (ca.strategy == ScanStrategy::Literal || ca.strategy == ScanStrategy::LiteralWindow) {
      if (!Upload(ca.literal, &needle, error)) return false;
      needle_len = (uint32_t)ca.literal.size();
      for (uint32_t i = 0; i < 4 && i < needle_len; ++i) {
        p4 |= (uint32_t)ca.literal[i] << (8 * i);
        pmask |= 0xFFu << (8 * i);
      }
    }
    if (ca.strategy == ScanStrategy::DfaFixed) {
      const ScanDfa& d = ca.dfa;
      if (!Upload(ft.dfa_next, &dfa.next, error) || !Upload(ft.dfa_class, &dfa.byte_class, error)) return false;
      dfa.n_states = d.n_states;
      dfa.n_classes = d.n_classes;
      dfa.first_accept_scaled = d.first_accept * d.n_classes;
      dfa.match_len = (uint32_t)d.match_len;
      dfa_smem = (((size_t)d.n_states * d.n_classes * 2 + 15) & ~(size_t)15) + 256;
    }
    return true;
  }
};

// ---------------------------------------------------------------------------
namespace {
std::mutex g_ctx_mu;
DeviceContext* g_ctx[16] = {nullptr};
int g_device_count = -2;
}  // namespace

int DeviceCount() {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  if (g_device_count == -2) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
    g_device_count = std::min(n, 16);
  }
  return g_device_count;
}

bool CudaOk(std::string* error) {
  if (DeviceCount() > 0) return true;
  if (error) *error = "rejit_b200: no CUDA device available (the sm_100a engine has no CPU fallback)";
  return false;
}

DeviceContext* ContextFor(int device, std::string* error) {
  if (!CudaOk(error)) return nullptr;
  if (device < 0 || device >= DeviceCount()) {
    if (error) *error = "rejit_b200: bad device index";
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  if (!g_ctx[device]) {
    DeviceContext* c = new DeviceContext();
    if (!c->Init(device, error)) { delete c; return nullptr; }
    g_ctx[device] = c;
  }
  return g_ctx[device];
}

void* DeviceAlloc(int device, size_t bytes, std::string* error) {
  if (!CudaOk(error)) return nullptr;
  void* p = nullptr;
  if (!Check(cudaSetDevice(device), "cudaSetDevice", error)) return nullptr;
  // 64 bytes of slack so that 16-byte vector loads near the end never leave
  // the allocation
  if (!Check(cudaMalloc(&p, bytes + 64), "cudaMalloc", error)) return nullptr;
  return p;
}
void DeviceFree(int device, void* p) { if (p) { cudaSetDevice(device); cudaFree(p); } }
void* PinnedAlloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
void PinnedFree(void* p) { if (p) cudaFreeHost(p); }

bool CopyToDevice(int device, void* dst, const void* src, size_t bytes, std::string* error) {
  RJ_TRY(cudaSetDevice(device));
  RJ_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return true;
}
bool CopyFromDevice(int device, void* dst, const void* src, size_t bytes, std::string* error) {
  RJ_TRY(cudaSetDevice(device));
  RJ_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return true;
}

void FlushL2(int device) {
  std::string err;
  DeviceContext* c = ContextFor(device, &err);
  if (!c) return;
  cudaSetDevice(device);
  const size_t bytes = 256u << 20;    // > 126 MB L2
  if (!c->flush.Reserve(bytes, &err)) return;
  k_fill_u32<<<c->sm_count * 4, 256, 0, c->stream>>>(c->flush.as<uint32_t>(), bytes / 4, 0x5a5a5a5au);
  cudaStreamSynchronize(c->stream);
}

// ---------------------------------------------------------------------------
Program* Program::Create(const LoweredRegexp& lr, std::string* error) {
  Program* p = new Program();
  if (!BuildAutomaton(lr, &p->automaton_, error)) { delete p; return nullptr; }
  return p;
}
Program::~Program() { for (auto* d : per_device_) delete d; }

DeviceProgram* Program::OnDevice(int device, std::string* error) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!per_device_[device]) {
    if (!Check(cudaSetDevice(device), "cudaSetDevice", error)) return nullptr;
    DeviceProgram* d = new DeviceProgram();
    if (!d->Build(automaton_, error)) { delete d; return nullptr; }
    per_device_[device] = d;
  }
  return per_device_[device];
}

// ===========================================================================
// the device pipeline
// ===========================================================================
namespace {

struct Slab {                 // how a launch maps local offsets to the whole text
  ScanRange own;
  uint64_t base_offset;       // added to every reported offset
};

constexpr int kFaithfulWalkers = 2048;

uint64_t FaithfulStride(const NfaTables& nfa) {
  uint64_t P = nfa.n_pos > 0 ? nfa.n_pos : 1;
  return ((2 * P * 8 + 3 * (uint64_t)nfa.words * 4) + 15) & ~15ull;
}

bool RunLargeResolve(DeviceContext* c, uint64_t m, const Carry& carry_in, uint64_t base_offset,
                     uint64_t* d_out, uint64_t out_cap, FaithfulArgs fa, RunStats* stats, std::string* error) {
  cudaStream_t s = c->stream;
  if (!c->sorted_b.Reserve(m * 8, error) || !c->sorted_e.Reserve(m * 8, error) ||
      !c->reach.Reserve(m * 8, error) || !c->take.Reserve(m * 4, error) ||
      !c->wide.Reserve(m * 8, error) || !c->slot.Reserve(m * 8, error)) return false;
  size_t tmp_sort = 0, tmp_scan = 0, tmp_sum = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, c->cand_b.as<uint64_t>(), c->sorted_b.as<uint64_t>(),
                                  c->cand_e.as<uint64_t>(), c->sorted_e.as<uint64_t>(), (int64_t)m, 0, 64, s);
  cub::DeviceScan::ExclusiveScan(nullptr, tmp_scan, c->sorted_e.as<uint64_t>(), c->reach.as<uint64_t>(),
                                 MaxOp(), (uint64_t)carry_in.cur, (int64_t)m, s);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_sum, c->wide.as<uint64_t>(), c->slot.as<uint64_t>(), (int64_t)m, s);
  size_t tmp = std::max(tmp_sort, std::max(tmp_scan, tmp_sum));
  if (!c->cub_tmp.Reserve(tmp, error)) return false;
  RJ_TRY(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, c->cand_b.as<uint64_t>(), c->sorted_b.as<uint64_t>(),
                                         c->cand_e.as<uint64_t>(), c->sorted_e.as<uint64_t>(), (int64_t)m, 0, 64, s));
  RJ_TRY(cub::DeviceScan::ExclusiveScan(c->cub_tmp.p, tmp, c->sorted_e.as<uint64_t>(), c->reach.as<uint64_t>(),
                                        MaxOp(), (uint64_t)carry_in.cur, (int64_t)m, s));
  int blocks = (int)std::min<uint64_t>((m + 255) / 256, (uint64_t)c->sm_count * 8);
  const uint64_t* final_e = c->sorted_e.as<uint64_t>();
  if (fa.enabled) {
    if (!c->fin_end.Reserve(m * 8, error)) return false;
    if (!c->fscratch.Reserve(fa.scratch_stride * kFaithfulWalkers, error)) return false;
    fa.scratch = c->fscratch.as<uint8_t>();
    fa.take = c->take.as<uint32_t>();
    fa.fin_end = c->fin_end.as<uint64_t>();
    final_e = fa.fin_end;
    k_segment_faithful<<<kFaithfulWalkers / 64, 64, 0, s>>>(c->sorted_b.as<uint64_t>(), c->sorted_e.as<uint64_t>(),
                                                            c->reach.as<uint64_t>(), m, fa);
  } else {
    k_segment_chain<<<blocks, 256, 0, s>>>(c->sorted_b.as<uint64_t>(), c->sorted_e.as<uint64_t>(),
                                           c->reach.as<uint64_t>(), m, carry_in, c->take.as<uint32_t>());
  }
  k_widen_flags<<<blocks, 256, 0, s>>>(c->take.as<uint32_t>(), c->wide.as<uint64_t>(), m);
  RJ_TRY(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->wide.as<uint64_t>(), c->slot.as<uint64_t>(), (int64_t)m, s));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  RJ_TRY(cudaMemsetAsync(ctr + 2, 0, 16, s));
  k_scatter_matches<<<blocks, 256, 0, s>>>(c->sorted_b.as<uint64_t>(), final_e,
                                           c->take.as<uint32_t>(), c->slot.as<uint64_t>(), m, base_offset,
                                           d_out, out_cap, ctr + 2, ctr + 3);
  k_finish_large<<<1, 32, 0, s>>>(c->sorted_b.as<uint64_t>(), final_e, c->take.as<uint32_t>(),
                                  c->slot.as<uint64_t>(), m, carry_in, ctr + 2, ctr + 3,
                                  c->status.as<PipelineStatus>());
  if (stats) { stats->launches += 9; stats->large_path = 1; }
  RJ_TRY(cudaGetLastError());
  return true;
}

// Runs scan (+verify) + resolve for one slab whose text is at d_text[0..n).
// Matches are written to d_out as global offsets.
bool RunPipeline(DeviceContext* c, Program* prog, DeviceProgram* dp, const uint8_t* d_text, uint64_t n,
                 const Slab& slab, const Carry& carry_in, uint64_t* d_out, uint64_t out_cap,
                 PipelineStatus* result, RunStats* stats, std::string* error) {
  const CompiledAutomaton& ca = prog->automaton();
  cudaStream_t s = c->stream;
  RJ_TRY(cudaSetDevice(c->device));
  if ((reinterpret_cast<uintptr_t>(d_text) & 15) != 0) {
    if (error) *error = "rejit_b200: device text must be 16-byte aligned";
    return false;
  }
  if (c->cand_cap == 0 && !c->ReserveCandidates(1u << 16, error)) return false;
  if (ca.strategy == ScanStrategy::LiteralWindow && c->hit_cap == 0 && !c->ReserveHits(1u << 16, error)) return false;

  for (int attempt = 0; attempt < 40; ++attempt) {
    unsigned long long* ctr = c->counters.as<unsigned long long>();
    PipelineStatus* d_status = c->status.as<PipelineStatus>();
    RJ_TRY(cudaMemsetAsync(ctr, 0, 32, s));
    RJ_TRY(cudaMemsetAsync(d_status, 0, sizeof(PipelineStatus), s));
    CandBuf cand{c->cand_b.as<uint64_t>(), c->cand_e.as<uint64_t>(), ctr, c->cand_cap};
    CandBuf hits{c->hit_b.as<uint64_t>(), c->hit_e.as<uint64_t>(), ctr + 1, c->hit_cap};
    uint64_t* outp = d_out ? d_out : c->out_pairs.as<uint64_t>();
    uint64_t ocap = d_out ? out_cap : c->cand_cap;

    if (stats) cudaEventRecord(c->ev[0], s);
    const int grid_full = c->sm_count * 8;
    switch (ca.strategy) {
      case ScanStrategy::Literal: {
        uint64_t pieces = (n + 511) / 512;
        int blocks = (int)std::min<uint64_t>((pieces + 8 * 4 - 1) / (8 * 4) + 1, (uint64_t)grid_full);
        k_lit_scan<4><<<blocks, 256, 0, s>>>(d_text, n, dp->needle, dp->needle_len, dp->p4, dp->pmask,
                                             slab.own, cand);
        if (stats) stats->launches += 1;
        break;
      }
      case ScanStrategy::LiteralWindow: {
        // the needle may sit up to window_hi bytes after an owned start
        ScanRange hit_range{slab.own.own_begin, slab.own.own_end + ca.window_hi + 1};
        uint64_t pieces = (n + 511) / 512;
        int blocks = (int)std::min<uint64_t>((pieces + 8 * 4 - 1) / (8 * 4) + 1, (uint64_t)grid_full);
        k_lit_scan<4><<<blocks, 256, 0, s>>>(d_text, n, dp->needle, dp->needle_len, dp->p4, dp->pmask,
                                             hit_range, hits);
        if (stats) cudaEventRecord(c->ev[1], s);
        k_window_verify<<<c->sm_count * 2, 256, 0, s>>>(d_text, n, dp->nfa, hits, ca.window_lo,
                                                        ca.window_hi, slab.own, cand);
        if (stats) stats->launches += 2;
        break;
      }
      case ScanStrategy::DfaFixed: {
        const uint32_t stream_bytes = 512;
        uint64_t streams = (n + stream_bytes - 1) / stream_bytes;
        int blocks = (int)std::min<uint64_t>((streams + 255) / 256, (uint64_t)grid_full);
        if (blocks < 1) blocks = 1;
        if (dp->dfa_smem > 48 * 1024) {
          RJ_TRY(cudaFuncSetAttribute(k_dfa_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp->dfa_smem));
        }
        k_dfa_scan<<<blocks, 256, dp->dfa_smem, s>>>(d_text, n, dp->dfa, stream_bytes, slab.own, cand);
        if (stats) stats->launches += 1;
        break;
      }
      case ScanStrategy::Generic: {
        uint64_t offsets = n + 1;
        int blocks = (int)std::min<uint64_t>((offsets + 255) / 256, (uint64_t)grid_full);
        k_generic_scan<<<blocks, 256, 0, s>>>(d_text, n, dp->nfa, slab.own, cand);
        if (stats) stats->launches += 1;
        break;
      }
    }
    if (stats && ca.strategy != ScanStrategy::LiteralWindow) cudaEventRecord(c->ev[1], s);
    FaithfulArgs fa{};
    fa.enabled = ca.reentrant ? 1 : 0;
    if (fa.enabled) {
      fa.nfa = dp->nfa;
      fa.text = d_text;
      fa.n = n;
      fa.scratch_stride = FaithfulStride(dp->nfa);
      if (!c->fscratch.Reserve(fa.scratch_stride * kFaithfulWalkers, error) ||
          !c->take.Reserve(kSmallResolveMax * 4, error) || !c->fin_end.Reserve(kSmallResolveMax * 8, error)) return false;
      fa.scratch = c->fscratch.as<uint8_t>();
      fa.take = c->take.as<uint32_t>();
      fa.fin_end = c->fin_end.as<uint64_t>();
    }
    RJ_TRY(cudaFuncSetAttribute(k_resolve_small, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallResolveMax * 16));
    k_resolve_small<<<1, 1024, kSmallResolveMax * 16, s>>>(cand, carry_in, slab.base_offset, outp, ocap, fa, d_status);
    if (stats) stats->launches += 1;
    RJ_TRY(cudaGetLastError());
    RJ_TRY(cudaMemcpyAsync(c->h_status, d_status, sizeof(PipelineStatus), cudaMemcpyDeviceToHost, s));
    // the hit counter is needed to detect a hit-buffer overflow
    unsigned long long h_hits = 0;
    if (ca.strategy == ScanStrategy::LiteralWindow)
      RJ_TRY(cudaMemcpyAsync(&h_hits, ctr + 1, 8, cudaMemcpyDeviceToHost, s));
    RJ_TRY(cudaStreamSynchronize(s));
    PipelineStatus st = *c->h_status;

    bool rerun = false;
    if (ca.strategy == ScanStrategy::LiteralWindow && h_hits > c->hit_cap) {
      if (!c->ReserveHits(h_hits + h_hits / 8 + 1024, error)) return false;
      rerun = true;
    }
    if (st.overflow || st.n_candidates > c->cand_cap) {
      uint64_t want = std::max<uint64_t>(st.n_candidates + st.n_candidates / 8 + 1024, c->cand_cap * 2);
      if (!c->ReserveCandidates(want, error)) return false;
      rerun = true;
    }
    if (rerun) {
      if (stats) stats->reruns += 1;
      continue;
    }
    if (st.need_large) {
      outp = d_out ? d_out : c->out_pairs.as<uint64_t>();
      if (!RunLargeResolve(c, st.n_candidates, carry_in, slab.base_offset, outp, ocap, fa, stats, error)) return false;
      RJ_TRY(cudaMemcpyAsync(c->h_status, d_status, sizeof(PipelineStatus), cudaMemcpyDeviceToHost, s));
      RJ_TRY(cudaStreamSynchronize(s));
      unsigned long long keep = st.n_candidates;
      st = *c->h_status;
      st.n_candidates = keep;
    }
    if (stats) {
      cudaEventRecord(c->ev[2], s);
      cudaEventSynchronize(c->ev[2]);
      cudaEventElapsedTime(&stats->scan_ms, c->ev[0], c->ev[1]);
      cudaEventElapsedTime(&stats->total_ms, c->ev[0], c->ev[2]);
      stats->candidates = st.n_candidates;
      stats->matches = st.n_matches;
      stats->strategy = (int)ca.strategy;
    }
    *result = st;
    return true;
  }
  if (error) *error = "rejit_b200: candidate buffers kept overflowing";
  return false;
}

}  // namespace

int64_t MatchAllDevice(int device, Program* prog, const uint8_t* d_text, uint64_t n, uint64_t* d_out,
                       uint64_t out_cap, const Carry& in, Carry* out, RunStats* stats, std::string* error,
                       const SlabView* own) {
  DeviceContext* c = ContextFor(device, error);
  if (!c) return -1;
  DeviceProgram* dp = prog->OnDevice(device, error);
  if (!dp) return -1;
  std::lock_guard<std::mutex> lk(c->mu);
  Slab slab{{0, n + 1}, 0};
  if (own) {
    slab.own.own_begin = std::min<uint64_t>(own->own_begin, n + 1);
    slab.own.own_end = std::min<uint64_t>(own->own_end, n + 1);
    slab.base_offset = own->base_offset;
  }
  PipelineStatus st;
  if (!RunPipeline(c, prog, dp, d_text, n, slab, in, d_out, out_cap, &st, stats, error)) return -1;
  if (out) { out->cur = st.carry_cur; out->tail = st.carry_tail; }
  return (int64_t)st.n_matches;
}

namespace {
// Device work for one slab of a host text: copy [lo, hi) of the text in, scan
// the owned starts, copy the matches out.
bool MatchSlabFromHost(DeviceContext* c, Program* prog, DeviceProgram* dp, const uint8_t* text, uint64_t n,
                       uint64_t own_lo, uint64_t own_hi, bool last, const Carry& in, Carry* out,
                       std::vector<uint64_t>* pairs, RunStats* stats, std::string* error) {
  const CompiledAutomaton& ca = prog->automaton();
  std::lock_guard<std::mutex> lk(c->mu);
  RJ_TRY(cudaSetDevice(c->device));
  // the slab is copied together with a 16-byte aligned left halo (>= 1 byte
  // when not at the text start, for the ^ context) and a right halo long
  // enough for any match begun inside the slab (everything up to the end of
  // the text when the pattern has no length bound)
  uint64_t lo = own_lo >= 16 ? ((own_lo - 1) & ~15ull) : 0;
  uint64_t hi = n;
  if (!last && ca.nfa.max_len != kInfLen) hi = std::min<uint64_t>(n, own_hi + ca.nfa.max_len + 2);
  uint64_t len = hi - lo;
  if (!c->text.Reserve(len + 64, error)) return false;
  if (len) RJ_TRY(cudaMemcpyAsync(c->text.p, text + lo, len, cudaMemcpyHostToDevice, c->stream));
  Slab slab;
  slab.own.own_begin = own_lo - lo;
  slab.own.own_end = (last ? n + 1 : own_hi) - lo;
  slab.base_offset = lo;
  // contexts at the slab's copy boundaries are only exact at the true text
  // boundaries; interior copies are shielded by the halos above, except that
  // "offset == end of copy" must not look like the end of the text:
  // RunPipeline treats n_local as the text length, so a copy that stops before
  // the real end is only legal when no owned run can reach it (bounded max_len).
  Carry local_in{in.cur > lo ? in.cur - lo : 0, (in.tail != kNoMatch && in.tail >= lo) ? in.tail - lo : kNoMatch};
  PipelineStatus st;
  if (!RunPipeline(c, prog, dp, c->text.as<uint8_t>(), len, slab, local_in, nullptr, 0, &st, stats, error)) return false;
  uint64_t cnt = st.n_matches;
  pairs->resize(cnt * 2);
  if (cnt) {
    RJ_TRY(cudaMemcpyAsync(pairs->data(), c->out_pairs.p, cnt * 16, cudaMemcpyDeviceToHost, c->stream));
    RJ_TRY(cudaStreamSynchronize(c->stream));
  }
  if (out) {
    out->cur = st.carry_cur + lo;
    out->tail = (st.carry_tail == kNoMatch) ? kNoMatch : st.carry_tail + lo;
  }
  return true;
}
}  // namespace

int64_t MatchAllHost(int device, Program* prog, const uint8_t* text, uint64_t n, uint64_t** pairs,
                     RunStats* stats, std::string* error) {
  DeviceContext* c = ContextFor(device, error);
  if (!c) return -1;
  DeviceProgram* dp = prog->OnDevice(device, error);
  if (!dp) return -1;
  std::vector<uint64_t> v;
  Carry in, out;
  if (!MatchSlabFromHost(c, prog, dp, text, n, 0, n, true, in, &out, &v, stats, error)) return -1;
  uint64_t cnt = v.size() / 2;
  *pairs = static_cast<uint64_t*>(malloc(std::max<size_t>(v.size() * 8, 8)));
  if (cnt) memcpy(*pairs, v.data(), v.size() * 8);
  return (int64_t)cnt;
}

int64_t MatchAllHostMultiGpu(Program* prog, const uint8_t* text, uint64_t n, int n_gpus, uint64_t** pairs,
                             RunStats* stats, std::string* error) {
  if (!CudaOk(error)) return -1;
  int g = std::max(1, std::min(n_gpus, DeviceCount()));
  if (n < (uint64_t)g * 4096) g = 1;
  // the label replay of re-entrant patterns cannot be cut at a slab edge
  if (prog->automaton().reentrant) g = 1;
  std::vector<std::vector<uint64_t>> part(g);
  std::vector<Carry> carry_out(g);
  std::vector<std::string> errs(g);
  std::vector<char> ok(g, 1);
  std::vector<RunStats> st(g);
  auto bounds = [&](int i) { return (n / g) * (uint64_t)i; };
  // round 1: every slab resolved as if no match from the left neighbour reached into it
  {
    std::vector<std::thread> pool;
    for (int i = 0; i < g; ++i)
      pool.emplace_back([&, i]() {
        DeviceContext* c = ContextFor(i, &errs[i]);
        DeviceProgram* dp = c ? prog->OnDevice(i, &errs[i]) : nullptr;
        if (!c || !dp) { ok[i] = 0; return; }
        uint64_t lo = bounds(i), hi = (i + 1 == g) ? n : bounds(i + 1);
        Carry in{lo, kNoMatch};
        ok[i] = MatchSlabFromHost(c, prog, dp, text, n, lo, hi, i + 1 == g, in, &carry_out[i], &part[i], &st[i], &errs[i]);
      });
    for (auto& t : pool) t.join();
  }
  for (int i = 0; i < g; ++i) if (!ok[i]) { if (error) *error = errs[i]; return -1; }
  // stitch: slab i must be redone when the chain arriving from the left differs
  // from the assumption (cur == slab start, no abutting non-empty match)
  Carry running = carry_out[0];
  for (int i = 1; i < g; ++i) {
    uint64_t lo = bounds(i), hi = (i + 1 == g) ? n : bounds(i + 1);
    bool differs = running.cur > lo || running.tail == lo;
    if (differs) {
      DeviceContext* c = ContextFor(i, error);
      DeviceProgram* dp = c ? prog->OnDevice(i, error) : nullptr;
      if (!c || !dp) return -1;
      Carry in = running;
      if (in.cur < lo) in.cur = lo;
      if (!MatchSlabFromHost(c, prog, dp, text, n, lo, hi, i + 1 == g, in, &carry_out[i], &part[i], &st[i], error)) return -1;
      if (stats) stats->reruns += 1;
    }
    running = carry_out[i];
  }
  size_t total = 0;
  for (auto& p : part) total += p.size();
  *pairs = static_cast<uint64_t*>(malloc(std::max<size_t>(total * 8, 8)));
  size_t at = 0;
  for (auto& p : part) { if (!p.empty()) memcpy(*pairs + at, p.data(), p.size() * 8); at += p.size(); }
  if (stats) {
    for (int i = 0; i < g; ++i) {
      stats->scan_ms = std::max(stats->scan_ms, st[i].scan_ms);
      stats->total_ms = std::max(stats->total_ms, st[i].total_ms);
      stats->launches += st[i].launches;
      stats->candidates += st[i].candidates;
    }
    stats->matches = total / 2;
    stats->strategy = (int)prog->automaton().strategy;
  }
  return (int64_t)(total / 2);
}

int MatchFullHost(int device, Program* prog, const uint8_t* text, uint64_t n, std::string* error) {
  DeviceContext* c = ContextFor(device, error);
  if (!c) return -1;
  DeviceProgram* dp = prog->OnDevice(device, error);
  if (!dp) return -1;
  std::lock_guard<std::mutex> lk(c->mu);
  if (!Check(cudaSetDevice(c->device), "cudaSetDevice", error)) return -1;
  if (!c->text.Reserve(n + 64, error)) return -1;
  if (n && !Check(cudaMemcpyAsync(c->text.p, text, n, cudaMemcpyHostToDevice, c->stream), "H2D", error)) return -1;
  PipelineStatus* d_status = c->status.as<PipelineStatus>();
  cudaMemsetAsync(d_status, 0, sizeof(PipelineStatus), c->stream);
  k_match_full<<<1, 32, 0, c->stream>>>(c->text.as<uint8_t>(), n, dp->nfa, d_status);
  if (!Check(cudaMemcpyAsync(c->h_status, d_status, sizeof(PipelineStatus), cudaMemcpyDeviceToHost, c->stream), "D2H", error)) return -1;
  if (!Check(cudaStreamSynchronize(c->stream), "sync", error)) return -1;
  return c->h_status->full_result ? 1 : 0;
}

}  // namespace rejit_b200
