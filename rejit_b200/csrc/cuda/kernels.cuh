// rejit_b200 — the sm_100a kernels.
//
// What replaces what:
//   loop A, the fast-forward scan      /root/reference/src/x64/codegen-x64.cc:1102-1403
//        -> k_lit_scan   literal / required-literal scan: 16-byte vector loads,
//                        funnel-shift compares, shuffles for straddling bytes
//        -> k_dfa_tma    exact table-driven scan for fixed-length alternations:
//                        text rows staged into shared memory with TMA bulk copies
//                        (cp.async.bulk + mbarrier), per-lane replicated
//                        transition rows (bank-conflict free), one DFA chain per lane
//        -> k_dfa_scan   same automaton, plain vector loads (fallback for big tables
//                        / dense matches)
//   loop B, the NFA active-state advance   codegen-x64.cc:535-677
//        -> NfaRun (device_program.h) driven by k_window_verify / k_generic_scan
//   match selection   codegen-x64.cc:401-522 + /root/reference/src/codegen.cc:36-86
//        -> k_resolve_ordered (one CTA, parallel: gather, max-scan, restart
//           points, segment chains, compaction) and the multi-CTA large path
//
// Candidate order.  Every scanning warp owns a contiguous SUB-REGION of start
// offsets and appends its candidates to that sub-region's slot range in
// increasing order of begin, so the concatenation over sub-regions is already
// sorted: no sort is needed before the chain is resolved.  (The unordered
// append + sort path survives only as the fallback of k_dfa_scan.)
#ifndef REJIT_B200_CUDA_KERNELS_CUH_
#define REJIT_B200_CUDA_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "device_program.h"
#include "engine.h"
#include "stitch.cuh"

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

constexpr uint32_t kLaneListOverflow = 0xFFFFFFFFu;

struct SubStore {                      // ordered store: slot range per sub-region
  uint64_t* begin;                     // [nsub * cap]
  uint64_t* end;                       // [nsub * cap]
  uint32_t* count;                     // [nsub]; may exceed cap (overflow), or kLaneListOverflow
  uint32_t cap;
  uint32_t pad;
  uint64_t nsub;
};

struct DenseList {                     // gathered (sorted) candidates
  uint64_t* begin;
  uint64_t* end;
  unsigned long long* count;
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
  unsigned int overflow;               // a slot range / buffer was too small
  unsigned int need_cap;               // largest per-sub-region count seen
  unsigned int need_large;             // too many candidates for the one-CTA resolve
  unsigned int dense;                  // lane hit list overflowed: use the fallback kernel
  unsigned int full_result;            // MatchFull answer
  unsigned int seq;                    // call id, written last (host polls the mapped copy)
};

struct CarrySet { Carry c[32]; };

// What the in-kernel finish reports per pattern, in mapped host memory.  Two
// 16-byte halves, each written with ONE store (one PCIe write, so each half is
// seen whole) and each carrying the call's sequence number: the host waits until
// both match — no system-scope fence on the device.
struct __align__(16) FinRecord {
  unsigned long long n_matches;
  unsigned int flags;                  // kFin* bits
  unsigned int seq0;
  unsigned long long last_end;         // end of the last match (0 if none)
  unsigned int need_cap;
  unsigned int seq1;
  unsigned long long last_nonempty;    // end of the last non-empty match (0 if none)
  unsigned int pad;
  unsigned int seq2;
};

// In-kernel finish of the fixed-length scans (k_dfa_tma, k_set_tma): the scan
// grid is one CTA per SM, launched cooperatively, so that after a grid barrier
// the same CTAs concatenate the slot ranges straight into the output.
struct FinishArgs {
  int enabled;
  uint32_t nseg, seg_subs;             // the sub-regions are cut into nseg segments of seg_subs
  uint32_t* segcount;                  // [K][nseg] candidates per (pattern, segment); zeroed by the host
  unsigned int* sync;                  // [0] arrive, [1] done, [2] flags, [3] need_cap; zeroed by the host
  unsigned long long* last_end;        // [K] end of the last candidate; zeroed by the host
  unsigned long long* totals;          // [K] candidates per pattern (written by the warp of the last segment)
  unsigned long long* last_ne;         // [K] end of the last non-empty candidate; zeroed by the host
  int local_pred;                      // geometry: a candidate can only touch a predecessor that lives in
                                       // one of the 32 preceding sub-regions (bounded match length)
  uint64_t* out_pairs;                 // pattern j's pairs start at out_pairs + j * 2 * out_stride
  uint64_t out_stride, out_cap;
  uint64_t base_offset;
  FinRecord* host_records;             // [K] mapped host memory (see FinRecord)
  unsigned int seq;
};
constexpr unsigned int kFinOverlap = 1u, kFinDense = 2u, kFinOverflow = 4u;
constexpr uint32_t kFinScratchWords = 4096;   // dynamic shared memory of the non-TMA scans (16 KB)

struct DfaTables {
  const uint16_t* next;                // [n_states * n_classes], entries pre-multiplied by n_classes
  const uint8_t* byte_class;           // [256]
  const uint32_t* pair;                // [n_states * n_classes^2]: state after two bytes | bit31 mid-accept
  int n_states, n_classes;
  int first_accept_scaled;             // first accepting state * n_classes
  int first_accept;                    // first accepting state
  uint32_t match_len;
};

struct FaithfulArgs {                  // only used for re-entrant patterns
  int enabled;
  NfaTables nfa;
  const uint8_t* text;
  uint64_t n;
  uint8_t* scratch;                    // per-walker label scratch
  uint64_t scratch_stride;             // bytes per walker
};

struct ResolveScratch {                // global scratch of the resolve kernels
  uint64_t* reach;                     // [cap]
  uint32_t* take;                      // [cap]
  uint64_t* fin_end;                   // [cap]
  uint64_t* slot;                      // [cap]
};

constexpr unsigned kFullMask = 0xFFFFFFFFu;
constexpr int kSmallResolveMax = 4096;       // sort-based fallback resolve
constexpr int kOrderedResolveMax = 16384;    // one-CTA ordered resolve
constexpr uint32_t kLitSubBytes = 16384;     // sub-region of the literal scan (32 pieces)
constexpr uint32_t kGenSubBytes = 16384;     // sub-region of the generic scan
constexpr uint32_t kWinSubHits = 8;          // needle hits per sub-region of the window verify
constexpr uint32_t kDfaStreamBytes = 272;    // bytes per lane sub-stream (k_dfa_tma): 17 * 16, so that the
                                             // lanes' 16-byte shared loads from a DENSE tile are conflict free
constexpr uint32_t kDfaSubBytes = 32 * kDfaStreamBytes;   // sub-region of one warp (8704 bytes)
constexpr uint32_t kDfaTileBytes = kDfaSubBytes + 16 + 112;   // + the 16 bytes before it, padded to 128

// ===========================================================================
// small device helpers
// ===========================================================================
__device__ __forceinline__ void AppendAggregated(const CandBuf& buf, uint64_t b, uint64_t e) {
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

// Ordered append by a full warp: lanes with `has` write in lane order.
__device__ __forceinline__ void EmitOrdered(const SubStore& st, uint64_t sub, uint32_t& k, bool has,
                                            uint64_t b, uint64_t e) {
  unsigned m = __ballot_sync(kFullMask, has);
  if (!m) return;
  uint32_t idx = k + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
  if (has && idx < st.cap) {
    st.begin[sub * st.cap + idx] = b;
    st.end[sub * st.cap + idx] = e;
  }
  k += __popc(m);
}

__device__ __forceinline__ uint32_t WarpInclusiveScan(uint32_t v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(kFullMask, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

__device__ __forceinline__ uint4 LoadText16(const uint8_t* __restrict__ text, uint64_t n, uint64_t at) {
  if (at + 16 <= n) return __ldg(reinterpret_cast<const uint4*>(text + at));
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

// ---- mbarrier / TMA bulk copy (PTX) ----------------------------------------
__device__ __forceinline__ uint32_t SmemAddr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void MbarInit(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool MbarTryWait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(SmemAddr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t parity) {
  while (!MbarTryWait(bar, parity)) {
  }
}
// global -> shared bulk copy (TMA, 1-D): size and both addresses multiples of 16
__device__ __forceinline__ void TmaLoad1D(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          SmemAddr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(SmemAddr(bar))
      : "memory");
}

// ===========================================================================
// K1: literal scan (ordered).  One warp owns a 16 KB sub-region = 32 pieces of
// 512 bytes; in a piece lane l holds bytes [16l, 16l+16) in four registers; the
// first min(m,4) needle bytes are compared at all 16 alignments with funnel
// shifts, the word straddling into the next lane comes from a shuffle.
// Survivors (rare) compare the rest of the needle from global memory.
// Algorithmic traffic: N bytes read + 16 bytes written per occurrence.
// ===========================================================================
// Hits of k_lit_scan are recorded in the hot loop as one shared-memory word per 16-byte lane group with a
// survivor ({group index in the sub-region, 16 alignment bits}: a ballot and a store, no call -- a call from inside
// the loop waits for every prefetched piece to land, which made dense texts latency-bound) and validated here, once
// per sub-region: bounds, ownership, the rest of the needle; then appended in offset order.  Called by the whole warp.
constexpr uint32_t kLitEntCap = kLitSubBytes / 16;      // every group of the sub-region may hold a survivor

__device__ __noinline__ uint32_t LitFlush(const uint8_t* __restrict__ text, uint64_t n, const uint8_t* __restrict__ needle,
                                          uint32_t m, const ScanRange& range, const SubStore& out, uint64_t sub,
                                          uint64_t sub_lo, const uint32_t* my_ent, uint32_t n_ent, uint32_t k) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  for (uint32_t base = 0; base < n_ent; base += 32) {
    const uint32_t ent = base + lane < n_ent ? my_ent[base + lane] : 0u;
    uint32_t valid = ent & 0xFFFFu;
    const uint64_t my = sub_lo + (uint64_t)(ent >> 16) * 16;
    uint32_t hh = valid;
    while (hh) {
      int j = __ffs(hh) - 1;
      hh &= hh - 1;
      uint64_t pos = my + j;
      bool ok = pos >= range.own_begin && pos < range.own_end && pos + m <= n;
      for (uint32_t i = 4; i < m && ok; ++i) ok = (text[pos + i] == needle[i]);
      if (!ok) valid &= ~(1u << j);
    }
    __syncwarp();
    uint32_t c = __popc(valid);
    uint32_t incl = WarpInclusiveScan(c);
    uint32_t total = __shfl_sync(kFullMask, incl, 31);
    uint32_t idx = k + incl - c;
    while (valid) {
      int j = __ffs(valid) - 1;
      valid &= valid - 1;
      if (idx < out.cap) {
        out.begin[sub * out.cap + idx] = my + j;
        out.end[sub * out.cap + idx] = my + j + m;
      }
      ++idx;
    }
    k += total;
  }
  __syncwarp();
  return k;
}

// the 16 alignment bits of one lane group: bit j = the first min(m, 4) needle bytes match at byte j
__device__ __forceinline__ uint32_t LitMask(const uint4& v, uint32_t nx, uint32_t p4, uint32_t pmask) {
  const uint32_t w[5] = {v.x, v.y, v.z, v.w, nx};
  uint32_t valid = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    uint32_t x = __funnelshift_r(w[j >> 2], w[(j >> 2) + 1], 8 * (j & 3));
    if (((x ^ p4) & pmask) == 0) valid |= 1u << j;
  }
  return valid;
}

template <bool kFull4>
__device__ __forceinline__ bool LitAny(const uint4& v, uint32_t nx, uint32_t p4, uint32_t pmask) {
  const uint32_t w[5] = {v.x, v.y, v.z, v.w, nx};
  bool any = false;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    uint32_t x = (j & 3) ? __funnelshift_r(w[j >> 2], w[(j >> 2) + 1], 8 * (j & 3)) : w[j >> 2];
    if (kFull4) any |= (x == p4);
    else any |= (((x ^ p4) & pmask) == 0);
  }
  return any;
}

// (defined below, next to the DFA scans)
__device__ __forceinline__ void FinishNote(const FinishArgs& fin, int j, uint64_t sub, uint32_t total, bool over,
                                           uint32_t cap);
__device__ __forceinline__ void FinishOrdered(const SubStore& st, uint64_t nsub_pat, int K, const FinishArgs& fin,
                                              const Carry* carries, uint32_t* scratch, uint32_t scratch_words);

template <bool kFull4>
__global__ void __launch_bounds__(256, 4)
k_lit_scan(const uint8_t* __restrict__ text, uint64_t n, const uint8_t* __restrict__ needle,
           uint32_t m, uint32_t p4, uint32_t pmask, ScanRange range, SubStore out, FinishArgs fin, Carry carry0) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  constexpr uint32_t kPieces = kLitSubBytes / 512;
  __shared__ uint32_t s_lit_ent[8 * kLitEntCap];           // per warp: the lane groups with a survivor, in offset order
  uint32_t* my_ent = s_lit_ent + (threadIdx.x >> 5) * kLitEntCap;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (uint64_t sub = warp; sub < out.nsub; sub += nwarps) {
    const uint64_t sub_lo = sub * kLitSubBytes;
    uint32_t k = 0, n_ent = 0;
    const bool live = sub_lo < n && sub_lo + kLitSubBytes > range.own_begin && sub_lo < range.own_end;
    if (live && sub_lo + kLitSubBytes + 16 <= n) {
      // ---- interior sub-region: unguarded 16-byte loads, 4 pieces per step, the
      // next step's loads are issued before the current step is examined --------
      const uint4* base = reinterpret_cast<const uint4*>(text + sub_lo) + lane;
      uint4 nxt[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) nxt[u] = __ldg(base + u * 32);
#pragma unroll 1
      for (uint32_t pc = 0; pc < kPieces; pc += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = nxt[u];
        if (pc + 4 < kPieces) {
#pragma unroll
          for (int u = 0; u < 4; ++u) nxt[u] = __ldg(base + (pc + 4 + u) * 32);
        } else {
          // word that follows the sub-region (in bounds: sub_lo + 16 KB + 16 <= n)
          nxt[0].x = __ldg(reinterpret_cast<const uint32_t*>(text + sub_lo + kLitSubBytes));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          // lane l needs the first word of lane l+1; lane 31 that of lane 0 of the next piece
          uint32_t up = __shfl_down_sync(kFullMask, v[u].x, 1);
          uint32_t wrap = __shfl_sync(kFullMask, (u < 3) ? v[u + 1].x : nxt[0].x, 0);
          uint32_t nx = (lane == 31) ? wrap : up;
          const bool any = LitAny<kFull4>(v[u], nx, p4, pmask);
          const uint32_t bal = __ballot_sync(kFullMask, any);
          if (bal) {
            if (any) my_ent[n_ent + __popc(bal & lt_mask)] = (((pc + u) * 32u + lane) << 16) | LitMask(v[u], nx, p4, pmask);
            n_ent += __popc(bal);
          }
        }
      }
    } else if (live) {
      // ---- first / last sub-regions of the text: guarded loads ------------------
      for (uint32_t pc = 0; pc < kPieces; ++pc) {
        const uint64_t my = sub_lo + (uint64_t)pc * 512 + (uint64_t)lane * 16;
        if (sub_lo + (uint64_t)pc * 512 >= n) break;
        uint4 v = (my < n) ? LoadText16(text, n, my) : make_uint4(0, 0, 0, 0);
        uint32_t nx = __shfl_down_sync(kFullMask, v.x, 1);
        if (lane == 31) nx = (my + 16 < n) ? LoadText4(text, n, my + 16) : 0u;
        const bool any = LitAny<kFull4>(v, nx, p4, pmask);
        const uint32_t bal = __ballot_sync(kFullMask, any);
        if (bal) {
          if (any) my_ent[n_ent + __popc(bal & lt_mask)] = ((pc * 32u + lane) << 16) | LitMask(v, nx, p4, pmask);
          n_ent += __popc(bal);
        }
      }
    }
    if (n_ent) k = LitFlush(text, n, needle, m, range, out, sub, sub_lo, my_ent, n_ent, k);
    if (lane == 0) {
      out.count[sub] = k;
      FinishNote(fin, 0, sub, k, false, out.cap);
    }
  }
  // whole-literal patterns finish in-kernel like the DFA scans (cooperative launch)
  if (fin.enabled) {
    extern __shared__ __align__(16) uint32_t lit_scratch[];
    FinishOrdered(out, out.nsub, 1, fin, &carry0, lit_scratch, kFinScratchWords);
  }
}

// ===========================================================================
// K2: exact DFA scan for fixed-length, anchor-free patterns (regex-dna).
// A warp owns a sub-region of 32 x 272 bytes, handed out dynamically in order.
// The sub-region and the 16 bytes before it are staged into shared memory with
// ONE TMA bulk copy (cp.async.bulk, completion on the warp's mbarrier); lane l
// walks bytes [272 l, 272 l + 272) of it.  272 = 17 x 16, so the lanes' 16-byte
// shared loads from the dense tile are conflict free.
//
// Each lane runs TWO independent automaton chains (144 + 128 bytes of its
// sub-stream, each entered 16 bytes early: a fixed-length-L automaton entered
// >= L-1 bytes early is in the same state as one sequential pass), and every
// chain advances TWO bytes per table lookup: the pair table
//     entry(state, c1*C + c2) = row address of the state after both bytes,
//                               bit 31 = "the state in between accepts"
// is replicated per lane (entry e of lane l lives in bank l: conflict free) and
// holds shared-memory addresses, so one step is
//     p = class[b0] * C + class[b1];   row = *(row + p*128)
// Accepting rows are the highest addresses; a 16-byte group is replayed byte by
// byte (plain 1-byte table) only when its running maximum crosses that
// threshold.  Match ends are kept in registers (3 per chain) and placed with a
// warp scan at the end of the sub-region, so candidates come out sorted.
// Algorithmic traffic: N bytes read + 16 bytes per match.
// ===========================================================================
// ---------------------------------------------------------------------------
// Finish of a fixed-length scan (see FinishArgs).  Matches of one fixed-length
// pattern come out of the scan sorted and can only interact with their direct
// neighbour: when no candidate begins before its predecessor ends, the
// candidates ARE the matches (the common case) and this copy is the whole
// resolve; otherwise kFinOverlap is raised and the host runs the general resolve
// kernel on the same slot ranges.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void GridBarrier(unsigned int* ctr, unsigned int expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v < expected) {                       // back off: hundreds of CTAs poll one L2 line
        const long long t0 = clock64();         // (no __nanosleep: ptxas reserves static shared memory for it)
        while (clock64() - t0 < 256) {}
      }
    } while (v < expected);
    __threadfence();
  }
  __syncthreads();
}

// the scan's bookkeeping for one (pattern, sub-region) count
__device__ __forceinline__ void FinishNote(const FinishArgs& fin, int j, uint64_t sub, uint32_t total, bool over,
                                           uint32_t cap) {
  if (!fin.enabled) return;
  if (over) { atomicOr(&fin.sync[2], kFinDense); return; }
  if (total > cap) { atomicOr(&fin.sync[2], kFinOverflow); atomicMax(&fin.sync[3], total); total = cap; }
  if (total) atomicAdd(&fin.segcount[(uint64_t)j * fin.nseg + (uint32_t)(sub / fin.seg_subs)], total);
}

__device__ __forceinline__ uint64_t WarpMax64(uint64_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const uint64_t o = __shfl_xor_sync(kFullMask, v, d);
    v = o > v ? o : v;
  }
  return v;
}

// End + 1 of the last candidate stored in the sub-regions before `sub_end` of one
// pattern's slice (cnt = its counts, row0 = index of its first sub-region), 0 if none.
__device__ __forceinline__ uint64_t FinPrevBefore(const SubStore& st, const uint32_t* cnt, uint64_t row0,
                                                  uint64_t sub_end, int local_pred, int lane) {
  uint64_t s2 = sub_end;
  while (s2 > 0) {
    const uint64_t base = s2 >= 32 ? s2 - 32 : 0;
    const uint64_t sub = base + lane;
    uint32_t c = (sub < s2) ? __ldcg(&cnt[sub]) : 0u;
    if (c == kLaneListOverflow) c = 0;
    if (c > st.cap) c = st.cap;
    const unsigned has = __ballot_sync(kFullMask, c != 0);
    if (has) {
      const int top = 31 - __clz(has);
      const uint32_t ct = __shfl_sync(kFullMask, c, top);
      return __ldcg(&st.end[(row0 + base + top) * st.cap + ct - 1]) + 1;
    }
    if (local_pred) return 0;
    s2 = base;
  }
  return 0;
}

// One candidate against its predecessor (P = predecessor's end + 1, 0 = none):
// the chain takes it iff it restarts the chain (ChainTake, device_program.h).
__device__ __forceinline__ bool FinTaken(uint64_t b, uint64_t e, uint64_t P, const Carry& cin) {
  bool ok = P == 0 || P <= b || (P == b + 1 && e > b);
  if (b < cin.cur || (e == b && b == cin.tail)) ok = false;
  return ok;
}

// ---------------------------------------------------------------------------
// FinishOrdered: the in-kernel finish of an ordered candidate store.
//   grid barrier -> every CTA takes one segment of sub-regions; its warps share
//   the (pattern, chunk of 32 sub-regions) items:
//     pass 1: counts and last ends of the item -> shared scratch
//     pass 2: offset of the item (segment prefix from the scan's atomic counters
//             + the chunk sums before it), its predecessor (last end before it),
//             then the copy to the output with the "every candidate restarts the
//             chain" check.
//   The last warp of the grid publishes the records.
// Latency matters here, not bandwidth: loads that do not depend on each other are
// issued together; sparse chunks are copied lane-per-sub-region, dense ones
// warp-per-sub-region (coalesced).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void FinishOrdered(const SubStore& st, uint64_t nsub_pat, int K, const FinishArgs& fin,
                                              const Carry* carries, uint32_t* scratch, uint32_t scratch_words) {
  GridBarrier(&fin.sync[0], gridDim.x);
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const unsigned int flags0 = __ldcg(&fin.sync[2]);      // consumed after the loads below are in flight
  for (uint32_t seg = blockIdx.x; seg < fin.nseg; seg += gridDim.x) {
    const uint64_t sub0 = (uint64_t)seg * fin.seg_subs;
    const uint64_t sub1 = (sub0 + fin.seg_subs < nsub_pat) ? sub0 + fin.seg_subs : nsub_pat;
    const uint32_t nch = (uint32_t)((sub1 - sub0 + 31) / 32);
    const uint32_t items = (uint32_t)K * nch;
    const uint32_t sum_words = (items + 1) & ~1u;
    const uint32_t seg_len = (uint32_t)(sub1 - sub0);
    const uint32_t cells = (uint32_t)K * seg_len;                  // (pattern, sub-region) pairs of the segment
    const uint32_t cell_words = (cells + 1) & ~1u;
    // scratch: [chunk sums u32][cell counts u32][chunk last P u64][cell offsets u64][cell predecessors u64]
    uint32_t* s_sum = scratch;
    uint32_t* s_cnt = scratch + sum_words;
    uint64_t* s_last = reinterpret_cast<uint64_t*>(scratch + sum_words + cell_words);
    uint64_t* s_at = s_last + items;
    uint64_t* s_prev = s_at + cells;
    unsigned long long* s_ne = reinterpret_cast<unsigned long long*>(s_prev + cells);   // [K] last non-empty end
    const bool fits = (uint64_t)sum_words + cell_words + 2ull * items + 4ull * cells + 2ull * K <= scratch_words;
    if (fits) for (int j = threadIdx.x; j < K; j += blockDim.x) s_ne[j] = 0;
    if (!fits && threadIdx.x == 0) atomicOr(&fin.sync[2], kFinOverlap);      // the host resolves instead
    const bool usable = fits && !(flags0 & (kFinDense | kFinOverflow));
    // ---- pass 1: counts and last ends of every item ------------------------------
    uint32_t c_keep = 0;
    uint64_t P_keep = 0;
    unsigned long long pre_keep = 0, mine_keep = 0;
    for (uint32_t it = warp_in_cta, k = 0; it < items; it += nwarps, ++k) {
      const uint32_t j = it / nch, ci = it - j * nch;
      const uint64_t sub = sub0 + (uint64_t)ci * 32 + lane;
      const uint32_t* cnt = st.count + (uint64_t)j * nsub_pat;
      uint32_t c = (sub < sub1) ? __ldcg(&cnt[sub]) : 0u;
      unsigned long long pre = 0, mine = 0;
      if (k == 0) {
        for (uint32_t s2 = lane; s2 < fin.nseg; s2 += 32) {
          const uint32_t v = __ldcg(&fin.segcount[(uint64_t)j * fin.nseg + s2]);
          if (s2 < seg) pre += v;
          if (s2 == seg) mine = v;
        }
      }
      if (c == kLaneListOverflow) c = 0;
      if (c > st.cap) c = st.cap;
      const uint64_t P = c ? __ldcg(&st.end[((uint64_t)j * nsub_pat + sub) * st.cap + c - 1]) + 1 : 0;
      uint32_t sum = c;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(kFullMask, sum, d);
      const uint64_t mx = WarpMax64(P);
      if (fits && lane == 0) { s_sum[it] = sum; s_last[it] = mx; }
      if (k == 0) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          pre += __shfl_xor_sync(kFullMask, pre, d);
          mine += __shfl_xor_sync(kFullMask, mine, d);
        }
        c_keep = c; P_keep = P; pre_keep = pre; mine_keep = mine;
      }
    }
    __syncthreads();
    // ---- pass 2: where every sub-region's candidates go, and who precedes them ---
    for (uint32_t it = warp_in_cta, k = 0; it < items; it += nwarps, ++k) {
      const uint32_t j = it / nch, ci = it - j * nch;
      const uint64_t row0 = (uint64_t)j * nsub_pat;
      const uint64_t sub = sub0 + (uint64_t)ci * 32 + lane;
      const uint32_t* cnt = st.count + row0;
      uint32_t c = c_keep;
      uint64_t P = P_keep;
      unsigned long long pre = pre_keep, mine = mine_keep;
      if (k != 0) {
        c = (sub < sub1) ? __ldcg(&cnt[sub]) : 0u;
        pre = 0; mine = 0;
        for (uint32_t s2 = lane; s2 < fin.nseg; s2 += 32) {
          const uint32_t v = __ldcg(&fin.segcount[(uint64_t)j * fin.nseg + s2]);
          if (s2 < seg) pre += v;
          if (s2 == seg) mine = v;
        }
        if (c == kLaneListOverflow) c = 0;
        if (c > st.cap) c = st.cap;
        P = c ? __ldcg(&st.end[(row0 + sub) * st.cap + c - 1]) + 1 : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          pre += __shfl_xor_sync(kFullMask, pre, d);
          mine += __shfl_xor_sync(kFullMask, mine, d);
        }
      }
      if (ci + 1 == nch && seg + 1 == fin.nseg && lane == 0) fin.totals[j] = pre + mine;
      if (!usable) continue;
      const uint32_t cell = j * seg_len + ci * 32 + lane;
      if (mine == 0) { if (sub < sub1) s_cnt[cell] = 0; continue; }
      unsigned long long before = 0;
      uint64_t prevP = 0;
      for (uint32_t c2 = lane; c2 < ci; c2 += 32) {
        before += s_sum[j * nch + c2];
        const uint64_t q = s_last[j * nch + c2];
        prevP = q > prevP ? q : prevP;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(kFullMask, before, d);
      prevP = WarpMax64(prevP);
      const uint32_t incl = WarpInclusiveScan(c);
      const uint32_t total = __shfl_sync(kFullMask, incl, 31);
      if (total != 0 && prevP == 0) prevP = FinPrevBefore(st, cnt, row0, sub0, fin.local_pred, lane);
      // predecessor of my sub-region's first candidate: the last end among the lanes before me
      uint64_t inclP = P;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint64_t o = __shfl_up_sync(kFullMask, inclP, d);
        if (lane >= d && o > inclP) inclP = o;
      }
      uint64_t myprev = __shfl_up_sync(kFullMask, inclP, 1);
      if (lane == 0) myprev = 0;
      if (prevP > myprev) myprev = prevP;
      if (sub < sub1) {
        s_cnt[cell] = c;
        s_at[cell] = pre + before + incl - c;
        s_prev[cell] = myprev;
      }
    }
    __syncthreads();
    // the segment's last end per pattern: one global atomic per (CTA, pattern)
    if (usable)
      for (int j = threadIdx.x; j < K; j += blockDim.x) {
        uint64_t m2 = 0;
        for (uint32_t c2 = 0; c2 < nch; ++c2) { const uint64_t q = s_last[j * nch + c2]; m2 = q > m2 ? q : m2; }
        if (m2) atomicMax(&fin.last_end[j], (unsigned long long)(m2 - 1));
      }
    // ---- pass 3: the copy.  Cells with at most 4 candidates: one LANE per cell (all
    // loads of 32 cells in flight at once); larger cells: the whole warp per cell,
    // 128 candidates in flight.
    if (usable) {
      bool bad = false;
      for (uint32_t base = warp_in_cta * 32; base < cells; base += nwarps * 32) {
        const uint32_t cell = base + lane;
        const uint32_t cs = cell < cells ? s_cnt[cell] : 0u;
        if (cs > 0 && cs <= 4) {
          const uint32_t j = cell / seg_len;
          const uint64_t slot0 = ((uint64_t)j * nsub_pat + sub0 + (cell - j * seg_len)) * st.cap;
          const unsigned long long ats = s_at[cell];
          uint64_t P = s_prev[cell];
          const Carry cin = carries[j];
          ulonglong2* outp = reinterpret_cast<ulonglong2*>(fin.out_pairs + (uint64_t)j * 2 * fin.out_stride);
          uint64_t bb[4], ee[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            bb[u] = ((uint32_t)u < cs) ? __ldcg(&st.begin[slot0 + u]) : 0;
            ee[u] = ((uint32_t)u < cs) ? __ldcg(&st.end[slot0 + u]) : 0;
          }
          uint64_t ne = 0;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if ((uint32_t)u < cs) {
              bad |= !FinTaken(bb[u], ee[u], P, cin);
              P = ee[u] + 1;
              if (ee[u] > bb[u]) ne = ee[u];
              if (ats + u < fin.out_cap) outp[ats + u] = make_ulonglong2(bb[u] + fin.base_offset, ee[u] + fin.base_offset);
            }
          }
          if (ne) atomicMax(&s_ne[j], (unsigned long long)ne);
        }
        unsigned big = __ballot_sync(kFullMask, cs > 4);
        while (big) {
          const int src = __ffs(big) - 1;
          big &= big - 1;
          const uint32_t bcell = base + src;
          const uint32_t bcs = __shfl_sync(kFullMask, cs, src);
          const uint32_t j = bcell / seg_len;
          const uint64_t slot0 = ((uint64_t)j * nsub_pat + sub0 + (bcell - j * seg_len)) * st.cap;
          const unsigned long long ats = s_at[bcell];
          uint64_t carryP = s_prev[bcell];
          const Carry cin = carries[j];
          ulonglong2* outp = reinterpret_cast<ulonglong2*>(fin.out_pairs + (uint64_t)j * 2 * fin.out_stride);
          uint64_t ne = 0;
          for (uint32_t i0 = 0; i0 < bcs; i0 += 128) {
            uint64_t bb[4], ee[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t i = i0 + u * 32 + lane;
              bb[u] = (i < bcs) ? __ldcg(&st.begin[slot0 + i]) : 0;
              ee[u] = (i < bcs) ? __ldcg(&st.end[slot0 + i]) : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint32_t ib = i0 + u * 32;
              if (ib >= bcs) break;
              const uint32_t i = ib + lane;
              uint64_t Pp = __shfl_up_sync(kFullMask, ee[u] + 1, 1);
              if (lane == 0) Pp = carryP;
              if (i < bcs) {
                bad |= !FinTaken(bb[u], ee[u], Pp, cin);
                if (ee[u] > bb[u]) ne = ee[u];
                if (ats + i < fin.out_cap) outp[ats + i] = make_ulonglong2(bb[u] + fin.base_offset, ee[u] + fin.base_offset);
              }
              const uint32_t nvalid = bcs - ib > 32 ? 32u : bcs - ib;
              carryP = __shfl_sync(kFullMask, ee[u] + 1, nvalid - 1);
            }
          }
          ne = WarpMax64(ne);
          if (lane == 0 && ne) atomicMax(&s_ne[j], (unsigned long long)ne);
        }
      }
      if (__any_sync(kFullMask, bad) && lane == 0) atomicOr(&fin.sync[2], kFinOverlap);
    }
    __syncthreads();
    if (usable)
      for (int j = threadIdx.x; j < K; j += blockDim.x)
        if (s_ne[j]) atomicMax(&fin.last_ne[j], s_ne[j]);
    __syncthreads();
  }
  // the last WARP of the grid to get here publishes the records (every warp
  // counts once; its atomics above are ordered before the count by the fence)
  __syncwarp();
  int is_last = 0;
  if (lane == 0) {
    __threadfence();
    is_last = (atomicAdd(&fin.sync[1], 1u) == gridDim.x * (unsigned)nwarps - 1u) ? 1 : 0;
  }
  if (!__shfl_sync(kFullMask, is_last, 0)) return;
  __threadfence();
  const unsigned int flags = __ldcg(&fin.sync[2]);
  const unsigned int need_cap = __ldcg(&fin.sync[3]);
  for (int j = lane; j < K; j += 32) {
    const unsigned long long tot = __ldcg(&fin.totals[j]);
    const unsigned long long le = __ldcg(&fin.last_end[j]);
    const unsigned long long ln = __ldcg(&fin.last_ne[j]);
    volatile uint4* dst = reinterpret_cast<volatile uint4*>(fin.host_records + j);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"((unsigned int)tot),
                 "r"((unsigned int)(tot >> 32)), "r"(flags), "r"(fin.seq) : "memory");
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 1), "r"((unsigned int)le),
                 "r"((unsigned int)(le >> 32)), "r"(need_cap), "r"(fin.seq) : "memory");
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 2), "r"((unsigned int)ln),
                 "r"((unsigned int)(ln >> 32)), "r"(0u), "r"(fin.seq) : "memory");
  }
}

constexpr int kDfaChainHits = 3;

__device__ __forceinline__ uint32_t Lds32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t Lds8(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 Lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// byte-wise replay of one 16-byte group with the 1-byte table; records the
// match ends (relative to sub_lo) that fall in (lo_excl, hi_incl]
__device__ __forceinline__ void DfaReplay(const uint4& v, uint32_t st1, const uint16_t* s_next, const uint8_t* s_class,
                                          uint32_t acc1, uint64_t p0, uint64_t limit, uint32_t L, uint64_t sub_lo,
                                          const ScanRange& range, uint32_t* hit, uint32_t& cnt) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  for (int i = 0; i < 16 && p0 + i < limit; ++i) {
    uint32_t c = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
    st1 = s_next[st1 + s_class[c]];
    if (st1 >= acc1) {
      uint64_t e = p0 + i + 1;
      if (e >= L) {
        uint64_t s = e - L;
        if (s >= range.own_begin && s < range.own_end) {
#pragma unroll
          for (int q = 0; q < kDfaChainHits; ++q)
            if (cnt == (uint32_t)q) hit[q] = (uint32_t)(e - sub_lo);
          ++cnt;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(576, 1)
k_dfa_tma(const uint8_t* __restrict__ text, uint64_t n, DfaTables dfa, ScanRange range, SubStore out,
          unsigned int* dense_flag, unsigned long long* work_counter, FinishArgs fin, CarrySet carries) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const uint32_t C = (uint32_t)dfa.n_classes;
  const uint32_t C2 = C * C;
  const int pair_entries = dfa.n_states * (int)C2;
  const int next_entries = dfa.n_states * (int)C;
  // layout: [pair rows: pair_entries*128][1-byte table u16][class map 256][barriers][tiles]
  uint32_t* s_rows = reinterpret_cast<uint32_t*>(smem_raw);
  uint16_t* s_next = reinterpret_cast<uint16_t*>(smem_raw + (size_t)pair_entries * 128);
  uint8_t* s_class = reinterpret_cast<uint8_t*>(s_next) + (((size_t)next_entries * 2 + 15) & ~(size_t)15);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_class + 256);
  uint8_t* s_tiles = reinterpret_cast<uint8_t*>(s_bar + warps_per_cta);
  s_tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_tiles) + 127) & ~(uintptr_t)127);
  const uint32_t rows_base = SmemAddr(s_rows);
  const uint32_t row_stride = C2 * 128u;                               // bytes between the rows of two states
  // the compact pair table is first copied (coalesced, one round trip) into the
  // not-yet-used tile area, then expanded into the per-lane replicated rows
  uint32_t* s_tmp = reinterpret_cast<uint32_t*>(s_tiles);
  for (int i = threadIdx.x; i < pair_entries; i += blockDim.x) s_tmp[i] = dfa.pair[i];
  for (int i = threadIdx.x; i < next_entries; i += blockDim.x) s_next[i] = dfa.next[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_class[i] = dfa.byte_class[i];
  uint64_t* bar = s_bar + warp_in_cta;
  if (lane == 0) MbarInit(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < pair_entries * 32; i += blockDim.x) {
    uint32_t v = s_tmp[i >> 5];
    s_rows[i] = (rows_base + (v & 0x7FFFFFFFu) * row_stride + (uint32_t)(i & 31) * 4u) | (v & 0x80000000u);
  }
  __syncthreads();

  // tile of this warp: [16 bytes before the sub-region][32 x 272 bytes]; lane l's
  // sub-stream starts at byte 16 + 272*l, its warm-up chunk at 272*l
  uint8_t* tile = s_tiles + (size_t)warp_in_cta * kDfaTileBytes;
  const uint32_t my_warm_addr = SmemAddr(tile) + (uint32_t)lane * kDfaStreamBytes;
  const uint32_t lane_base = rows_base + (uint32_t)lane * 4u;          // row of state 0 for this lane
  const uint32_t acc_addr = lane_base + (uint32_t)dfa.first_accept * row_stride;
  const uint32_t acc1 = (uint32_t)dfa.first_accept_scaled;
  const uint32_t class_base = SmemAddr(s_class);
  const uint32_t L = dfa.match_len;
  uint32_t phase = 0;

  for (;;) {
    // dynamic scheduling: sub-regions are handed out in order, one per request
    unsigned long long sub = 0;
    if (lane == 0) sub = atomicAdd(work_counter, 1ull);
    sub = __shfl_sync(kFullMask, sub, 0);
    if (sub >= out.nsub) break;
    const uint64_t sub_lo = sub * kDfaSubBytes;
    const bool live = sub_lo < n && sub_lo + kDfaSubBytes + L > range.own_begin && sub_lo < range.own_end + L;
    uint32_t cntA = 0, cntB = 0;
    uint32_t hitA[kDfaChainHits] = {0, 0, 0}, hitB[kDfaChainHits] = {0, 0, 0};
    if (live) {
      // ---- stage the tile: ONE bulk copy per warp ---------------------------
      const bool sub_warm = sub_lo >= 16;
      if (lane == 0) {
        const uint64_t src = sub_warm ? sub_lo - 16 : 0;
        uint64_t end = sub_lo + kDfaSubBytes;
        if (end > n) end = (n + 15) & ~15ull;                              // stays inside the last 16-byte block
        const uint32_t bytes = (uint32_t)(end - src);
        MbarExpectTx(bar, bytes);
        TmaLoad1D(tile + (sub_warm ? 0 : 16), text + src, bytes, bar);
      }
      MbarWait(bar, phase);
      phase ^= 1;
      const uint64_t a = sub_lo + (uint64_t)lane * kDfaStreamBytes;        // my sub-stream [a, b)
      const uint64_t b = (a + kDfaStreamBytes < n) ? a + kDfaStreamBytes : n;
      if (a < n) {
        // ---- two chains (144 + 128 bytes), two bytes per lookup ----------------
        // bytes past the end of the text are stale shared memory: they can only
        // produce "ends" beyond b, which the replay discards
        const bool warm = a >= 16;
        uint32_t rowA = lane_base, rowB = lane_base;
#pragma unroll 1
        for (uint32_t it = 0; it < 10; ++it) {
          const uint4 vA = Lds128(my_warm_addr + it * 16);
          const uint4 vB = Lds128(my_warm_addr + (it < 9 ? (9 + it) * 16 : 0));
          const uint32_t wA[4] = {vA.x, vA.y, vA.z, vA.w};
          const uint32_t wB[4] = {vB.x, vB.y, vB.z, vB.w};
          uint32_t pA[8], pB[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            uint32_t a0 = Lds8(class_base + __byte_perm(wA[k >> 1], 0, 0x4440 + 2 * (k & 1)));
            uint32_t a1 = Lds8(class_base + __byte_perm(wA[k >> 1], 0, 0x4441 + 2 * (k & 1)));
            uint32_t b0 = Lds8(class_base + __byte_perm(wB[k >> 1], 0, 0x4440 + 2 * (k & 1)));
            uint32_t b1 = Lds8(class_base + __byte_perm(wB[k >> 1], 0, 0x4441 + 2 * (k & 1)));
            pA[k] = (a0 * C + a1) * 128u;
            pB[k] = (b0 * C + b1) * 128u;
          }
          const uint32_t rowA0 = rowA, rowB0 = rowB;
          uint32_t peakA = 0, peakB = 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            uint32_t eA = Lds32(rowA + pA[k]);
            uint32_t eB = Lds32(rowB + pB[k]);
            peakA = max(peakA, eA);
            peakB = max(peakB, eB);
            rowA = eA & 0x7FFFFFFFu;
            rowB = eB & 0x7FFFFFFFu;
          }
          if (it == 0) {
            if (!warm) rowA = lane_base;              // no bytes before the text: chain A starts at its data
          } else {
            if (peakA >= acc_addr)
              DfaReplay(vA, ((rowA0 - lane_base) / row_stride) * C, s_next, s_class, acc1,
                        a - 16 + (uint64_t)it * 16, b, L, sub_lo, range, hitA, cntA);
            if (it < 9 && peakB >= acc_addr)
              DfaReplay(vB, ((rowB0 - lane_base) / row_stride) * C, s_next, s_class, acc1,
                        a + 128 + (uint64_t)it * 16, b, L, sub_lo, range, hitB, cntB);
          }
        }
      }
      __syncwarp();      // everyone is done with the tile before the next TMA overwrites it
    }
    // ---- ordered emission: lanes cover increasing offsets, A before B ---------
    const uint32_t my_cnt = cntA + cntB;
    if (__any_sync(kFullMask, my_cnt != 0)) {
      bool over = __any_sync(kFullMask, cntA > kDfaChainHits || cntB > kDfaChainHits);
      uint32_t cA = cntA > kDfaChainHits ? kDfaChainHits : cntA;
      uint32_t cB = cntB > kDfaChainHits ? kDfaChainHits : cntB;
      uint32_t c = cA + cB;
      uint32_t incl = WarpInclusiveScan(c);
      uint32_t total = __shfl_sync(kFullMask, incl, 31);
      uint32_t idx = incl - c;
#pragma unroll
      for (int q = 0; q < kDfaChainHits; ++q) {
        if (q < (int)cA && idx + q < out.cap) {
          uint64_t e = sub_lo + hitA[q];
          out.begin[sub * out.cap + idx + q] = e - L;
          out.end[sub * out.cap + idx + q] = e;
        }
      }
#pragma unroll
      for (int q = 0; q < kDfaChainHits; ++q) {
        if (q < (int)cB && idx + cA + q < out.cap) {
          uint64_t e = sub_lo + hitB[q];
          out.begin[sub * out.cap + idx + cA + q] = e - L;
          out.end[sub * out.cap + idx + cA + q] = e;
        }
      }
      if (lane == 0) {
        out.count[sub] = over ? kLaneListOverflow : total;
        if (over) *dense_flag = 1u;
        FinishNote(fin, 0, sub, total, over, out.cap);
      }
    } else if (lane == 0) {
      out.count[sub] = 0;
    }
  }
  if (fin.enabled) FinishOrdered(out, out.nsub, 1, fin, carries.c, reinterpret_cast<uint32_t*>(s_tiles),
                                 (uint32_t)warps_per_cta * (kDfaTileBytes / 4));
}

// ===========================================================================
// K2-set: the same scan for a SET of fixed-length patterns fused into one
// automaton (regex-dna counts nine patterns over one text): one pass over the
// text advances all of them.  Same tile staging, chain geometry and two-byte
// steps as k_dfa_tma; the union automaton is too large to replicate per lane, so
// its pair table is shared (entry = byte offset of the next state's row, bit 31
// = accept in between).  An accepting state carries a bitmask of the patterns
// ending there; the replay fans the ends out per pattern, and the emission
// writes every pattern's ends, in order, into that pattern's own slot range
// (slot index = pattern * nsub + sub-region).
// Algorithmic traffic: N bytes read (once for all patterns) + 16 bytes per match.
// ===========================================================================
struct SetTables {
  const uint16_t* t1;                  // [R*C] next state * C           (R rows = states + shadow rows)
  const uint32_t* t2;                  // [R << (row_shift-2)] row offset after two bytes
  const uint8_t* byte_class;           // [256]
  const uint32_t* accept_mask;         // [R]
  uint32_t match_len[32];
  int n_patterns;
  int n_rows, n_classes, first_accept;
  int row_shift;                       // pair-table rows are 2^row_shift bytes
};

constexpr int kSetChainHits = 4;

// records the ends of the patterns accepted in `state` at text offset `e`
__device__ __forceinline__ void SetRecord(uint32_t state, uint64_t e, uint64_t limit, const uint32_t* s_mask,
                                          const SetTables& tb, uint64_t sub_lo, const ScanRange& range,
                                          uint32_t* hit, uint32_t& cnt) {
  if (e > limit) return;
  uint32_t m = s_mask[state];
  while (m) {
    int j = __ffs(m) - 1;
    m &= m - 1;
    uint32_t L = tb.match_len[j];
    if (e >= L) {
      uint64_t s = e - L;
      if (s >= range.own_begin && s < range.own_end) {
#pragma unroll
        for (int q = 0; q < kSetChainHits; ++q)
          if (cnt == (uint32_t)q) hit[q] = (uint32_t)(e - sub_lo) | ((uint32_t)j << 16);
        ++cnt;
      }
    }
  }
}

// Entries of the pair table are shared-memory ADDRESSES of the next row (the
// kernel adds the table base while staging it); rows at or above the first
// accepting row mean "an accept happened in this pair" (shadow rows: only in
// between), so one step of a chain is   row = *(row + hi[b0] + lo[b1])   with
// hi = class*C*4 and lo = class*4 looked up per byte.
__global__ void __launch_bounds__(576, 1)
k_set_tma(const uint8_t* __restrict__ text, uint64_t n, SetTables tb, ScanRange range, SubStore out,
          uint64_t nsub_pat, unsigned int* dense_flag, unsigned long long* work_counter, FinishArgs fin,
          CarrySet carries) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const uint32_t C = (uint32_t)tb.n_classes;
  const int t2_entries = tb.n_rows << (tb.row_shift - 2);
  const int t1_entries = tb.n_rows * (int)C;
  // layout: [t2 u32, padded rows][accept masks u32][t1 u16][class 256][hi 256][lo 256][barriers][tiles]
  uint32_t* s_t2 = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* s_mask = s_t2 + t2_entries;
  uint16_t* s_t1 = reinterpret_cast<uint16_t*>(s_mask + tb.n_rows);
  uint8_t* s_class = reinterpret_cast<uint8_t*>(s_t1) + (((size_t)t1_entries * 2 + 15) & ~(size_t)15);
  uint8_t* s_hi = s_class + 256;
  uint8_t* s_lo = s_hi + 256;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_lo + 256) + 7) & ~(uintptr_t)7);
  uint8_t* s_tiles = reinterpret_cast<uint8_t*>(s_bar + warps_per_cta);
  s_tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_tiles) + 127) & ~(uintptr_t)127);
  const uint32_t t2_base = SmemAddr(s_t2);
  for (int i = threadIdx.x; i < t2_entries; i += blockDim.x) s_t2[i] = tb.t2[i] + t2_base;
  for (int i = threadIdx.x; i < tb.n_rows; i += blockDim.x) s_mask[i] = tb.accept_mask[i];
  for (int i = threadIdx.x; i < t1_entries; i += blockDim.x) s_t1[i] = tb.t1[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const uint32_t c = tb.byte_class[i];
    s_class[i] = (uint8_t)c;
    s_hi[i] = (uint8_t)(c * C * 4u);
    s_lo[i] = (uint8_t)(c * 4u);
  }
  uint64_t* bar = s_bar + warp_in_cta;
  if (lane == 0) MbarInit(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  uint8_t* tile = s_tiles + (size_t)warp_in_cta * kDfaTileBytes;
  const uint32_t my_warm_addr = SmemAddr(tile) + (uint32_t)lane * kDfaStreamBytes;
  const int row_shift = tb.row_shift;
  const uint32_t acc_row = t2_base + ((uint32_t)tb.first_accept << row_shift);
  const uint32_t hi_base = SmemAddr(s_hi);
  const uint32_t lo_base = SmemAddr(s_lo);
  const int K = tb.n_patterns;
  uint32_t phase = 0;

  for (;;) {
    unsigned long long sub = 0;
    if (lane == 0) sub = atomicAdd(work_counter, 1ull);
    sub = __shfl_sync(kFullMask, sub, 0);
    if (sub >= nsub_pat) break;
    const uint64_t sub_lo = sub * kDfaSubBytes;
    const bool live = sub_lo < n && sub_lo + kDfaSubBytes + 32 > range.own_begin && sub_lo < range.own_end + 32;
    uint32_t cntA = 0, cntB = 0;
    uint32_t hitA[kSetChainHits] = {0, 0, 0, 0}, hitB[kSetChainHits] = {0, 0, 0, 0};
    if (live) {
      const bool sub_warm = sub_lo >= 16;
      if (lane == 0) {
        const uint64_t src = sub_warm ? sub_lo - 16 : 0;
        uint64_t end = sub_lo + kDfaSubBytes;
        if (end > n) end = (n + 15) & ~15ull;
        const uint32_t bytes = (uint32_t)(end - src);
        MbarExpectTx(bar, bytes);
        TmaLoad1D(tile + (sub_warm ? 0 : 16), text + src, bytes, bar);
      }
      MbarWait(bar, phase);
      phase ^= 1;
      const uint64_t a = sub_lo + (uint64_t)lane * kDfaStreamBytes;
      const uint64_t b = (a + kDfaStreamBytes < n) ? a + kDfaStreamBytes : n;
      if (a < n) {
        const bool warm = a >= 16;
        uint32_t rowA = t2_base, rowB = t2_base;        // shared-memory address of the state's row
#pragma unroll 1
        for (uint32_t it = 0; it < 10; ++it) {
          const uint4 vA = Lds128(my_warm_addr + it * 16);
          const uint4 vB = Lds128(my_warm_addr + (it < 9 ? (9 + it) * 16 : 0));
          const uint32_t wA[4] = {vA.x, vA.y, vA.z, vA.w};
          const uint32_t wB[4] = {vB.x, vB.y, vB.z, vB.w};
          uint32_t pA[8], pB[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            uint32_t a0 = Lds8(hi_base + __byte_perm(wA[k >> 1], 0, 0x4440 + 2 * (k & 1)));
            uint32_t a1 = Lds8(lo_base + __byte_perm(wA[k >> 1], 0, 0x4441 + 2 * (k & 1)));
            uint32_t b0 = Lds8(hi_base + __byte_perm(wB[k >> 1], 0, 0x4440 + 2 * (k & 1)));
            uint32_t b1 = Lds8(lo_base + __byte_perm(wB[k >> 1], 0, 0x4441 + 2 * (k & 1)));
            pA[k] = a0 + a1;
            pB[k] = b0 + b1;
          }
          const uint32_t rowA0 = rowA, rowB0 = rowB;
          uint32_t peakA = 0, peakB = 0;
          uint32_t eA[8], eB[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            eA[k] = Lds32(rowA + pA[k]);
            eB[k] = Lds32(rowB + pB[k]);
            peakA = max(peakA, eA[k]);
            peakB = max(peakB, eB[k]);
            rowA = eA[k];
            rowB = eB[k];
          }
          if (it == 0) {
            if (!warm) rowA = t2_base;
          } else if (peakA >= acc_row || (it < 9 && peakB >= acc_row)) {
            // rare: some pair of this group accepted — walk the saved entries
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
              if (ch == 0 ? (peakA < acc_row) : (it >= 9 || peakB < acc_row)) continue;
              uint32_t prev = ch ? rowB0 : rowA0;
              const uint64_t p0 = (ch ? a + 128 : a - 16) + (uint64_t)it * 16;
              uint32_t cnt = ch ? cntB : cntA;
              uint32_t hit[kSetChainHits];
#pragma unroll
              for (int q = 0; q < kSetChainHits; ++q) hit[q] = ch ? hitB[q] : hitA[q];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint32_t e = ch ? eB[k] : eA[k];
                if (e >= acc_row) {
                  const uint32_t w = ch ? wB[k >> 1] : wA[k >> 1];
                  const uint32_t c1 = s_class[__byte_perm(w, 0, 0x4440 + 2 * (k & 1))];
                  const uint32_t mid = s_t1[((prev - t2_base) >> row_shift) * C + c1] / C;
                  SetRecord(mid, p0 + 2 * k + 1, b, s_mask, tb, sub_lo, range, hit, cnt);
                  SetRecord((e - t2_base) >> row_shift, p0 + 2 * k + 2, b, s_mask, tb, sub_lo, range, hit, cnt);
                }
                prev = e;
              }
              if (ch) { cntB = cnt; } else { cntA = cnt; }
#pragma unroll
              for (int q = 0; q < kSetChainHits; ++q) { if (ch) hitB[q] = hit[q]; else hitA[q] = hit[q]; }
            }
          }
        }
      }
      __syncwarp();
    }
    // ---- ordered emission, one slot range per pattern --------------------------
    const uint32_t cA = cntA > kSetChainHits ? kSetChainHits : cntA;
    const uint32_t cB = cntB > kSetChainHits ? kSetChainHits : cntB;
    uint32_t pm = 0;
#pragma unroll
    for (int q = 0; q < kSetChainHits; ++q) {
      if (q < (int)cA) pm |= 1u << (hitA[q] >> 16);
      if (q < (int)cB) pm |= 1u << (hitB[q] >> 16);
    }
    uint32_t warp_pm = pm;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) warp_pm |= __shfl_xor_sync(kFullMask, warp_pm, d);
    if (lane < K && !((warp_pm >> lane) & 1u)) out.count[(uint64_t)lane * nsub_pat + sub] = 0;
    if (warp_pm) {
      bool over = __any_sync(kFullMask, cntA > kSetChainHits || cntB > kSetChainHits);
      uint32_t todo = warp_pm;
      while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        uint32_t c = 0;
#pragma unroll
        for (int q = 0; q < kSetChainHits; ++q) {
          if (q < (int)cA && (int)(hitA[q] >> 16) == j) ++c;
          if (q < (int)cB && (int)(hitB[q] >> 16) == j) ++c;
        }
        uint32_t incl = WarpInclusiveScan(c);
        uint32_t total = __shfl_sync(kFullMask, incl, 31);
        uint32_t idx = incl - c;
        const uint64_t slot0 = ((uint64_t)j * nsub_pat + sub) * out.cap;
        const uint32_t L = tb.match_len[j];
#pragma unroll
        for (int q = 0; q < kSetChainHits; ++q) {
          if (q < (int)cA && (int)(hitA[q] >> 16) == j) {
            if (idx < out.cap) { uint64_t e = sub_lo + (hitA[q] & 0xFFFFu); out.begin[slot0 + idx] = e - L; out.end[slot0 + idx] = e; }
            ++idx;
          }
        }
#pragma unroll
        for (int q = 0; q < kSetChainHits; ++q) {
          if (q < (int)cB && (int)(hitB[q] >> 16) == j) {
            if (idx < out.cap) { uint64_t e = sub_lo + (hitB[q] & 0xFFFFu); out.begin[slot0 + idx] = e - L; out.end[slot0 + idx] = e; }
            ++idx;
          }
        }
        if (lane == 0) {
          out.count[(uint64_t)j * nsub_pat + sub] = over ? kLaneListOverflow : total;
          FinishNote(fin, j, sub, total, over, out.cap);
        }
      }
      if (over && lane == 0) *dense_flag = 1u;
    }
  }
  if (fin.enabled) FinishOrdered(out, nsub_pat, K, fin, carries.c, reinterpret_cast<uint32_t*>(s_tiles),
                                 (uint32_t)warps_per_cta * (kDfaTileBytes / 4));
}

// ===========================================================================
// K2-kmer: the set scan for members of at most 8 bytes over at most four live
// byte values (SetDfa::Kmer in host/automaton.h; the nine regex-dna variants of
// sample/regexdna.cc:52-62).  "Does a member end at e" is a function of the
// eight 2-bit codes before e, so there is no automaton state and no dependent
// chain.
//   tables   every CTA copies the set's small tables (match lengths, the list of
//            accepted 8-mers and their member masks) to shared memory with one
//            coalesced load issued before anything else, and builds from the
//            list the bitmap (index = 7 + R codes, R = 3 consecutive ends per
//            lookup: 2^20 bits = 128 KB) and a 512-slot hash 8-mer -> member mask.
//            Sets that accept more than kKmerListMax 8-mers fetch the bitmap with
//            one TMA bulk copy and read the masks from a global 65536-entry table.
//   scan     every CTA owns a contiguous segment of 512-byte rows, every warp a
//            contiguous run of them.  A lane loads 16 bytes of a row (coalesced
//            uint4, four rows in flight, the loop unrolled by four so that the
//            pipeline registers are renamed, not moved), packs them into 16 codes
//            (one AND + one multiply per four bytes), takes the seven codes before
//            them from its neighbour (shuffle) and does 16 / R lookups.  A row with
//            a hit (one in four on regex-dna) costs one ballot and one shared-memory
//            store per lane with a hit: {group index, lookup bits}.  No branch on
//            data, no atomic, no call (a __noinline__ slow path drained the
//            prefetched rows on every call), nothing written to global memory.
//   finish   (same CTA, no second kernel) the warp expands its groups into hits in
//            position order (one warp scan), checks every hit exactly against the
//            text (one 64-bit read; a byte whose code aliases a live byte is not
//            that byte) and fans it out per member through the hash; per-member
//            indices by ballot.
//   exchange every CTA publishes, per member, {count, first end, last end} as two
//            16-byte records carrying the call's sequence number and reads the
//            records of the CTAs before it (cooperative launch: all resident): this
//            is the only grid-wide step.  It then writes its matches at their final
//            place; the last CTA also checks the seams and reports to the host.
// Candidates that overlap (or a carry reaching into the slab) raise kFinOverlap and
// the host repeats the call with k_set_tma + the general resolve.
// Kernel parameters stay small on purpose (the arrays live in `tab`): parameters
// are read through the constant cache and are best touched before the text streams.
// Algorithmic traffic: N bytes read (once for all members) + 16 bytes per match.
// ===========================================================================
constexpr uint32_t kKmerListMax = 128;  // accepted 8-mers up to which every CTA builds its tables itself
constexpr uint32_t kKmerHashSlots = 512;
// the set's small tables, one global array of words that every CTA copies to shared memory first thing (ONE
// coalesced load; kernel parameters are read through the constant cache, and a constant-cache miss issued
// while the text streams was measured to wait for microseconds)
constexpr uint32_t kKmerTabMatchLen = 0;      // [32]
constexpr uint32_t kKmerTabLenLe = 32;        // [9]
constexpr uint32_t kKmerTabListX = 48;        // [kKmerListMax] accepted 8-mers (codes, oldest in the low bits)
constexpr uint32_t kKmerTabListMask = kKmerTabListX + kKmerListMax;   // [kKmerListMax] their member masks
constexpr uint32_t kKmerTabWords = kKmerTabListMask + kKmerListMax;

struct KmerTables {
  const uint32_t* bitmap;              // [2^(kIdxBits - 5)]   (n_list == 0)
  const uint32_t* mask16;              // [65536]              (n_list == 0)
  const uint32_t* tab;                 // [kKmerTabWords]
  uint32_t field_mask;                 // 0x03030303 << shift
  uint32_t mult;                       // 0x01041040 >> shift: packs four fields into the top byte
  uint32_t shift;
  uint32_t canon;                      // byte c: the live byte with code c (or a byte with another code)
  // a short list of accepted 8-mers (the nine regex-dna variants accept 50): every CTA builds the bitmap and a
  // hash of the member masks in shared memory itself instead of fetching 128 KB + one L2 miss per hit
  uint32_t n_list;                     // 0: use bitmap / mask16 from global memory
};

struct __align__(16) KmerXchg {        // one per (CTA, member); each half is one 16-byte store
  unsigned long long first_end;
  unsigned int count, seq0;
  unsigned long long last_end;
  unsigned int flags, seq1;
};

struct KmerRun {
  uint64_t row_lo;                     // first 512-byte row that can hold an owned end
  uint32_t rows_per_cta, rows_per_warp;
  KmerXchg* xchg;                      // [gridDim.x][32]
  uint2* stage;                        // [gridDim.x * 32][stage_cap]: matches of a warp that no longer fit its shared list
  uint32_t stage_cap;
  unsigned int* gsync;                 // [0] flags, [1] CTAs done; zero between calls (the reporting CTA clears them)
  unsigned long long* gfinal;          // [32][2]: per member {matches, last end}, written by the CTA with the highest index;
                                       // [64]: the call's flags (for the stitch kernel that may follow, scan_emit.cuh)
  uint64_t* out_pairs;                 // member j's pairs start at out_pairs + j * 2 * out_stride
  uint64_t out_stride, out_cap, base_offset;
  FinRecord* host_records;
  unsigned int seq;
  int has_carry;                       // some member's chain arrives from the left (CarrySet is read only then)
  StitchLink stitch;                   // one process per GPU: the reporting CTA exchanges the chain states with the
                                       // neighbouring ranks' GPUs before it reports (stitch.cuh); enabled = 0 otherwise
#ifdef RJ_KMER_PROBE
  unsigned long long* probe;           // tuning builds only: [grid][32][2] globaltimer at scan start / end, then [grid] at exit
#endif
};

#ifdef RJ_KMER_PROBE
#define RJ_KMER_STAMP(slot)                                                                          \
  do {                                                                                               \
    if (run.probe && lane == 0) {                                                                    \
      unsigned long long t_;                                                                         \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                         \
      run.probe[((size_t)blockIdx.x * 32 + warp) * 2 + (slot)] = t_;                                 \
    }                                                                                                \
  } while (0)
#else
#define RJ_KMER_STAMP(slot) do {} while (0)
#endif

// R consecutive ends per lookup: a window of 7 + R codes indexes the bitmap
constexpr int kKmerEnds = kKmerR;                                   // host/automaton.h
constexpr int kKmerTests = (16 + kKmerEnds - 1) / kKmerEnds;        // lookups per 16-byte group
constexpr int kKmerIdxBits = 2 * (7 + kKmerEnds);
constexpr int kKmerWordBits = kKmerIdxBits - 5;
constexpr uint32_t kKmerBitmapBytes = 4u << kKmerWordBits;          // 32 KB (R = 2), 128 KB (R = 3)
constexpr uint32_t kKmerThreads = 1024;
constexpr uint32_t kKmerWarps = kKmerThreads / 32;
constexpr uint32_t kKmerMaxRowsPerCta = 1u << 19;  // offsets inside a CTA's rows fit 28 bits (256 MB of text per CTA)
constexpr uint32_t kKmerWarpRaw = 128;             // 16-byte groups with a hit / hits a warp collects before it checks them
constexpr uint32_t kKmerFlushAt = 96;              // ... it stops streaming and checks them at this many groups
constexpr uint32_t kKmerStageSm = 128;             // checked matches a warp keeps in shared memory
constexpr uint32_t kKmerSmemFixed = kKmerBitmapBytes + kKmerWarps * (kKmerWarpRaw * 8 + kKmerStageSm * 8) +
                                    4 * kKmerWarps * 32 * 4 + kKmerHashSlots * 8 + kKmerTabWords * 4 + 2048;
// first letter (0..15) of the ends lookup t answers: x, x+1, .., x+R-1; the last lookup is pulled back into the group
__host__ __device__ constexpr int KmerTestX(int t) { return t * kKmerEnds < 16 - kKmerEnds ? t * kKmerEnds : 16 - kKmerEnds; }

__device__ __forceinline__ uint32_t KmerPack(const uint4& v, uint32_t fm, uint32_t mult) {
  const uint32_t p0 = (v.x & fm) * mult, p1 = (v.y & fm) * mult, p2 = (v.z & fm) * mult, p3 = (v.w & fm) * mult;
  return __byte_perm(__byte_perm(p0, p1, 0x0073), __byte_perm(p2, p3, 0x0073), 0x5410);
}

// The members that end at text offset e (exact): the eight bytes before e give the codes x16 (oldest in the low
// bits) and the run of live bytes that ends at e; the member mask of x16 (hash in shared memory, or the global
// 65536-entry table) is cut to the members no longer than that run.  The eight loads do not depend on each other:
// one round trip to L2.
__device__ __forceinline__ uint32_t KmerVerify(const uint8_t* __restrict__ text, uint64_t e, uint32_t fm, uint32_t mult,
                                               uint32_t shift, uint32_t canon, bool from_list,
                                               const uint32_t* __restrict__ mask16, const uint32_t* s_hkey,
                                               const uint32_t* s_hval, const uint32_t* s_lenle) {
  // w = text[e-8 .. e), byte 0 the oldest; `missing`: 0xFF in the bytes that lie before the text
  unsigned long long w, missing = 0;
  if (e >= 8) {
    const uint8_t* p = text + e - 8;
    const unsigned long long* a = reinterpret_cast<const unsigned long long*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)7);
    const unsigned long long w0 = __ldg(a), w1 = __ldg(a + 1);       // device texts are padded: a + 16 <= e + 8
    const uint32_t sh = ((uint32_t)reinterpret_cast<uintptr_t>(p) & 7u) * 8u;
    w = sh ? (w0 >> sh) | (w1 << (64u - sh)) : w0;
  } else {
    w = 0;
    for (uint32_t i = 0; i < 8; ++i) {
      if (e >= 8 - i) w |= (unsigned long long)__ldg(text + e - 8 + i) << (8 * i);
      else missing |= 0xFFull << (8 * i);
    }
  }
  const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
  const uint32_t x16 = (((lo & fm) * mult) >> 24) | ((((hi & fm) * mult) >> 24) << 8);   // codes, oldest in the low bits
  // the byte each code stands for (PRMT with the codes as selector nibbles) against the byte itself
  const uint32_t c0 = (lo >> shift) & 0x03030303u, c1 = (hi >> shift) & 0x03030303u;
  const uint32_t sel0 = ((c0 | (c0 >> 4)) & 0xFFu) | (((c0 >> 8) | (c0 >> 12)) & 0xFF00u);
  const uint32_t sel1 = ((c1 | (c1 >> 4)) & 0xFFu) | (((c1 >> 8) | (c1 >> 12)) & 0xFF00u);
  const unsigned long long diff =
      ((unsigned long long)(__byte_perm(canon, 0u, sel1) ^ hi) << 32 | (__byte_perm(canon, 0u, sel0) ^ lo)) | missing;
  const uint32_t v = diff ? (uint32_t)__clzll((long long)diff) >> 3 : 8u;     // live bytes counted back from e
  uint32_t members = 0;
  if (from_list) {
    const uint32_t key = x16 | 0x10000u;
    uint32_t slot = ((x16 * 0x9E3Bu) >> 5) & (kKmerHashSlots - 1u);
    for (uint32_t kk; (kk = s_hkey[slot]) != 0u; slot = (slot + 1u) & (kKmerHashSlots - 1u))
      if (kk == key) { members = s_hval[slot]; break; }
  } else {
    members = __ldg(mask16 + x16);
  }
  return members & s_lenle[v];
}

// Round 2: a warp owns a contiguous run of rows of ANY length.  It streams them until its list of groups with a
// hit is half full, leaves the streaming loop, checks those hits (exact check, per-member counts, overlap test) and
// keeps the checked matches {end, members} — in shared memory while they fit, else in a per-warp staging area in
// global memory — and goes on streaming.  Nothing but the final exchange is grid-wide, so the fixed costs (tables,
// exchange, report) are paid once per call whatever the text size.
__global__ void __launch_bounds__(kKmerThreads, 1)
k_set_kmer(const uint8_t* __restrict__ text, uint64_t n, int n_patterns, KmerTables km, ScanRange range, KmerRun run,
           CarrySet carries) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int R = kKmerEnds;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int K = n_patterns;
  // the set's small tables: requested before anything else, while the memory system is still idle
  const uint32_t tab_word = threadIdx.x < kKmerTabWords ? __ldg(km.tab + threadIdx.x) : 0u;
  // layout: [bitmap][per warp: hits u32 [64], groups with a hit u32 [64], checked matches uint2 [64]]
  //         [per warp and member: count, offset, first end, last end u32 [warp][32]][hash keys, values]
  //         [small tables][misc]
  uint32_t* s_bitmap = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* s_raw = s_bitmap + kKmerBitmapBytes / 4;
  uint32_t* s_ent = s_raw + kKmerWarps * kKmerWarpRaw;
  uint2* s_stage = reinterpret_cast<uint2*>(s_ent + kKmerWarps * kKmerWarpRaw);
  uint32_t* s_wcnt = reinterpret_cast<uint32_t*>(s_stage + kKmerWarps * kKmerStageSm);
  uint32_t* s_woff = s_wcnt + kKmerWarps * 32;
  uint32_t* s_wfirst = s_woff + kKmerWarps * 32;
  uint32_t* s_wlast = s_wfirst + kKmerWarps * 32;
  uint32_t* s_hkey = s_wlast + kKmerWarps * 32;
  uint32_t* s_hval = s_hkey + kKmerHashSlots;
  uint32_t* s_tab = s_hval + kKmerHashSlots;
  const uint32_t* s_mlen = s_tab + kKmerTabMatchLen;                                  // [32] match length per member
  const uint32_t* s_lenle = s_tab + kKmerTabLenLe;                                    // [9] members no longer than v
  uint32_t* s_misc = s_tab + kKmerTabWords;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_misc);                              // 8 bytes
  uint32_t* s_flags = s_misc + 2;
  uint32_t* s_base = s_misc + 4;                                                      // [32] matches of the CTAs before me
  uint32_t* s_count = s_base + 32;                                                    // [32] my matches per member
  unsigned long long* s_prevlast = reinterpret_cast<unsigned long long*>(s_count + 32);   // [32] last end before my CTA
  unsigned long long* s_cfirst = s_prevlast + 32;                                     // [32] my first / last end
  unsigned long long* s_clast = s_cfirst + 32;

  const bool from_list = km.n_list != 0;
  if (threadIdx.x == 0) {
    *s_flags = 0;
    if (!from_list) {
      MbarInit(s_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      MbarExpectTx(s_bar, kKmerBitmapBytes);
      TmaLoad1D(s_bitmap, km.bitmap, kKmerBitmapBytes, s_bar);
    }
  }
  if (threadIdx.x < 32) { s_base[threadIdx.x] = 0; s_count[threadIdx.x] = 0; s_prevlast[threadIdx.x] = 0; }

  // ---- my warp's rows -------------------------------------------------------------
  const uint64_t cta_row0 = run.row_lo + (uint64_t)blockIdx.x * run.rows_per_cta;
  const uint64_t cta_base = cta_row0 << 9;
  const uint64_t n16 = (n + 15) & ~15ull;                   // device texts are padded: whole groups can be read
  const uint64_t total_rows = (n16 + 511) >> 9;
  uint32_t cta_rows = 0;
  if (cta_row0 < total_rows)
    cta_rows = total_rows - cta_row0 < run.rows_per_cta ? (uint32_t)(total_rows - cta_row0) : run.rows_per_cta;
  const uint32_t w_row0 = (uint32_t)warp * run.rows_per_warp;
  const uint32_t w_row1 = w_row0 + run.rows_per_warp < cta_rows ? w_row0 + run.rows_per_warp : cta_rows;
  // rows [w_row0, w_load1) have my 16 bytes inside the text (only the text's last row can be cut)
  uint32_t w_load1 = w_row1;
  if (w_row1 > w_row0 && cta_base + ((uint64_t)(w_row1 - 1) << 9) + (uint64_t)lane * 16 >= n16) w_load1 = w_row1 - 1;
  const uint32_t fm = km.field_mask, mult = km.mult;
  const uint4* src = reinterpret_cast<const uint4*>(text + cta_base + ((uint64_t)w_row0 << 9)) + lane;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  // the 16 bytes before my warp's first row (lane 31's codes of "the row before"), then four rows in flight;
  // issued before the tables are built so that they arrive meanwhile
  const bool has_prev = w_row0 < w_row1 && cta_row0 + w_row0 > 0;
  const uint4 prev16 = has_prev ? __ldg(reinterpret_cast<const uint4*>(text + cta_base + ((uint64_t)w_row0 << 9) - 16)) : zero4;
  uint4 v0 = w_row0 < w_load1 ? __ldg(src) : zero4;
  uint4 v1 = w_row0 + 1 < w_load1 ? __ldg(src + 32) : zero4;
  uint4 v2 = w_row0 + 2 < w_load1 ? __ldg(src + 64) : zero4;
  uint4 v3 = w_row0 + 3 < w_load1 ? __ldg(src + 96) : zero4;
  if (from_list) {
    // the bitmap: index x = 7 + R codes; bit set iff one of its R 8-mers (x >> 2k) & 0xFFFF is on the list
    uint4* b4 = reinterpret_cast<uint4*>(s_bitmap);
#pragma unroll
    for (uint32_t i = 0; i < kKmerBitmapBytes / 16 / kKmerThreads; ++i) b4[i * kKmerThreads + threadIdx.x] = zero4;
    if (threadIdx.x < kKmerHashSlots) s_hkey[threadIdx.x] = 0;
  }
  if (threadIdx.x < kKmerTabWords) s_tab[threadIdx.x] = tab_word;
  __syncthreads();                                          // tables copied, bitmap zeroed / the barrier is initialised
  if (from_list) {
    // 8-mer a at codes [k, k + 8) of x: the k codes below it (fl) pick the word together with a's low bits, the
    // R-1-k codes above it (fh) only pick bits of that word: one atomic per (a, k, fl), 1 + 4 + .. + 4^(R-1) per a
    constexpr uint32_t kPerA = ((1u << (2 * R)) - 1u) / 3u;
    const uint32_t total = km.n_list * kPerA;
    for (uint32_t i = threadIdx.x; i < total; i += kKmerThreads) {
      const uint32_t a = s_tab[kKmerTabListX + i / kPerA];
      uint32_t rem = i % kPerA, k = 0;
      while (rem >= (1u << (2 * k))) { rem -= 1u << (2 * k); ++k; }
      const uint32_t fl = rem;
      const uint32_t xl = fl | (a << (2 * k));             // x without the codes above the 8-mer
      const uint32_t sh = 16 + 2 * k - kKmerWordBits;      // where those codes sit in x >> kKmerWordBits
      uint32_t bits = 0;
      for (uint32_t fh = 0; fh < (1u << (2 * (R - 1 - k))); ++fh) bits |= 1u << (31 - ((xl >> kKmerWordBits) | (fh << sh)));
      atomicOr(&s_bitmap[xl & ((1u << kKmerWordBits) - 1u)], bits);
    }
    if (threadIdx.x < km.n_list) {
      const uint32_t a = s_tab[kKmerTabListX + threadIdx.x];
      uint32_t slot = ((a * 0x9E3Bu) >> 5) & (kKmerHashSlots - 1u);
      while (atomicCAS(&s_hkey[slot], 0u, a | 0x10000u) != 0u) slot = (slot + 1u) & (kKmerHashSlots - 1u);
      s_hval[slot] = s_tab[kKmerTabListMask + threadIdx.x];
    }
    __syncthreads();
  } else {
    MbarWait(s_bar, 0);
  }
  uint32_t prevQ = has_prev ? KmerPack(prev16, fm, mult) : 0u;   // lane 31: codes of the 16 bytes before the row
  uint32_t* my_raw = s_raw + warp * kKmerWarpRaw;
  uint32_t* my_ent = s_ent + warp * kKmerWarpRaw;
  uint2* my_stage = s_stage + warp * kKmerStageSm;
  uint2* my_gstage = run.stage + ((size_t)blockIdx.x * kKmerWarps + warp) * run.stage_cap;
  uint32_t n_ent = 0;                                       // 16-byte groups of my warp with a hit so far (uniform)
  uint32_t n_sm = 0, n_gl = 0;                              // checked matches in my shared list / in the staging area
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int from = (lane + 31) & 31;
  unsigned int flags = 0;
  uint32_t cj = 0, firstj = 0, lastj = 0;                   // member `lane`: count, first end, last end (relative to the CTA)
  // one row: pack my 16 bytes, refill the register they came from with the row four ahead, look the 16 ends up
  const uint4* fetch = src + 4 * 32;                        // the row four ahead of the one being looked up
  auto row = [&](uint4& v, uint32_t r) {
    const uint32_t Q = KmerPack(v, fm, mult);
    v = r + 4 < w_load1 ? __ldg(fetch) : zero4;
    fetch += 32;
    // the codes before mine: my left neighbour's; lane 0 gets lane 31's of the row before
    const uint32_t P = __shfl_sync(kFullMask, lane == 31 ? prevQ : Q, from);
    prevQ = Q;
    uint32_t acc = 0;
#pragma unroll
    for (int t = 0; t < kKmerTests; ++t) {
      // bits [16 + 2x, ...) of (Q:P): 7 + R codes from bit 2 on = the ends after letters x .. x+R-1
      const int s0 = 16 + 2 * KmerTestX(t);
      const uint32_t w = s0 < 32 ? __funnelshift_r(P, Q, s0) : (Q >> (s0 - 32));
      const uint32_t word = *reinterpret_cast<const uint32_t*>(smem_raw + (w & ((4u << kKmerWordBits) - 4u)));   // (the bitmap is first)
      acc = __funnelshift_l(__funnelshift_l(0u, word, w >> (kKmerWordBits + 2)), acc, 1);   // bit kKmerTests-1-t: lookup t
    }
    // a row with hits (about one in four on regex-dna): one entry per lane with a hit, in lane = position order
    const uint32_t bal = __ballot_sync(kFullMask, acc != 0);
    if (bal) {
      const uint32_t at = n_ent + __popc(bal & lt_mask);
      if (acc && at < kKmerWarpRaw) my_ent[at] = (((r << 5) + lane) << 8) | acc;
      n_ent += __popc(bal);
    }
  };
  uint32_t r = w_row0;
  RJ_KMER_STAMP(0);
#pragma unroll 1
  for (;;) {
    // ---- stream until the list of groups with a hit is half full ------------------------------------
#pragma unroll 1
    for (; r + 4 <= w_row1 && n_ent < kKmerFlushAt; r += 4) { row(v0, r); row(v1, r + 1); row(v2, r + 2); row(v3, r + 3); }
    if (r + 4 > w_row1) {
      if (r < w_row1) row(v0, r);
      if (r + 1 < w_row1) row(v1, r + 1);
      if (r + 2 < w_row1) row(v2, r + 2);
      r = w_row1;
    }
    // (v0 .. v3 already hold the four rows that follow — every row() refilled its register — and stay in flight
    // while the hits are checked)
    // ---- my hits so far: exact check, per-member counts ---------------------------------------------
    if (n_ent > kKmerWarpRaw) { flags |= kFinDense; n_ent = 0; }
    __syncwarp();
    // groups -> hits, in position order: one entry of my_raw per lookup that hit
    uint32_t n_raw = 0;
    for (uint32_t base = 0; base < n_ent; base += 32) {
      const uint32_t ent = base + lane < n_ent ? my_ent[base + lane] : 0u;
      const uint32_t acc = ent & 0xFFu, pos16 = (ent >> 8) * 16u;     // bit kKmerTests-1-t: lookup t hit
      const uint32_t cnt = __popc(acc);
      const uint32_t incl = WarpInclusiveScan(cnt);
      uint32_t at = n_raw + incl - cnt;
      uint32_t x = __brev(acc) >> (32 - kKmerTests);                  // bit t: lookup t
      while (x) {
        const int t = __ffs(x) - 1;
        x &= x - 1;
        const int xt = t * R < 16 - R ? t * R : 16 - R;
        // first of the R ends (relative to the CTA's first byte) | the first end that is this lookup's own << 30
        if (at < kKmerWarpRaw) my_raw[at] = (pos16 + xt + 1u) | ((uint32_t)(t * R - xt) << 30);
        ++at;
      }
      n_raw += __shfl_sync(kFullMask, incl, 31);
    }
    if (n_raw > kKmerWarpRaw) { flags |= kFinDense; n_raw = 0; }
    n_ent = 0;
    __syncwarp();
    const uint32_t n_cand = R * n_raw;
    // candidate R h + k = end k of hit h.  Two batches of 32 are checked at a time (their text reads are in flight
    // together), then accounted for one after the other, in position order.
    for (uint32_t i0 = 0; i0 < n_cand; i0 += 64) {
      // room for 64 more checked matches in my shared list, else it moves to the staging area
      if (n_sm + 64 > kKmerStageSm) {
        for (uint32_t q = lane; q < n_sm; q += 32)
          if (n_gl + q < run.stage_cap) my_gstage[n_gl + q] = my_stage[q];
        n_gl += n_sm;
        n_sm = 0;
        __syncwarp();
      }
      uint32_t mm2[2] = {0, 0}, er2[2] = {0, 0};
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t i = i0 + 32 * half + lane;
        if (i < n_cand) {
          const uint32_t h = i / R, k = i - h * R;
          const uint32_t raw = my_raw[h];
          er2[half] = (raw & 0x3FFFFFFFu) + k;
          const uint64_t e = cta_base + er2[half];
          if (k >= (raw >> 30) && e <= n)
            mm2[half] = KmerVerify(text, e, fm, mult, km.shift, km.canon, from_list, km.mask16, s_hkey, s_hval, s_lenle);
        }
      }
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const uint32_t erel = half ? er2[1] : er2[0];
        uint32_t m = 0;
        for (uint32_t mm = half ? mm2[1] : mm2[0]; mm; mm &= mm - 1) {
          const int j = __ffs(mm) - 1;
          const uint64_t b = cta_base + erel - s_mlen[j];
          if (b >= range.own_begin && b < range.own_end) {
            m |= 1u << j;
            if (run.has_carry && b < carries.c[j].cur) flags |= kFinOverlap;    // the chain arriving from the left reaches past it
          }
        }
        // a candidate overlaps an earlier one of the same member iff that one ends less than a match length before
        // it: its predecessor in this pass, the last one of the passes before; other warps and CTAs: the seam checks
        for (uint32_t todo = __reduce_or_sync(kFullMask, m); todo; todo &= todo - 1) {
          const int j = __ffs(todo) - 1;
          const uint32_t bal = __ballot_sync(kFullMask, (m >> j) & 1u);
          const uint32_t L = s_mlen[j];
          const int lo = __ffs(bal) - 1, hi = 31 - __clz(bal);
          const uint32_t e_lo = __shfl_sync(kFullMask, erel, lo), e_hi = __shfl_sync(kFullMask, erel, hi);
          const uint32_t before = bal & ((1u << lane) - 1u);
          const int pl = before ? 31 - __clz(before) : lane;
          const uint32_t e_prev = __shfl_sync(kFullMask, erel, pl);
          if (((m >> j) & 1u) && before && e_prev + L > erel) flags |= kFinOverlap;
          if (lane == j) {
            if (cj && lastj + L > e_lo) flags |= kFinOverlap;
            if (!cj) firstj = e_lo;
            lastj = e_hi;
            cj += __popc(bal);
          }
        }
        const uint32_t balm = __ballot_sync(kFullMask, m != 0);
        if (m) my_stage[n_sm + __popc(balm & lt_mask)] = make_uint2(erel, m);
        n_sm += __popc(balm);
        __syncwarp();
      }
    }
    if (r >= w_row1) break;
  }
  RJ_KMER_STAMP(1);
  if (n_gl > run.stage_cap) flags |= kFinOverflow;
  s_wcnt[warp * 32 + lane] = cj;
  s_wfirst[warp * 32 + lane] = firstj;
  s_wlast[warp * 32 + lane] = lastj;
  if (flags) atomicOr(s_flags, flags);
  __syncthreads();

  // ---- warp j: member j over the 32 warps (lane = warp): offsets, seams, my CTA's record --------
  if (warp < K) {
    const int j = warp;
    const uint32_t L = s_mlen[j];
    const uint32_t c = s_wcnt[lane * 32 + j], fe = s_wfirst[lane * 32 + j], le = s_wlast[lane * 32 + j];
    const uint32_t incl = WarpInclusiveScan(c);
    s_woff[lane * 32 + j] = incl - c;
    const uint32_t total = __shfl_sync(kFullMask, incl, 31);
    // last end among the warps before me (ends grow with the warp index)
    uint32_t run_max = c ? le : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(kFullMask, run_max, d);
      if (lane >= d && o > run_max) run_max = o;
    }
    uint32_t before = __shfl_up_sync(kFullMask, run_max, 1);
    if (lane == 0) before = 0;
    const bool bad = c && before && before + L > fe;
    const unsigned has = __ballot_sync(kFullMask, c != 0);
    const bool any_bad = __any_sync(kFullMask, bad);
    const uint32_t cta_first = __shfl_sync(kFullMask, fe, has ? __ffs(has) - 1 : 0);
    const uint32_t cta_last = __shfl_sync(kFullMask, run_max, 31);
    if (lane == 0) {
      s_count[j] = total;
      const unsigned long long f64 = total ? cta_base + cta_first : 0ull, l64 = total ? cta_base + cta_last : 0ull;
      s_cfirst[j] = f64;
      s_clast[j] = l64;
      // every flag this CTA can raise is known now (the seams BETWEEN CTAs are checked by the top CTA, which reads
      // every record anyway): it travels in the record, nothing is left to say after the exchange
      const unsigned int fl = *s_flags | (any_bad ? kFinOverlap : 0u);
      if (any_bad) atomicOr(s_flags, kFinOverlap);
      volatile uint4* dst = reinterpret_cast<volatile uint4*>(run.xchg + (size_t)blockIdx.x * 32 + j);
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"((unsigned int)f64),
                   "r"((unsigned int)(f64 >> 32)), "r"(total), "r"(run.seq) : "memory");
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 1), "r"((unsigned int)l64),
                   "r"((unsigned int)(l64 >> 32)), "r"(fl), "r"(run.seq) : "memory");
    }
  }
  // nobody polls before this CTA's own records are out: warps spinning on system-scope loads keep the
  // load/store queue full and starved the publishing warps of their shared-memory and shuffle slots
  __syncthreads();
  // ---- exchange: the CTAs before me.  Warp w reads CTA w, w + 32, ... (lane j = member j), five CTAs' records
  // in flight at once: the number of matches before me.  The CTA with the highest index (the "top" CTA) also keeps
  // every record in shared memory — the bitmap's space is free now — for the seams and the report.
  static_assert(kKmerBitmapBytes >= 96 * 1024 + 160 * 32 * 4 && 160 * 32 * 16 <= 96 * 1024, "the top CTA's record copy lives in the bitmap's space");
  const bool top = blockIdx.x + 1 == gridDim.x;              // (the host launches at most 160 CTAs)
  uint4* s_rec = reinterpret_cast<uint4*>(smem_raw);                 // [CTA][32]: {first end lo, hi, last end lo, hi}, top CTA only
  uint32_t* s_rcnt = reinterpret_cast<uint32_t*>(smem_raw + 96 * 1024);   // [CTA][32]: matches
  {
    unsigned int seen_flags = 0;
    for (uint32_t c0 = warp; c0 < blockIdx.x; c0 += 160) {
      uint4 a[5], b[5];
      unsigned pending = 0;
      uint32_t sum = 0;
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        a[u] = b[u] = zero4;
        if (c0 + 32 * u < blockIdx.x && lane < K) pending |= 1u << u;
      }
      while (pending) {
#pragma unroll
        for (int u = 0; u < 5; ++u)
          if ((pending >> u) & 1u) {
            const volatile uint4* p = reinterpret_cast<const volatile uint4*>(run.xchg + (size_t)(c0 + 32 * u) * 32 + lane);
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[u].x), "=r"(a[u].y), "=r"(a[u].z), "=r"(a[u].w) : "l"(p) : "memory");
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b[u].x), "=r"(b[u].y), "=r"(b[u].z), "=r"(b[u].w) : "l"(p + 1) : "memory");
          }
#pragma unroll
        for (int u = 0; u < 5; ++u)
          if (((pending >> u) & 1u) && a[u].w == run.seq && b[u].w == run.seq) {
            pending &= ~(1u << u);
            sum += a[u].z;
            seen_flags |= b[u].z;
            if (top) {
              s_rec[(c0 + 32 * u) * 32 + lane] = make_uint4(a[u].x, a[u].y, b[u].x, b[u].y);
              s_rcnt[(c0 + 32 * u) * 32 + lane] = a[u].z;
            }
          }
      }
      if (sum) atomicAdd(&s_base[lane], sum);
    }
    if (top) {
      seen_flags = __reduce_or_sync(kFullMask, seen_flags);
      if (seen_flags && lane == 0) atomicOr(s_flags, seen_flags);
    }
  }
  __syncthreads();
  // ---- the top CTA: the seams between the CTAs (warp j = member j, 32 CTAs per step, in order), then the report ----
  if (top) {
    if (warp < K) {
      const int j = warp;
      const unsigned long long L = s_mlen[j];
      unsigned long long carry = 0;                         // last end of member j so far (0: none)
      bool bad = false;
      for (uint32_t cb = 0; cb < gridDim.x; cb += 32) {
        const uint32_t cc = cb + lane;
        unsigned long long fe = 0, le = 0;
        uint32_t cnt = 0;
        if (cc < blockIdx.x) {
          const uint4 r = s_rec[cc * 32 + j];
          cnt = s_rcnt[cc * 32 + j];
          fe = (unsigned long long)r.y << 32 | r.x;
          le = (unsigned long long)r.w << 32 | r.z;
        } else if (cc == blockIdx.x) {
          cnt = s_count[j]; fe = s_cfirst[j]; le = s_clast[j];
        }
        unsigned long long run_max = cnt ? le : 0ull;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned long long o = __shfl_up_sync(kFullMask, run_max, d);
          if (lane >= d && o > run_max) run_max = o;
        }
        unsigned long long before = __shfl_up_sync(kFullMask, run_max, 1);
        if (lane == 0 || before < carry) before = carry;
        if (cnt && before && before + L > fe) bad = true;
        const unsigned long long last = __shfl_sync(kFullMask, run_max, 31);
        if (last > carry) carry = last;
      }
      bad = __any_sync(kFullMask, bad);
      if (lane == 0) {
        s_prevlast[j] = carry;                              // (here: the last end of member j in the whole call)
        if (bad) atomicOr(s_flags, kFinOverlap);
      }
    }
    // (a named barrier for the K + 1 warps involved would do; the other warps only have their matches left to write)
  }
  // ---- my warp's matches, at their final place: first the ones that moved to the staging area, then my shared list
  {
    uint32_t done = 0;                                      // member `lane`: matches of my warp already written
    const uint32_t n_gl_ok = n_gl < run.stage_cap ? n_gl : run.stage_cap;
    const uint32_t n_all = n_gl_ok + n_sm;
    for (uint32_t i0 = 0; i0 < n_all; i0 += 32) {
      const uint32_t i = i0 + lane;
      uint2 ent = make_uint2(0, 0);
      if (i < n_gl_ok) ent = __ldcg(my_gstage + i);
      else if (i < n_all) ent = my_stage[i - n_gl_ok];
      const uint32_t m = ent.y;
      const uint64_t e = cta_base + ent.x;
      for (uint32_t todo = __reduce_or_sync(kFullMask, m); todo; todo &= todo - 1) {
        const int j = __ffs(todo) - 1;
        const uint32_t bal = __ballot_sync(kFullMask, (m >> j) & 1u);
        const uint32_t dj = __shfl_sync(kFullMask, done, j);
        if ((m >> j) & 1u) {
          const unsigned long long at = (unsigned long long)s_base[j] + s_woff[warp * 32 + j] + dj + __popc(bal & ((1u << lane) - 1u));
          if (at < run.out_cap)
            reinterpret_cast<ulonglong2*>(run.out_pairs + (uint64_t)j * 2 * run.out_stride)[at] =
                make_ulonglong2(e - s_mlen[j] + run.base_offset, e + run.base_offset);
        }
        if (lane == j) done += __popc(bal);
      }
    }
  }
  // ---- the top CTA reports: totals, last ends and flags are all known to it (its matches are on their way; the
  // host reads the pairs after the kernel has ended) ------------------------------------------------------------
  if (top) {
    __syncthreads();                                        // the seams above
    if (warp == 0) {
      const unsigned int flg = *s_flags;
      unsigned long long total = 0, le = 0;
      if (lane < K) {
        total = (unsigned long long)s_base[lane] + s_count[lane];
        le = s_prevlast[lane];
        run.gfinal[2 * lane] = total;
        run.gfinal[2 * lane + 1] = le;
      }
      if (lane == 0) run.gfinal[64] = flg;
      if (run.stitch.enabled) {
        // scan + stitch in one kernel: the chain states leave for the right neighbour's HBM now; a call that the
        // host will repeat (flags, output too small) is sent as invalid and sent again by the repeat
        unsigned long long cur = 0;
        uint32_t has = 0, invalid = flg ? 1u : 0u;
        if (lane < K) {
          if (total > run.out_cap) invalid = 1;
          if (total) { cur = le + run.base_offset; has = 1; }
        }
        invalid = __any_sync(kFullMask, invalid != 0) ? 1u : 0u;
        StitchExchangeWarp(run.stitch, K, cur, has, has, invalid);
      }
      if (lane < K) {
        volatile uint4* dst = reinterpret_cast<volatile uint4*>(run.host_records + lane);
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"((unsigned int)total),
                     "r"((unsigned int)(total >> 32)), "r"(flg), "r"(run.seq) : "memory");
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 1), "r"((unsigned int)le),
                     "r"((unsigned int)(le >> 32)), "r"(0u), "r"(run.seq) : "memory");
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 2), "r"((unsigned int)le),
                     "r"((unsigned int)(le >> 32)), "r"(0u), "r"(run.seq) : "memory");
      }
    }
  }
#ifdef RJ_KMER_PROBE
  if (run.probe && threadIdx.x == 0) {
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    run.probe[(size_t)gridDim.x * 64 + blockIdx.x] = t_;
  }
#endif
}

// ---------------------------------------------------------------------------
// K2 fallback: same automaton, 16-byte global loads, non-replicated table,
// unordered append (dense matches / tables too large for k_dfa_tma).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dfa_scan(const uint8_t* __restrict__ text, uint64_t n, DfaTables dfa, uint32_t stream_bytes,
           ScanRange range, CandBuf out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
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
    if (b <= range.own_begin || a >= range.own_end + L) continue;
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
        state = (p + i < b) ? st : state;
        peak = max(peak, state);
      }
      if (peak >= (uint32_t)acc) {
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
// K3: generic scan (ordered).  A warp owns 16 KB of start offsets; per 512-byte
// piece a lane holds 16 bytes (one vector load) and tests each of its 16 offsets
// against a 256-bit start filter — "can a match begin with this byte in this
// line context?" (the context bit sol comes from the previous byte, eol is a
// function of the byte itself, so two bitmaps suffice).  Only offsets that pass
// run the per-start NFA.  Empty-match patterns (^, x*) pass through the same
// filter (their bitmaps include every byte where the empty match is possible).
// ===========================================================================
struct GenFilter {
  uint32_t t[2][8];      // [sol][byte >> 5] bit (byte & 31)
};

__device__ __noinline__ void GenEmit(const uint8_t* __restrict__ text, uint64_t n, const NfaTables& nfa,
                                     const ScanRange& range, const SubStore& out, uint64_t sub, uint32_t& k,
                                     uint64_t my, uint32_t cand) {
  uint64_t ends[16];
  uint32_t valid = 0;
  uint32_t hh = cand;
  while (hh) {
    int j = __ffs(hh) - 1;
    hh &= hh - 1;
    uint64_t s = my + j;
    if (s < range.own_begin || s >= range.own_end || s > n) continue;
    uint64_t e = NfaRunAny(nfa, text, n, s);
    if (e != kNoMatch) { ends[j] = e; valid |= 1u << j; }
  }
  __syncwarp();
  uint32_t c = __popc(valid);
  uint32_t incl = WarpInclusiveScan(c);
  uint32_t total = __shfl_sync(kFullMask, incl, 31);
  uint32_t idx = k + incl - c;
  while (valid) {
    int j = __ffs(valid) - 1;
    valid &= valid - 1;
    if (idx < out.cap) {
      out.begin[sub * out.cap + idx] = my + j;
      out.end[sub * out.cap + idx] = ends[j];
    }
    ++idx;
  }
  __syncwarp();
  k += total;
}

__global__ void __launch_bounds__(256)
k_generic_scan(const uint8_t* __restrict__ text, uint64_t n, NfaTables nfa, GenFilter flt, ScanRange range,
               SubStore out, FinishArgs fin, Carry carry0) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  constexpr uint32_t kPieces = kGenSubBytes / 512;
  for (uint64_t sub = warp; sub < out.nsub; sub += nwarps) {
    const uint64_t sub_lo = sub * kGenSubBytes;
    uint32_t k = 0;
    if (sub_lo <= n && sub_lo + kGenSubBytes > range.own_begin && sub_lo < range.own_end) {
      for (uint32_t pc = 0; pc < kPieces; ++pc) {
        const uint64_t piece_lo = sub_lo + (uint64_t)pc * 512;
        if (piece_lo > n) break;
        const uint64_t my = piece_lo + (uint64_t)lane * 16;
        uint4 v = (my < n) ? LoadText16(text, n, my) : make_uint4(0, 0, 0, 0);
        // byte before my chunk: last byte of the previous lane, or of the previous piece
        uint32_t prev = __shfl_up_sync(kFullMask, v.w >> 24, 1);
        if (lane == 0) prev = (my == 0) ? 0x0Au : (uint32_t)text[my - 1];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t cand = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint32_t c = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
          uint32_t sol = (prev == 0x0Au || prev == 0x0Du) ? 1u : 0u;
          uint32_t bit = (flt.t[sol][c >> 5] >> (c & 31)) & 1u;
          cand |= bit << j;
          prev = c;
        }
        // offsets at or beyond n: only the offset n itself can still hold an (empty) match
        if (my + 16 > n) {
          uint32_t keep = (my >= n) ? 0u : ((1u << (uint32_t)(n - my)) - 1u);
          cand &= keep;
          if (n >= my && n < my + 16) {
            uint32_t pb = (n == 0) ? 0x0Au : (uint32_t)text[n - 1];
            int ctx = ((pb == 0x0Au || pb == 0x0Du) ? 1 : 0) | 2;
            if (nfa.accept_empty[nfa.has_anchor ? ctx : 0]) cand |= 1u << (uint32_t)(n - my);
          }
        }
        if (__any_sync(kFullMask, cand != 0)) GenEmit(text, n, nfa, range, out, sub, k, my, cand);
      }
    }
    if (lane == 0) {
      out.count[sub] = k;
      FinishNote(fin, 0, sub, k, false, out.cap);
    }
  }
  if (fin.enabled) {
    extern __shared__ __align__(16) uint32_t gen_scratch[];
    FinishOrdered(out, out.nsub, 1, fin, &carry0, gen_scratch, kFinScratchWords);
  }
}

// ===========================================================================
// K4: verify the window of possible starts in front of every needle hit
// (ordered).  A warp owns 8 consecutive hits; windows are clipped against the
// previous hit's window so that every start is tried once and in order.
// ===========================================================================
__global__ void __launch_bounds__(256)
k_window_verify(const uint8_t* __restrict__ text, uint64_t n, NfaTables nfa, DenseList hits, uint32_t lo,
                uint32_t hi, ScanRange range, SubStore out, FinishArgs fin, Carry carry0) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long nh = *hits.count;
  if (nh > hits.cap) nh = hits.cap;
  for (uint64_t sub = warp; sub < out.nsub; sub += nwarps) {
    uint32_t k = 0;
    for (uint32_t q = 0; q < kWinSubHits; ++q) {
      uint64_t hidx = sub * kWinSubHits + q;
      if (hidx >= nh) break;
      const uint64_t h = hits.begin[hidx];
      if (h < lo) continue;
      uint64_t s_max = h - lo;                               // inclusive
      uint64_t s_min = h >= hi ? h - hi : 0;
      if (hidx > 0) {
        uint64_t prev = hits.begin[hidx - 1];
        if (prev >= lo && prev - lo + 1 > s_min) s_min = prev - lo + 1;   // already covered
      }
      for (uint64_t base = s_min; base <= s_max; base += 32) {
        uint64_t s = base + lane;
        uint64_t e = kNoMatch;
        if (s <= s_max && s >= range.own_begin && s < range.own_end && s < n) {
          int ctx = nfa.has_anchor ? ContextAt(text, n, s) : 0;
          if (nfa.start_ok[ctx * 256 + text[s]]) e = NfaRunAny(nfa, text, n, s);
        }
        __syncwarp();
        EmitOrdered(out, sub, k, e != kNoMatch, s, e);
      }
    }
    if (lane == 0) {
      out.count[sub] = k;
      FinishNote(fin, 0, sub, k, false, out.cap);
    }
  }
  if (fin.enabled) {
    extern __shared__ __align__(16) uint32_t win_scratch[];
    FinishOrdered(out, out.nsub, 1, fin, &carry0, win_scratch, kFinScratchWords);
  }
}

// ===========================================================================
// MatchFull: one sequential run from offset 0 (a latency-bound sibling of the
// hot path).
// ===========================================================================
__global__ void k_match_full(const uint8_t* __restrict__ text, uint64_t n, NfaTables nfa,
                             PipelineStatus* status) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint64_t e = NfaRunAny(nfa, text, n, 0, /*full_only=*/true);
    status->full_result = (e == n) ? 1u : 0u;
  }
}

// ===========================================================================
// block-wide scans used by the one-CTA kernels (1024 threads)
// ===========================================================================
__device__ __forceinline__ uint32_t BlockExclusiveSum(uint32_t v, uint32_t* total, uint32_t* s_warp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = WarpInclusiveScan(v);
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t x = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
    uint32_t xi = WarpInclusiveScan(x);
    s_warp[lane] = xi - x;
    if (lane == 31) s_warp[32] = xi;
  }
  __syncthreads();
  uint32_t r = s_warp[wid] + incl - v;
  *total = s_warp[32];
  __syncthreads();
  return r;
}

__device__ __forceinline__ uint64_t BlockExclusiveMax(uint64_t v, uint64_t init, uint64_t* s_warp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t t = __shfl_up_sync(kFullMask, incl, d);
    if (lane >= d && t > incl) incl = t;
  }
  uint64_t excl = __shfl_up_sync(kFullMask, incl, 1);
  if (lane == 0) excl = 0;
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint64_t x = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
    uint64_t xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint64_t t = __shfl_up_sync(kFullMask, xi, d);
      if (lane >= d && t > xi) xi = t;
    }
    uint64_t xe = __shfl_up_sync(kFullMask, xi, 1);
    if (lane == 0) xe = 0;
    s_warp[lane] = xe;
  }
  __syncthreads();
  uint64_t r = s_warp[wid] > excl ? s_warp[wid] : excl;
  if (init > r) r = init;
  __syncthreads();
  return r;
}

// Concatenates the sub-regions' slot ranges into a dense (sorted) list.
// Returns false (uniformly) on overflow / dense marker; *m_out = total.
// Per batch of kGatherBatch x 1024 sub-regions: (1) all counts are loaded up
// front, (2) the block scans run back to back on registers, (3) every thread
// then issues the loads of all its entries before storing any of them, so the
// global-memory latency is paid about once per batch instead of once per entry.
// If `spec_out` is non-null the pairs are also written there (offset by
// `base_offset`) — the speculative output of k_resolve_ordered's fast path.
constexpr int kGatherBatch = 4;

__device__ __forceinline__ bool GatherSubStore(const SubStore& st, const DenseList& dense, PipelineStatus* status,
                                               uint32_t* s_warp, unsigned long long* m_out,
                                               unsigned long long gather_limit,
                                               uint64_t* spec_out = nullptr, uint64_t spec_cap = 0,
                                               uint64_t base_offset = 0) {
  __shared__ unsigned int s_flags[2];          // [0] max count, [1] dense marker
  __shared__ unsigned long long s_total;
  if (threadIdx.x < 2) s_flags[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  // pass 1: total and overflow markers only (the counts stay hot in L2)
  {
    unsigned long long mine = 0;
    unsigned int mx = 0, marker = 0;
    for (uint64_t blk0 = 0; blk0 < st.nsub; blk0 += 8ull * blockDim.x) {
      uint32_t cc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {                  // eight loads in flight per thread
        uint64_t sub = blk0 + (uint64_t)u * blockDim.x + threadIdx.x;
        cc[u] = (sub < st.nsub) ? st.count[sub] : 0u;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t c = cc[u];
        if (c == kLaneListOverflow) marker = 1;
        else { if (c > st.cap) { mx = max(mx, c); c = st.cap; } mine += c; }
      }
    }
    if (marker) atomicOr(&s_flags[1], 1u);
    if (mx) atomicMax(&s_flags[0], mx);
    atomicAdd(&s_total, mine);
    __syncthreads();
    *m_out = s_total;
    bool ok1 = true;
    if (s_flags[1]) { if (threadIdx.x == 0) status->dense = 1; ok1 = false; }
    if (s_flags[0]) { if (threadIdx.x == 0) { status->overflow = 1; status->need_cap = s_flags[0]; } ok1 = false; }
    if (s_total > dense.cap) { if (threadIdx.x == 0) status->overflow = 1; ok1 = false; }
    if (threadIdx.x == 0) *dense.count = s_total;
    if (!ok1 || s_total > gather_limit) return ok1;      // too many for one CTA: the host runs the multi-CTA gather
  }
  unsigned long long base = 0;
  for (uint64_t blk0 = 0; blk0 < st.nsub; blk0 += (uint64_t)kGatherBatch * blockDim.x) {
    uint32_t c[kGatherBatch];
    unsigned long long at[kGatherBatch];
#pragma unroll
    for (int u = 0; u < kGatherBatch; ++u) {
      uint64_t sub = blk0 + (uint64_t)u * blockDim.x + threadIdx.x;
      c[u] = (sub < st.nsub) ? st.count[sub] : 0u;
    }
#pragma unroll
    for (int u = 0; u < kGatherBatch; ++u) {
      uint32_t total;
      uint32_t off = BlockExclusiveSum(c[u], &total, s_warp);
      at[u] = base + off;
      base += total;
    }
    // first two entries of every sub-region: loads issued together
    uint64_t b0[kGatherBatch], e0[kGatherBatch], b1[kGatherBatch], e1[kGatherBatch];
#pragma unroll
    for (int u = 0; u < kGatherBatch; ++u) {
      uint64_t sub = blk0 + (uint64_t)u * blockDim.x + threadIdx.x;
      if (c[u] > 0) { b0[u] = st.begin[sub * st.cap]; e0[u] = st.end[sub * st.cap]; }
      if (c[u] > 1) { b1[u] = st.begin[sub * st.cap + 1]; e1[u] = st.end[sub * st.cap + 1]; }
    }
    auto put = [&](unsigned long long k, uint64_t bb, uint64_t ee) {
      if (k < dense.cap) { dense.begin[k] = bb; dense.end[k] = ee; }
      if (spec_out && k < spec_cap) { spec_out[2 * k] = bb + base_offset; spec_out[2 * k + 1] = ee + base_offset; }
    };
#pragma unroll
    for (int u = 0; u < kGatherBatch; ++u) {
      uint64_t sub = blk0 + (uint64_t)u * blockDim.x + threadIdx.x;
      if (c[u] > 0) put(at[u], b0[u], e0[u]);
      if (c[u] > 1) put(at[u] + 1, b1[u], e1[u]);
      for (uint32_t i = 2; i < c[u]; ++i) put(at[u] + i, st.begin[sub * st.cap + i], st.end[sub * st.cap + i]);
    }
  }
  __syncthreads();
  *m_out = base;
  return true;
}

// Stage boundary of the literal+window pipeline: dense list of needle hits.
__global__ void __launch_bounds__(512, 1)
k_gather_hits(SubStore st, DenseList dense, PipelineStatus* status, FinRecord* host_rec, unsigned int seq) {
  __shared__ uint32_t s_warp[33];
  unsigned long long m;
  GatherSubStore(st, dense, status, s_warp, &m, ~0ull);
  __syncthreads();
  if (threadIdx.x == 0) {
    status->n_hits = m;
    if (host_rec) {
      // the fused pipeline reads the hit stage's outcome from mapped memory (see FinRecord)
      const unsigned int flags = (status->overflow ? kFinOverflow : 0u) | (status->dense ? kFinDense : 0u);
      const unsigned int need_cap = status->need_cap;
      volatile uint4* dst = reinterpret_cast<volatile uint4*>(host_rec);
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"((unsigned int)m),
                   "r"((unsigned int)(m >> 32)), "r"(flags), "r"(seq) : "memory");
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 1), "r"(0u), "r"(0u), "r"(need_cap),
                   "r"(seq) : "memory");
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst + 2), "r"(0u), "r"(0u), "r"(0u),
                   "r"(seq) : "memory");
    }
  }
}

__device__ __forceinline__ bool IsRestart(const uint64_t* b, const uint64_t* e, const uint64_t* reach, uint64_t i) {
  return reach[i] < b[i] || (reach[i] == b[i] && e[i] > b[i]);
}

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

// ===========================================================================
// Resolve (ordered input), one CTA of 1024 threads, everything parallel:
//   gather -> reach = exclusive max of ends -> restart points -> every restart
//   point walks its segment (ChainTake; label replay for re-entrant patterns)
//   -> exclusive sum of take flags -> pairs out.
// Restart rule: candidate i restarts the chain when no earlier candidate can
// influence it: reach[i] < begin[i], or reach[i] == begin[i] and it is non-empty
// (strictly smaller only, for re-entrant patterns).
// ===========================================================================
__device__ __forceinline__ void ResolveOrderedBody(const SubStore& st, const DenseList& dense, const ResolveScratch& rs,
                                                   const Carry& carry_in, uint64_t base_offset,
                                                   uint64_t* __restrict__ out_pairs, uint64_t out_cap,
                                                   const FaithfulArgs& fa, PipelineStatus* status) {
  __shared__ uint32_t s_warp[33];
  __shared__ uint64_t s_warp64[32];
  __shared__ unsigned long long s_last[2];
  unsigned long long m;
  // the gather also writes the candidates straight to the output: if the fast
  // path below holds they ARE the matches; otherwise the general path
  // overwrites the output
  bool ok = GatherSubStore(st, dense, status, s_warp, &m, (unsigned long long)kOrderedResolveMax,
                           fa.enabled ? nullptr : out_pairs, out_cap, base_offset);
  if (threadIdx.x == 0) status->n_candidates = m;
  if (!ok) return;
  if (m > (unsigned long long)kOrderedResolveMax) {
    if (threadIdx.x == 0) status->need_large = 1;
    return;
  }
  if (threadIdx.x < 2) s_last[threadIdx.x] = 0;
  const uint32_t M = (uint32_t)m;
  const uint32_t ipt = (M + blockDim.x - 1) / blockDim.x;      // items per thread (blocked)
  const uint32_t i0 = min(M, threadIdx.x * ipt), i1 = min(M, i0 + ipt);
  const uint64_t* b = dense.begin;
  const uint64_t* e = dense.end;
  if (!fa.enabled) {
    // Fast path: when no candidate begins before its predecessor ends (and none
    // is empty), every candidate restarts the chain, i.e. the matches ARE the
    // candidates.  This is the common case (sparse, fixed-length matches).
    int bad = 0;
    for (uint32_t i = i0; i < i1; ++i) {
      uint64_t prev_end = (i == 0) ? carry_in.cur : e[i - 1];
      bad |= (prev_end > b[i]) | (e[i] <= b[i]);
      if (i + 1 < M && e[i] > e[i + 1]) bad = 1;               // ends not monotone: need the max-scan
    }
    if (!__syncthreads_or(bad)) {
      if (threadIdx.x == 0) {
        status->n_matches = M;
        status->carry_cur = M ? e[M - 1] : carry_in.cur;
        status->carry_tail = M ? e[M - 1] : carry_in.tail;
      }
      return;
    }
  }
  // reach
  uint64_t local = 0;
  for (uint32_t i = i0; i < i1; ++i) local = e[i] > local ? e[i] : local;
  uint64_t run = BlockExclusiveMax(local, carry_in.cur, s_warp64);
  for (uint32_t i = i0; i < i1; ++i) { rs.reach[i] = run; run = e[i] > run ? e[i] : run; }
  __syncthreads();
  // chains
  for (uint32_t i = i0; i < i1; ++i) {
    bool head = fa.enabled ? (i == 0 || rs.reach[i] < b[i]) : (i == 0 || IsRestart(b, e, rs.reach, i));
    if (!head) continue;
    if (fa.enabled) {
      uint32_t j = i + 1;
      while (j < M && !(rs.reach[j] < b[j])) ++j;
      FaithfulSegment(fa.nfa, fa.text, fa.n, b, e, i, j, WalkerScratch(fa, threadIdx.x), rs.take, rs.fin_end);
    } else {
      ChainState cs;
      if (i == 0) { cs.cur = carry_in.cur; cs.tail = carry_in.tail; }
      else { cs.cur = 0; cs.tail = kNoMatch; }
      for (uint32_t j = i; j < M; ++j) {
        if (j > i && IsRestart(b, e, rs.reach, j)) break;
        rs.take[j] = ChainTake(&cs, b[j], e[j]) ? 1u : 0u;
        rs.fin_end[j] = e[j];
      }
    }
  }
  __syncthreads();
  // compaction
  uint32_t mine = 0;
  for (uint32_t i = i0; i < i1; ++i) mine += rs.take[i];
  uint32_t total;
  uint32_t off = BlockExclusiveSum(mine, &total, s_warp);
  for (uint32_t i = i0; i < i1; ++i) {
    if (!rs.take[i]) continue;
    if (off < out_cap) {
      out_pairs[2 * (uint64_t)off] = b[i] + base_offset;
      out_pairs[2 * (uint64_t)off + 1] = rs.fin_end[i] + base_offset;
    }
    ++off;
    atomicMax(&s_last[0], (unsigned long long)i + 1);
    if (rs.fin_end[i] > b[i]) atomicMax(&s_last[1], (unsigned long long)i + 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    status->n_matches = total;
    uint64_t cur = carry_in.cur, tail = carry_in.tail;
    if (s_last[0]) {
      uint64_t i = s_last[0] - 1;
      cur = rs.fin_end[i] > b[i] ? rs.fin_end[i] : b[i] + 1;
    }
    if (s_last[1]) tail = rs.fin_end[s_last[1] - 1];
    status->carry_cur = cur;
    status->carry_tail = tail;
  }
}

// The kernel proper: resolve, then publish the status block to mapped host
// memory (the host spins on `seq` instead of paying a stream synchronisation).
__global__ void __launch_bounds__(512, 1)
k_resolve_ordered(SubStore st, DenseList dense, ResolveScratch rs, Carry carry_in, uint64_t base_offset,
                  uint64_t* __restrict__ out_pairs, uint64_t out_cap, FaithfulArgs fa, PipelineStatus* status,
                  volatile PipelineStatus* host_status, unsigned int seq) {
  ResolveOrderedBody(st, dense, rs, carry_in, base_offset, out_pairs, out_cap, fa, status);
  __syncthreads();
  if (threadIdx.x == 0 && host_status) {
    __threadfence();
    PipelineStatus v = *status;
    host_status->n_candidates = v.n_candidates;
    host_status->n_hits = v.n_hits;
    host_status->n_matches = v.n_matches;
    host_status->carry_cur = v.carry_cur;
    host_status->carry_tail = v.carry_tail;
    host_status->overflow = v.overflow;
    host_status->need_cap = v.need_cap;
    host_status->need_large = v.need_large;
    host_status->dense = v.dense;
    host_status->full_result = v.full_result;
    __threadfence_system();
    host_status->seq = seq;
    __threadfence_system();
  }
}

// One CTA per pattern of a fused set: the same resolve, on that pattern's slice
// of the stores / scratch / output / status arrays.
__global__ void __launch_bounds__(512, 1)
k_resolve_set(SubStore st, uint64_t nsub_pat, DenseList dense, ResolveScratch rs, uint64_t per_cap,
              uint64_t* __restrict__ out_pairs, PipelineStatus* status, volatile PipelineStatus* host_status,
              unsigned long long* dense_counts, unsigned int seq, CarrySet carries, uint64_t base_offset) {
  const uint64_t j = blockIdx.x;
  SubStore mine = st;
  mine.begin += j * nsub_pat * st.cap;
  mine.end += j * nsub_pat * st.cap;
  mine.count += j * nsub_pat;
  mine.nsub = nsub_pat;
  DenseList d = dense;
  d.begin += j * per_cap;
  d.end += j * per_cap;
  d.count = dense_counts + j;
  d.cap = per_cap;
  ResolveScratch r = rs;
  r.reach += j * per_cap;
  r.take += j * per_cap;
  r.fin_end += j * per_cap;
  r.slot += j * per_cap;
  FaithfulArgs fa{};
  fa.enabled = 0;
  const Carry carry = carries.c[j];
  PipelineStatus* my_status = status + j;
  ResolveOrderedBody(mine, d, r, carry, base_offset, out_pairs + j * 2 * per_cap, per_cap, fa, my_status);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    PipelineStatus v = *my_status;
    volatile PipelineStatus* h = host_status + j;
    h->n_candidates = v.n_candidates;
    h->n_matches = v.n_matches;
    h->carry_cur = v.carry_cur;
    h->carry_tail = v.carry_tail;
    h->overflow = v.overflow;
    h->need_cap = v.need_cap;
    h->need_large = v.need_large;
    h->dense = v.dense;
    __threadfence_system();
    h->seq = seq;
    __threadfence_system();
  }
}

// ===========================================================================
// Resolve, sort-based fallback (unordered candidates from k_dfa_scan): one CTA
// sorts by begin (bitonic, shared memory) and walks the chain.
// ===========================================================================
__global__ void __launch_bounds__(1024)
k_resolve_small(CandBuf cand, Carry carry_in, uint64_t base_offset, uint64_t* __restrict__ out_pairs,
                uint64_t out_cap, PipelineStatus* status) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
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
  if (threadIdx.x == 0) {
    ChainState st{carry_in.cur, carry_in.tail};
    unsigned long long taken = 0;
    uint64_t prev_b = ~0ull;
    for (int i = 0; i < count; ++i) {
      uint64_t b = kb[i], e = ke[i];
      if (b == prev_b) continue;
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

// Multi-CTA gather for large candidate counts: offsets = exclusive sum of the
// sub-region counts (computed by cub::DeviceScan), one lane per sub-region.
__global__ void k_clamp_counts(const uint32_t* __restrict__ count, uint32_t cap, uint64_t nsub, uint64_t* __restrict__ wide) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = tid; i < nsub; i += nthreads) {
    uint32_t c = count[i];
    wide[i] = (c == kLaneListOverflow) ? 0 : (c > cap ? cap : c);
  }
}

// One warp per sub-region: the slot range is copied with coalesced accesses.
__global__ void k_gather_multi(SubStore st, const uint64_t* __restrict__ offsets, DenseList dense) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t sub = warp; sub < st.nsub; sub += nwarps) {
    uint32_t c = st.count[sub];
    if (c == kLaneListOverflow) c = 0;
    if (c > st.cap) c = st.cap;
    const uint64_t at = offsets[sub];
    for (uint32_t i = lane; i < c; i += 32) {
      if (at + i < dense.cap) {
        dense.begin[at + i] = st.begin[sub * st.cap + i];
        dense.end[at + i] = st.end[sub * st.cap + i];
      }
    }
  }
}

// ===========================================================================
// Large path (multi-CTA) on a dense, sorted candidate list.
// ===========================================================================
struct MaxOp {
  __host__ __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a > b ? a : b; }
};

__global__ void k_segment_chain(const uint64_t* __restrict__ b, const uint64_t* __restrict__ e,
                                const uint64_t* __restrict__ reach, uint64_t m, Carry carry_in,
                                uint32_t* __restrict__ take, uint64_t* __restrict__ fin_end) {
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
      fin_end[j] = e[j];
      if (b[j] == prev_b) { take[j] = 0; continue; }
      prev_b = b[j];
      take[j] = ChainTake(&st, b[j], e[j]) ? 1u : 0u;
    }
  }
}

__global__ void k_segment_faithful(const uint64_t* __restrict__ b, const uint64_t* __restrict__ e,
                                   const uint64_t* __restrict__ reach, uint64_t m, FaithfulArgs fa,
                                   uint32_t* take, uint64_t* fin_end) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  FaithfulScratch sc = WalkerScratch(fa, tid);
  for (uint64_t i = tid; i < m; i += nthreads) {
    bool head = (i == 0) || (reach[i] < b[i]);
    if (!head) continue;
    uint64_t j = i + 1;
    while (j < m && !(reach[j] < b[j])) ++j;
    FaithfulSegment(fa.nfa, fa.text, fa.n, b, e, i, j, sc, take, fin_end);
  }
}

__global__ void k_scatter_matches(const uint64_t* __restrict__ b, const uint64_t* __restrict__ e,
                                  const uint32_t* __restrict__ take, const uint64_t* __restrict__ slot,
                                  uint64_t m, uint64_t base_offset, uint64_t* __restrict__ out_pairs,
                                  uint64_t out_cap, unsigned long long* last_any,
                                  unsigned long long* last_nonempty) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // the last taken candidate is the one whose slot is total-1 (one writer, no
  // atomic); the last NON-EMPTY one needs a maximum, reduced per warp first
  const uint64_t total = m ? slot[m - 1] + take[m - 1] : 0;
  unsigned long long best = 0;
  for (uint64_t i0 = tid - lane; i0 < m; i0 += nthreads) {
    const uint64_t i = i0 + lane;
    if (i < m && take[i]) {
      const uint64_t at = slot[i];
      const uint64_t bi = b[i], ei = e[i];
      if (at < out_cap) {
        out_pairs[2 * at] = bi + base_offset;
        out_pairs[2 * at + 1] = ei + base_offset;
      }
      if (at + 1 == total) *last_any = (unsigned long long)(i + 1);
      if (ei > bi) best = (unsigned long long)(i + 1);      // i grows along the loop
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(kFullMask, best, d);
    best = o > best ? o : best;
  }
  if (lane == 0 && best) atomicMax(last_nonempty, best);
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
// ReplaceAll on the device (reference: Regej::ReplaceAll
// src/rejit.cc:221-226 and Replace :97-112): the matches are already in device
// memory, so the rebuilt text is a prefix sum plus a gather.
//   removed[i] = sum of the lengths of the matches before match i (CUB scan);
// an input byte p outside every match lands at
//   p - (removed bytes before p) + with_len * (matches beginning at or before p)
// and the replacement of match i right before the byte that follows it.
// One CTA per 4 KB input tile: two binary searches find the tile's matches,
// their begins/ends/prefixes are staged in shared memory, every thread then
// places its own 16 input bytes (replace.cuh: k_replace_stage).
// ===========================================================================
__global__ void k_match_lengths(const uint64_t* __restrict__ pairs, uint64_t m, uint64_t* __restrict__ len) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = tid; i < m; i += nthreads) len[i] = pairs[2 * i + 1] - pairs[2 * i];
}

}  // namespace rejit_b200

#endif  // REJIT_B200_CUDA_KERNELS_CUH_
