// rejit_b200 — device-side stitch of slab-sharded texts (one process per GPU).
//
// Every rank sends the chain state that leaves its slab — per pattern: where the next match may begin — straight
// into the right neighbour's device memory (a peer store over NVLink; the inbox is mapped through CUDA IPC) and
// checks the state that arrives from the left: only when that chain reaches into the slab does the host repeat the
// slab call with the real carry.  No host-to-host hop, no collective.  Two users of the same device function:
//   k_set_kmer   the CTA that reports the call's result does the exchange itself, before it reports (the step's
//                device time then contains the neighbour's answer: one kernel, scan + stitch)
//   k_stitch     one warp on the engine's stream, for every other scan path (the host hands it the carry)
//   inbox slot (step & 63): [32] uint4 {cur lo, cur hi | ne << 31 | has << 30, step, invalid}
//   a step whose scan must be repeated (a buffer was too small) is sent as `invalid` and sent again afterwards
//   flow control: a rank is at most 32 steps ahead of its right neighbour (ack word written back by the receiver)
#ifndef REJIT_B200_CUDA_STITCH_CUH_
#define REJIT_B200_CUDA_STITCH_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace rejit_b200 {

constexpr uint32_t kStitchSlots = 64;
constexpr uint32_t kStitchAckWord = kStitchSlots * 32 * 4;               // index (in words) of the ack word
constexpr uint32_t kStitchInboxBytes = kStitchSlots * 32 * 16 + 64;

// Mapped host memory, one per device context.  Every 16-byte record carries the step it belongs to and is written with
// ONE store, so the host needs no ordering between them and the kernel no system-scope fence (two of those — one after
// the peer store, one before the report — were 3-4 us of every step).
struct StitchReport {
  uint4 rec[32];                        // [j] = {cur lo, cur hi, step, 0}: global offset where the chain arriving from the
                                        // left lets a match of pattern j begin (0: none)
  uint4 head;                           // {redo, status, arrived_ne, step}: bit j of redo: that chain reaches into the slab;
                                        // status 1: sent as invalid (send again), 2: a neighbour did not answer;
                                        // bit j of arrived_ne: that chain's last match was non-empty
};

struct StitchLink {                     // where to send, where to listen (kernel parameter)
  int enabled, rank;
  uint64_t slab_begin;                  // first owned start, global
  uint4* inbox;                         // mine
  uint4* right_inbox;                   // the right neighbour's (NULL on the last rank)
  unsigned int* left_ack;               // the ack word in the left neighbour's inbox (NULL on rank 0)
  unsigned int step;
  StitchReport* report;
};

__device__ __forceinline__ void StitchStore16(uint4* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 StitchLoad16(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// Called by one whole warp; lane j < K carries pattern j's leaving state (cur: global offset, 0 with has == 0).
__device__ __forceinline__ void StitchExchangeWarp(const StitchLink& a, int K, unsigned long long cur, uint32_t ne, uint32_t has,
                                                   uint32_t invalid) {
  const int lane = threadIdx.x & 31;
  uint32_t status = invalid ? 1u : 0u;
  const uint32_t slot = (a.step & (kStitchSlots - 1)) * 32;
  // ---- send to the right (not more than 32 steps ahead of what the neighbour has consumed) ------------
  if (a.right_inbox) {
    const volatile unsigned int* ack = reinterpret_cast<const volatile unsigned int*>(a.inbox) + kStitchAckWord;
    for (uint32_t polls = 0; (int)(a.step - *ack) > 32; ++polls)
      if (polls > (1u << 22)) { status |= 2u; break; }
    if (lane < K)
      StitchStore16(a.right_inbox + slot + lane, (uint32_t)cur, (uint32_t)(cur >> 32) | (ne << 31) | (has << 30), a.step, invalid);
    // (no fence: a record is one 16-byte store that carries its own step number)
  }
  // ---- what arrives from the left (an `invalid` record is followed by a valid one for the same step) ------
  unsigned long long arr = 0;
  uint32_t arr_ne = 0, arr_has = 0;
  if (a.rank > 0) {
    if (lane < K) {
      for (uint32_t polls = 0;; ++polls) {
        const uint4 v = StitchLoad16(a.inbox + slot + lane);
        if (v.z == a.step && v.w == 0) {
          arr = (unsigned long long)(v.y & 0x3FFFFFFFu) << 32 | v.x;
          arr_ne = v.y >> 31;
          arr_has = (v.y >> 30) & 1u;
          break;
        }
        if (polls > (1u << 23)) { status |= 2u; break; }
        const long long t0 = clock64();
        while (clock64() - t0 < 128) {}
      }
    }
    __syncwarp();
    if (lane == 0 && a.left_ack) *reinterpret_cast<volatile unsigned int*>(a.left_ack) = a.step;
  }
  const bool redo = arr_has && (arr > a.slab_begin || (arr_ne && arr == a.slab_begin));
  const uint32_t redo_mask = __ballot_sync(0xFFFFFFFFu, redo), arr_ne_mask = __ballot_sync(0xFFFFFFFFu, arr_ne != 0);
  status = __reduce_or_sync(0xFFFFFFFFu, status);
  const unsigned long long rep = arr_has ? arr : 0ull;
  StitchStore16(a.report->rec + lane, (uint32_t)rep, (uint32_t)(rep >> 32), a.step, 0u);
  if (lane == 0) StitchStore16(&a.report->head, redo_mask, status, arr_ne_mask, a.step);
}

struct StitchArgs {                     // k_stitch: the host hands over the leaving states
  int K;
  unsigned long long sent_cur[32];      // global offsets
  unsigned int sent_ne, sent_has;
  StitchLink link;
};

__global__ void __launch_bounds__(32, 1) k_stitch(StitchArgs a) {
  const int lane = threadIdx.x;
  StitchExchangeWarp(a.link, a.K, lane < a.K ? a.sent_cur[lane] : 0ull, (a.sent_ne >> lane) & 1u, (a.sent_has >> lane) & 1u, 0u);
}

}  // namespace rejit_b200

#endif  // REJIT_B200_CUDA_STITCH_CUH_
