// rejit_b200 — ReplaceAll on the device, round 2 (reference: Regej::ReplaceAll,
// /root/reference/src/rejit.cc:221-226, and Replace, :97-112).
//
//   k_replace_stage   the rebuild of round 1 (k_replace_tiles: prefix sum over the match lengths, then every
//                     4 KB input tile places its own bytes and replacements) with both ends staged in shared
//                     memory: the tile is fetched with coalesced 16-byte loads, its output is assembled in
//                     shared memory and leaves with coalesced 16-byte stores.  Round 1 read and wrote single
//                     bytes from global memory: 0.2 TB/s.
//   k_translate_*     a SET of patterns that each match exactly one byte (regex-dna's eleven IUB codes,
//                     /root/reference/sample/regexdna.cc:69-85) is one byte -> string table: one counting pass,
//                     one prefix sum over the tiles, one writing pass — instead of eleven scan + rebuild passes.
//                     Sequential ReplaceAll calls and the table give the same text iff no replacement holds a
//                     byte that a LATER pattern matches (checked on the host, engine.cu).
// Algorithmic traffic: rebuild N + 16 M read, N' written; translate 2 N read (the second time from L2 when the
// text fits it), N' written.
#ifndef REJIT_B200_CUDA_REPLACE_CUH_
#define REJIT_B200_CUDA_REPLACE_CUH_

#include "kernels.cuh"

namespace rejit_b200 {

constexpr uint32_t kRepStageBytes = 12288;          // output of one 4 KB tile that is assembled in shared memory

// coalesced copy of `len` bytes staged at s_out[(dst address) & 15 ...] to dst (whole CTA)
__device__ __forceinline__ void StageCopyOut(uint8_t* __restrict__ dst, const uint8_t* s_out, uint32_t len) {
  const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);       // s_out + a holds dst[0]
  uint32_t head = (16u - a) & 15u;
  if (head > len) head = len;
  if (threadIdx.x < head) dst[threadIdx.x] = s_out[a + threadIdx.x];
  const uint32_t body = (len - head) >> 4;
  const uint4* sv = reinterpret_cast<const uint4*>(s_out + a + head);
  uint4* dv = reinterpret_cast<uint4*>(dst + head);
  for (uint32_t i = threadIdx.x; i < body; i += blockDim.x) dv[i] = sv[i];
  const uint32_t done = head + (body << 4);
  if (threadIdx.x < len - done) dst[done + threadIdx.x] = s_out[a + done + threadIdx.x];
}

// ReplacePlace (device_program.h) with the tile's input and output in shared memory: s_in[pos] = text[tile_lo + pos],
// the output byte with global index q goes to s_out[q - out_base + align].
__device__ __forceinline__ void ReplacePlaceStaged(uint32_t thread, const uint8_t* s_in, uint32_t span, bool last,
                                                   uint32_t head_skip, uint32_t cnt, const uint16_t* s_b, const uint16_t* s_e,
                                                   const uint16_t* s_r, const uint8_t* __restrict__ with, uint32_t w,
                                                   uint8_t* s_out) {
  const uint32_t p = thread * 16;
  if (p > span) return;
  uint32_t lo = 0, hi = cnt;                // i = own matches beginning before p
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (s_b[mid] < p) lo = mid + 1; else hi = mid; }
  uint32_t i = lo;
  uint32_t rem = head_skip < p ? head_skip : p;        // removed inside the tile before p
  uint32_t cur_skip = head_skip;
  if (i > 0) {
    const uint32_t pe = s_e[i - 1];
    rem += s_r[i - 1] + ((pe < p ? pe : p) - s_b[i - 1]);
    if (pe > cur_skip) cur_skip = pe;
  }
  uint32_t q = p - rem + w * i;
  const uint32_t stop = (p + 16 < span) ? p + 16 : span + ((last && p + 16 > span) ? 1u : 0u);
  for (uint32_t pos = p; pos < stop; ++pos) {
    if (i < cnt && s_b[i] == pos) {
      for (uint32_t k = 0; k < w; ++k) s_out[q + k] = __ldg(with + k);
      q += w;
      if (s_e[i] > cur_skip) cur_skip = s_e[i];
      ++i;
    }
    if (pos < span && pos >= cur_skip) s_out[q++] = s_in[pos];
  }
}

__global__ void __launch_bounds__(256)
k_replace_stage(const uint8_t* __restrict__ text, uint64_t n, const uint64_t* __restrict__ pairs,
                const uint64_t* __restrict__ removed, uint64_t m, const uint8_t* __restrict__ with, uint32_t w,
                uint8_t* __restrict__ out, uint64_t n_tiles) {
  __shared__ uint16_t s_b[kReplaceTile + 2];        // begin - tile_lo of the tile's own matches
  __shared__ uint16_t s_e[kReplaceTile + 2];        // end - tile_lo, clipped to the tile
  __shared__ uint16_t s_r[kReplaceTile + 2];        // removed[m0 + i] - removed[m0]
  __shared__ __align__(16) uint8_t s_in[kReplaceTile + 16];
  __shared__ __align__(16) uint8_t s_out[kRepStageBytes + 32];
  __shared__ ReplaceTileHead s_head, s_next;
  const uint64_t n16 = (n + 15) & ~15ull;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t tile_lo = tile * kReplaceTile;
    const bool last = tile + 1 == n_tiles;
    const uint64_t tile_hi = last ? n : tile_lo + kReplaceTile;
    // the tile's bytes: one coalesced 16-byte load per thread (device texts are padded to whole groups)
    {
      const uint64_t at = tile_lo + (uint64_t)threadIdx.x * 16;
      const uint4 v = at < n16 ? __ldg(reinterpret_cast<const uint4*>(text + at)) : make_uint4(0, 0, 0, 0);
      reinterpret_cast<uint4*>(s_in)[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) s_head.m0 = ReplaceLowerBound(pairs, m, tile_lo);
    if (threadIdx.x == 32) s_next.m0 = last ? m : ReplaceLowerBound(pairs, m, tile_hi);
    __syncthreads();
    if (threadIdx.x == 0) { s_head.m1 = s_next.m0; ReplaceHead(pairs, removed, m, tile_lo, &s_head); }
    if (threadIdx.x == 32) ReplaceHead(pairs, removed, m, tile_hi, &s_next);
    __syncthreads();
    const ReplaceTileHead h = s_head;
    const uint32_t cnt = (uint32_t)(h.m1 - h.m0);
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
      const uint64_t b = pairs[2 * (h.m0 + i)], e = pairs[2 * (h.m0 + i) + 1];
      s_b[i] = (uint16_t)(b - tile_lo);
      s_e[i] = (uint16_t)((e < tile_hi ? e : tile_hi) - tile_lo);
      s_r[i] = (uint16_t)(removed[h.m0 + i] - h.r0);
    }
    const uint64_t out_base = tile_lo - h.removed_before + (uint64_t)w * h.m0;
    const uint64_t out_next = tile_hi - s_next.removed_before + (uint64_t)w * s_next.m0;     // where the next tile begins
    const uint64_t out_len = out_next - out_base;
    __syncthreads();
    if (out_len <= kRepStageBytes) {
      const uint32_t span = (uint32_t)(tile_hi - tile_lo);
      const uint32_t head_skip = (uint32_t)((h.skip_end < tile_hi ? h.skip_end : tile_hi) - tile_lo);
      const uint32_t align = (uint32_t)(reinterpret_cast<uintptr_t>(out + out_base) & 15u);
      ReplacePlaceStaged(threadIdx.x, s_in, span, last, head_skip, cnt, s_b, s_e, s_r, with, w, s_out + align);
      __syncthreads();
      StageCopyOut(out + out_base, s_out, (uint32_t)out_len);
    } else {
      // many long replacements in one tile: straight to global memory, as in round 1
      ReplacePlace(threadIdx.x, text, tile_lo, tile_hi, last, h, cnt, s_b, s_e, s_r, with, w, out);
    }
    __syncthreads();
  }
}

// ===========================================================================
// byte -> string table (a set of one-byte patterns)
// ===========================================================================
constexpr uint32_t kTransTile = 4096;
constexpr uint32_t kTransStage = 4096 * 8;          // a tile whose output is longer goes straight to global memory
constexpr uint32_t kTransMaxBytes = 4096;           // all replacement strings together
constexpr uint8_t kTransNone = 0xFF;

struct TranslateTable {                 // device memory
  uint16_t len[256];                    // output bytes of input byte b (1 = copied)
  uint16_t off[256];                    // its replacement starts at bytes[off[b]]
  uint8_t pat[256];                     // the pattern that matches it, kTransNone = none
  uint8_t bytes[kTransMaxBytes];
};

// pass 1: per tile the number of output bytes, per pattern the number of matches
__global__ void __launch_bounds__(256)
k_translate_count(const uint8_t* __restrict__ text, uint64_t n, const TranslateTable* __restrict__ tab,
                  uint64_t* __restrict__ tile_len, unsigned long long* __restrict__ counts, uint64_t n_tiles) {
  __shared__ uint16_t s_len[256];
  __shared__ uint8_t s_pat[256];
  __shared__ uint32_t s_hist[8][32];
  __shared__ uint32_t s_sum[8];
  s_len[threadIdx.x] = tab->len[threadIdx.x];
  s_pat[threadIdx.x] = tab->pat[threadIdx.x];
  s_hist[threadIdx.x >> 5][threadIdx.x & 31] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t n16 = (n + 15) & ~15ull;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t at = tile * kTransTile + (uint64_t)threadIdx.x * 16;
    const uint4 v = at < n16 ? __ldg(reinterpret_cast<const uint4*>(text + at)) : make_uint4(0, 0, 0, 0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t mine = 0;
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      if (at + p >= n) break;
      const uint32_t b = (w[p >> 2] >> (8 * (p & 3))) & 0xFFu;
      mine += s_len[b];
      const uint32_t j = s_pat[b];
      if (j != kTransNone) atomicAdd(&s_hist[warp][j & 31], 1u);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(kFullMask, mine, d);
    if (lane == 0) s_sum[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t t = 0;
      for (int q = 0; q < 8; ++q) t += s_sum[q];
      tile_len[tile] = t;
    }
    __syncthreads();
  }
  if (threadIdx.x < 32) {
    unsigned long long t = 0;
    for (int q = 0; q < 8; ++q) t += s_hist[q][threadIdx.x];
    if (t) atomicAdd(&counts[threadIdx.x], t);
  }
}

// pass 2: tile_off = exclusive prefix sum of tile_len; the tile's output is assembled in shared memory
__global__ void __launch_bounds__(256)
k_translate_write(const uint8_t* __restrict__ text, uint64_t n, const TranslateTable* __restrict__ tab,
                  const uint64_t* __restrict__ tile_off, uint8_t* __restrict__ out, uint64_t n_tiles) {
  extern __shared__ __align__(16) uint8_t s_dyn[];              // [kTransStage + 32] output, then the table's strings
  __shared__ uint16_t s_len[256], s_off[256];
  __shared__ uint32_t s_warp[33];
  uint8_t* s_out = s_dyn;
  uint8_t* s_bytes = s_dyn + kTransStage + 32;
  s_len[threadIdx.x] = tab->len[threadIdx.x];
  s_off[threadIdx.x] = tab->off[threadIdx.x];
  for (uint32_t i = threadIdx.x; i < kTransMaxBytes; i += blockDim.x) s_bytes[i] = tab->bytes[i];
  __syncthreads();
  const uint64_t n16 = (n + 15) & ~15ull;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t at = tile * kTransTile + (uint64_t)threadIdx.x * 16;
    const uint4 v = at < n16 ? __ldg(reinterpret_cast<const uint4*>(text + at)) : make_uint4(0, 0, 0, 0);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t mine = 0;
#pragma unroll
    for (int p = 0; p < 16; ++p)
      if (at + p < n) mine += s_len[(w[p >> 2] >> (8 * (p & 3))) & 0xFFu];
    uint32_t total;
    const uint32_t before = BlockExclusiveSum(mine, &total, s_warp);
    const uint64_t out_base = tile_off[tile];
    const bool staged = total <= kTransStage;
    const uint32_t align = (uint32_t)(reinterpret_cast<uintptr_t>(out + out_base) & 15u);
    uint8_t* dst = staged ? s_out + align + before : out + out_base + before;
#pragma unroll 1
    for (int p = 0; p < 16; ++p) {
      if (at + p >= n) break;
      const uint32_t b = (w[p >> 2] >> (8 * (p & 3))) & 0xFFu;
      const uint32_t l = s_len[b];
      if (l == 1 && s_off[b] == 0xFFFFu) { *dst++ = (uint8_t)b; continue; }
      const uint8_t* src = s_bytes + s_off[b];
      for (uint32_t k = 0; k < l; ++k) dst[k] = src[k];
      dst += l;
    }
    __syncthreads();
    if (staged) StageCopyOut(out + out_base, s_out, total);
    __syncthreads();
  }
}

}  // namespace rejit_b200

#endif  // REJIT_B200_CUDA_REPLACE_CUH_
