// rejit_b200 — ReplaceAll on the device, round 2 (reference: Regej::ReplaceAll,
// /root/reference/src/rejit.cc:221-226, and Replace, :97-112).
//
//   k_replace_stage   the rebuild of round 1 (k_replace_tiles: prefix sum over the match lengths, then every
//                     4 KB input tile places its own bytes and replacements) with both ends staged in shared
//                     memory: the tile is fetched with coalesced 16-byte loads, its output is assembled in
//                     shared memory and leaves with coalesced 16-byte stores.  Round 1 read and wrote single
//                     bytes from global memory: 0.2 TB/s.
//   k_translate_*2    a SET of patterns that each match exactly one byte (regex-dna's eleven IUB codes,
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

// first[t] = index of the first match that begins at or after tile t's first byte (first[n_tiles] = m): one thread
// per tile, so that the ~log2(m) dependent loads of a million searches overlap.  (k_replace_stage used to search
// twice per tile on one thread of the CTA: with 80 M matches that was 25 us of latency per 4 KB tile, 3/4 of the
// rebuild of a 5 GB FASTA file.)
__global__ void __launch_bounds__(256)
k_replace_index(const uint64_t* __restrict__ pairs, uint64_t m, uint64_t n_tiles, uint64_t* __restrict__ first) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  first[t] = t == n_tiles ? m : ReplaceLowerBound(pairs, m, t * kReplaceTile);
}

__global__ void __launch_bounds__(256)
k_replace_stage(const uint8_t* __restrict__ text, uint64_t n, const uint64_t* __restrict__ pairs,
                const uint64_t* __restrict__ removed, uint64_t m, const uint8_t* __restrict__ with, uint32_t w,
                uint8_t* __restrict__ out, uint64_t n_tiles, const uint64_t* __restrict__ first) {
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
    if (threadIdx.x == 0) s_head.m0 = first[tile];
    if (threadIdx.x == 32) s_next.m0 = first[tile + 1];
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
constexpr uint32_t kTransMaxBytes = 4096;           // all replacement strings together
constexpr uint8_t kTransNone = 0xFF;

struct TranslateTable {                 // device memory
  uint32_t pre_mask, pre_value;         // every replaced byte b has (b & pre_mask) == pre_value (both x 0x01010101)
  uint16_t len[256];                    // output bytes of input byte b (1 = copied)
  uint16_t off[256];                    // its replacement starts at bytes[off[b]]
  uint8_t pat[256];                     // the pattern that matches it, kTransNone = none
  uint8_t bytes[kTransMaxBytes];
};

// ===========================================================================
// byte -> string table, round 2b: warp-autonomous streaming passes.
//   The first version worked a 4 KB tile per CTA iteration with one load per thread in flight, two block barriers
//   per tile, shared-memory atomics per replaced byte and byte-wide stores with four-way bank conflicts for EVERY
//   byte: 0.5 TB/s of traffic.  Here a WARP owns a 16 KB tile (32 rows of 512 bytes, four rows in flight, no block
//   barrier after the tables are loaded):
//   count   one class lookup per byte (0: copied, j + 1: replaced by pattern j); only lanes that hold a replaced
//           byte look again (lane-private histogram bins in shared memory: no atomics, bank = lane).
//   write   a row without replaced bytes (the common case) is COPIED: the output offset is not 16-byte aligned in
//           general, so every lane builds the aligned 16-byte chunk that covers its place from its left
//           neighbour's bytes and its own (shuffle + byte permute) and the warp stores 512 aligned bytes; the
//           chunk cut by the row's end waits for the next row.  A row with replaced bytes is assembled in the
//           warp's staging area and leaves with aligned 16-byte stores.
// Algorithmic traffic: 2 N read, N' written.
// ===========================================================================
constexpr uint32_t kTr2Rows = 32;                                   // rows of 512 bytes per tile
constexpr uint32_t kTr2Tile = kTr2Rows * 512;                       // 16 KB, one warp
constexpr uint32_t kTr2Warps = 8;
constexpr uint32_t kTr2Stage = 2048;                                // output of one row that is assembled in shared memory
constexpr uint32_t kTr2MaxLen = 4095;                               // longest replacement (16 of them fit 16 bits)
constexpr int kTr2Depth = 8;                                        // rows in flight per warp, counting pass
__device__ __forceinline__ uint4 Tr2Load(const uint4* p) {          // read once: no L1 allocation
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

struct Tr2Tables {                      // shared memory, loaded once per CTA
  uint8_t cls[256];                     // 0: the byte is copied; j + 1: pattern j replaces it
  uint16_t len[34];                     // [cls]: output bytes (len[0] = 1)
  uint16_t off[34];                     // [cls]: the replacement starts at bytes[off]
};

__device__ __forceinline__ void Tr2LoadTables(const TranslateTable* __restrict__ tab, Tr2Tables* t) {
  for (uint32_t b = threadIdx.x; b < 256; b += blockDim.x) {
    const uint8_t pj = tab->pat[b];
    t->cls[b] = pj == kTransNone ? 0 : (uint8_t)(pj + 1);
    if (pj != kTransNone) { t->len[pj + 1] = tab->len[b]; t->off[pj + 1] = tab->off[b]; }     // (same value from every byte of the pattern)
  }
  if (threadIdx.x == 0) { t->len[0] = 1; t->off[0] = 0; }
}

// class ids of my 16 bytes OR-ed together (0: nothing to replace).  First a SWAR test on the bits that all replaced
// bytes have in common (a word without a candidate byte costs 4 instructions instead of 16): "some byte of x is zero"
// may also fire on the byte above a zero byte, which only sends the word to the exact test.
__device__ __forceinline__ uint32_t Tr2AnySpecial(const uint4& v, const uint8_t* cls, uint32_t pm, uint32_t pv) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t any = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t x = (w[j] & pm) ^ pv;
    if ((x - 0x01010101u) & ~x & 0x80808080u) {
#pragma unroll
      for (int k = 0; k < 4; ++k) any |= cls[__byte_perm(w[j], 0u, 0x4440 | k)];
    }
  }
  return any;
}

__global__ void __launch_bounds__(kTr2Warps * 32)
k_translate_count2(const uint8_t* __restrict__ text, uint64_t n, const TranslateTable* __restrict__ tab, int K,
                   uint64_t* __restrict__ tile_len, unsigned long long* __restrict__ counts, uint64_t n_tiles) {
  extern __shared__ __align__(16) uint8_t tr_smem[];
  Tr2Tables* t = reinterpret_cast<Tr2Tables*>(tr_smem);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(tr_smem + ((sizeof(Tr2Tables) + 15) & ~15u));      // [warp][K][32 lanes]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Tr2LoadTables(tab, t);
  for (uint32_t i = threadIdx.x; i < kTr2Warps * (uint32_t)K * 32; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  uint32_t* my_hist = s_hist + (uint32_t)warp * K * 32 + lane;
  const uint32_t pm = tab->pre_mask, pv = tab->pre_value;
  const uint64_t n16 = (n + 15) & ~15ull;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  const uint64_t n_warps = (uint64_t)gridDim.x * kTr2Warps;
  for (uint64_t tile = (uint64_t)blockIdx.x * kTr2Warps + warp; tile < n_tiles; tile += n_warps) {
    const uint64_t tile_lo = tile * kTr2Tile;
    const uint64_t mine = tile_lo + (uint64_t)lane * 16;
    const uint32_t rows = n - tile_lo >= kTr2Tile ? kTr2Rows : (uint32_t)((n - tile_lo + 511) >> 9);
    const uint32_t rows_ld = mine < n16 ? (uint32_t)(((n16 - mine + 511) >> 9) < kTr2Rows ? ((n16 - mine + 511) >> 9) : kTr2Rows) : 0u;
    const uint4* src = reinterpret_cast<const uint4*>(text + mine);
    uint4 v[kTr2Depth];                                     // rows in flight (see k_scan_emit: four were latency-bound)
#pragma unroll
    for (int u = 0; u < kTr2Depth; ++u) v[u] = (uint32_t)u < rows_ld ? Tr2Load(src + u * 32) : zero4;
    uint32_t extra = 0;                                     // output bytes beyond one per input byte (may be "negative")
    auto row = [&](uint4& v, uint32_t r) {
      const uint4 cur = v;
      v = r + kTr2Depth < rows_ld ? Tr2Load(src + (r + kTr2Depth) * 32) : zero4;
      if (Tr2AnySpecial(cur, t->cls, pm, pv)) {
        const uint64_t at = mine + (uint64_t)r * 512;
        const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const uint32_t c = t->cls[__byte_perm(w[p >> 2], 0u, 0x4440 | (p & 3))];
          if (c && at + p < n) { extra += (uint32_t)t->len[c] - 1u; my_hist[(c - 1) * 32] += 1; }
        }
      }
    };
    uint32_t r = 0;
#pragma unroll 1
    for (; r + kTr2Depth <= rows; r += kTr2Depth) {
#pragma unroll
      for (int u = 0; u < kTr2Depth; ++u) row(v[u], r + u);
    }
#pragma unroll
    for (int u = 0; u < kTr2Depth - 1; ++u)
      if (r + u < rows) row(v[u], r + u);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) extra += __shfl_xor_sync(kFullMask, extra, d);
    if (lane == 0) {
      const uint64_t in_bytes = n - tile_lo < kTr2Tile ? n - tile_lo : kTr2Tile;
      tile_len[tile] = in_bytes + (uint64_t)(int64_t)(int32_t)extra;
    }
  }
  __syncwarp();
  for (int j = 0; j < K; ++j) {
    uint32_t c = my_hist[j * 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(kFullMask, c, d);
    if (lane == 0 && c) atomicAdd(&counts[j], (unsigned long long)c);
  }
}

// the aligned 16-byte chunk that begins `o` bytes (1..15) into the 32-byte window p:v
__device__ __forceinline__ uint4 Tr2Window(const uint4& p, const uint4& v, uint32_t o) {
  const uint32_t w[8] = {p.x, p.y, p.z, p.w, v.x, v.y, v.z, v.w};
  const bool s2 = (o & 8u) != 0, s1 = (o & 4u) != 0;
  uint32_t a[6], b[5];
#pragma unroll
  for (int i = 0; i < 6; ++i) a[i] = s2 ? w[i + 2] : w[i];
#pragma unroll
  for (int i = 0; i < 5; ++i) b[i] = s1 ? a[i + 1] : a[i];
  const uint32_t sel = 0x3210u + 0x1111u * (o & 3u);
  return make_uint4(__byte_perm(b[0], b[1], sel), __byte_perm(b[1], b[2], sel), __byte_perm(b[2], b[3], sel),
                    __byte_perm(b[3], b[4], sel));
}

__global__ void __launch_bounds__(kTr2Warps * 32, 4)
k_translate_write2(const uint8_t* __restrict__ text, uint64_t n, const TranslateTable* __restrict__ tab,
                   const uint64_t* __restrict__ tile_off, uint8_t* __restrict__ out, uint64_t n_tiles) {
  __shared__ Tr2Tables s_t;
  __shared__ __align__(16) uint8_t s_bytes[kTransMaxBytes];
  __shared__ __align__(16) uint8_t s_stage_all[kTr2Warps][kTr2Stage + 32];
  __shared__ uint32_t s_list_all[kTr2Warps][512];             // replaced bytes of the row in work: place | class << 16
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Tr2LoadTables(tab, &s_t);
  for (uint32_t i = threadIdx.x; i < kTransMaxBytes; i += blockDim.x) s_bytes[i] = tab->bytes[i];
  __syncthreads();
  const Tr2Tables* t = &s_t;
  const uint32_t pm = tab->pre_mask, pv = tab->pre_value;
  uint8_t* stage = s_stage_all[warp];
  uint32_t* list = s_list_all[warp];
  const uint64_t n16 = (n + 15) & ~15ull;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  const uint64_t n_warps = (uint64_t)gridDim.x * kTr2Warps;
  for (uint64_t tile = (uint64_t)blockIdx.x * kTr2Warps + warp; tile < n_tiles; tile += n_warps) {
    const uint64_t tile_lo = tile * kTr2Tile;
    const uint64_t mine = tile_lo + (uint64_t)lane * 16;
    const uint32_t rows = n - tile_lo >= kTr2Tile ? kTr2Rows : (uint32_t)((n - tile_lo + 511) >> 9);
    const uint32_t rows_ld = mine < n16 ? (uint32_t)(((n16 - mine + 511) >> 9) < kTr2Rows ? ((n16 - mine + 511) >> 9) : kTr2Rows) : 0u;
    const uint4* src = reinterpret_cast<const uint4*>(text + mine);
    uint4 v0 = 0 < rows_ld ? __ldg(src) : zero4, v1 = 1 < rows_ld ? __ldg(src + 32) : zero4,
          v2 = 2 < rows_ld ? __ldg(src + 64) : zero4, v3 = 3 < rows_ld ? __ldg(src + 96) : zero4;
    uint64_t pos = tile_off[tile];                          // where the next row's output begins (uniform)
    bool pend = false;                                      // the last sh bytes of the row before wait in `last` (lane 31)
    uint4 last = zero4;
    auto flush = [&]() {
      if (pend) {
        const uint32_t sh = (uint32_t)pos & 15u;            // (pos moved by whole rows since the run began)
        if (lane == 31) {
          const uint32_t w[4] = {last.x, last.y, last.z, last.w};
          for (uint32_t k = 16 - sh; k < 16; ++k) out[pos - 16 + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
        }
        pend = false;
      }
    };
    auto row = [&](uint4& v, uint32_t r) {
      const uint4 cur = v;
      v = r + 4 < rows_ld ? __ldg(src + (r + 4) * 32) : zero4;
      const uint64_t at = mine + (uint64_t)r * 512;
      const bool whole = tile_lo + (uint64_t)(r + 1) * 512 <= n;                           // every byte of the row is text
      const uint32_t any = Tr2AnySpecial(cur, t->cls, pm, pv);
      if (whole && !__any_sync(kFullMask, any != 0)) {
        // ---- copy -----------------------------------------------------------------------------------
        const uint32_t sh = (uint32_t)pos & 15u;
        if (sh == 0) {
          *reinterpret_cast<uint4*>(out + pos + (uint64_t)lane * 16) = cur;
        } else {
          uint4 give = lane == 31 ? last : cur, p;
          const int from = (lane + 31) & 31;
          p.x = __shfl_sync(kFullMask, give.x, from); p.y = __shfl_sync(kFullMask, give.y, from);
          p.z = __shfl_sync(kFullMask, give.z, from); p.w = __shfl_sync(kFullMask, give.w, from);
          if (lane > 0 || pend) {
            *reinterpret_cast<uint4*>(out + (pos - sh) + (uint64_t)lane * 16) = Tr2Window(p, cur, 16 - sh);
          } else {
            const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
            for (uint32_t k = 0; k < 16 - sh; ++k) out[pos + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
          }
          pend = true;
          last = cur;
        }
        pos += 512;
        return;
      }
      // ---- a row with replaced bytes (or the text's last row): assembled in shared memory ----------------
      // Two phases, so that lanes copying a replacement string do not hold up lanes copying single bytes (one
      // divergent loop over the 16 bytes ran at 0.13 TB/s on the IUB section): every lane places its copied bytes
      // and leaves {place, class} of its replaced bytes in the warp's list; then the list's entries are spread
      // evenly over the lanes.
      flush();
      const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
      uint32_t packed = 0;                                  // output bytes | replaced bytes << 22
      const uint32_t nb = whole ? 16u : (at < n ? (n - at < 16 ? (uint32_t)(n - at) : 16u) : 0u);      // my bytes that are text
#pragma unroll
      for (int p = 0; p < 16; ++p)
        if ((uint32_t)p < nb) {
          const uint32_t c = t->cls[__byte_perm(w[p >> 2], 0u, 0x4440 | (p & 3))];
          packed += (uint32_t)t->len[c] + (c ? 1u << 22 : 0u);
        }
      const uint32_t incl = WarpInclusiveScan(packed);
      const uint32_t all = __shfl_sync(kFullMask, incl, 31);
      const uint32_t total = all & 0x3FFFFFu, n_spec = all >> 22;
      const uint32_t excl = incl - packed;
      const uint32_t a = (uint32_t)pos & 15u;
      const bool staged = total <= kTr2Stage;
      if (staged) {
        uint32_t q = a + (excl & 0x3FFFFFu), li = excl >> 22;
#pragma unroll
        for (int p = 0; p < 16; ++p)
          if ((uint32_t)p < nb) {
            const uint32_t b = __byte_perm(w[p >> 2], 0u, 0x4440 | (p & 3));
            const uint32_t c = t->cls[b];
            if (!c) { stage[q++] = (uint8_t)b; }
            else { list[li++] = q | (c << 16); q += t->len[c]; }
          }
        __syncwarp();
        for (uint32_t e = lane; e < n_spec; e += 32) {
          const uint32_t ent = list[e];
          const uint32_t c = ent >> 16;
          uint8_t* d = stage + (ent & 0xFFFFu);
          const uint8_t* from = s_bytes + t->off[c];
          const uint32_t l = t->len[c];
          for (uint32_t k = 0; k < l; ++k) d[k] = from[k];
        }
      } else {
        // (replacements longer than the staging area: byte by byte to global memory)
        uint8_t* dst = out + pos + (excl & 0x3FFFFFu);
#pragma unroll 1
        for (int p = 0; p < 16; ++p) {
          if (at + p >= n) break;
          const uint32_t b = (w[p >> 2] >> (8 * (p & 3))) & 0xFFu;
          const uint32_t c = t->cls[b];
          if (!c) { *dst++ = (uint8_t)b; continue; }
          const uint32_t l = t->len[c];
          const uint8_t* from = s_bytes + t->off[c];
          for (uint32_t k = 0; k < l; ++k) dst[k] = from[k];
          dst += l;
        }
      }
      __syncwarp();
      if (staged) {
        uint8_t* o = out + pos;
        uint32_t head = (16u - a) & 15u;
        if (head > total) head = total;
        if ((uint32_t)lane < head) o[lane] = stage[a + lane];
        const uint32_t body = (total - head) >> 4;
        const uint4* sv = reinterpret_cast<const uint4*>(stage + a + head);
        uint4* dv = reinterpret_cast<uint4*>(o + head);
        for (uint32_t i = lane; i < body; i += 32) dv[i] = sv[i];
        const uint32_t done = head + (body << 4);
        if ((uint32_t)lane < total - done) o[done + lane] = stage[a + done + lane];
        __syncwarp();
      }
      pos += total;
    };
    uint32_t r = 0;
#pragma unroll 1
    for (; r + 4 <= rows; r += 4) { row(v0, r); row(v1, r + 1); row(v2, r + 2); row(v3, r + 3); }
    if (r < rows) row(v0, r);
    if (r + 1 < rows) row(v1, r + 1);
    if (r + 2 < rows) row(v2, r + 2);
    flush();
  }
}

}  // namespace rejit_b200

#endif  // REJIT_B200_CUDA_REPLACE_CUH_
