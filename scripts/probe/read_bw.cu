// Tuning aid: read-only bandwidth of this GPU under the access patterns the scan kernels use, next to a copy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe/read_bw scripts/probe/read_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldna(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// (A) linear: the whole grid sweeps the text together; U independent 16-byte loads per thread and step
template <int U>
__global__ void __launch_bounds__(256) k_linear(const uint4* __restrict__ src, size_t n16, unsigned* sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldna(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345679u) *sink = acc;
}

// (B) a warp owns a tile of `rows` rows of 512 bytes (tickets), four rows in flight: the scan kernels' pattern
__global__ void __launch_bounds__(256, 4) k_tiles(const uint4* __restrict__ src, size_t n16, uint32_t rows, unsigned long long* ticket,
                                                 unsigned* sink, int depth) {
  const int lane = threadIdx.x & 31;
  const size_t ntiles = (n16 * 16 + (size_t)rows * 512 - 1) / ((size_t)rows * 512);
  uint32_t acc = 0;
  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1ull);
    t = __shfl_sync(0xFFFFFFFFu, t, 0);
    if (t >= ntiles) break;
    const uint4* p = src + t * rows * 32 + lane;
    const size_t left = n16 - (t * rows * 32 + lane);            // (the last tile may be short: rows that start inside)
    uint32_t r_max = rows;
    if ((size_t)rows * 32 > left + lane) r_max = (uint32_t)((left + 31) / 32);
    if (depth == 4) {
      uint4 v0 = ldna(p), v1 = ldna(p + 32), v2 = ldna(p + 64), v3 = ldna(p + 96);
      for (uint32_t r = 0; r + 4 <= r_max; r += 4) {
        const uint4 a = v0, b = v1, c = v2, d = v3;
        if (r + 8 <= r_max) { v0 = ldna(p + (r + 4) * 32); v1 = ldna(p + (r + 5) * 32); v2 = ldna(p + (r + 6) * 32); v3 = ldna(p + (r + 7) * 32); }
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
      }
    } else {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ldna(p + u * 32);
      for (uint32_t r = 0; r + 8 <= r_max; r += 8) {
        uint4 w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = v[u];
        if (r + 16 <= r_max) {
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = ldna(p + (r + 8 + u) * 32);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc ^= w[u].x ^ w[u].y ^ w[u].z ^ w[u].w;
      }
    }
  }
  if (acc == 0x12345679u) *sink = acc;
}

__global__ void k_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
}

int main() {
  const size_t bytes = 2000000000ull, n16 = bytes / 16;
  uint4 *a, *b;
  unsigned* sink;
  unsigned long long* ticket;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 4); cudaMalloc(&ticket, 8);
  cudaMemset(a, 1, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto time = [&](const char* name, auto launch, double moved) {
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
      cudaMemset(ticket, 0, 8);
      cudaEventRecord(e0);
      launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("%-44s %8.3f ms  %7.1f GB/s  (%s)\n", name, best, moved / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  time("copy (read + write bytes)", [&] { k_copy<<<sms * 8, 256>>>(a, b, n16); }, 2.0 * bytes);
  time("linear read, 1 load in flight, 8 CTAs/SM", [&] { k_linear<1><<<sms * 8, 256>>>(a, n16, sink); }, bytes);
  time("linear read, 4 loads in flight, 8 CTAs/SM", [&] { k_linear<4><<<sms * 8, 256>>>(a, n16, sink); }, bytes);
  time("linear read, 8 loads in flight, 8 CTAs/SM", [&] { k_linear<8><<<sms * 8, 256>>>(a, n16, sink); }, bytes);
  time("linear read, 4 loads in flight, 4 CTAs/SM", [&] { k_linear<4><<<sms * 4, 256>>>(a, n16, sink); }, bytes);
  for (uint32_t rows : {16u, 32u, 64u, 128u, 256u}) {
    char name[96];
    snprintf(name, sizeof name, "warp tiles of %u rows, 4 in flight, 4 CTAs/SM", rows);
    time(name, [&] { k_tiles<<<sms * 4, 256>>>(a, n16, rows, ticket, sink, 4); }, bytes);
    snprintf(name, sizeof name, "warp tiles of %u rows, 8 in flight, 4 CTAs/SM", rows);
    time(name, [&] { k_tiles<<<sms * 4, 256>>>(a, n16, rows, ticket, sink, 8); }, bytes);
  }
  // the same tiling with dynamic shared memory allocated (not used): does the L1 that is left bound the loads in flight?
  for (int kb : {0, 16, 32, 48, 55}) {
    char name[96];
    snprintf(name, sizeof name, "warp tiles of 64 rows, 4 in flight, %d KB smem/CTA", kb);
    cudaFuncSetAttribute(k_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
    time(name, [&] { k_tiles<<<sms * 4, 256, kb * 1024>>>(a, n16, 64, ticket, sink, 4); }, bytes);
  }
  for (int kb : {48}) {
    char name[96];
    snprintf(name, sizeof name, "warp tiles of 64 rows, 8 in flight, %d KB smem/CTA, 3 CTAs/SM", kb);
    time(name, [&] { k_tiles<<<sms * 3, 256, kb * 1024>>>(a, n16, 64, ticket, sink, 8); }, bytes);
  }
  // the same at 500 MB (the C3 text): three tiles of 64 rows per warp
  const size_t small16 = 500000000ull / 16;
  time("500 MB: linear read, 4 loads in flight", [&] { k_linear<4><<<sms * 8, 256>>>(a, small16, sink); }, 5e8);
  time("500 MB: warp tiles of 64 rows, 4 in flight", [&] { k_tiles<<<sms * 4, 256>>>(a, small16, 64, ticket, sink, 4); }, 5e8);
  time("500 MB: warp tiles of 64 rows, 8 in flight", [&] { k_tiles<<<sms * 4, 256>>>(a, small16, 64, ticket, sink, 8); }, 5e8);
  return 0;
}
