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
#include <atomic>
#include <condition_variable>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "device_program.h"
#include "kernels.cuh"
#include "scan_emit.cuh"
#include "replace.cuh"

namespace rejit_b200 {

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

// the same for functions that return a count (-1 = failed): `false` would read as "0 matches, success"
#define RJ_TRY_COUNT(call)                           \
  do {                                               \
    if (!Check((call), #call, error)) return -1;     \
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
  template <class T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

// ---------------------------------------------------------------------------
// A small pool of persistent worker threads: parallel-for over task indices.  Used to stage pageable host text
// into pinned chunks (several memcpy streams keep PCIe busy) and to drive one slab per device in MatchAllParallel
// (round 1 spawned threads per call).
// ---------------------------------------------------------------------------
class WorkerPool {
 public:
  static WorkerPool& Staging() { static WorkerPool* pool = new WorkerPool(31); return *pool; }   // memcpy streams
  static WorkerPool& Devices() { static WorkerPool* pool = new WorkerPool(15); return *pool; }   // one task per GPU
  int size() const { return (int)threads_.size(); }
  // fn(i) for i in [0, n), on up to `width` threads (the caller works too); returns when all are done
  void Run(int n, int width, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    if (n == 1 || width <= 1 || threads_.empty()) { for (int i = 0; i < n; ++i) fn(i); return; }
    std::unique_lock<std::mutex> gate(run_mu_);              // one parallel-for at a time
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn; next_ = 0; total_ = n; done_ = 0;
      helpers_ = std::min<int>(width - 1, (int)threads_.size());
      ++epoch_;
    }
    cv_.notify_all();
    Work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return done_ == total_ && active_ == 0; });
    fn_ = nullptr;
  }

 private:
  explicit WorkerPool(unsigned cap) {
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int)std::max(1u, std::min(cap, hw > 3 ? hw - 2 : 1u));
    for (int i = 0; i < n; ++i) threads_.emplace_back([this, i] { Loop(i); });
    for (auto& t : threads_) t.detach();
  }
  void Work() {
    for (;;) {
      int i;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (next_ >= total_) return;
        i = next_++;
      }
      (*fn_)(i);
      std::lock_guard<std::mutex> lk(mu_);
      if (++done_ == total_) done_cv_.notify_all();
    }
  }
  void Loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (id >= helpers_ || !fn_) continue;
        ++active_;
      }
      Work();
      std::lock_guard<std::mutex> lk(mu_);
      if (--active_ == 0) done_cv_.notify_all();
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, total_ = 0, done_ = 0, helpers_ = 0, active_ = 0;
  uint64_t epoch_ = 0;
};

// finish area inside DeviceContext::status
constexpr size_t kFinMaxSegments = 1024;
constexpr size_t kFinSyncOffset = 128, kFinLastOffset = 160, kFinTotalOffset = 168, kFinLastNeOffset = 176,
                 kFinSegOffset = 184;

class DeviceContext {
 public:
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 227 * 1024;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::mutex mu;                      // one MatchAll at a time per device context
  Buffer text;                        // staging for host-text calls
  // ordered stores (slot ranges per sub-region): candidates and needle hits
  Buffer sub_b, sub_e, sub_count, hsub_b, hsub_e, hsub_count;
  // dense (gathered, sorted) lists
  Buffer dense_b, dense_e, hits_b, hits_e, out_pairs, with_buf;
  uint64_t dense_cap = 0, hits_cap = 0;
  // unordered fallback (k_dfa_scan)
  Buffer cand_b, cand_e;
  uint64_t cand_cap = 0;
  Buffer counters;                    // u64: [0] unordered count [1] dense count [2] hits count [3] last_any [4] last_nonempty
  Buffer status;
  Buffer sorted_b, sorted_e, reach, take, wide, slot, fin_end, cub_tmp;
  uint64_t scratch_cap = 0;
  Buffer fscratch;                    // label scratch for re-entrant patterns
  // fused pattern sets
  Buffer set_sub_b, set_sub_e, set_sub_count, set_dense_b, set_dense_e, set_reach, set_take, set_fin, set_slot,
      set_out, set_status, set_counts, kmer_xchg, kmer_stage, kmer_probe;
  PipelineStatus* h_set_status = nullptr;      // pinned + mapped, 32 entries
  PipelineStatus* h_set_status_dev = nullptr;
  Buffer flush;
  // pinned staging ring for pageable host texts (UploadHostText)
  uint8_t* arena = nullptr;            // pinned, grows to the longest pageable text seen (<= 256 MB)
  size_t arena_bytes = 0;
  bool arena_used = false;
  cudaEvent_t arena_ev = nullptr;      // recorded after the last copy out of the arena
  Buffer trans_tab, trans_len, trans_off, trans_counts;   // byte -> string table of ReplaceAllSetDevice and its tile sums
  // device-side stitch (one process per GPU): my inbox, the neighbours' inboxes mapped through CUDA IPC
  void* stitch_inbox = nullptr;
  void* stitch_right = nullptr;
  void* stitch_left = nullptr;
  int stitch_rank = 0, stitch_world = 1;
  unsigned int stitch_step = 0;
  StitchReport* h_stitch = nullptr;
  StitchReport* h_stitch_dev = nullptr;
  Buffer em_records;                  // look-back records of the single-pass scans (scan_emit.cuh)
  int em_blocks_per_sm = 0;           // resident CTAs of k_scan_emit per SM (asked once)
  bool emit = true;                   // single-pass scan + emit available (RJ_NO_EMIT=1: the round-1 pipelines only)
  PipelineStatus* h_status = nullptr; // pinned + mapped: the resolve kernel writes it, the host spins on seq
  PipelineStatus* h_status_dev = nullptr;   // device view of h_status
  FinRecord* h_fin = nullptr;               // mapped: records of the in-kernel finish (one per pattern)
  FinRecord* h_fin_dev = nullptr;
  unsigned int call_seq = 0;
  bool attr_done = false;
  bool coop = false;                  // cooperative launch available: scans finish in-kernel
  int lit_blocks_per_sm = 0;          // co-resident CTAs of k_lit_scan (occupancy query, cached)
  int gen_blocks_per_sm = 0, win_blocks_per_sm = 0;

  bool Init(int dev, std::string* error) {
    device = dev;
    RJ_TRY(cudaSetDevice(dev));
    cudaDeviceProp prop;
    RJ_TRY(cudaGetDeviceProperties(&prop, dev));
    sm_count = prop.multiProcessorCount;
    smem_optin = prop.sharedMemPerBlockOptin;
    coop = prop.cooperativeLaunch != 0 && getenv("RJ_NO_FUSED_FINISH") == nullptr;
    emit = getenv("RJ_NO_EMIT") == nullptr && getenv("RJ_NO_FUSED_FINISH") == nullptr;
    RJ_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (auto& e : ev) RJ_TRY(cudaEventCreate(&e));
    RJ_TRY(cudaHostAlloc(&h_status, sizeof(PipelineStatus), cudaHostAllocMapped));
    memset(h_status, 0, sizeof(PipelineStatus));
    RJ_TRY(cudaHostGetDevicePointer(&h_status_dev, h_status, 0));
    RJ_TRY(cudaHostAlloc(&h_fin, 32 * sizeof(FinRecord), cudaHostAllocMapped));
    memset(h_fin, 0, 32 * sizeof(FinRecord));
    RJ_TRY(cudaHostGetDevicePointer(&h_fin_dev, h_fin, 0));
    RJ_TRY(cudaHostAlloc(&h_set_status, 32 * sizeof(PipelineStatus), cudaHostAllocMapped));
    memset(h_set_status, 0, 32 * sizeof(PipelineStatus));
    RJ_TRY(cudaHostGetDevicePointer(&h_set_status_dev, h_set_status, 0));
    // status and the counters share one allocation so that one memset clears both
    // layout: [PipelineStatus][counters 40 B, at +64][finish: sync 8 x u32 at +128, last_end at +160,
    //          total at +168, last non-empty end at +176, segcount at +184 (one u32 per segment)]
    if (!status.Reserve(kFinSegOffset + 4 * (size_t)kFinMaxSegments + 64, error)) return false;
    counters.p = static_cast<uint8_t*>(status.p) + ((sizeof(PipelineStatus) + 15) & ~size_t(15));
    counters.bytes = 0;                 // not owned
    return true;
  }
  bool ReserveDense(uint64_t cap, std::string* error) {
    if (cap <= dense_cap) return true;
    if (!dense_b.Reserve(cap * 8, error) || !dense_e.Reserve(cap * 8, error) ||
        !out_pairs.Reserve(cap * 16, error)) return false;
    dense_cap = cap;
    return ReserveScratch(cap, error);
  }
  bool ReserveScratch(uint64_t cap, std::string* error) {
    if (cap <= scratch_cap) return true;
    if (!reach.Reserve(cap * 8, error) || !take.Reserve(cap * 4, error) || !fin_end.Reserve(cap * 8, error) ||
        !slot.Reserve(cap * 8, error) || !wide.Reserve(cap * 8, error)) return false;
    scratch_cap = cap;
    return true;
  }
  bool ReserveHits(uint64_t cap, std::string* error) {
    if (cap <= hits_cap) return true;
    if (!hits_b.Reserve(cap * 8, error) || !hits_e.Reserve(cap * 8, error)) return false;
    hits_cap = cap;
    return true;
  }
  bool ReserveUnordered(uint64_t cap, std::string* error) {
    if (cap <= cand_cap) return true;
    if (!cand_b.Reserve(cap * 8, error) || !cand_e.Reserve(cap * 8, error) || !out_pairs.Reserve(cap * 16, error))
      return false;
    cand_cap = cap;
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
  // adaptive capacities (remembered across calls so that steady state never reruns)
  GenFilter gen_filter{};             // start filter of the generic scan
  EmFilter em_filter{};               // the same for the single-pass scan (SWAR form when the start set is small)
  bool emit_off = false;              // candidates too dense for the single-pass scan's lists: round-1 pipeline
  uint32_t cand_sub_cap = 16;         // slots per sub-region, candidate store
  uint32_t hit_sub_cap = 16;          // slots per sub-region, needle-hit store
  bool dense_mode = false;            // k_dfa_tma's lane lists overflowed once: use k_dfa_scan
  uint32_t fuse_misses = 0, fuse_skips = 0;   // in-kernel finish: consecutive misses / calls skipped since
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

    // start filter: byte c can begin a match (or the empty match is possible) in
    // the context (sol, eol(c))
    for (int sol = 0; sol < 2; ++sol)
      for (int b = 0; b < 256; ++b) {
        int eol = (b == '\n' || b == '\r') ? 1 : 0;
        int ctx = ca.nfa.has_anchor ? (sol | (eol << 1)) : 0;
        if (ca.start_ok[ctx][b] || ca.nfa.accept_empty[ctx]) gen_filter.t[sol][b >> 5] |= 1u << (b & 31);
      }
    {
      // SWAR form of the start filter: S0 = start bytes that need no context, S1 = start bytes right after a
      // line break.  Exact when S0 has at most four bytes and S1 adds at most four more, or every byte.
      EmFilter& f = em_filter;
      memcpy(f.t, gen_filter.t, sizeof f.t);
      std::vector<int> s0, extra;
      bool subset = true;
      for (int b = 0; b < 256; ++b) {
        const bool in0 = (f.t[0][b >> 5] >> (b & 31)) & 1u, in1 = (f.t[1][b >> 5] >> (b & 31)) & 1u;
        if (in0) s0.push_back(b);
        if (in0 && !in1) subset = false;
        if (in1 && !in0) extra.push_back(b);
      }
      f.use_sol = extra.empty() ? 0u : 1u;
      f.all1 = (s0.size() + extra.size() == 256) ? 1u : 0u;
      if (subset && s0.size() <= 4 && (extra.size() <= 4 || f.all1)) {
        f.swar = 1;
        f.n0 = (uint32_t)s0.size();
        for (size_t i = 0; i < s0.size(); ++i) f.b0 |= (uint32_t)s0[i] << (8 * i);
        if (!f.all1) {
          f.n1 = (uint32_t)extra.size();
          for (size_t i = 0; i < extra.size(); ++i) f.b1 |= (uint32_t)extra[i] << (8 * i);
        }
      }
    }
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
      if (!Upload(ft.dfa_next, &dfa.next, error) || !Upload(ft.dfa_class, &dfa.byte_class, error) ||
          !Upload(ft.dfa_pair, &dfa.pair, error)) return false;
      dfa.first_accept = d.first_accept;
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

// Text-sized device buffers come and go in chains of ReplaceAll calls and in
// upload / search / free loops; cudaMalloc + cudaFree cost about a millisecond
// a pair, so freed buffers are kept (a few per device) and handed out again when
// the size fits.
namespace {
struct PooledBuffer { void* p; size_t bytes; };
std::mutex g_pool_mu;
std::vector<PooledBuffer> g_pool[16];
std::vector<PooledBuffer> g_live[16];          // sizes of the buffers handed out
constexpr size_t kPoolBuffers = 6;
}  // namespace

void* DeviceAlloc(int device, size_t bytes, std::string* error) {
  if (!CudaOk(error)) return nullptr;
  if (device < 0 || device >= 16) { if (error) *error = "rejit_b200: bad device"; return nullptr; }
  // 64 bytes of slack so that 16-byte vector loads near the end never leave
  // the allocation
  const size_t want = bytes + 64;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto& pool = g_pool[device];
    size_t best = pool.size();
    for (size_t i = 0; i < pool.size(); ++i)
      if (pool[i].bytes >= want && pool[i].bytes <= 2 * want + (1u << 20) &&
          (best == pool.size() || pool[i].bytes < pool[best].bytes)) best = i;
    if (best != pool.size()) {
      PooledBuffer b = pool[best];
      pool.erase(pool.begin() + best);
      g_live[device].push_back(b);
      return b.p;
    }
  }
  void* p = nullptr;
  if (!Check(cudaSetDevice(device), "cudaSetDevice", error)) return nullptr;
  const size_t grow = want + want / 8;           // room for a text that grows a little (IUB substitutions)
  if (cudaMalloc(&p, grow) != cudaSuccess) {
    cudaGetLastError();
    // out of memory: drop the pool and try the exact size
    {
      std::lock_guard<std::mutex> lk(g_pool_mu);
      for (auto& b : g_pool[device]) cudaFree(b.p);
      g_pool[device].clear();
    }
    if (!Check(cudaMalloc(&p, want), "cudaMalloc", error)) return nullptr;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_live[device].push_back({p, want});
    return p;
  }
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_live[device].push_back({p, grow});
  return p;
}
void DeviceFree(int device, void* p) {
  if (!p || device < 0 || device >= 16) return;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  auto& live = g_live[device];
  size_t bytes = 0;
  for (size_t i = 0; i < live.size(); ++i)
    if (live[i].p == p) { bytes = live[i].bytes; live.erase(live.begin() + i); break; }
  auto& pool = g_pool[device];
  if (bytes && pool.size() < kPoolBuffers) { pool.push_back({p, bytes}); return; }
  if (bytes && !pool.empty()) {
    // keep the larger buffers
    size_t smallest = 0;
    for (size_t i = 1; i < pool.size(); ++i) if (pool[i].bytes < pool[smallest].bytes) smallest = i;
    if (pool[smallest].bytes < bytes) std::swap(pool[smallest].p, p), std::swap(pool[smallest].bytes, bytes);
  }
  cudaSetDevice(device);
  cudaFree(p);
}
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
bool CopyOnDevice(int device, void* dst, const void* src, size_t bytes, std::string* error) {
  RJ_TRY(cudaSetDevice(device));
  RJ_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
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

// Waits until the in-kernel finish has published its K records for call `seq`
// (both halves of every record), then turns record j into a PipelineStatus.
bool WaitFinRecords(DeviceContext* c, int K, unsigned int seq, std::string* error) {
  uint64_t spins = 0;
  for (int j = 0; j < K; ++j) {
    volatile FinRecord* r = c->h_fin + j;
    while (r->seq0 != seq || r->seq1 != seq || r->seq2 != seq) {
      if ((++spins & 0x3FFF) == 0) {
        cudaError_t q = cudaStreamQuery(c->stream);
        if (q == cudaSuccess) {
          if (r->seq0 == seq && r->seq1 == seq && r->seq2 == seq) break;
          if (error) *error = "rejit_b200: the scan kernel did not report";
          return false;
        }
        if (q != cudaErrorNotReady) { Check(q, "cudaStreamQuery", error); return false; }
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return true;
}

PipelineStatus StatusFromRecord(const FinRecord& r, const Carry& carry_in) {
  PipelineStatus st{};
  st.n_candidates = r.n_matches;
  st.n_matches = r.n_matches;
  // the chain state after the last match (ChainTake): a non-empty match [b,e) leaves (e, e); an empty
  // one at b leaves (b + 1, end of the last non-empty match)
  st.carry_cur = carry_in.cur;
  st.carry_tail = carry_in.tail;
  if (r.n_matches) {
    const bool last_is_empty = (r.flags & kFinLastEmpty) != 0 || r.last_nonempty != r.last_end;
    st.carry_cur = last_is_empty ? r.last_end + 1 : r.last_end;
    if (r.last_nonempty) st.carry_tail = r.last_nonempty;
  }
  st.overflow = (r.flags & kFinOverflow) ? 1u : 0u;
  st.need_cap = r.need_cap;
  st.need_large = (r.flags & kFinOverlap) ? 1u : 0u;
  st.dense = (r.flags & kFinDense) ? 1u : 0u;
  return st;
}

// Kernels that use more dynamic shared memory than the default limit (once per device context).
bool EnsureKernelAttributes(DeviceContext* c, std::string* error) {
  if (c->attr_done) return true;
  RJ_TRY(cudaFuncSetAttribute(k_resolve_small, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallResolveMax * 16));
  RJ_TRY(cudaFuncSetAttribute(k_dfa_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin));
  RJ_TRY(cudaFuncSetAttribute(k_dfa_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  RJ_TRY(cudaFuncSetAttribute(k_set_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin));
  RJ_TRY(cudaFuncSetAttribute(k_set_kmer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin));
  RJ_TRY(cudaFuncSetAttribute(k_scan_emit<kEmLiteral, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmSmemBytes));
  RJ_TRY(cudaFuncSetAttribute(k_scan_emit<kEmLiteral, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmSmemBytes));
  RJ_TRY(cudaFuncSetAttribute(k_scan_emit<kEmWindow, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmSmemBytes));
  RJ_TRY(cudaFuncSetAttribute(k_scan_emit<kEmGeneric, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmSmemBytes));
  RJ_TRY(cudaFuncSetAttribute(k_scan_emit<kEmGeneric, true, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmSmemBytes));
  c->attr_done = true;
  return true;
}

struct Slab {                 // how a launch maps local offsets to the whole text
  ScanRange own;
  uint64_t base_offset;       // added to every reported offset
};

constexpr int kFaithfulWalkers = 2048;

uint64_t FaithfulStride(const NfaTables& nfa) {
  uint64_t P = nfa.n_pos > 0 ? nfa.n_pos : 1;
  return ((2 * P * 8 + 3 * (uint64_t)nfa.words * 4) + 15) & ~15ull;
}

// Multi-CTA resolve of m candidates.  `b`/`e` are sorted by begin unless
// `needs_sort` (then they are the unordered buffers and get radix-sorted).
bool RunLargeResolve(DeviceContext* c, const uint64_t* b_in, const uint64_t* e_in, uint64_t m, uint64_t n_text,
                     bool needs_sort, const Carry& carry_in, uint64_t base_offset, uint64_t* d_out,
                     uint64_t out_cap, FaithfulArgs fa, RunStats* stats, std::string* error) {
  cudaStream_t s = c->stream;
  if (!c->ReserveScratch(m, error)) return false;
  const uint64_t* b = b_in;
  const uint64_t* e = e_in;
  size_t tmp_sort = 0, tmp_scan = 0, tmp_sum = 0;
  int end_bit = 1;
  while (end_bit < 64 && (n_text >> end_bit) != 0) ++end_bit;
  if (needs_sort) {
    if (!c->sorted_b.Reserve(m * 8, error) || !c->sorted_e.Reserve(m * 8, error)) return false;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, b_in, c->sorted_b.as<uint64_t>(), e_in,
                                    c->sorted_e.as<uint64_t>(), (int64_t)m, 0, end_bit, s);
  }
  cub::DeviceScan::ExclusiveScan(nullptr, tmp_scan, e_in, c->reach.as<uint64_t>(), MaxOp(), (uint64_t)carry_in.cur,
                                 (int64_t)m, s);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_sum, c->wide.as<uint64_t>(), c->slot.as<uint64_t>(), (int64_t)m, s);
  size_t tmp = std::max(tmp_sort, std::max(tmp_scan, tmp_sum));
  if (!c->cub_tmp.Reserve(tmp, error)) return false;
  if (needs_sort) {
    RJ_TRY(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, b_in, c->sorted_b.as<uint64_t>(), e_in,
                                           c->sorted_e.as<uint64_t>(), (int64_t)m, 0, end_bit, s));
    b = c->sorted_b.as<uint64_t>();
    e = c->sorted_e.as<uint64_t>();
    if (stats) stats->launches += 6;
  }
  RJ_TRY(cub::DeviceScan::ExclusiveScan(c->cub_tmp.p, tmp, e, c->reach.as<uint64_t>(), MaxOp(),
                                        (uint64_t)carry_in.cur, (int64_t)m, s));
  int blocks = (int)std::min<uint64_t>((m + 255) / 256, (uint64_t)c->sm_count * 8);
  if (fa.enabled) {
    if (!c->fscratch.Reserve(fa.scratch_stride * kFaithfulWalkers, error)) return false;
    fa.scratch = c->fscratch.as<uint8_t>();
    k_segment_faithful<<<kFaithfulWalkers / 64, 64, 0, s>>>(b, e, c->reach.as<uint64_t>(), m, fa,
                                                            c->take.as<uint32_t>(), c->fin_end.as<uint64_t>());
  } else {
    k_segment_chain<<<blocks, 256, 0, s>>>(b, e, c->reach.as<uint64_t>(), m, carry_in, c->take.as<uint32_t>(),
                                           c->fin_end.as<uint64_t>());
  }
  k_widen_flags<<<blocks, 256, 0, s>>>(c->take.as<uint32_t>(), c->wide.as<uint64_t>(), m);
  RJ_TRY(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->wide.as<uint64_t>(), c->slot.as<uint64_t>(), (int64_t)m, s));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  RJ_TRY(cudaMemsetAsync(ctr + 3, 0, 16, s));
  k_scatter_matches<<<blocks, 256, 0, s>>>(b, c->fin_end.as<uint64_t>(), c->take.as<uint32_t>(),
                                           c->slot.as<uint64_t>(), m, base_offset, d_out, out_cap, ctr + 3, ctr + 4);
  k_finish_large<<<1, 32, 0, s>>>(b, c->fin_end.as<uint64_t>(), c->take.as<uint32_t>(), c->slot.as<uint64_t>(), m,
                                  carry_in, ctr + 3, ctr + 4, c->status.as<PipelineStatus>());
  if (stats) { stats->launches += 9; stats->large_path = 1; }
  RJ_TRY(cudaGetLastError());
  return true;
}

struct StoreDims { uint64_t nsub; uint32_t cap; };

bool ReserveStore(Buffer* b, Buffer* e, Buffer* cnt, const StoreDims& d, std::string* error) {
  return b->Reserve(d.nsub * d.cap * 8, error) && e->Reserve(d.nsub * d.cap * 8, error) &&
         cnt->Reserve(d.nsub * 4 + 16, error);
}

// how many warps of k_dfa_tma fit next to the replicated table
size_t DfaTmaFixedSmem(const DfaTables& dfa) {
  return (size_t)dfa.n_states * dfa.n_classes * dfa.n_classes * 128 + (size_t)dfa.n_states * dfa.n_classes * 2 + 16 +
         256 + 8 * 32 + 256;
}
int DfaTmaWarps(const DeviceContext* c, const DfaTables& dfa) {
  size_t fixed = DfaTmaFixedSmem(dfa);
  if (fixed + 4 * kDfaTileBytes > c->smem_optin) return 0;
  size_t w = (c->smem_optin - fixed) / kDfaTileBytes;
  int cap = 18;
  if (const char* env = getenv("RJ_DFA_WARPS")) cap = std::max(4, std::min(18, atoi(env)));
  return (int)std::min<size_t>(w, (size_t)cap);
}

// Runs scan (+verify) + resolve for one slab whose text is at d_text[0..n).
// Matches are written to d_out (or the context's out_pairs) as global offsets.
// `rebuild` (optional, ReplaceAll): the single-pass generic scan writes the rebuilt text itself (scan_emit.cuh,
// EmitArgs::rep_out); done = it did (else the matches are in out_pairs as usual and the caller rebuilds).
struct FusedRebuild {
  uint8_t* d_out = nullptr;            // >= n + 64 bytes
  const uint8_t* d_with = nullptr;     // the replacement, device memory
  uint32_t with_len = 0;
  bool done = false;
  uint64_t removed = 0;                // bytes inside the matches
};

bool RunPipeline(DeviceContext* c, Program* prog, DeviceProgram* dp, const uint8_t* d_text, uint64_t n,
                 const Slab& slab, const Carry& carry_in, uint64_t* d_out, uint64_t out_cap,
                 PipelineStatus* result, RunStats* stats, std::string* error, FusedRebuild* rebuild = nullptr) {
  const CompiledAutomaton& ca = prog->automaton();
  cudaStream_t s = c->stream;
  RJ_TRY(cudaSetDevice(c->device));
  if ((reinterpret_cast<uintptr_t>(d_text) & 15) != 0) {
    if (error) *error = "rejit_b200: device text must be 16-byte aligned";
    return false;
  }
  if (!EnsureKernelAttributes(c, error)) return false;
  if (c->dense_cap == 0 && !c->ReserveDense(1u << 16, error)) return false;
  const int tma_warps = (ca.strategy == ScanStrategy::DfaFixed) ? DfaTmaWarps(c, dp->dfa) : 0;
  const uint32_t wsize = ca.window_hi - ca.window_lo + 1;

  for (int attempt = 0; attempt < 48; ++attempt) {
    unsigned long long* ctr = c->counters.as<unsigned long long>();
    PipelineStatus* d_status = c->status.as<PipelineStatus>();
    RJ_TRY(cudaMemsetAsync(d_status, 0, kFinSegOffset + 4 * (size_t)kFinMaxSegments, s));
    const bool use_fallback = ca.strategy == ScanStrategy::DfaFixed && (dp->dense_mode || tma_warps < 4);
    if (use_fallback && c->cand_cap == 0 && !c->ReserveUnordered(1u << 16, error)) return false;
    uint64_t ocap = d_out ? out_cap : c->out_pairs.bytes / 16;
    uint64_t* outp = d_out ? d_out : c->out_pairs.as<uint64_t>();

    FaithfulArgs fa{};
    fa.enabled = ca.reentrant ? 1 : 0;
    if (fa.enabled) {
      fa.nfa = dp->nfa;
      fa.text = d_text;
      fa.n = n;
      fa.scratch_stride = FaithfulStride(dp->nfa);
      if (!c->fscratch.Reserve(fa.scratch_stride * kFaithfulWalkers, error)) return false;
      fa.scratch = c->fscratch.as<uint8_t>();
    }
    DenseList dense{c->dense_b.as<uint64_t>(), c->dense_e.as<uint64_t>(), ctr + 1, c->dense_cap};
    ResolveScratch rs{c->reach.as<uint64_t>(), c->take.as<uint32_t>(), c->fin_end.as<uint64_t>(), c->slot.as<uint64_t>()};
    SubStore cand{};
    cand.cap = dp->cand_sub_cap;

    if (stats) cudaEventRecord(c->ev[0], s);
    const int grid_full = c->sm_count * 8;
    bool ordered = true;
    // The in-kernel finish only pays off when the candidates turn out to be the matches.  A pattern whose
    // candidates overlap by nature (several starts reach the same end) keeps failing the test: after a
    // miss the finish is skipped, and probed again every 32nd call.
    const bool try_fuse = c->coop && !fa.enabled && (dp->fuse_misses == 0 || (dp->fuse_skips & 31) == 31);
    if (c->coop && !fa.enabled && !try_fuse) dp->fuse_skips++;
    bool fused = false;                 // the scan kernel also produced the matches and the status
    bool rebuild_launched = false;      // ... and (ReplaceAll) the rebuilt text instead of the match list
    unsigned int fused_seq = 0, hits_seq = 0;
    // largest co-resident grid of a 256-thread kernel with the finish scratch (cached occupancy query)
    auto coop_blocks = [&](const void* fn, int* cache) -> int {
      if (*cache == 0) {
        int nb = 0;
        if (!Check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, 256, kFinScratchWords * 4), "occupancy", error)) return 0;
        *cache = std::max(1, nb);
      }
      return *cache * c->sm_count;
    };
    // arguments of the in-kernel finish (FinishOrdered) for a scan grid of `blocks` CTAs
    auto make_fin = [&](int blocks, uint64_t nsub) {
      FinishArgs fin{};
      uint8_t* base = static_cast<uint8_t*>(c->status.p);
      fin.enabled = 1;
      const int segs = std::min(blocks, (int)kFinMaxSegments);   // one segment per CTA
      fin.seg_subs = (uint32_t)((nsub + segs - 1) / segs);
      fin.nseg = (uint32_t)((nsub + fin.seg_subs - 1) / fin.seg_subs);
      fin.sync = reinterpret_cast<unsigned int*>(base + kFinSyncOffset);
      fin.last_end = reinterpret_cast<unsigned long long*>(base + kFinLastOffset);
      fin.totals = reinterpret_cast<unsigned long long*>(base + kFinTotalOffset);
      fin.last_ne = reinterpret_cast<unsigned long long*>(base + kFinLastNeOffset);
      fin.segcount = reinterpret_cast<uint32_t*>(base + kFinSegOffset);
      // a candidate can only touch a predecessor in the 32 preceding sub-regions when matches are
      // shorter than one sub-region (all scans here cut sub-regions by text offset)
      fin.local_pred = (ca.nfa.max_len != kInfLen && ca.nfa.max_len + 64 < kDfaSubBytes) ? 1 : 0;
      fin.out_pairs = outp;
      fin.out_stride = 0;
      fin.out_cap = ocap;
      fin.base_offset = slab.base_offset;
      fin.host_records = c->h_fin_dev;
      fin.seq = fused_seq = ++c->call_seq ? c->call_seq : ++c->call_seq;
      return fin;
    };
    // ---- single-pass scan + emit (scan_emit.cuh): the text is read once, the matches are written once ----
    const bool window_fits = ca.strategy != ScanStrategy::LiteralWindow || (uint64_t)ca.window_hi + 64 <= kEmBias;
    const bool use_emit = c->emit && try_fuse && !dp->emit_off && window_fits &&
                          (ca.strategy == ScanStrategy::Literal || ca.strategy == ScanStrategy::LiteralWindow ||
                           ca.strategy == ScanStrategy::Generic);
    if (use_emit) {
      EmitArgs em{};
      const uint64_t first_start = std::min<uint64_t>(slab.own.own_begin, n);
      uint64_t last_pos = slab.own.own_end ? slab.own.own_end - 1 : 0;            // last owned start
      if (ca.strategy == ScanStrategy::LiteralWindow) last_pos += (uint64_t)ca.window_hi + 1;   // ... or needle hit
      if (ca.strategy != ScanStrategy::Generic) last_pos += 3;      // a literal is tested by the lane that holds its fourth byte
      last_pos = std::min<uint64_t>(last_pos, n);
      // tile size: 32 KB for sparse matches; patterns whose every start is a candidate (generic scans, one-byte
      // literals) fill the per-tile candidate lists: 24 / 16 KB tiles measured 15-25 % faster there and 4-15 % slower on
      // sparse literals (gpurun_out/r2o_ab_rows*.txt)
      em.rows = ca.strategy == ScanStrategy::Generic ? 48u : (ca.strategy == ScanStrategy::Literal && dp->needle_len == 1 ? 32u : kEmRows);
      // (Tiles up to a quarter smaller so that the last round of tiles is full — 500 MB in 32 KB tiles is 3.22 rounds —
      // measured no better on the literal scans and 12 % worse on the generic ones: r2v against r2p.)
      const uint64_t tile_bytes = (uint64_t)em.rows * 512;
      em.tile0 = first_start / tile_bytes;
      em.ntiles = std::max<uint64_t>(last_pos / tile_bytes, em.tile0) - em.tile0 + 1;
      const uint64_t n_groups = (em.ntiles + 31) / 32;
      if (!c->em_records.Reserve((em.ntiles + n_groups) * 32 + 64, error)) return false;
      uint8_t* base = static_cast<uint8_t*>(c->status.p);
      em.records = c->em_records.as<uint4>();
      em.group_records = em.records + 2 * em.ntiles;
      em.sync = reinterpret_cast<unsigned int*>(base + kFinSyncOffset);
      em.final_state = reinterpret_cast<unsigned long long*>(base + kFinLastOffset);
      em.out_pairs = outp;
      em.out_cap = ocap;
      em.base_offset = slab.base_offset;
      em.host_records = c->h_fin_dev;
      em.seq = fused_seq = ++c->call_seq ? c->call_seq : ++c->call_seq;
      em.carry_in = carry_in;
      const bool fuse_rebuild = rebuild && rebuild->d_out && ca.strategy == ScanStrategy::Generic;
      if (fuse_rebuild) { em.rep_out = rebuild->d_out; em.rep_with = rebuild->d_with; em.rep_w = rebuild->with_len; }
      rebuild_launched = fuse_rebuild;
      EmLit lit{};
      lit.needle = dp->needle;
      lit.m = dp->needle_len; lit.p4 = dp->p4; lit.pmask = dp->pmask;
      lit.win_lo = ca.window_lo; lit.win_hi = ca.window_hi;
      // (eight rows in flight — 80 registers, three CTAs per SM — measured 4.1 TB/s against 4.5 TB/s for four rows and four
      // CTAs on a 2 GB text: the literal filter is bound by the integer pipe, not by the loads in flight)
      const void* fn = ca.strategy == ScanStrategy::Literal
                           ? (dp->needle_len >= 4 ? (const void*)k_scan_emit<kEmLiteral, true, 4> : (const void*)k_scan_emit<kEmLiteral, false, 4>)
                           : ca.strategy == ScanStrategy::LiteralWindow ? (const void*)k_scan_emit<kEmWindow, true, 4>
                           : fuse_rebuild ? (const void*)k_scan_emit<kEmGeneric, true, 4, true>
                                          : (const void*)k_scan_emit<kEmGeneric, true, 4>;
      // CTAs that are resident together (the round-robin deal needs all of them; four per SM by construction, asked once)
      if (c->em_blocks_per_sm == 0) {
        int nb = 4;
        const void* all[5] = {(const void*)k_scan_emit<kEmLiteral, true, 4>, (const void*)k_scan_emit<kEmLiteral, false, 4>,
                              (const void*)k_scan_emit<kEmWindow, true, 4>, (const void*)k_scan_emit<kEmGeneric, true, 4>,
                              (const void*)k_scan_emit<kEmGeneric, true, 4, true>};
        for (const void* f : all) {
          int k = 0;
          RJ_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, f, (int)kEmThreads, kEmSmemBytes));
          nb = std::min(nb, k);
        }
        c->em_blocks_per_sm = std::max(1, nb);
      }
      const int blocks = (int)std::min<uint64_t>((em.ntiles + kEmWarps - 1) / kEmWarps, (uint64_t)c->sm_count * c->em_blocks_per_sm);
      // Round robin needs every CTA of the grid resident at once: the grid is never larger than what the occupancy
      // calculator says fits.  A plain launch, not a cooperative one: cudaLaunchCooperativeKernel costs this kernel
      // 20 us (0.130 vs 0.110 ms for 500 MB), and a look-back that waits in vain (another client of the GPU holding SMs
      // for good) times out, raises kFinStuck and sends the call to the general path.  RJ_EM_TICKETS=1: the ticket counter.
      static const bool tickets = getenv("RJ_EM_TICKETS") != nullptr;
      em.static_stride = tickets ? 0u : (uint32_t)blocks * kEmWarps;
      uint64_t n_arg = n;
      ScanRange own_arg = slab.own;
      void* kargs[] = {(void*)&d_text, (void*)&n_arg, (void*)&lit, (void*)&dp->nfa, (void*)&dp->em_filter, (void*)&own_arg, (void*)&em};
      RJ_TRY(cudaLaunchKernel(fn, dim3(blocks), dim3(kEmThreads), kargs, kEmSmemBytes, s));
      fused = true;
      if (stats) stats->launches += 1;
    } else
    switch (ca.strategy) {
      case ScanStrategy::Literal: {
        cand.nsub = (n + kLitSubBytes - 1) / kLitSubBytes;
        if (cand.nsub == 0) cand.nsub = 1;
        if (!ReserveStore(&c->sub_b, &c->sub_e, &c->sub_count, {cand.nsub, cand.cap}, error)) return false;
        cand.begin = c->sub_b.as<uint64_t>(); cand.end = c->sub_e.as<uint64_t>(); cand.count = c->sub_count.as<uint32_t>();
        int blocks = (int)std::min<uint64_t>((cand.nsub + 7) / 8, (uint64_t)grid_full);
        const bool full4 = dp->needle_len >= 4;
        FinishArgs fin{};
        if (try_fuse && dp->needle_len + 64 < kLitSubBytes) {
          // finish in-kernel: the grid must be co-resident for the barrier
          if (c->lit_blocks_per_sm == 0) {
            int nb = 0;
            RJ_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_lit_scan<true>, 256, kFinScratchWords * 4));
            int nb2 = 0;
            RJ_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, k_lit_scan<false>, 256, kFinScratchWords * 4));
            c->lit_blocks_per_sm = std::max(1, std::min(nb, nb2));
          }
          blocks = std::min(blocks, c->lit_blocks_per_sm * c->sm_count);
          fin = make_fin(blocks, cand.nsub);
          fused = true;
          const uint8_t* needle = dp->needle;
          uint32_t nl = dp->needle_len, p4 = dp->p4, pmask = dp->pmask;
          ScanRange own = slab.own;
          Carry c0 = carry_in;
          void* args[] = {(void*)&d_text, (void*)&n, (void*)&needle, (void*)&nl, (void*)&p4, (void*)&pmask,
                          (void*)&own, (void*)&cand, (void*)&fin, (void*)&c0};
          const void* fn = full4 ? (const void*)k_lit_scan<true> : (const void*)k_lit_scan<false>;
          RJ_TRY(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(256), args, kFinScratchWords * 4, s));
        } else if (full4) {
          k_lit_scan<true><<<blocks, 256, 0, s>>>(d_text, n, dp->needle, dp->needle_len, dp->p4, dp->pmask, slab.own, cand, fin, carry_in);
        } else {
          k_lit_scan<false><<<blocks, 256, 0, s>>>(d_text, n, dp->needle, dp->needle_len, dp->p4, dp->pmask, slab.own, cand, fin, carry_in);
        }
        if (stats) stats->launches += 1;
        break;
      }
      case ScanStrategy::LiteralWindow: {
        SubStore hs{};
        hs.cap = dp->hit_sub_cap;
        hs.nsub = std::max<uint64_t>(1, (n + kLitSubBytes - 1) / kLitSubBytes);
        if (!ReserveStore(&c->hsub_b, &c->hsub_e, &c->hsub_count, {hs.nsub, hs.cap}, error)) return false;
        hs.begin = c->hsub_b.as<uint64_t>(); hs.end = c->hsub_e.as<uint64_t>(); hs.count = c->hsub_count.as<uint32_t>();
        if (c->hits_cap == 0 && !c->ReserveHits(1u << 14, error)) return false;
        DenseList hits{c->hits_b.as<uint64_t>(), c->hits_e.as<uint64_t>(), ctr + 2, c->hits_cap};
        // the needle may sit up to window_hi bytes after an owned start
        ScanRange hit_range{slab.own.own_begin, slab.own.own_end + ca.window_hi + 1};
        int blocks = (int)std::min<uint64_t>((hs.nsub + 7) / 8, (uint64_t)grid_full);
        k_lit_scan<true><<<blocks, 256, 0, s>>>(d_text, n, dp->needle, dp->needle_len, dp->p4, dp->pmask, hit_range, hs,
                                                FinishArgs{}, Carry());
        if (stats) cudaEventRecord(c->ev[1], s);
        const bool win_fused = try_fuse;
        if (win_fused) hits_seq = ++c->call_seq ? c->call_seq : ++c->call_seq;
        k_gather_hits<<<1, 512, 0, s>>>(hs, hits, d_status, win_fused ? c->h_fin_dev + 31 : nullptr, hits_seq);
        cand.cap = kWinSubHits * wsize;
        cand.nsub = (c->hits_cap + kWinSubHits - 1) / kWinSubHits;
        if (!ReserveStore(&c->sub_b, &c->sub_e, &c->sub_count, {cand.nsub, cand.cap}, error)) return false;
        cand.begin = c->sub_b.as<uint64_t>(); cand.end = c->sub_e.as<uint64_t>(); cand.count = c->sub_count.as<uint32_t>();
        int wblocks = (int)std::min<uint64_t>((cand.nsub + 7) / 8, (uint64_t)c->sm_count * 4);
        FinishArgs fin{};
        if (try_fuse) {
          wblocks = std::min(wblocks, coop_blocks((const void*)k_window_verify, &c->win_blocks_per_sm));
          if (wblocks < 1) return false;
          fin = make_fin(wblocks, cand.nsub);
          fin.local_pred = 0;                    // sub-regions of this store are cut by hit index, not by offset
          fused = true;
          uint32_t wlo = ca.window_lo, whi = ca.window_hi;
          ScanRange own = slab.own;
          Carry c0 = carry_in;
          void* args[] = {(void*)&d_text, (void*)&n, (void*)&dp->nfa, (void*)&hits, (void*)&wlo, (void*)&whi, (void*)&own,
                          (void*)&cand, (void*)&fin, (void*)&c0};
          RJ_TRY(cudaLaunchCooperativeKernel((const void*)k_window_verify, dim3(wblocks), dim3(256), args, kFinScratchWords * 4, s));
        } else {
          k_window_verify<<<wblocks, 256, 0, s>>>(d_text, n, dp->nfa, hits, ca.window_lo, ca.window_hi, slab.own, cand, fin, carry_in);
        }
        if (stats) stats->launches += 3;
        break;
      }
      case ScanStrategy::DfaFixed: {
        if (!use_fallback) {
          cand.nsub = std::max<uint64_t>(1, (n + kDfaSubBytes - 1) / kDfaSubBytes);
          if (!ReserveStore(&c->sub_b, &c->sub_e, &c->sub_count, {cand.nsub, cand.cap}, error)) return false;
          cand.begin = c->sub_b.as<uint64_t>(); cand.end = c->sub_e.as<uint64_t>(); cand.count = c->sub_count.as<uint32_t>();
          size_t smem = DfaTmaFixedSmem(dp->dfa) + (size_t)tma_warps * kDfaTileBytes;
          int blocks = (int)std::min<uint64_t>((cand.nsub + tma_warps - 1) / tma_warps, (uint64_t)c->sm_count);
          FinishArgs fin{};
          CarrySet cset{};
          cset.c[0] = carry_in;
          unsigned int* dense_flag = &d_status->dense;
          unsigned long long* work = ctr + 0;
          if (try_fuse) {
            // the scan grid finishes the job itself (see FinishFixed)
            fin = make_fin(blocks, cand.nsub);
            dense_flag = fin.sync + 4;
            ScanRange own = slab.own;
            void* args[] = {(void*)&d_text, (void*)&n, (void*)&dp->dfa, (void*)&own, (void*)&cand, (void*)&dense_flag,
                            (void*)&work, (void*)&fin, (void*)&cset};
            RJ_TRY(cudaLaunchCooperativeKernel((const void*)k_dfa_tma, dim3(blocks), dim3(tma_warps * 32), args, smem, s));
            fused = true;
          } else {
            k_dfa_tma<<<blocks, tma_warps * 32, smem, s>>>(d_text, n, dp->dfa, slab.own, cand, dense_flag, work, fin, cset);
          }
        } else {
          ordered = false;
          CandBuf un{c->cand_b.as<uint64_t>(), c->cand_e.as<uint64_t>(), ctr, c->cand_cap};
          const uint32_t stream_bytes = 256;
          uint64_t streams = (n + stream_bytes - 1) / stream_bytes;
          int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((streams + 255) / 256, (uint64_t)grid_full));
          k_dfa_scan<<<blocks, 256, dp->dfa_smem, s>>>(d_text, n, dp->dfa, stream_bytes, slab.own, un);
        }
        if (stats) stats->launches += 1;
        break;
      }
      case ScanStrategy::Generic: {
        cand.nsub = (n + 1 + kGenSubBytes - 1) / kGenSubBytes;
        cand.cap = std::min<uint32_t>(std::max<uint32_t>(cand.cap, 64), kGenSubBytes);
        if (!ReserveStore(&c->sub_b, &c->sub_e, &c->sub_count, {cand.nsub, cand.cap}, error)) return false;
        cand.begin = c->sub_b.as<uint64_t>(); cand.end = c->sub_e.as<uint64_t>(); cand.count = c->sub_count.as<uint32_t>();
        int blocks = (int)std::min<uint64_t>((cand.nsub + 7) / 8, (uint64_t)grid_full);
        FinishArgs fin{};
        if (try_fuse) {
          blocks = std::min(blocks, coop_blocks((const void*)k_generic_scan, &c->gen_blocks_per_sm));
          if (blocks < 1) return false;
          fin = make_fin(blocks, cand.nsub);
          fin.local_pred = (ca.nfa.max_len != kInfLen && ca.nfa.max_len + 64 < kGenSubBytes) ? 1 : 0;
          fused = true;
          ScanRange own = slab.own;
          Carry c0 = carry_in;
          void* args[] = {(void*)&d_text, (void*)&n, (void*)&dp->nfa, (void*)&dp->gen_filter, (void*)&own, (void*)&cand,
                          (void*)&fin, (void*)&c0};
          RJ_TRY(cudaLaunchCooperativeKernel((const void*)k_generic_scan, dim3(blocks), dim3(256), args, kFinScratchWords * 4, s));
        } else {
          k_generic_scan<<<blocks, 256, 0, s>>>(d_text, n, dp->nfa, dp->gen_filter, slab.own, cand, fin, carry_in);
        }
        if (stats) stats->launches += 1;
        break;
      }
    }
    if (stats && (use_emit || ca.strategy != ScanStrategy::LiteralWindow)) cudaEventRecord(c->ev[1], s);
    PipelineStatus st;
    // spin on the mapped status block; fall back to the stream state if the
    // kernel cannot have run (launch failure, sticky error)
    auto wait_status = [&](unsigned int seq) -> bool {
      volatile unsigned int* vseq = &c->h_status->seq;
      uint64_t spins = 0;
      while (*vseq != seq) {
        if ((++spins & 0x3FFF) == 0) {
          cudaError_t q = cudaStreamQuery(s);
          if (q == cudaSuccess) { if (*vseq == seq) break; if (error) *error = "rejit_b200: resolve kernel did not report"; return false; }
          if (q != cudaErrorNotReady) { Check(q, "cudaStreamQuery", error); return false; }
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      st = *c->h_status;
      return true;
    };
    auto resolve_ordered = [&]() -> bool {
      const unsigned int seq = ++c->call_seq ? c->call_seq : ++c->call_seq;
      k_resolve_ordered<<<1, 512, 0, s>>>(cand, dense, rs, carry_in, slab.base_offset, outp, ocap, fa, d_status,
                                          c->h_status_dev, seq);
      if (stats) stats->launches += 1;
      RJ_TRY(cudaGetLastError());
      return wait_status(seq);
    };
    if (ordered && fused) {
      RJ_TRY(cudaGetLastError());
      if (!WaitFinRecords(c, rebuild_launched ? 2 : 1, fused_seq, error)) return false;
      st = StatusFromRecord(c->h_fin[0], carry_in);
      if (use_emit) {
        // the single-pass scan reports through the same record; when a chain crossed a tile edge (need_large) or a
        // tile was too dense for its lists, the round-1 pipeline runs instead (next attempt)
        if (st.dense) { dp->emit_off = true; if (stats) stats->reruns += 1; continue; }
        if (st.need_large) { dp->fuse_misses++; dp->fuse_skips = 0; if (stats) stats->reruns += 1; continue; }
        dp->fuse_misses = 0;
        if (rebuild_launched && !st.overflow) {
          // the rebuilt text is in place: no match list to make room for, nothing more to run
          rebuild->done = true;
          rebuild->removed = c->h_fin[1].n_matches;
          if (stats) {
            cudaEventRecord(c->ev[2], s);
            cudaEventSynchronize(c->ev[2]);
            cudaEventElapsedTime(&stats->scan_ms, c->ev[0], c->ev[1]);
            cudaEventElapsedTime(&stats->total_ms, c->ev[0], c->ev[2]);
            stats->candidates = st.n_candidates;
            stats->matches = st.n_matches;
            stats->strategy = (int)ca.strategy;
          } else {
            RJ_TRY(cudaStreamSynchronize(s));           // (the report leaves a few microseconds before the last warps do)
          }
          *result = st;
          return true;
        }
      } else if (ca.strategy == ScanStrategy::LiteralWindow) {
        // outcome of the hit stage (k_gather_hits), published the same way in slot 31
        volatile FinRecord* hr = c->h_fin + 31;
        uint64_t spins = 0;
        while (hr->seq0 != hits_seq || hr->seq1 != hits_seq || hr->seq2 != hits_seq)
          if ((++spins & 0x3FFFFF) == 0 && cudaStreamQuery(s) != cudaErrorNotReady) break;
        if (hr->seq0 != hits_seq) { if (error) *error = "rejit_b200: the hit stage did not report"; return false; }
        st.n_hits = hr->n_matches;
        if (hr->flags & kFinOverflow) { st.overflow = 1; st.need_cap = std::max(st.need_cap, (unsigned int)hr->need_cap); }
      }
      if (!use_emit && !st.overflow && !st.dense) {
        if (st.need_large) { dp->fuse_misses++; dp->fuse_skips = 0; } else { dp->fuse_misses = 0; }
      }
      if (!use_emit && st.need_large && !st.overflow && !st.dense) {
        // neighbouring candidates overlap: the general resolve decides
        RJ_TRY(cudaMemsetAsync(d_status, 0, sizeof(PipelineStatus), s));
        if (!resolve_ordered()) return false;
      }
    } else if (ordered) {
      if (!resolve_ordered()) return false;
    } else {
      CandBuf un{c->cand_b.as<uint64_t>(), c->cand_e.as<uint64_t>(), ctr, c->cand_cap};
      k_resolve_small<<<1, 1024, kSmallResolveMax * 16, s>>>(un, carry_in, slab.base_offset, outp, ocap, d_status);
      if (stats) stats->launches += 1;
      RJ_TRY(cudaGetLastError());
      RJ_TRY(cudaMemcpyAsync(c->h_status, d_status, sizeof(PipelineStatus), cudaMemcpyDeviceToHost, s));
      RJ_TRY(cudaStreamSynchronize(s));
      st = *c->h_status;
    }

    // ---- capacity protocol: grow what overflowed and run again -----------------
    bool rerun = false;
    if (st.dense && !dp->dense_mode) { dp->dense_mode = true; rerun = true; }
    if (ca.strategy == ScanStrategy::LiteralWindow && st.n_hits > c->hits_cap) {
      if (!c->ReserveHits(st.n_hits + st.n_hits / 4 + 1024, error)) return false;
      rerun = true;
    }
    if (st.overflow && st.need_cap) {
      uint32_t want = std::max<uint32_t>(st.need_cap, 2 * (ca.strategy == ScanStrategy::LiteralWindow ? dp->hit_sub_cap : dp->cand_sub_cap));
      if (ca.strategy == ScanStrategy::LiteralWindow) dp->hit_sub_cap = want; else dp->cand_sub_cap = want;
      rerun = true;
    }
    if (ordered && st.n_candidates > c->dense_cap) {
      if (!c->ReserveDense(st.n_candidates + st.n_candidates / 8 + 1024, error)) return false;
      rerun = true;
    }
    if (!ordered && (st.overflow || st.n_candidates > c->cand_cap)) {
      uint64_t want = std::max<uint64_t>(st.n_candidates + st.n_candidates / 8 + 1024, c->cand_cap * 2);
      if (!c->ReserveUnordered(want, error) || !c->ReserveDense(want, error)) return false;
      rerun = true;
    }
    if (st.overflow && !rerun) rerun = true;
    if (rerun) {
      if (stats) stats->reruns += 1;
      continue;
    }
    if (st.need_large && ordered) {
      // many candidates: gather the slot ranges with the whole GPU
      if (!c->wide.Reserve(cand.nsub * 8 + 64, error) || !c->slot.Reserve(cand.nsub * 8 + 64, error)) return false;
      size_t tmp = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->wide.as<uint64_t>(), c->slot.as<uint64_t>(), (int64_t)cand.nsub, s);
      if (!c->cub_tmp.Reserve(tmp, error)) return false;
      int gblocks = (int)std::min<uint64_t>((cand.nsub + 255) / 256, (uint64_t)c->sm_count * 8);
      k_clamp_counts<<<gblocks, 256, 0, s>>>(cand.count, cand.cap, cand.nsub, c->wide.as<uint64_t>());
      RJ_TRY(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->wide.as<uint64_t>(), c->slot.as<uint64_t>(), (int64_t)cand.nsub, s));
      int wblocks = (int)std::min<uint64_t>((cand.nsub + 7) / 8, (uint64_t)c->sm_count * 8);
      k_gather_multi<<<wblocks, 256, 0, s>>>(cand, c->slot.as<uint64_t>(), dense);
      if (stats) stats->launches += 4;
    }
    if (st.need_large) {
      outp = d_out ? d_out : c->out_pairs.as<uint64_t>();
      const uint64_t* bb = ordered ? c->dense_b.as<uint64_t>() : c->cand_b.as<uint64_t>();
      const uint64_t* ee = ordered ? c->dense_e.as<uint64_t>() : c->cand_e.as<uint64_t>();
      if (!RunLargeResolve(c, bb, ee, st.n_candidates, n, !ordered, carry_in, slab.base_offset, outp, ocap, fa, stats, error))
        return false;
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
// Host text -> device, on the context's stream.  Pinned (or registered) memory goes in one asynchronous copy.
// Pageable memory — what a caller of Regej::MatchAll(const char*, size_t, ...) hands over — is staged through a
// pinned arena: worker threads only memcpy (1 MB chunks, claimed in order), and whichever worker finishes the next
// chunk in line issues the DMA for every consecutive chunk that is ready — one thread in the driver at a time, copies
// in order, merged when several chunks are ready.  (Round 2a let every worker wait on an event, memcpy and enqueue
// its own copy: past four threads they queued on the driver's lock and the upload got slower.)
bool UploadHostText(DeviceContext* c, void* d_dst, const uint8_t* src, size_t len, std::string* error) {
  if (!len) return true;
  cudaPointerAttributes attr{};
  const bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (pinned || len < (4u << 20)) {
    RJ_TRY(cudaMemcpyAsync(d_dst, src, len, cudaMemcpyHostToDevice, c->stream));
    return true;
  }
  constexpr size_t kChunk = 1u << 20;
  constexpr size_t kArenaMax = 256u << 20;                  // longer texts go through the arena in rounds
  const size_t want = std::min(kArenaMax, (len + kChunk - 1) / kChunk * kChunk);
  if (c->arena_bytes < want) {
    if (c->arena) { RJ_TRY(cudaStreamSynchronize(c->stream)); cudaFreeHost(c->arena); c->arena = nullptr; c->arena_bytes = 0; }
    RJ_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->arena), want, cudaHostAllocDefault));
    c->arena_bytes = want;
    if (!c->arena_ev) RJ_TRY(cudaEventCreateWithFlags(&c->arena_ev, cudaEventDisableTiming));
  }
  static const int width = getenv("RJ_STAGE_WIDTH") ? std::max(1, atoi(getenv("RJ_STAGE_WIDTH"))) : 4;     // (measured: 2: 21, 4: 30, 6: 25, 8: 21, 14: 17 GB/s on a 16-core host)
  const int device = c->device;
  for (size_t round_lo = 0; round_lo < len; round_lo += c->arena_bytes) {
    const size_t round_len = std::min(c->arena_bytes, len - round_lo);
    // the arena's previous contents have left for the device
    if (c->arena_used) RJ_TRY(cudaEventSynchronize(c->arena_ev));
    const int n_chunks = (int)((round_len + kChunk - 1) / kChunk);
    std::unique_ptr<std::atomic<uint8_t>[]> ready(new std::atomic<uint8_t>[n_chunks]);
    for (int i = 0; i < n_chunks; ++i) ready[i].store(0, std::memory_order_relaxed);
    std::mutex issue_mu;
    int cursor = 0;                                         // next chunk to hand to the DMA engine (under issue_mu)
    std::atomic<int> failed{0};
    auto issue = [&]() {                                    // (issue_mu held)
      while (cursor < n_chunks && ready[cursor].load(std::memory_order_acquire)) {
        int j = cursor + 1;
        while (j < n_chunks && j - cursor < 8 && ready[j].load(std::memory_order_acquire)) ++j;
        const size_t off = (size_t)cursor * kChunk, n = std::min((size_t)j * kChunk, round_len) - off;
        if (cudaMemcpyAsync(static_cast<uint8_t*>(d_dst) + round_lo + off, c->arena + off, n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
          failed = 1;
        cursor = j;
      }
    };
    WorkerPool::Staging().Run(n_chunks, width, [&](int i) {
      const size_t off = (size_t)i * kChunk, n = std::min(kChunk, round_len - off);
      memcpy(c->arena + off, src + round_lo + off, n);
      ready[i].store(1, std::memory_order_release);
      // hand over what is ready, unless another thread is doing just that (it will see this chunk too: it re-checks
      // `ready` under the lock, and the last chunk is always followed by the drain below)
      if (issue_mu.try_lock()) {
        cudaSetDevice(device);
        issue();
        issue_mu.unlock();
      }
    });
    {
      std::lock_guard<std::mutex> lk(issue_mu);
      issue();
    }
    c->arena_used = true;
    if (failed.load() || cursor != n_chunks || cudaEventRecord(c->arena_ev, c->stream) != cudaSuccess) {
      Check(cudaGetLastError(), "staged upload", error);
      if (error && error->empty()) *error = "rejit_b200: staged upload failed";
      return false;
    }
  }
  return true;
}

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
  if (!UploadHostText(c, c->text.p, text + lo, len, error)) return false;
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

// MatchFirst / MatchAnywhere with early exit (the reference
// returns at the first registered match, x64/codegen-x64.cc:427-431, 481-486):
// the text is searched in growing slabs — only the slab (plus halo) is copied to
// the device and scanned — and the search stops at the first slab that holds a
// match.  With no match before a slab, the first match found in it is
// MatchAll()[0] (the first candidate of a chain is always selected).
int MatchFirstHost(int device, Program* prog, const uint8_t* text, uint64_t n, uint64_t pair[2], std::string* error) {
  DeviceContext* c = ContextFor(device, error);
  if (!c) return -1;
  DeviceProgram* dp = prog->OnDevice(device, error);
  if (!dp) return -1;
  const CompiledAutomaton& ca = prog->automaton();
  uint64_t step = 1ull << 18;
  if (ca.reentrant || ca.nfa.max_len == kInfLen) step = n + 1;      // one pass (the halo would be the whole text)
  uint64_t lo = 0;
  for (;;) {
    const uint64_t hi = (n - lo <= step) ? n : lo + step;
    const bool last = hi == n;
    Carry in{lo, kNoMatch}, out;
    std::vector<uint64_t> found;
    if (!MatchSlabFromHost(c, prog, dp, text, n, lo, hi, last, in, &out, &found, nullptr, error)) return -1;
    if (!found.empty()) {
      if (pair) { pair[0] = found[0]; pair[1] = found[1]; }
      return 1;
    }
    if (last) return 0;
    lo = hi;
    step *= 8;
  }
}

int64_t MatchAllResident(int device, Program* prog, const uint8_t* d_text, uint64_t n, uint64_t** pairs,
                         RunStats* stats, std::string* error) {
  DeviceContext* c = ContextFor(device, error);
  if (!c) return -1;
  DeviceProgram* dp = prog->OnDevice(device, error);
  if (!dp) return -1;
  std::lock_guard<std::mutex> lk(c->mu);
  Slab slab{{0, n + 1}, 0};
  PipelineStatus st;
  Carry in;
  if (!RunPipeline(c, prog, dp, d_text, n, slab, in, nullptr, 0, &st, stats, error)) return -1;
  uint64_t cnt = st.n_matches;
  *pairs = static_cast<uint64_t*>(malloc(std::max<size_t>(cnt * 16, 8)));
  if (cnt) {
    if (!Check(cudaMemcpyAsync(*pairs, c->out_pairs.p, cnt * 16, cudaMemcpyDeviceToHost, c->stream), "D2H", error) ||
        !Check(cudaStreamSynchronize(c->stream), "sync", error)) {
      free(*pairs);
      *pairs = nullptr;
      return -1;
    }
  }
  return (int64_t)cnt;
}

// ===========================================================================
// fused pattern sets
// ===========================================================================
class DeviceSet {
 public:
  SetTables tb{};
  KmerTables km{};
  bool kmer = false;                  // k_set_kmer instead of k_set_tma (SetDfa::Kmer)
  bool kmer_off = false;              // ... until hits were too dense for it twice in a row
  int kmer_dense = 0;
  std::vector<void*> allocs;
  size_t fixed_smem = 0;
  uint32_t sub_cap = 16;
  uint32_t stage_cap = 64;            // k_set_kmer: staged matches per warp beyond its shared list
  uint64_t per_cap = 1u << 14;        // dense / output capacity per pattern
  bool dense_mode = false;
  ~DeviceSet() { for (void* p : allocs) cudaFree(p); }
  template <class T>
  bool Upload(const T* host, size_t count, const T** dev, std::string* error) {
    void* p = nullptr;
    RJ_TRY(cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16)));
    allocs.push_back(p);
    RJ_TRY(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<const T*>(p);
    return true;
  }
};

SetProgram* SetProgram::Create(const std::vector<Program*>& members) {
  SetProgram* sp = new SetProgram();
  sp->members_ = members;
  std::vector<const CompiledAutomaton*> autos;
  for (Program* p : members) autos.push_back(&p->automaton());
  sp->fused_ = BuildSetDfa(autos, &sp->dfa_);
  if (sp->fused_) {
    sp->describe_ = "fused set: " + std::to_string(members.size()) + " patterns, one DFA of " +
                    std::to_string(sp->dfa_.n_states) + " states (+" + std::to_string(sp->dfa_.n_rows - sp->dfa_.n_states) +
                    " shadow rows) x " + std::to_string(sp->dfa_.n_classes) + " classes";
    if (sp->dfa_.kmer.ok) sp->describe_ += "; k-mer index (2-bit codes, (byte >> " + std::to_string(sp->dfa_.kmer.shift) + ") & 3)";
  } else {
    sp->describe_ = "set of " + std::to_string(members.size()) + " patterns run one by one (not fusable)";
  }
  return sp;
}
SetProgram::~SetProgram() { for (auto* d : per_device_) delete d; }

DeviceSet* SetProgram::OnDevice(int device, std::string* error) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!per_device_[device]) {
    if (!Check(cudaSetDevice(device), "cudaSetDevice", error)) return nullptr;
    DeviceSet* d = new DeviceSet();
    const SetDfa& f = dfa_;
    if (!d->Upload(f.t1.data(), f.t1.size(), &d->tb.t1, error) || !d->Upload(f.t2.data(), f.t2.size(), &d->tb.t2, error) ||
        !d->Upload(f.byte_class.data(), 256, &d->tb.byte_class, error) ||
        !d->Upload(f.accept_mask.data(), f.accept_mask.size(), &d->tb.accept_mask, error)) { delete d; return nullptr; }
    for (int j = 0; j < f.n_patterns; ++j) d->tb.match_len[j] = f.match_len[j];
    d->tb.n_patterns = f.n_patterns;
    d->tb.n_rows = f.n_rows;
    d->tb.n_classes = f.n_classes;
    d->tb.first_accept = f.first_accept;
    d->tb.row_shift = f.row_shift;
    d->fixed_smem = f.t2.size() * 4 + f.accept_mask.size() * 4 + f.t1.size() * 2 + 16 + 3 * 256 + 8 * 32 + 512;
    static const bool no_kmer = getenv("RJ_NO_KMER") != nullptr;
    if (f.kmer.ok && !no_kmer) {
      if (!d->Upload(f.kmer.bitmap.data(), f.kmer.bitmap.size(), &d->km.bitmap, error) ||
          !d->Upload(f.kmer.mask16.data(), f.kmer.mask16.size(), &d->km.mask16, error)) { delete d; return nullptr; }
      d->km.shift = f.kmer.shift;
      d->km.field_mask = 0x03030303u << f.kmer.shift;
      d->km.mult = 0x01041040u >> f.kmer.shift;
      // a code without a live byte gets a byte that has ANOTHER code: never equal to a text byte with this code
      d->km.canon = f.kmer.canon;
      for (uint32_t cd = 0; cd < 4; ++cd)
        if (!((f.kmer.canon_ok >> cd) & 1u)) d->km.canon |= ((((cd + 1) & 3u) << f.kmer.shift) & 0xFFu) << (8 * cd);
      // the small tables as one array of words (kKmerTab*); few accepted 8-mers: every CTA builds the bitmap and
      // the member-mask hash from the list
      std::vector<uint32_t> tab(kKmerTabWords, 0);
      for (int j = 0; j < f.n_patterns; ++j) tab[kKmerTabMatchLen + j] = f.match_len[j];
      for (int v = 0; v <= 8; ++v) tab[kKmerTabLenLe + v] = f.kmer.len_le[v];
      static const bool no_list = getenv("RJ_KMER_NO_LIST") != nullptr;
      std::vector<uint32_t> accepted;
      for (uint32_t x = 0; x < 65536 && accepted.size() <= kKmerListMax; ++x) if (f.kmer.mask16[x]) accepted.push_back(x);
      if (!no_list && !accepted.empty() && accepted.size() <= kKmerListMax) {
        d->km.n_list = (uint32_t)accepted.size();
        for (size_t i = 0; i < accepted.size(); ++i) {
          tab[kKmerTabListX + i] = accepted[i];
          tab[kKmerTabListMask + i] = f.kmer.mask16[accepted[i]];
        }
      }
      if (!d->Upload(tab.data(), tab.size(), &d->km.tab, error)) { delete d; return nullptr; }
      d->kmer = true;
    }
    per_device_[device] = d;
  }
  return per_device_[device];
}

int MatchAllSetResident(int device, SetProgram* set, const uint8_t* d_text, uint64_t n, int64_t* counts,
                        uint64_t** pairs, RunStats* stats, std::string* error, const SlabView* own_view,
                        const Carry* carry_in, Carry* carry_out, StitchCall* stitch) {
  DeviceContext* c = ContextFor(device, error);
  if (!c) return -1;
  const int K = set->size();
  auto one_by_one = [&]() -> int {
    RunStats acc;
    for (int j = 0; j < K; ++j) {
      RunStats rs;
      uint64_t* pj = nullptr;
      int64_t r;
      if (own_view || carry_in || carry_out) {
        Carry in = carry_in ? carry_in[j] : Carry();
        Carry out;
        r = MatchAllDevice(device, set->members()[j], d_text, n, nullptr, 0, in, &out, stats ? &rs : nullptr, error, own_view);
        if (carry_out) carry_out[j] = out;
        pj = nullptr;
      } else {
        r = MatchAllResident(device, set->members()[j], d_text, n, &pj, stats ? &rs : nullptr, error);
      }
      if (r < 0) return -1;
      counts[j] = r;
      if (pairs) pairs[j] = pj; else free(pj);
      acc.scan_ms += rs.scan_ms; acc.total_ms += rs.total_ms; acc.launches += rs.launches;
      acc.candidates += rs.candidates; acc.matches += rs.matches; acc.reruns += rs.reruns;
    }
    if (stats) { *stats = acc; stats->strategy = -1; }
    return 0;
  };
  if (!set->fused()) return one_by_one();
  DeviceSet* ds = set->OnDevice(device, error);
  if (!ds) return -1;
  if (ds->dense_mode) return one_by_one();
  if (ds->fixed_smem + 4 * kDfaTileBytes > c->smem_optin) { ds->dense_mode = true; return one_by_one(); }
  {
    std::lock_guard<std::mutex> lk(c->mu);
    if (!Check(cudaSetDevice(c->device), "cudaSetDevice", error)) return -1;
    if ((reinterpret_cast<uintptr_t>(d_text) & 15) != 0) { if (error) *error = "rejit_b200: device text must be 16-byte aligned"; return -1; }
    cudaStream_t s = c->stream;
    if (!EnsureKernelAttributes(c, error)) return -1;
    // what a finished call hands back (statuses in c->h_set_status, pairs in c->set_out)
    auto deliver = [&](uint64_t per_cap, int strategy) -> int {
      uint64_t total_m = 0, total_c = 0;
      for (int j = 0; j < K; ++j) {
        const PipelineStatus& h = c->h_set_status[j];
        counts[j] = (int64_t)h.n_matches;
        if (carry_out) { carry_out[j].cur = h.carry_cur; carry_out[j].tail = h.carry_tail; }
        total_m += h.n_matches;
        total_c += h.n_candidates;
      }
      if (stats) {
        cudaEventRecord(c->ev[2], s);
        cudaEventSynchronize(c->ev[2]);
        cudaEventElapsedTime(&stats->scan_ms, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&stats->total_ms, c->ev[0], c->ev[2]);
        stats->candidates = total_c;
        stats->matches = total_m;
        stats->strategy = strategy;
      }
      if (pairs) {
        for (int j = 0; j < K; ++j) {
          pairs[j] = static_cast<uint64_t*>(malloc(std::max<size_t>((size_t)counts[j] * 16, 8)));
          if (counts[j] && !Check(cudaMemcpyAsync(pairs[j], c->set_out.as<uint64_t>() + (uint64_t)j * 2 * per_cap,
                                                  (size_t)counts[j] * 16, cudaMemcpyDeviceToHost, s), "D2H", error)) return -1;
        }
        if (!Check(cudaStreamSynchronize(s), "sync", error)) return -1;
      }
      return 0;
    };
    // ---- k-mer index: one kernel, scan + finish (k_set_kmer) ---------------------------------------
    if (ds->kmer && !ds->kmer_off && c->coop) {
      ScanRange own{0, n + 1};
      uint64_t base_offset = 0;
      if (own_view) {
        own.own_begin = std::min<uint64_t>(own_view->own_begin, n + 1);
        own.own_end = std::min<uint64_t>(own_view->own_end, n + 1);
        base_offset = own_view->base_offset;
      }
      const uint64_t n16 = (n + 15) & ~15ull;
      const uint64_t row_lo = own.own_begin >> 9;
      const uint64_t last_end = std::min<uint64_t>(n, own.own_end + 8);          // an owned match ends at most here
      const uint64_t row_hi = std::max<uint64_t>(std::min<uint64_t>((last_end + 511) >> 9, (n16 + 511) >> 9), row_lo + 1);
      const uint64_t rows = row_hi - row_lo;
      // one CTA per SM (cooperative: every CTA waits for the records of the CTAs before it); a warp owns a
      // contiguous run of rows of any length
      const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((rows + 31) / 32, std::min<uint64_t>((uint64_t)c->sm_count, 160)));
      const size_t kmer_smem = kKmerSmemFixed;
      const uint64_t rows_per_warp = (rows + (uint64_t)blocks * kKmerWarps - 1) / ((uint64_t)blocks * kKmerWarps);
      const uint64_t rows_per_cta = rows_per_warp * kKmerWarps;
      if (rows_per_cta < kKmerMaxRowsPerCta && kmer_smem <= c->smem_optin) {
        for (int attempt = 0; attempt < 8; ++attempt) {
          const uint64_t per_cap = ds->per_cap;
          if (!c->set_out.Reserve(K * per_cap * 16, error)) return -1;
          if (!c->kmer_xchg.p) {
            // exchange records, then the two sync words and the per-member totals (all zero between calls)
            const size_t bytes = (size_t)c->sm_count * 32 * sizeof(KmerXchg) + 1024;
            if (!c->kmer_xchg.Reserve(bytes, error) || !Check(cudaMemsetAsync(c->kmer_xchg.p, 0, bytes, s), "memset", error)) return -1;
          }
          // staging area of the warps whose shared lists overflow: sized for the text's density (grows on demand)
          if (!c->kmer_stage.Reserve((size_t)blocks * kKmerWarps * ds->stage_cap * sizeof(uint2), error)) return -1;
          KmerRun run{};
          run.row_lo = row_lo;
          run.rows_per_cta = (uint32_t)rows_per_cta;
          run.rows_per_warp = (uint32_t)rows_per_warp;
          run.xchg = c->kmer_xchg.as<KmerXchg>();
          uint8_t* tail = c->kmer_xchg.as<uint8_t>() + (size_t)c->sm_count * 32 * sizeof(KmerXchg);
          run.gsync = reinterpret_cast<unsigned int*>(tail);
          run.gfinal = reinterpret_cast<unsigned long long*>(tail + 64);
          run.stage = c->kmer_stage.as<uint2>();
          run.stage_cap = ds->stage_cap;
          run.out_pairs = c->set_out.as<uint64_t>();
          run.out_stride = per_cap;
          run.out_cap = per_cap;
          run.base_offset = base_offset;
          run.host_records = c->h_fin_dev;
          run.seq = ++c->call_seq ? c->call_seq : ++c->call_seq;
          if (stitch && c->stitch_inbox) {
            run.stitch.enabled = 1;
            run.stitch.rank = c->stitch_rank;
            run.stitch.slab_begin = stitch->slab_begin;
            run.stitch.inbox = static_cast<uint4*>(c->stitch_inbox);
            run.stitch.right_inbox = static_cast<uint4*>(c->stitch_right);
            run.stitch.left_ack = c->stitch_left ? reinterpret_cast<unsigned int*>(c->stitch_left) + kStitchAckWord : nullptr;
            run.stitch.step = stitch->step;
            run.stitch.report = c->h_stitch_dev;
          }
#ifdef RJ_KMER_PROBE
          if (!c->kmer_probe.Reserve((size_t)blocks * 65 * 8, error)) return -1;
          cudaMemsetAsync(c->kmer_probe.p, 0, (size_t)blocks * 65 * 8, s);
          run.probe = c->kmer_probe.as<unsigned long long>();
#endif
          CarrySet carries;
          for (int j = 0; j < 32; ++j) carries.c[j] = (carry_in && j < K) ? carry_in[j] : Carry();
          uint64_t n_arg = n;
          if (stats) cudaEventRecord(c->ev[0], s);
          run.has_carry = 0;
          for (int j = 0; j < K; ++j) if (carries.c[j].cur != 0) run.has_carry = 1;
          int k_arg = K;
          void* kargs[] = {(void*)&d_text, (void*)&n_arg, (void*)&k_arg, (void*)&ds->km, (void*)&own, (void*)&run, (void*)&carries};
          // (a plain launch was measured too: no difference for this kernel, unlike k_scan_emit)
          if (!Check(cudaLaunchCooperativeKernel((const void*)k_set_kmer, dim3(blocks), dim3(kKmerThreads), kargs, kmer_smem, s),
                     "cooperative launch", error)) return -1;
          if (stats) { cudaEventRecord(c->ev[1], s); stats->launches += 1; }
          if (!Check(cudaGetLastError(), "launch", error) || !WaitFinRecords(c, K, run.seq, error)) return -1;
#ifdef RJ_KMER_PROBE
          {
            // tuning builds only: when did the warps start / stop streaming, when did the CTAs leave (ns after the first start)
            std::vector<unsigned long long> pr((size_t)blocks * 65);
            cudaStreamSynchronize(s);
            cudaMemcpy(pr.data(), run.probe, pr.size() * 8, cudaMemcpyDeviceToHost);
            unsigned long long t0 = ~0ull;
            for (int b2 = 0; b2 < blocks * 32; ++b2) if (pr[(size_t)b2 * 2]) t0 = std::min(t0, pr[(size_t)b2 * 2]);
            std::vector<unsigned long long> ends, cta_scan, cta_exit, starts;
            std::vector<std::vector<unsigned long long>> by_wid(32);
            for (int b2 = 0; b2 < blocks; ++b2) {
              unsigned long long mx = 0;
              for (int w = 0; w < 32; ++w) {
                const unsigned long long a = pr[((size_t)b2 * 32 + w) * 2], e = pr[((size_t)b2 * 32 + w) * 2 + 1];
                if (!a || !e) continue;
                starts.push_back(a - t0); ends.push_back(e - t0); by_wid[w].push_back(e - t0);
                mx = std::max(mx, e - t0);
              }
              cta_scan.push_back(mx);
              cta_exit.push_back(pr[(size_t)blocks * 64 + b2] - t0);
            }
            auto q = [](std::vector<unsigned long long> v, double f) { if (v.empty()) return 0ull; std::sort(v.begin(), v.end()); return v[(size_t)(f * (v.size() - 1))]; };
            fprintf(stderr, "[kmer probe] n=%llu rows/warp=%llu | warp scan start med %llu | warp scan end min %llu med %llu p90 %llu max %llu | CTA scan end (slowest warp) min %llu med %llu max %llu | CTA exit med %llu max %llu (ns)\n",
                    (unsigned long long)n, (unsigned long long)rows_per_warp, q(starts, 0.5), q(ends, 0.0), q(ends, 0.5), q(ends, 0.9), q(ends, 1.0),
                    q(cta_scan, 0.0), q(cta_scan, 0.5), q(cta_scan, 1.0), q(cta_exit, 0.5), q(cta_exit, 1.0));
            std::string line = "[kmer probe] median scan end by warp id:";
            for (int w = 0; w < 32; ++w) line += " " + std::to_string(q(by_wid[w], 0.5));
            fprintf(stderr, "%s\n", line.c_str());
          }
#endif
          bool redo = false, give_up = false, dense = false, grow_stage = false;
          uint64_t need = 0;
          for (int j = 0; j < K; ++j) {
            c->h_set_status[j] = StatusFromRecord(c->h_fin[j], carries.c[j]);
            const PipelineStatus& h = c->h_set_status[j];
            if (h.dense || h.need_large) give_up = true;
            if (h.dense) dense = true;
            if (h.overflow) grow_stage = true;
            if (h.n_matches > per_cap) { need = std::max<uint64_t>(need, h.n_matches); redo = true; }
          }
          if (!dense) ds->kmer_dense = 0;
          if (give_up) {                                         // overlapping or too dense: the general path decides
            if (dense && ++ds->kmer_dense >= 2) ds->kmer_off = true;
            break;
          }
          if (grow_stage) { ds->stage_cap *= 4; redo = true; }
          if (need) ds->per_cap = need + need / 4 + 1024;
          if (redo) { if (stats) stats->reruns += 1; continue; }
          if (run.stitch.enabled) stitch->sent = true;          // the kernel has sent this step's states (valid)
          return deliver(per_cap, 4);
        }
      }
    }
    int warps = (int)std::min<size_t>((c->smem_optin - ds->fixed_smem) / kDfaTileBytes, 18);
    if (const char* env = getenv("RJ_DFA_WARPS")) warps = std::max(4, std::min(warps, atoi(env)));
    const uint64_t nsub = std::max<uint64_t>(1, (n + kDfaSubBytes - 1) / kDfaSubBytes);
    bool fallback = false;
    for (int attempt = 0; attempt < 40 && !fallback; ++attempt) {
      const uint64_t per_cap = ds->per_cap;
      if (!c->set_sub_b.Reserve((uint64_t)K * nsub * ds->sub_cap * 8, error) ||
          !c->set_sub_e.Reserve((uint64_t)K * nsub * ds->sub_cap * 8, error) ||
          !c->set_sub_count.Reserve((uint64_t)K * nsub * 4 + 16, error) ||
          !c->set_dense_b.Reserve(K * per_cap * 8, error) || !c->set_dense_e.Reserve(K * per_cap * 8, error) ||
          !c->set_reach.Reserve(K * per_cap * 8, error) || !c->set_take.Reserve(K * per_cap * 4, error) ||
          !c->set_fin.Reserve(K * per_cap * 8, error) || !c->set_slot.Reserve(K * per_cap * 8, error) ||
          !c->set_out.Reserve(K * per_cap * 16, error) || !c->set_status.Reserve(32 * sizeof(PipelineStatus), error)) return -1;
      PipelineStatus* d_status = c->set_status.as<PipelineStatus>();
      int blocks = (int)std::min<uint64_t>((nsub + warps - 1) / warps, (uint64_t)c->sm_count);
      // set_counts: [0..K) dense counts, [40] work counter, then the finish area
      // (sync 8 x u32 at +512, last_end[32] at +544, totals[32] at +800, last_ne[32] at +1056,
      //  segcount[K][nseg] at +1312)
      const bool fuse_finish = c->coop;
      const uint32_t seg_subs = (uint32_t)((nsub + blocks - 1) / blocks);
      const uint32_t nseg = (uint32_t)((nsub + seg_subs - 1) / seg_subs);
      const size_t counts_bytes = 1312 + (fuse_finish ? (size_t)K * nseg * 4 : 0);
      if (!c->set_counts.Reserve(counts_bytes, error)) return -1;
      unsigned long long* d_counts = c->set_counts.as<unsigned long long>();
      if (!Check(cudaMemsetAsync(d_counts, 0, counts_bytes, s), "memset", error)) return -1;
      if (!fuse_finish && !Check(cudaMemsetAsync(d_status, 0, 32 * sizeof(PipelineStatus), s), "memset", error)) return -1;
      SubStore st{};
      st.begin = c->set_sub_b.as<uint64_t>(); st.end = c->set_sub_e.as<uint64_t>(); st.count = c->set_sub_count.as<uint32_t>();
      st.cap = ds->sub_cap; st.nsub = (uint64_t)K * nsub;
      DenseList dense{c->set_dense_b.as<uint64_t>(), c->set_dense_e.as<uint64_t>(), d_counts, per_cap};
      ResolveScratch rs{c->set_reach.as<uint64_t>(), c->set_take.as<uint32_t>(), c->set_fin.as<uint64_t>(), c->set_slot.as<uint64_t>()};
      ScanRange own{0, n + 1};
      uint64_t base_offset = 0;
      if (own_view) {
        own.own_begin = std::min<uint64_t>(own_view->own_begin, n + 1);
        own.own_end = std::min<uint64_t>(own_view->own_end, n + 1);
        base_offset = own_view->base_offset;
      }
      CarrySet carries;
      for (int j = 0; j < 32; ++j) carries.c[j] = (carry_in && j < K) ? carry_in[j] : Carry();
      if (stats) cudaEventRecord(c->ev[0], s);
      size_t smem = ds->fixed_smem + (size_t)warps * kDfaTileBytes;
      auto wait_all = [&](unsigned int seq) -> bool {
        uint64_t spins = 0;
        for (int j = 0; j < K; ++j) {
          volatile unsigned int* vseq = &c->h_set_status[j].seq;
          while (*vseq != seq) {
            if ((++spins & 0x3FFF) == 0) {
              cudaError_t q = cudaStreamQuery(s);
              if (q == cudaSuccess) { if (*vseq == seq) break; if (error) *error = "rejit_b200: set resolve did not report"; return false; }
              if (q != cudaErrorNotReady) { Check(q, "cudaStreamQuery", error); return false; }
            }
          }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        return true;
      };
      auto resolve_set = [&]() -> bool {
        const unsigned int seq = ++c->call_seq ? c->call_seq : ++c->call_seq;
        k_resolve_set<<<K, 512, 0, s>>>(st, nsub, dense, rs, per_cap, c->set_out.as<uint64_t>(), d_status,
                                        c->h_set_status_dev, d_counts, seq, carries, base_offset);
        if (stats) stats->launches += 1;
        if (!Check(cudaGetLastError(), "launch", error)) return false;
        return wait_all(seq);
      };
      FinishArgs fin{};
      unsigned int* dense_flag = &d_status->dense;
      unsigned long long* work = d_counts + 40;
      if (fuse_finish) {
        uint8_t* base = c->set_counts.as<uint8_t>();
        fin.enabled = 1;
        fin.seg_subs = seg_subs;
        fin.nseg = nseg;
        fin.sync = reinterpret_cast<unsigned int*>(base + 512);
        fin.last_end = reinterpret_cast<unsigned long long*>(base + 544);
        fin.totals = reinterpret_cast<unsigned long long*>(base + 800);
        fin.last_ne = reinterpret_cast<unsigned long long*>(base + 1056);
        fin.segcount = reinterpret_cast<uint32_t*>(base + 1312);
        fin.local_pred = 1;                      // members are at most 17 bytes long
        fin.out_pairs = c->set_out.as<uint64_t>();
        fin.out_stride = per_cap;
        fin.out_cap = per_cap;
        fin.base_offset = base_offset;
        fin.host_records = c->h_fin_dev;
        fin.seq = ++c->call_seq ? c->call_seq : ++c->call_seq;
        dense_flag = fin.sync + 4;
        uint64_t n_arg = n, nsub_arg = nsub;
        void* args[] = {(void*)&d_text, (void*)&n_arg, (void*)&ds->tb, (void*)&own, (void*)&st, (void*)&nsub_arg,
                        (void*)&dense_flag, (void*)&work, (void*)&fin, (void*)&carries};
        if (!Check(cudaLaunchCooperativeKernel((const void*)k_set_tma, dim3(blocks), dim3(warps * 32), args, smem, s),
                   "cooperative launch", error)) return -1;
        if (stats) { cudaEventRecord(c->ev[1], s); stats->launches += 1; }
        if (!Check(cudaGetLastError(), "launch", error) || !WaitFinRecords(c, K, fin.seq, error)) return -1;
        for (int j = 0; j < K; ++j) c->h_set_status[j] = StatusFromRecord(c->h_fin[j], carries.c[j]);
        bool overlap = false, clean = true;
        for (int j = 0; j < K; ++j) {
          const PipelineStatus& h = c->h_set_status[j];
          if (h.overflow || h.dense) clean = false;
          if (h.need_large) overlap = true;
        }
        if (overlap && clean) {
          // some pattern's neighbouring candidates overlap: the general resolve decides
          if (!Check(cudaMemsetAsync(d_status, 0, 32 * sizeof(PipelineStatus), s), "memset", error)) return -1;
          if (!resolve_set()) return -1;
        }
      } else {
        k_set_tma<<<blocks, warps * 32, smem, s>>>(d_text, n, ds->tb, own, st, nsub, dense_flag, work, fin, carries);
        if (stats) { cudaEventRecord(c->ev[1], s); stats->launches += 1; }
        if (!resolve_set()) return -1;
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      bool rerun = false;
      uint32_t need_cap = 0;
      uint64_t need_dense = 0;
      for (int j = 0; j < K; ++j) {
        const PipelineStatus& h = c->h_set_status[j];
        if (h.dense || h.need_large) fallback = true;
        if (h.overflow && h.need_cap) { need_cap = std::max(need_cap, h.need_cap); rerun = true; }
        if (h.n_candidates > per_cap) { need_dense = std::max<uint64_t>(need_dense, h.n_candidates); rerun = true; }
        else if (h.overflow && !h.need_cap) rerun = true;
      }
      if (fallback) break;
      if (rerun) {
        if (need_cap) ds->sub_cap = std::max<uint32_t>(need_cap, 2 * ds->sub_cap);
        if (need_dense) ds->per_cap = need_dense + need_dense / 4 + 1024;
        if (stats) stats->reruns += 1;
        continue;
      }
      return deliver(per_cap, 4);
    }
    ds->dense_mode = true;
  }
  return one_by_one();
}

int64_t MatchAllHostMultiGpu(Program* prog, const uint8_t* text, uint64_t n, int n_gpus, uint64_t** pairs,
                             RunStats* stats, std::string* error) {
  if (!CudaOk(error)) return -1;
  int g = std::max(1, std::min(n_gpus, DeviceCount()));
  if (n < (uint64_t)g * 4096) g = 1;
  // the label replay of re-entrant patterns cannot be cut at a slab edge: one device, and the caller is told
  // (stats->large_path bit 1; rejit_b200_program_is_shardable answers the same question beforehand)
  if (prog->automaton().reentrant && g > 1) { g = 1; if (stats) stats->large_path |= 2; }
  std::vector<std::vector<uint64_t>> part(g);
  std::vector<Carry> carry_out(g);
  std::vector<std::string> errs(g);
  std::vector<char> ok(g, 1);
  std::vector<RunStats> st(g);
  auto bounds = [&](int i) { return (n / g) * (uint64_t)i; };
  // round 1: every slab resolved as if no match from the left neighbour reached into it
  WorkerPool::Devices().Run(g, g, [&](int i) {               // persistent threads, one slab per device
    DeviceContext* c = ContextFor(i, &errs[i]);
    DeviceProgram* dp = c ? prog->OnDevice(i, &errs[i]) : nullptr;
    if (!c || !dp) { ok[i] = 0; return; }
    uint64_t lo = bounds(i), hi = (i + 1 == g) ? n : bounds(i + 1);
    Carry in{lo, kNoMatch};
    ok[i] = MatchSlabFromHost(c, prog, dp, text, n, lo, hi, i + 1 == g, in, &carry_out[i], &part[i], &st[i], &errs[i]);
  });
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

// ===========================================================================
// ReplaceAll
// ===========================================================================
int64_t ReplaceAllDevice(int device, Program* prog, const uint8_t* d_text, uint64_t n, const uint8_t* with,
                         uint64_t with_len, void** d_out, uint64_t* out_len, uint64_t* out_capacity,
                         RunStats* stats, std::string* error) {
  if (d_out) *d_out = nullptr;
  if (out_len) *out_len = 0;
  DeviceContext* c = ContextFor(device, error);
  if (!c) return -1;
  DeviceProgram* dp = prog->OnDevice(device, error);
  if (!dp) return -1;
  if (with_len > 0xFFFFFFFFull) { if (error) *error = "rejit_b200: replacement too long"; return -1; }
  std::lock_guard<std::mutex> lk(c->mu);
  Slab slab{{0, n + 1}, 0};
  PipelineStatus st;
  Carry in;
  cudaStream_t s = c->stream;
  // Generic scans whose every match is at least as long as the replacement (the strip of a FASTA file: `>.*\n|\n` -> "")
  // write the rebuilt text from the scan kernel itself: one launch, no match list, no lengths / prefix sum / index /
  // staging passes.  The output is never longer than the text, so it can be allocated before the matches are known.
  // A call the single-pass scan cannot finish (chains across tile edges, tiles too dense for their lists) comes back
  // with the matches as usual and is rebuilt below.  RJ_NO_FUSED_REBUILD=1: always the separate passes.
  static const bool no_fused_rebuild = getenv("RJ_NO_FUSED_REBUILD") != nullptr;
  const CompiledAutomaton& ca = prog->automaton();
  if (!no_fused_rebuild && c->emit && !dp->emit_off && !ca.reentrant && ca.strategy == ScanStrategy::Generic &&
      with_len <= ca.nfa.min_len && n > 0) {
    FusedRebuild fr;
    RJ_TRY_COUNT(cudaSetDevice(c->device));            // (Buffer::Reserve allocates on the current device)
    if (!c->with_buf.Reserve(with_len + 16, error)) return -1;
    if (with_len) RJ_TRY_COUNT(cudaMemcpyAsync(c->with_buf.p, with, with_len, cudaMemcpyHostToDevice, s));
    void* out = DeviceAlloc(device, n + 64, error);
    if (!out) return -1;
    fr.d_out = static_cast<uint8_t*>(out);
    fr.d_with = c->with_buf.as<uint8_t>();
    fr.with_len = (uint32_t)with_len;
    if (!RunPipeline(c, prog, dp, d_text, n, slab, in, nullptr, 0, &st, stats, error, &fr)) { DeviceFree(device, out); return -1; }
    if (fr.done) {
      if (fr.removed > n + st.n_matches * with_len) {
        DeviceFree(device, out);
        if (error) *error = "rejit_b200: the fused rebuild reported an impossible length";
        return -1;
      }
      *d_out = out;
      *out_len = n - fr.removed + st.n_matches * with_len;
      if (out_capacity) *out_capacity = n + 64;
      return (int64_t)st.n_matches;
    }
    DeviceFree(device, out);          // (back to the pool; the separate rebuild below asks for the exact size)
  } else if (!RunPipeline(c, prog, dp, d_text, n, slab, in, nullptr, 0, &st, stats, error)) {
    return -1;
  }
  const uint64_t m = st.n_matches;
  const uint64_t* pairs = c->out_pairs.as<uint64_t>();
  if (stats) cudaEventRecord(c->ev[0], s);
  uint64_t total_removed = 0;
  if (m) {
    if (!c->ReserveScratch(m, error)) return -1;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->wide.as<uint64_t>(), c->slot.as<uint64_t>(), (int64_t)m, s);
    if (!c->cub_tmp.Reserve(tmp, error)) return -1;
    int blocks = (int)std::min<uint64_t>((m + 255) / 256, (uint64_t)c->sm_count * 8);
    k_match_lengths<<<blocks, 256, 0, s>>>(pairs, m, c->wide.as<uint64_t>());
    RJ_TRY_COUNT(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->wide.as<uint64_t>(), c->slot.as<uint64_t>(), (int64_t)m, s));
    uint64_t tail[2] = {0, 0};
    RJ_TRY_COUNT(cudaMemcpyAsync(&tail[0], c->slot.as<uint64_t>() + (m - 1), 8, cudaMemcpyDeviceToHost, s));
    RJ_TRY_COUNT(cudaMemcpyAsync(&tail[1], c->wide.as<uint64_t>() + (m - 1), 8, cudaMemcpyDeviceToHost, s));
    RJ_TRY_COUNT(cudaStreamSynchronize(s));
    total_removed = tail[0] + tail[1];
    if (stats) stats->launches += 3;
  }
  const uint64_t len = n - total_removed + m * with_len;
  const uint64_t cap = len + 64;
  void* out = DeviceAlloc(device, cap, error);
  if (!out) return -1;
  bool ok = true;
  if (m == 0) {
    if (n) ok = Check(cudaMemcpyAsync(out, d_text, n, cudaMemcpyDeviceToDevice, s), "D2D", error);
  } else {
    ok = c->with_buf.Reserve(with_len + 16, error);
    if (ok && with_len) ok = Check(cudaMemcpyAsync(c->with_buf.p, with, with_len, cudaMemcpyHostToDevice, s), "H2D", error);
    if (ok) {
      const uint64_t n_tiles = n / kReplaceTile + 1;
      ok = c->trans_off.Reserve((n_tiles + 1) * 8, error);      // (the translate path's tile array: free here)
      if (ok) {
        k_replace_index<<<(unsigned)((n_tiles + 256) / 256), 256, 0, s>>>(pairs, m, n_tiles, c->trans_off.as<uint64_t>());
        int blocks = (int)std::min<uint64_t>(n_tiles, (uint64_t)c->sm_count * 8);
        k_replace_stage<<<blocks, 256, 0, s>>>(d_text, n, pairs, c->slot.as<uint64_t>(), m, c->with_buf.as<uint8_t>(),
                                               (uint32_t)with_len, static_cast<uint8_t*>(out), n_tiles, c->trans_off.as<uint64_t>());
        ok = Check(cudaGetLastError(), "k_replace_stage", error);
        if (stats) stats->launches += 2;
      }
    }
  }
  if (ok && stats) {
    cudaEventRecord(c->ev[1], s);
    ok = Check(cudaEventSynchronize(c->ev[1]), "sync", error);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    stats->total_ms += ms;
  } else if (ok) {
    ok = Check(cudaStreamSynchronize(s), "sync", error);
  }
  if (!ok) { DeviceFree(device, out); return -1; }
  *d_out = out;
  *out_len = len;
  if (out_capacity) *out_capacity = cap;
  return (int64_t)m;
}

// ---------------------------------------------------------------------------
// A set of one-byte patterns as one byte -> string table (replace.cuh k_translate_*).
// ---------------------------------------------------------------------------
bool ReplaceSetFusable(const std::vector<Program*>& progs, const std::vector<std::string>& withs) {
  if (progs.empty() || progs.size() > 32 || progs.size() != withs.size()) return false;
  size_t bytes = 0;
  for (size_t i = 0; i < progs.size(); ++i) {
    const PositionNfa& a = progs[i]->automaton().nfa;
    if (a.n_pos != 1 || a.has_anchor || a.min_len != 1 || a.max_len != 1 || a.accept_empty[0]) return false;
    if (withs[i].size() > kTr2MaxLen) return false;
    bytes += withs[i].size();
  }
  if (bytes > kTransMaxBytes) return false;
  // the calls run one after the other in the reference: a replacement must not hold a byte a LATER pattern matches
  for (size_t i = 0; i < progs.size(); ++i)
    for (size_t j = i + 1; j < progs.size(); ++j) {
      const std::array<uint32_t, 8>& cls = progs[j]->automaton().nfa.cls[0];
      for (unsigned char ch : withs[i]) if ((cls[ch >> 5] >> (ch & 31)) & 1u) return false;
    }
  return true;
}

int64_t ReplaceAllSetDevice(int device, const std::vector<Program*>& progs, const uint8_t* d_text, uint64_t n,
                            const std::vector<std::string>& withs, void** d_out, uint64_t* out_len, uint64_t* out_capacity,
                            int64_t* counts, RunStats* stats, std::string* error) {
  if (d_out) *d_out = nullptr;
  if (out_len) *out_len = 0;
  DeviceContext* c = ContextFor(device, error);
  if (!c) return -1;
  if (!ReplaceSetFusable(progs, withs)) { if (error) *error = "rejit_b200: this set is not a byte -> string table"; return -1; }
  const int K = (int)progs.size();
  TranslateTable tab;
  memset(&tab, 0, sizeof tab);
  uint32_t at = 0;
  std::vector<uint32_t> off(K);
  for (int i = 0; i < K; ++i) {
    off[i] = at;
    memcpy(tab.bytes + at, withs[i].data(), withs[i].size());
    at += (uint32_t)withs[i].size();
  }
  for (int b = 0; b < 256; ++b) {
    tab.len[b] = 1; tab.off[b] = 0xFFFF; tab.pat[b] = kTransNone;
    for (int i = 0; i < K; ++i) {
      const std::array<uint32_t, 8>& cls = progs[i]->automaton().nfa.cls[0];
      if ((cls[b >> 5] >> (b & 31)) & 1u) {                 // the first pattern that matches the byte replaces it
        tab.len[b] = (uint16_t)withs[i].size(); tab.off[b] = (uint16_t)off[i]; tab.pat[b] = (uint8_t)i;
        break;
      }
    }
  }
  {
    // the bits all replaced bytes have in common (the kernels' SWAR pre-filter)
    uint32_t ones = 0xFF, zeros = 0xFF;
    for (int b = 0; b < 256; ++b) if (tab.pat[b] != kTransNone) { ones &= (uint32_t)b; zeros &= ~(uint32_t)b & 0xFFu; }
    tab.pre_mask = (ones | zeros) * 0x01010101u;
    tab.pre_value = ones * 0x01010101u;
  }
  std::lock_guard<std::mutex> lk(c->mu);
  RJ_TRY_COUNT(cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  const uint64_t n_tiles = (n + kTr2Tile - 1) / kTr2Tile;
  if (!c->trans_tab.Reserve(sizeof tab, error) || !c->trans_len.Reserve((n_tiles + 1) * 8, error) ||
      !c->trans_off.Reserve((n_tiles + 1) * 8, error) || !c->trans_counts.Reserve(32 * 8, error)) return -1;
  if (stats) cudaEventRecord(c->ev[0], s);
  RJ_TRY_COUNT(cudaMemcpyAsync(c->trans_tab.p, &tab, sizeof tab, cudaMemcpyHostToDevice, s));
  RJ_TRY_COUNT(cudaMemsetAsync(c->trans_counts.p, 0, 32 * 8, s));
  unsigned long long h_counts[32] = {0};
  uint64_t total = 0;
  // one warp per 16 KB tile, eight warps per CTA, at most four CTAs per SM (32 warps x four rows in flight)
  const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((n_tiles + kTr2Warps - 1) / kTr2Warps, (uint64_t)c->sm_count * 4));
  if (n_tiles) {
    const size_t count_smem = ((sizeof(Tr2Tables) + 15) & ~(size_t)15) + (size_t)kTr2Warps * K * 32 * 4;
    k_translate_count2<<<blocks, kTr2Warps * 32, count_smem, s>>>(d_text, n, c->trans_tab.as<TranslateTable>(), K,
                                                                  c->trans_len.as<uint64_t>(),
                                                                  c->trans_counts.as<unsigned long long>(), n_tiles);
    RJ_TRY_COUNT(cudaGetLastError());
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->trans_len.as<uint64_t>(), c->trans_off.as<uint64_t>(), (int64_t)n_tiles, s);
    if (!c->cub_tmp.Reserve(tmp, error)) return -1;
    RJ_TRY_COUNT(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->trans_len.as<uint64_t>(), c->trans_off.as<uint64_t>(), (int64_t)n_tiles, s));
    uint64_t tail[2] = {0, 0};
    RJ_TRY_COUNT(cudaMemcpyAsync(&tail[0], c->trans_off.as<uint64_t>() + (n_tiles - 1), 8, cudaMemcpyDeviceToHost, s));
    RJ_TRY_COUNT(cudaMemcpyAsync(&tail[1], c->trans_len.as<uint64_t>() + (n_tiles - 1), 8, cudaMemcpyDeviceToHost, s));
    RJ_TRY_COUNT(cudaMemcpyAsync(h_counts, c->trans_counts.p, 32 * 8, cudaMemcpyDeviceToHost, s));
    RJ_TRY_COUNT(cudaStreamSynchronize(s));
    total = tail[0] + tail[1];
  }
  const uint64_t cap = total + 64;
  void* out = DeviceAlloc(device, cap, error);
  if (!out) return -1;
  bool ok = true;
  if (n_tiles) {
    k_translate_write2<<<blocks, kTr2Warps * 32, 0, s>>>(d_text, n, c->trans_tab.as<TranslateTable>(), c->trans_off.as<uint64_t>(),
                                                         static_cast<uint8_t*>(out), n_tiles);
    ok = Check(cudaGetLastError(), "k_translate_write2", error);
  }
  if (ok && stats) {
    cudaEventRecord(c->ev[1], s);
    ok = Check(cudaEventSynchronize(c->ev[1]), "sync", error);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    stats->total_ms += ms;
    stats->scan_ms += ms;
    stats->launches += n_tiles ? 5 : 0;
  } else if (ok) {
    ok = Check(cudaStreamSynchronize(s), "sync", error);
  }
  if (!ok) { DeviceFree(device, out); return -1; }
  int64_t all = 0;
  for (int i = 0; i < K; ++i) { if (counts) counts[i] = (int64_t)h_counts[i]; all += (int64_t)h_counts[i]; }
  *d_out = out;
  *out_len = total;
  if (out_capacity) *out_capacity = cap;
  return all;
}

int64_t ReplaceAllHost(int device, Program* prog, const uint8_t* text, uint64_t n, const uint8_t* with,
                       uint64_t with_len, uint8_t** out, uint64_t* out_len, RunStats* stats, std::string* error) {
  void* d_text = DeviceAlloc(device, n + 64, error);
  if (!d_text) return -1;
  if (n && !CopyToDevice(device, d_text, text, n, error)) { DeviceFree(device, d_text); return -1; }
  void* d_out = nullptr;
  uint64_t len = 0;
  int64_t m = ReplaceAllDevice(device, prog, static_cast<const uint8_t*>(d_text), n, with, with_len, &d_out, &len, nullptr,
                               stats, error);
  DeviceFree(device, d_text);
  if (m < 0) return -1;
  *out = static_cast<uint8_t*>(malloc(std::max<uint64_t>(len, 1)));
  bool ok = *out != nullptr && (len == 0 || CopyFromDevice(device, *out, d_out, len, error));
  DeviceFree(device, d_out);
  if (!ok) { free(*out); *out = nullptr; if (error && error->empty()) *error = "rejit_b200: out of memory"; return -1; }
  *out_len = len;
  return m;
}

// ===========================================================================
// device-side stitch (scan_emit.cuh: k_stitch)
// ===========================================================================
bool StitchOpen(int device, int rank, int world, void* handle64, std::string* error) {
  DeviceContext* c = ContextFor(device, error);
  if (!c) return false;
  std::lock_guard<std::mutex> lk(c->mu);
  RJ_TRY(cudaSetDevice(c->device));
  if (!c->stitch_inbox) {
    RJ_TRY(cudaMalloc(&c->stitch_inbox, kStitchInboxBytes));
    RJ_TRY(cudaHostAlloc(&c->h_stitch, sizeof(StitchReport), cudaHostAllocMapped));
    RJ_TRY(cudaHostGetDevicePointer(&c->h_stitch_dev, c->h_stitch, 0));
  }
  RJ_TRY(cudaMemset(c->stitch_inbox, 0, kStitchInboxBytes));
  memset(c->h_stitch, 0, sizeof(StitchReport));
  c->stitch_rank = rank;
  c->stitch_world = world;
  c->stitch_step = 0;
  cudaIpcMemHandle_t h;
  RJ_TRY(cudaIpcGetMemHandle(&h, c->stitch_inbox));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  return true;
}

bool StitchConnect(int device, const void* left_handle64, const void* right_handle64, std::string* error) {
  DeviceContext* c = ContextFor(device, error);
  if (!c) return false;
  std::lock_guard<std::mutex> lk(c->mu);
  RJ_TRY(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  if (left_handle64 && !c->stitch_left) {
    memcpy(&h, left_handle64, 64);
    RJ_TRY(cudaIpcOpenMemHandle(&c->stitch_left, h, cudaIpcMemLazyEnablePeerAccess));
  }
  if (right_handle64 && !c->stitch_right) {
    memcpy(&h, right_handle64, 64);
    RJ_TRY(cudaIpcOpenMemHandle(&c->stitch_right, h, cudaIpcMemLazyEnablePeerAccess));
  }
  return true;
}

void StitchClose(int device) {
  std::string err;
  DeviceContext* c = ContextFor(device, &err);
  if (!c) return;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->stitch_left) cudaIpcCloseMemHandle(c->stitch_left);
  if (c->stitch_right) cudaIpcCloseMemHandle(c->stitch_right);
  c->stitch_left = c->stitch_right = nullptr;
}

namespace {
bool WaitStitchReport(DeviceContext* c, unsigned int step, int K, Carry* arrived, uint32_t* redo_mask, std::string* error) {
  volatile StitchReport* r = c->h_stitch;
  // every record carries the step: the head and the K records are complete once each shows it (one 16-byte store each)
  auto complete = [&]() {
    if (r->head.w != step) return false;
    for (int j = 0; j < K; ++j) if (r->rec[j].z != step) return false;
    return true;
  };
  uint64_t spins = 0;
  while (!complete()) {
    if ((++spins & 0x3FFF) == 0) {
      cudaError_t q = cudaStreamQuery(c->stream);
      if (q == cudaSuccess) { if (complete()) break; if (error) *error = "rejit_b200: the stitch did not report"; return false; }
      if (q != cudaErrorNotReady) { Check(q, "cudaStreamQuery", error); return false; }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  const unsigned int redo = r->head.x, status = r->head.y, ne = r->head.z;
  if (status & 2u) { if (error) *error = "rejit_b200: a neighbouring rank did not answer the stitch"; return false; }
  for (int j = 0; j < K; ++j) {
    const uint64_t cur = (uint64_t)r->rec[j].y << 32 | r->rec[j].x;
    arrived[j].cur = cur;
    arrived[j].tail = ((ne >> j) & 1u) ? cur : kNoMatch;
  }
  *redo_mask = redo;
  return true;
}
}  // namespace

unsigned int StitchNextStep(int device) {
  std::string err;
  DeviceContext* c = ContextFor(device, &err);
  if (!c || !c->stitch_inbox) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  return ++c->stitch_step ? c->stitch_step : ++c->stitch_step;
}

bool StitchCollect(int device, unsigned int step, int K, Carry* arrived, uint32_t* redo_mask, std::string* error) {
  DeviceContext* c = ContextFor(device, error);
  if (!c) return false;
  std::lock_guard<std::mutex> lk(c->mu);
  return WaitStitchReport(c, step, K, arrived, redo_mask, error);
}

bool StitchExchange(int device, int K, const Carry* leaving, uint64_t slab_begin, Carry* arrived, uint32_t* redo_mask,
                    std::string* error, unsigned int step) {
  DeviceContext* c = ContextFor(device, error);
  if (!c) return false;
  if (!c->stitch_inbox || K < 1 || K > 32) { if (error) *error = "rejit_b200: stitch not opened (or more than 32 patterns)"; return false; }
  std::lock_guard<std::mutex> lk(c->mu);
  RJ_TRY(cudaSetDevice(c->device));
  StitchArgs a{};
  a.K = K;
  for (int j = 0; j < K; ++j) {
    // a chain that has not moved past the slab's first byte cannot reach the neighbour: nothing to say
    const bool has = leaving[j].cur > slab_begin || (leaving[j].tail != kNoMatch);
    a.sent_cur[j] = has ? leaving[j].cur : 0;
    if (has) a.sent_has |= 1u << j;
    if (has && leaving[j].tail == leaving[j].cur) a.sent_ne |= 1u << j;
  }
  a.link.enabled = 1;
  a.link.rank = c->stitch_rank;
  a.link.slab_begin = slab_begin;
  a.link.inbox = static_cast<uint4*>(c->stitch_inbox);
  a.link.right_inbox = static_cast<uint4*>(c->stitch_right);
  a.link.left_ack = c->stitch_left ? reinterpret_cast<unsigned int*>(c->stitch_left) + kStitchAckWord : nullptr;
  if (!step) step = ++c->stitch_step ? c->stitch_step : ++c->stitch_step;
  a.link.step = step;
  a.link.report = c->h_stitch_dev;
  c->h_stitch->head.w = 0;                                   // (a fused attempt of the same step may have reported already)
  for (int j = 0; j < 32; ++j) c->h_stitch->rec[j].z = 0;
  std::atomic_thread_fence(std::memory_order_release);
  k_stitch<<<1, 32, 0, c->stream>>>(a);
  RJ_TRY(cudaGetLastError());
  return WaitStitchReport(c, step, K, arrived, redo_mask, error);
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
