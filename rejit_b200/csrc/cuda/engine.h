// Internal C++ interface of the sm_100a matching engine (below the C ABI of
// include/rejit_b200.h, above the kernels of engine.cu).
#ifndef REJIT_B200_CUDA_ENGINE_H_
#define REJIT_B200_CUDA_ENGINE_H_

#include <stdint.h>
#include <string>
#include <vector>

#include "../host/automaton.h"

namespace rejit_b200 {

struct Carry {                  // chain state crossing a slab boundary (§8e)
  uint64_t cur = 0;             // smallest offset where the next match may begin
  uint64_t tail = ~0ull;        // end of the last selected non-empty match
};

struct RunStats {
  float scan_ms = 0;            // the text-scanning kernel(s) only (CUDA events)
  float total_ms = 0;           // whole device pipeline (scan + verify + resolve)
  uint32_t launches = 0;        // kernels launched by this call
  uint32_t reruns = 0;          // pipeline restarts after a capacity overflow
  uint64_t candidates = 0;      // (begin, E(begin)) pairs before selection
  uint64_t matches = 0;
  int strategy = 0;
  int large_path = 0;
};

class DeviceProgram;            // per-device tables of one compiled pattern
class DeviceContext;            // per-device stream and scratch buffers

// A compiled pattern: host automaton + lazily created per-device tables.
class Program {
 public:
  static Program* Create(const LoweredRegexp& lr, std::string* error);
  ~Program();
  const CompiledAutomaton& automaton() const { return automaton_; }
  DeviceProgram* OnDevice(int device, std::string* error);

 private:
  Program() {}
  CompiledAutomaton automaton_;
  DeviceProgram* per_device_[16] = {nullptr};
};

// ---- device plumbing ------------------------------------------------------
int DeviceCount();
bool CudaOk(std::string* error);                       // is a usable GPU present
DeviceContext* ContextFor(int device, std::string* error);
void* DeviceAlloc(int device, size_t bytes, std::string* error);
void DeviceFree(int device, void* p);
void* PinnedAlloc(size_t bytes);
void PinnedFree(void* p);
bool CopyToDevice(int device, void* dst, const void* src, size_t bytes, std::string* error);
bool CopyFromDevice(int device, void* dst, const void* src, size_t bytes, std::string* error);
bool CopyOnDevice(int device, void* dst, const void* src, size_t bytes, std::string* error);
void FlushL2(int device);                              // writes a buffer larger than L2

// ---- the hot path -----------------------------------------------------------
// MatchAll over text resident in device memory (16-byte aligned).  Writes up to
// out_cap (begin,end) offset pairs to d_out (device memory, may be null when
// out_cap == 0) and returns the number of matches, or -1 with *error set.
// `own` (optional) restricts the reported matches to starts in
// [own_begin, own_end) of the buffer (a slab with halos); base_offset is added
// to every reported offset.  Carries are in buffer coordinates.
struct SlabView {
  uint64_t own_begin = 0;
  uint64_t own_end = ~0ull;      // clamped to n + 1
  uint64_t base_offset = 0;
};
int64_t MatchAllDevice(int device, Program* prog, const uint8_t* d_text, uint64_t n,
                       uint64_t* d_out, uint64_t out_cap, const Carry& in, Carry* out,
                       RunStats* stats, std::string* error, const SlabView* own = nullptr);

// MatchAll over host text: H2D copy, device pipeline, D2H of the match list.
// *pairs is malloc'ed by the callee (count*2 uint64), caller frees.
int64_t MatchAllHost(int device, Program* prog, const uint8_t* text, uint64_t n,
                     uint64_t** pairs, RunStats* stats, std::string* error);

// MatchAll over text already resident on `device` (uploaded once, searched by
// several patterns); only the match list travels back.
int64_t MatchAllResident(int device, Program* prog, const uint8_t* d_text, uint64_t n, uint64_t** pairs,
                         RunStats* stats, std::string* error);

// Same, with the text cut into `n_gpus` contiguous slabs, one per device, each
// scanned with a right halo; chains are stitched at the slab edges (§8e).
int64_t MatchAllHostMultiGpu(Program* prog, const uint8_t* text, uint64_t n, int n_gpus,
                             uint64_t** pairs, RunStats* stats, std::string* error);

// A set of compiled patterns matched against the same text in ONE fused pass
// when every member is a fixed-length anchor-free DFA pattern (otherwise the
// members are simply run one after the other).
class DeviceSet;
class SetProgram {
 public:
  static SetProgram* Create(const std::vector<Program*>& members);
  ~SetProgram();
  int size() const { return (int)members_.size(); }
  bool fused() const { return fused_; }
  const std::string& describe() const { return describe_; }
  const std::vector<Program*>& members() const { return members_; }
  const SetDfa& dfa() const { return dfa_; }
  DeviceSet* OnDevice(int device, std::string* error);

 private:
  std::vector<Program*> members_;
  bool fused_ = false;
  SetDfa dfa_;
  std::string describe_;
  DeviceSet* per_device_[16] = {nullptr};
};

// MatchAll of every member over device-resident text.  counts[j] = number of
// matches of member j; if pairs != nullptr, pairs[j] receives a malloc'ed array
// of counts[j] (begin,end) pairs.  Returns 0, or -1 with *error set.
// `own`, `carry_in` (one per member) and `carry_out` are optional: slab sharding.
// `stitch` (optional, one process per GPU): the call is step `step` of the device-side stitch; when the k-mer scan
// ran, its reporting CTA has exchanged the chain states with the neighbouring GPUs itself (sent = true) and the
// answer waits in the context's report (StitchCollect); otherwise the caller sends with StitchExchange(step).
struct StitchCall {
  unsigned int step = 0;
  uint64_t slab_begin = 0;             // first owned start, global
  bool sent = false;
};
int MatchAllSetResident(int device, SetProgram* set, const uint8_t* d_text, uint64_t n, int64_t* counts,
                        uint64_t** pairs, RunStats* stats, std::string* error, const SlabView* own = nullptr,
                        const Carry* carry_in = nullptr, Carry* carry_out = nullptr, StitchCall* stitch = nullptr);

// Regej::ReplaceAll on the device (reference
// src/rejit.cc:221-226, 97-112): every match in d_text[0..n) is replaced by
// with[0..with_len) (host bytes).  *d_out receives a DeviceAlloc'ed buffer holding
// the *out_len rebuilt bytes; returns the number of matches, or -1.
int64_t ReplaceAllDevice(int device, Program* prog, const uint8_t* d_text, uint64_t n, const uint8_t* with,
                         uint64_t with_len, void** d_out, uint64_t* out_len, uint64_t* out_capacity,
                         RunStats* stats, std::string* error);
// A SET of patterns that each match exactly one byte, applied "one after the other" (regex-dna's eleven IUB
// substitutions, /root/reference/sample/regexdna.cc:69-85) as ONE byte -> string table: a counting pass, a prefix
// sum and a writing pass instead of one scan + rebuild per pattern.  ReplaceSetFusable says whether the table gives
// exactly what the sequential calls give (one-byte patterns, no replacement holds a byte that a later pattern
// matches); counts[i] = matches of pattern i.  Returns the total number of matches, or -1.
bool ReplaceSetFusable(const std::vector<Program*>& progs, const std::vector<std::string>& withs);
int64_t ReplaceAllSetDevice(int device, const std::vector<Program*>& progs, const uint8_t* d_text, uint64_t n,
                            const std::vector<std::string>& withs, void** d_out, uint64_t* out_len, uint64_t* out_capacity,
                            int64_t* counts, RunStats* stats, std::string* error);
// Host-pointer convenience: uploads, replaces, downloads (*out is malloc'ed).
int64_t ReplaceAllHost(int device, Program* prog, const uint8_t* text, uint64_t n, const uint8_t* with,
                       uint64_t with_len, uint8_t** out, uint64_t* out_len, RunStats* stats, std::string* error);

// MatchFirst (pair != nullptr) / MatchAnywhere with early exit: 1 / 0, or -1 on error.
int MatchFirstHost(int device, Program* prog, const uint8_t* text, uint64_t n, uint64_t pair[2], std::string* error);

// Device-side stitch for one-process-per-GPU sharding: StitchOpen creates this rank's inbox in
// device memory and returns its CUDA IPC handle (64 bytes) for the neighbours; StitchConnect maps the neighbours'
// inboxes (either may be null at the ends of the chain); StitchExchange sends the chain states that leave this
// rank's slab (global offsets) into the right neighbour's device memory over NVLink, waits for the states arriving
// from the left and says for which patterns (bit mask) the arriving chain reaches into the slab.
bool StitchOpen(int device, int rank, int world, void* handle64, std::string* error);
bool StitchConnect(int device, const void* left_handle64, const void* right_handle64, std::string* error);
void StitchClose(int device);
bool StitchExchange(int device, int K, const Carry* leaving, uint64_t slab_begin, Carry* arrived, uint32_t* redo_mask,
                    std::string* error, unsigned int step = 0);
unsigned int StitchNextStep(int device);                 // the step number of the next exchange (0: stitch not opened)
bool StitchCollect(int device, unsigned int step, int K, Carry* arrived, uint32_t* redo_mask, std::string* error);

// MatchFull: 1 / 0, or -1 on error.
int MatchFullHost(int device, Program* prog, const uint8_t* text, uint64_t n, std::string* error);

}  // namespace rejit_b200

#endif  // REJIT_B200_CUDA_ENGINE_H_
