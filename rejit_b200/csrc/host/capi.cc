// Implementation of the C ABI declared in include/rejit_b200.h.
#include "../../../include/rejit_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <memory>
#include <string>
#include <vector>

#include "../cuda/engine.h"
#include "automaton.h"
#include "ir.h"

using namespace rejit_b200;

namespace {

struct IrHolder {
  rejit_b200_ir ir;                 // must stay the first member
  std::vector<rejit_b200_edge> edges;
  std::vector<uint8_t> payload;
};

void SetErr(char* err, size_t len, const std::string& msg) {
  if (!err || !len) return;
  size_t k = msg.size() < len - 1 ? msg.size() : len - 1;
  memcpy(err, msg.data(), k);
  err[k] = 0;
}

// Nothing may leave an extern "C" function by exception (std::terminate in a C, cgo or ctypes caller): every entry
// point that can allocate is a function-try-block closed by this handler.  (The message is built without allocating
// when memory is what ran out.)
void SetErrNoAlloc(char* err, size_t len, const char* where, const char* what) {
  if (!err || !len) return;
  snprintf(err, len, "%s: %s", where, what ? what : "exception");
}
#define RJ_CATCH(where, fail)                                                                 \
  catch (const std::exception& e) { SetErrNoAlloc(err, err_length, where, e.what()); return fail; } \
  catch (...) { SetErrNoAlloc(err, err_length, where, nullptr); return fail; }

void PutU16(std::vector<uint8_t>* v, size_t x) {
  v->push_back(static_cast<uint8_t>(x & 0xFF));
  v->push_back(static_cast<uint8_t>((x >> 8) & 0xFF));
}

void Flatten(const LoweredRegexp& lr, IrHolder* h) {
  auto add = [&](const Edge& e) {
    rejit_b200_edge f;
    f.kind = e.kind;
    f.entry_state = e.entry;
    f.exit_state = e.exit;
    f.payload_offset = static_cast<int32_t>(h->payload.size());
    f.flags = 0;
    if (e.kind == kEdgeLiteral) {
      h->payload.insert(h->payload.end(), e.bytes.begin(), e.bytes.end());
    } else if (e.kind == kEdgeCharSet) {
      PutU16(&h->payload, e.singles.size());
      h->payload.insert(h->payload.end(), e.singles.begin(), e.singles.end());
      PutU16(&h->payload, e.ranges.size());
      for (const ByteRange& r : e.ranges) { h->payload.push_back(r.lo); h->payload.push_back(r.hi); }
      f.flags = e.negated ? 1 : 0;
    }
    f.payload_length = static_cast<int32_t>(h->payload.size()) - f.payload_offset;
    h->edges.push_back(f);
  };
  for (const Edge& e : lr.matching) add(e);
  for (const Edge& e : lr.control) add(e);
  h->ir.n_states = lr.n_states;
  h->ir.entry_state = lr.entry_state;
  h->ir.exit_state = lr.exit_state;
  h->ir.n_matching = static_cast<int32_t>(lr.matching.size());
  h->ir.n_control = static_cast<int32_t>(lr.control.size());
  h->ir.edges = h->edges.data();
  h->ir.payload = h->payload.data();
  h->ir.payload_length = h->payload.size();
}

// The IR comes from a foreign binding: everything that is later used as an index is checked here.
constexpr int32_t kMaxIrStates = 1 << 20;
constexpr int32_t kMaxIrEdges = 1 << 22;

bool Unflatten(const rejit_b200_ir* ir, LoweredRegexp* lr, std::string* error) {
  if (!ir) { *error = "rejit_b200_compile: no IR"; return false; }
  if (ir->n_states > kMaxIrStates || ir->n_matching > kMaxIrEdges || ir->n_control > kMaxIrEdges) {
    *error = "regular expression too large for the sm_100a engine (more than 2^20 states or 2^22 edges)";
    return false;
  }
  if (ir->n_states < 1 || ir->n_states > kMaxIrStates || ir->entry_state < 0 || ir->entry_state >= ir->n_states ||
      ir->exit_state < 0 || ir->exit_state >= ir->n_states || ir->n_matching < 0 || ir->n_control < 0 ||
      ir->n_matching > kMaxIrEdges || ir->n_control > kMaxIrEdges ||
      (ir->n_matching + ir->n_control > 0 && !ir->edges) || (ir->payload_length > 0 && !ir->payload)) {
    *error = "rejit_b200_compile: malformed IR";
    return false;
  }
  lr->n_states = ir->n_states;
  lr->entry_state = ir->entry_state;
  lr->exit_state = ir->exit_state;
  int total = ir->n_matching + ir->n_control;
  // the engine's position budget (automaton.cc), checked on the declared sizes before anything is copied
  uint64_t positions = 0;
  for (int i = 0; i < ir->n_matching; ++i) {
    const rejit_b200_edge& f = ir->edges[i];
    positions += (f.kind == kEdgeLiteral && f.payload_length > 0) ? static_cast<uint64_t>(f.payload_length) : 1;
  }
  if (positions > kMaxPatternPositions || static_cast<uint64_t>(ir->n_control) > 16 * kMaxPatternPositions) {
    *error = "regular expression too large for the sm_100a engine (more than 4096 byte positions)";
    return false;
  }
  for (int i = 0; i < total; ++i) {
    const rejit_b200_edge& f = ir->edges[i];
    Edge e;
    e.kind = f.kind;
    e.entry = f.entry_state;
    e.exit = f.exit_state;
    if (e.entry < 0 || e.exit < 0 || e.entry >= ir->n_states || e.exit >= ir->n_states ||
        f.payload_offset < 0 || f.payload_length < 0 ||
        static_cast<size_t>(f.payload_offset) + f.payload_length > ir->payload_length) {
      *error = "rejit_b200_compile: malformed IR";
      return false;
    }
    const uint8_t* p = ir->payload + f.payload_offset;
    if (f.kind == kEdgeLiteral) {
      if (f.payload_length == 0) { *error = "rejit_b200_compile: empty MultipleChar"; return false; }
      e.bytes.assign(p, p + f.payload_length);
    } else if (f.kind == kEdgeCharSet) {
      size_t at = 0;
      auto need = [&](size_t k) { return at + k <= static_cast<size_t>(f.payload_length); };
      if (!need(2)) { *error = "rejit_b200_compile: malformed Bracket"; return false; }
      size_t ns = p[at] | (p[at + 1] << 8); at += 2;
      if (!need(ns + 2)) { *error = "rejit_b200_compile: malformed Bracket"; return false; }
      e.singles.assign(p + at, p + at + ns); at += ns;
      size_t nr = p[at] | (p[at + 1] << 8); at += 2;
      if (!need(2 * nr)) { *error = "rejit_b200_compile: malformed Bracket"; return false; }
      for (size_t r = 0; r < nr; ++r) e.ranges.push_back({p[at + 2 * r], p[at + 2 * r + 1]});
      e.negated = (f.flags & 1) != 0;
    } else if (f.kind < 0 || f.kind > kEdgeEpsilon) {
      *error = "rejit_b200_compile: unknown edge kind";
      return false;
    }
    bool control = f.kind >= kEdgeLineStart;
    if (control != (i >= ir->n_matching)) { *error = "rejit_b200_compile: edge in the wrong list"; return false; }
    (control ? lr->control : lr->matching).push_back(e);
  }
  return true;
}

}  // namespace

struct rejit_b200_program {
  Program* prog;
};

struct rejit_b200_set {
  SetProgram* set;
};

struct rejit_b200_text {
  int device;
  void* d_ptr;
  size_t length;
  size_t capacity;
};

namespace {
// Freed text buffers are kept (two per process) so that a loop of
// upload / match / free does not pay cudaMalloc + cudaFree every time.
struct CachedText { int device; void* p; size_t capacity; };
std::mutex g_text_cache_mu;
std::vector<CachedText> g_text_cache;
}  // namespace

extern "C" {

int rejit_b200_parse(const char* pattern, size_t pattern_length, int parser_opt,
                     rejit_b200_ir** out_ir, char* err, size_t err_length) {
  if (out_ir) *out_ir = nullptr;
  // nothing may leave an extern "C" function by exception (std::terminate in a C caller)
  try {
    ParseOptions opt;
    opt.parser_opt = parser_opt != 0;
    std::string msg;
    NodePtr root = ParseERE(pattern, pattern_length, opt, &msg);
    if (!root || !WithinBudget(root.get(), &msg)) {
      SetErr(err, err_length, msg);
      return -1;
    }
    LoweredRegexp lr = Lower(root.get());
    IrHolder* h = new IrHolder();
    Flatten(lr, h);
    if (out_ir) *out_ir = &h->ir; else delete h;
    return 0;
  } catch (const std::exception& e) {
    SetErr(err, err_length, std::string("rejit_b200_parse: ") + e.what());
    return -1;
  }
}

void rejit_b200_ir_free(rejit_b200_ir* ir) {
  delete reinterpret_cast<IrHolder*>(ir);
}

size_t rejit_b200_ir_dump(const rejit_b200_ir* ir, char* buffer, size_t buffer_length) {
  LoweredRegexp lr;
  std::string error;
  std::string text = Unflatten(ir, &lr, &error) ? DumpLowered(lr) : ("error: " + error);
  if (buffer && buffer_length) {
    size_t k = text.size() < buffer_length - 1 ? text.size() : buffer_length - 1;
    memcpy(buffer, text.data(), k);
    buffer[k] = 0;
  }
  return text.size();
}

rejit_b200_program* rejit_b200_compile(const rejit_b200_ir* ir, char* err, size_t err_length) {
  try {
    LoweredRegexp lr;
    std::string error;
    if (!ir || !Unflatten(ir, &lr, &error)) {
      SetErr(err, err_length, ir ? error : "rejit_b200_compile: null IR");
      return nullptr;
    }
    Program* p = Program::Create(lr, &error);
    if (!p) {
      SetErr(err, err_length, error);
      return nullptr;
    }
    rejit_b200_program* h = new rejit_b200_program;
    h->prog = p;
    return h;
  } catch (const std::exception& e) {
    SetErr(err, err_length, std::string("rejit_b200_compile: ") + e.what());
    return nullptr;
  }
}

void rejit_b200_program_free(rejit_b200_program* program) {
  if (!program) return;
  delete program->prog;
  delete program;
}

int rejit_b200_program_is_shardable(const rejit_b200_program* program) {
  return program && !program->prog->automaton().reentrant ? 1 : 0;
}

// A text may be cut into slabs (ownership range / carries) only for patterns whose chain state at a cut is the
// (cur, tail) pair; the label replay of re-entrant patterns cannot be resumed from it.
static bool SlabAllowed(const Program* p, bool sliced, char* err, size_t err_length) {
  if (!sliced || !p->automaton().reentrant) return true;
  SetErr(err, err_length, "rejit_b200: this pattern is re-entrant (a running thread can re-enter its start) and cannot be "
                          "matched slab by slab; see rejit_b200_program_is_shardable");
  return false;
}

const char* rejit_b200_program_describe(const rejit_b200_program* program) {
  return program ? program->prog->automaton().describe.c_str() : "";
}

static void FillStats(const RunStats& s, rejit_b200_stats* out) {
  if (!out) return;
  out->scan_ms = s.scan_ms;
  out->total_ms = s.total_ms;
  out->launches = s.launches;
  out->reruns = s.reruns;
  out->candidates = s.candidates;
  out->matches = s.matches;
  out->strategy = s.strategy;
  out->large_path = s.large_path;
}

int64_t rejit_b200_match_all_alloc(rejit_b200_program* program, const char* text, size_t text_length,
                                   uint64_t** out_pairs, rejit_b200_stats* stats, char* err, size_t err_length) try {
  if (!program) { SetErr(err, err_length, "rejit_b200_match_all_alloc: null program"); return -1; }
  std::string error;
  RunStats rs;
  uint64_t* pairs = nullptr;
  int64_t r = MatchAllHost(0, program->prog, reinterpret_cast<const uint8_t*>(text), text_length, &pairs,
                           stats ? &rs : nullptr, &error);
  if (r < 0) { SetErr(err, err_length, error); return -1; }
  FillStats(rs, stats);
  if (out_pairs) *out_pairs = pairs; else free(pairs);
  return r;
} RJ_CATCH("rejit_b200_match_all_alloc", -1)

int64_t rejit_b200_match_all(rejit_b200_program* program, const char* text, size_t text_length,
                             uint64_t* out_pairs, size_t capacity, char* err, size_t err_length) try {
  if (!program) { SetErr(err, err_length, "rejit_b200_match_all: null program"); return -1; }
  uint64_t* pairs = nullptr;
  int64_t r = rejit_b200_match_all_alloc(program, text, text_length, &pairs, nullptr, err, err_length);
  if (r < 0) return r;
  size_t k = static_cast<size_t>(r) < capacity ? static_cast<size_t>(r) : capacity;
  if (k && out_pairs) memcpy(out_pairs, pairs, k * 16);
  free(pairs);
  return r;
} RJ_CATCH("rejit_b200_match_all", -1)

int rejit_b200_match_first(rejit_b200_program* program, const char* text, size_t text_length,
                           uint64_t out_pair[2], char* err, size_t err_length) try {
  if (!program) { SetErr(err, err_length, "rejit_b200_match_first: null program"); return -1; }
  // MatchFirst := first element of MatchAll, found slab by slab with early exit
  std::string error;
  if (!CudaOk(&error)) { SetErr(err, err_length, error); return -1; }
  uint64_t pair[2] = {0, 0};
  int r = MatchFirstHost(0, program->prog, reinterpret_cast<const uint8_t*>(text), text_length, pair, &error);
  if (r < 0) { SetErr(err, err_length, error); return -1; }
  if (r > 0 && out_pair) { out_pair[0] = pair[0]; out_pair[1] = pair[1]; }
  return r;
} RJ_CATCH("rejit_b200_match_first", -1)

int rejit_b200_match_full(rejit_b200_program* program, const char* text, size_t text_length,
                          char* err, size_t err_length) try {
  if (!program) { SetErr(err, err_length, "rejit_b200_match_full: null program"); return -1; }
  std::string error;
  int r = MatchFullHost(0, program->prog, reinterpret_cast<const uint8_t*>(text), text_length, &error);
  if (r < 0) SetErr(err, err_length, error);
  return r;
} RJ_CATCH("rejit_b200_match_full", -1)

int rejit_b200_match_anywhere(rejit_b200_program* program, const char* text, size_t text_length,
                              char* err, size_t err_length) try {
  if (!program) { SetErr(err, err_length, "rejit_b200_match_anywhere: null program"); return -1; }
  std::string error;
  if (!CudaOk(&error)) { SetErr(err, err_length, error); return -1; }
  int r = MatchFirstHost(0, program->prog, reinterpret_cast<const uint8_t*>(text), text_length, nullptr, &error);
  if (r < 0) SetErr(err, err_length, error);
  return r;
} RJ_CATCH("rejit_b200_match_anywhere", -1)

int64_t rejit_b200_match_all_multi_gpu(rejit_b200_program* program, const char* text, size_t text_length,
                                       int n_gpus, uint64_t** out_pairs, rejit_b200_stats* stats,
                                       char* err, size_t err_length) try {
  if (!program) { SetErr(err, err_length, "rejit_b200_match_all_multi_gpu: null program"); return -1; }
  std::string error;
  RunStats rs;
  uint64_t* pairs = nullptr;
  int64_t r = MatchAllHostMultiGpu(program->prog, reinterpret_cast<const uint8_t*>(text), text_length, n_gpus,
                                   &pairs, &rs, &error);
  if (r < 0) { SetErr(err, err_length, error); return -1; }
  FillStats(rs, stats);
  if (out_pairs) *out_pairs = pairs; else free(pairs);
  return r;
} RJ_CATCH("rejit_b200_match_all_multi_gpu", -1)

int rejit_b200_device_count(void) { return DeviceCount(); }

void* rejit_b200_device_alloc(int device, size_t bytes) {
  std::string error;
  void* p = DeviceAlloc(device, bytes, &error);
  if (!p) fprintf(stderr, "rejit_b200_device_alloc: %s\n", error.c_str());
  return p;
}
void rejit_b200_device_free(int device, void* ptr) { DeviceFree(device, ptr); }
void* rejit_b200_pinned_alloc(size_t bytes) { return PinnedAlloc(bytes); }
void rejit_b200_pinned_free(void* ptr) { PinnedFree(ptr); }

int rejit_b200_copy_to_device(int device, void* dst, const void* src, size_t bytes) {
  std::string error;
  if (CopyToDevice(device, dst, src, bytes, &error)) return 0;
  fprintf(stderr, "rejit_b200_copy_to_device: %s\n", error.c_str());
  return -1;
}
int rejit_b200_copy_from_device(int device, void* dst, const void* src, size_t bytes) {
  std::string error;
  if (CopyFromDevice(device, dst, src, bytes, &error)) return 0;
  fprintf(stderr, "rejit_b200_copy_from_device: %s\n", error.c_str());
  return -1;
}
void rejit_b200_flush_l2(int device) { FlushL2(device); }

int64_t rejit_b200_match_all_device(rejit_b200_program* program, int device, const void* d_text,
                                    size_t text_length, uint64_t* d_out_pairs, size_t capacity,
                                    const rejit_b200_carry* carry_in, rejit_b200_carry* carry_out,
                                    rejit_b200_stats* stats, char* err, size_t err_length) try {
  if (!program) { SetErr(err, err_length, "rejit_b200_match_all_device: null program"); return -1; }
  std::string error;
  Carry in, out;
  if (carry_in) { in.cur = carry_in->cur; in.tail = carry_in->tail; }
  if (!SlabAllowed(program->prog, carry_in && (carry_in->cur != 0 || carry_in->tail != ~0ull), err, err_length)) return -1;
  RunStats rs;
  int64_t r = MatchAllDevice(device, program->prog, static_cast<const uint8_t*>(d_text), text_length,
                             d_out_pairs, capacity, in, &out, stats ? &rs : nullptr, &error);
  if (r < 0) { SetErr(err, err_length, error); return -1; }
  if (carry_out) { carry_out->cur = out.cur; carry_out->tail = out.tail; }
  FillStats(rs, stats);
  return r;
} RJ_CATCH("rejit_b200_match_all_device", -1)

rejit_b200_text* rejit_b200_text_upload(int device, const char* text, size_t text_length, char* err,
                                        size_t err_length) try {
  std::string error;
  void* d = nullptr;
  size_t capacity = 0;
  {
    std::lock_guard<std::mutex> lk(g_text_cache_mu);
    for (size_t i = 0; i < g_text_cache.size(); ++i)
      if (g_text_cache[i].device == device && g_text_cache[i].capacity >= text_length) {
        d = g_text_cache[i].p;
        capacity = g_text_cache[i].capacity;
        g_text_cache.erase(g_text_cache.begin() + i);
        break;
      }
  }
  if (!d) {
    capacity = text_length ? text_length : 1;
    d = DeviceAlloc(device, capacity, &error);
    if (!d) { SetErr(err, err_length, error); return nullptr; }
  }
  if (text_length && !CopyToDevice(device, d, text, text_length, &error)) {
    DeviceFree(device, d);
    SetErr(err, err_length, error);
    return nullptr;
  }
  rejit_b200_text* t = new rejit_b200_text;
  t->device = device;
  t->d_ptr = d;
  t->length = text_length;
  t->capacity = capacity;
  return t;
} RJ_CATCH("rejit_b200_text_upload", nullptr)

rejit_b200_text* rejit_b200_text_from_device(int device, const void* d_text, size_t text_length, char* err,
                                             size_t err_length) try {
  std::string error;
  void* d = DeviceAlloc(device, text_length ? text_length : 1, &error);
  if (!d) { SetErr(err, err_length, error); return nullptr; }
  if (text_length && !CopyOnDevice(device, d, d_text, text_length, &error)) {
    DeviceFree(device, d);
    SetErr(err, err_length, error);
    return nullptr;
  }
  rejit_b200_text* t = new rejit_b200_text;
  t->device = device;
  t->d_ptr = d;
  t->length = text_length;
  t->capacity = text_length ? text_length : 1;
  return t;
} RJ_CATCH("rejit_b200_text_from_device", nullptr)

void rejit_b200_text_free(rejit_b200_text* text) {
  if (!text) return;
  {
    std::lock_guard<std::mutex> lk(g_text_cache_mu);
    if (g_text_cache.size() < 2) {
      g_text_cache.push_back({text->device, text->d_ptr, text->capacity});
      text->d_ptr = nullptr;
    }
  }
  if (text->d_ptr) DeviceFree(text->device, text->d_ptr);
  delete text;
}

int64_t rejit_b200_match_all_text(rejit_b200_program* program, const rejit_b200_text* text, uint64_t** out_pairs,
                                  rejit_b200_stats* stats, char* err, size_t err_length) try {
  if (!program) { SetErr(err, err_length, "rejit_b200_match_all_text: null program"); return -1; }
  if (!text) { SetErr(err, err_length, "rejit_b200_match_all_text: null text"); return -1; }
  std::string error;
  RunStats rs;
  uint64_t* pairs = nullptr;
  int64_t r = MatchAllResident(text->device, program->prog, static_cast<const uint8_t*>(text->d_ptr), text->length,
                               &pairs, stats ? &rs : nullptr, &error);
  if (r < 0) { SetErr(err, err_length, error); return -1; }
  FillStats(rs, stats);
  if (out_pairs) *out_pairs = pairs; else free(pairs);
  return r;
} RJ_CATCH("rejit_b200_match_all_text", -1)

rejit_b200_set* rejit_b200_set_create(rejit_b200_program* const* programs, int count) {
  if (!programs || count < 1) return nullptr;
  try {
    std::vector<Program*> members;
    for (int i = 0; i < count; ++i) {
      if (!programs[i]) return nullptr;
      members.push_back(programs[i]->prog);
    }
    std::unique_ptr<rejit_b200_set> s(new rejit_b200_set);
    s->set = SetProgram::Create(members);
    return s->set ? s.release() : nullptr;
  } catch (...) {
    return nullptr;
  }
}

void rejit_b200_set_free(rejit_b200_set* set) {
  if (!set) return;
  delete set->set;
  delete set;
}

const char* rejit_b200_set_describe(const rejit_b200_set* set) { return set ? set->set->describe().c_str() : ""; }

int rejit_b200_set_kmer_tables(const rejit_b200_set* set, uint32_t* info, uint32_t* bitmap, uint32_t* mask16) {
  if (!set || !set->set->fused() || !set->set->dfa().kmer.ok) return 0;
  const SetDfa::Kmer& km = set->set->dfa().kmer;
  if (info) {
    info[0] = km.shift; info[1] = km.canon; info[2] = km.canon_ok; info[3] = (uint32_t)set->set->dfa().n_patterns;
    for (int v = 0; v <= 8; ++v) info[4 + v] = km.len_le[v];
    info[13] = (uint32_t)kKmerR;
  }
  if (bitmap) memcpy(bitmap, km.bitmap.data(), km.bitmap.size() * 4);
  if (mask16) memcpy(mask16, km.mask16.data(), km.mask16.size() * 4);
  return 1;
}

int rejit_b200_match_all_set_device(rejit_b200_set* set, int device, const void* d_text, size_t text_length,
                                    int64_t* out_counts, rejit_b200_stats* stats, char* err, size_t err_length) try {
  if (!set) { SetErr(err, err_length, "rejit_b200_match_all_set_device: null set"); return -1; }
  std::string error;
  RunStats rs;
  int r = MatchAllSetResident(device, set->set, static_cast<const uint8_t*>(d_text), text_length, out_counts, nullptr,
                              stats ? &rs : nullptr, &error);
  if (r < 0) { SetErr(err, err_length, error); return -1; }
  FillStats(rs, stats);
  return 0;
} RJ_CATCH("rejit_b200_match_all_set_device", -1)

int rejit_b200_match_all_set_device_slab(rejit_b200_set* set, int device, const void* d_text, size_t text_length,
                                         uint64_t own_begin, uint64_t own_end, uint64_t base_offset,
                                         const rejit_b200_carry* carry_in, rejit_b200_carry* carry_out,
                                         int64_t* out_counts, rejit_b200_stats* stats, char* err, size_t err_length) try {
  if (!set) { SetErr(err, err_length, "rejit_b200_match_all_set_device_slab: null set"); return -1; }
  std::string error;
  RunStats rs;
  const int k = set->set->size();
  for (Program* member : set->set->members())
    if (!SlabAllowed(member, true, err, err_length)) return -1;
  std::vector<Carry> in(k), out(k);
  if (carry_in) for (int j = 0; j < k; ++j) { in[j].cur = carry_in[j].cur; in[j].tail = carry_in[j].tail; }
  SlabView view;
  view.own_begin = own_begin;
  view.own_end = own_end;
  view.base_offset = base_offset;
  int r = MatchAllSetResident(device, set->set, static_cast<const uint8_t*>(d_text), text_length, out_counts, nullptr,
                              stats ? &rs : nullptr, &error, &view, in.data(), out.data());
  if (r < 0) { SetErr(err, err_length, error); return -1; }
  if (carry_out) for (int j = 0; j < k; ++j) { carry_out[j].cur = out[j].cur; carry_out[j].tail = out[j].tail; }
  FillStats(rs, stats);
  return 0;
} RJ_CATCH("rejit_b200_match_all_set_device_slab", -1)

int rejit_b200_match_all_set_text(rejit_b200_set* set, const rejit_b200_text* text, int64_t* out_counts,
                                  uint64_t** out_pairs, rejit_b200_stats* stats, char* err, size_t err_length) try {
  if (!text) { SetErr(err, err_length, "rejit_b200_match_all_set_text: null text"); return -1; }
  if (!set) { SetErr(err, err_length, "rejit_b200_match_all_set_text: null set"); return -1; }
  std::string error;
  RunStats rs;
  int r = MatchAllSetResident(text->device, set->set, static_cast<const uint8_t*>(text->d_ptr), text->length,
                              out_counts, out_pairs, stats ? &rs : nullptr, &error);
  if (r < 0) { SetErr(err, err_length, error); return -1; }
  FillStats(rs, stats);
  return 0;
} RJ_CATCH("rejit_b200_match_all_set_text", -1)

int64_t rejit_b200_match_all_device_slab(rejit_b200_program* program, int device, const void* d_text,
                                         size_t text_length, uint64_t own_begin, uint64_t own_end,
                                         uint64_t base_offset, uint64_t* d_out_pairs, size_t capacity,
                                         const rejit_b200_carry* carry_in, rejit_b200_carry* carry_out,
                                         rejit_b200_stats* stats, char* err, size_t err_length) try {
  if (!program) { SetErr(err, err_length, "rejit_b200_match_all_device_slab: null program"); return -1; }
  std::string error;
  Carry in, out;
  if (carry_in) { in.cur = carry_in->cur; in.tail = carry_in->tail; }
  if (!SlabAllowed(program->prog, true, err, err_length)) return -1;
  SlabView view;
  view.own_begin = own_begin;
  view.own_end = own_end;
  view.base_offset = base_offset;
  RunStats rs;
  int64_t r = MatchAllDevice(device, program->prog, static_cast<const uint8_t*>(d_text), text_length,
                             d_out_pairs, capacity, in, &out, stats ? &rs : nullptr, &error, &view);
  if (r < 0) { SetErr(err, err_length, error); return -1; }
  if (carry_out) { carry_out->cur = out.cur; carry_out->tail = out.tail; }
  FillStats(rs, stats);
  return r;
} RJ_CATCH("rejit_b200_match_all_device_slab", -1)

int64_t rejit_b200_replace_all(rejit_b200_program* program, const char* text, size_t text_length, const char* with,
                               size_t with_length, char** out, size_t* out_length, rejit_b200_stats* stats,
                               char* err, size_t err_length) try {
  std::string error;
  if (!program || !out || !out_length) { SetErr(err, err_length, "rejit_b200: null argument"); return -1; }
  if (!CudaOk(&error)) { SetErr(err, err_length, error); return -1; }
  RunStats rs;
  uint8_t* rebuilt = nullptr;
  uint64_t len = 0;
  int64_t r = ReplaceAllHost(0, program->prog, reinterpret_cast<const uint8_t*>(text), text_length,
                             reinterpret_cast<const uint8_t*>(with), with_length, &rebuilt, &len, stats ? &rs : nullptr,
                             &error);
  if (r < 0) { SetErr(err, err_length, error); return -1; }
  FillStats(rs, stats);
  *out = reinterpret_cast<char*>(rebuilt);
  *out_length = len;
  return r;
} RJ_CATCH("rejit_b200_replace_all", -1)

rejit_b200_text* rejit_b200_replace_all_text(rejit_b200_program* program, const rejit_b200_text* text, const char* with,
                                             size_t with_length, int64_t* n_matches, rejit_b200_stats* stats,
                                             char* err, size_t err_length) try {
  std::string error;
  if (!program || !text) { SetErr(err, err_length, "rejit_b200: null argument"); return nullptr; }
  RunStats rs;
  void* d_out = nullptr;
  uint64_t len = 0, cap = 0;
  int64_t r = ReplaceAllDevice(text->device, program->prog, static_cast<const uint8_t*>(text->d_ptr), text->length,
                               reinterpret_cast<const uint8_t*>(with), with_length, &d_out, &len, &cap,
                               stats ? &rs : nullptr, &error);
  if (r < 0 || !d_out) { SetErr(err, err_length, r < 0 ? error : "rejit_b200: ReplaceAll produced no text"); return nullptr; }
  FillStats(rs, stats);
  if (n_matches) *n_matches = r;
  rejit_b200_text* t = new rejit_b200_text;
  t->device = text->device;
  t->d_ptr = d_out;
  t->length = len;
  t->capacity = cap;
  return t;
} RJ_CATCH("rejit_b200_replace_all_text", nullptr)

rejit_b200_text* rejit_b200_replace_all_set_text(rejit_b200_program* const* programs, int count, const rejit_b200_text* text,
                                                 const char* const* withs, const size_t* with_lengths, int64_t* n_matches,
                                                 rejit_b200_stats* stats, char* err, size_t err_length) try {
  std::string error;
  if (!programs || count < 1 || !text || !withs || !with_lengths) { SetErr(err, err_length, "rejit_b200: null argument"); return nullptr; }
  std::vector<Program*> progs;
  std::vector<std::string> with;
  for (int i = 0; i < count; ++i) {
    if (!programs[i]) { SetErr(err, err_length, "rejit_b200: null program"); return nullptr; }
    progs.push_back(programs[i]->prog);
    with.emplace_back(withs[i] ? withs[i] : "", with_lengths[i]);
  }
  RunStats rs;
  if (ReplaceSetFusable(progs, with)) {
    // one-byte patterns whose replacements no later pattern touches: ONE byte -> string table
    void* d_out = nullptr;
    uint64_t len = 0, cap = 0;
    std::vector<int64_t> counts(count, 0);
    int64_t r = ReplaceAllSetDevice(text->device, progs, static_cast<const uint8_t*>(text->d_ptr), text->length, with, &d_out,
                                    &len, &cap, counts.data(), stats ? &rs : nullptr, &error);
    if (r < 0 || !d_out) { SetErr(err, err_length, r < 0 ? error : "rejit_b200: ReplaceAll produced no text"); return nullptr; }
    if (n_matches) for (int i = 0; i < count; ++i) n_matches[i] = counts[i];
    FillStats(rs, stats);
    rejit_b200_text* t = new rejit_b200_text;
    t->device = text->device; t->d_ptr = d_out; t->length = len; t->capacity = cap;
    return t;
  }
  // anything else: the calls one after the other, as the reference does
  const rejit_b200_text* cur = text;
  rejit_b200_text* owned = nullptr;
  rejit_b200_stats acc{};
  for (int i = 0; i < count; ++i) {
    int64_t k = 0;
    rejit_b200_stats one{};
    rejit_b200_text* next = rejit_b200_replace_all_text(programs[i], cur, with[i].data(), with[i].size(), &k, &one, err, err_length);
    if (owned) rejit_b200_text_free(owned);
    if (!next) return nullptr;
    if (n_matches) n_matches[i] = k;
    acc.scan_ms += one.scan_ms; acc.total_ms += one.total_ms; acc.launches += one.launches; acc.reruns += one.reruns;
    acc.matches += one.matches; acc.candidates += one.candidates;
    owned = next;
    cur = next;
  }
  if (stats) { *stats = acc; stats->strategy = -1; }
  return owned;
} RJ_CATCH("rejit_b200_replace_all_set_text", nullptr)

size_t rejit_b200_text_length(const rejit_b200_text* text) { return text ? text->length : 0; }
const void* rejit_b200_text_device_ptr(const rejit_b200_text* text) { return text ? text->d_ptr : nullptr; }

int rejit_b200_text_download(const rejit_b200_text* text, char* dst, size_t capacity, char* err, size_t err_length) try {
  std::string error;
  if (!text || (!dst && text->length)) { SetErr(err, err_length, "rejit_b200: null argument"); return -1; }
  if (capacity < text->length) { SetErr(err, err_length, "rejit_b200: destination too small"); return -1; }
  if (text->length && !CopyFromDevice(text->device, dst, text->d_ptr, text->length, &error)) {
    SetErr(err, err_length, error);
    return -1;
  }
  return 0;
} RJ_CATCH("rejit_b200_text_download", -1)

int rejit_b200_stitch_open(int device, int rank, int world, void* handle_out, char* err, size_t err_length) try {
  std::string error;
  if (!handle_out) { SetErr(err, err_length, "rejit_b200: null argument"); return -1; }
  if (!StitchOpen(device, rank, world, handle_out, &error)) { SetErr(err, err_length, error); return -1; }
  return 0;
} RJ_CATCH("rejit_b200_stitch_open", -1)

int rejit_b200_stitch_connect(int device, const void* left_handle, const void* right_handle, char* err, size_t err_length) try {
  std::string error;
  if (!StitchConnect(device, left_handle, right_handle, &error)) { SetErr(err, err_length, error); return -1; }
  return 0;
} RJ_CATCH("rejit_b200_stitch_connect", -1)

void rejit_b200_stitch_close(int device) { StitchClose(device); }

int rejit_b200_stitch_exchange(int device, int count, const rejit_b200_carry* leaving, uint64_t slab_begin,
                               rejit_b200_carry* arrived, uint32_t* redo_mask, char* err, size_t err_length) try {
  std::string error;
  if (count < 1 || count > 32 || !leaving || !arrived || !redo_mask) { SetErr(err, err_length, "rejit_b200: bad argument"); return -1; }
  Carry out[32], in[32];
  for (int j = 0; j < count; ++j) { out[j].cur = leaving[j].cur; out[j].tail = leaving[j].tail; }
  if (!StitchExchange(device, count, out, slab_begin, in, redo_mask, &error)) { SetErr(err, err_length, error); return -1; }
  for (int j = 0; j < count; ++j) { arrived[j].cur = in[j].cur; arrived[j].tail = in[j].tail; }
  return 0;
} RJ_CATCH("rejit_b200_stitch_exchange", -1)

int rejit_b200_match_all_set_device_stitched(rejit_b200_set* set, int device, const void* d_text, size_t text_length,
                                             uint64_t own_begin, uint64_t own_end, uint64_t base_offset,
                                             rejit_b200_carry* carry_out, rejit_b200_carry* arrived, uint32_t* redo_mask,
                                             int64_t* out_counts, rejit_b200_stats* stats, char* err, size_t err_length) try {
  if (!set) { SetErr(err, err_length, "rejit_b200_match_all_set_device_stitched: null set"); return -1; }
  std::string error;
  RunStats rs;
  const int k = set->set->size();
  if (k > 32 || !arrived || !redo_mask) { SetErr(err, err_length, "rejit_b200: bad argument"); return -1; }
  for (Program* member : set->set->members())
    if (!SlabAllowed(member, true, err, err_length)) return -1;
  StitchCall sc;
  sc.step = StitchNextStep(device);
  if (!sc.step) { SetErr(err, err_length, "rejit_b200: stitch not opened"); return -1; }
  sc.slab_begin = base_offset + own_begin;
  std::vector<Carry> in(k), out(k);
  for (int j = 0; j < k; ++j) { in[j].cur = own_begin; in[j].tail = ~0ull; }      // nothing arrives (to be checked by the stitch)
  SlabView view;
  view.own_begin = own_begin;
  view.own_end = own_end;
  view.base_offset = base_offset;
  int r = MatchAllSetResident(device, set->set, static_cast<const uint8_t*>(d_text), text_length, out_counts, nullptr,
                              stats ? &rs : nullptr, &error, &view, in.data(), out.data(), &sc);
  Carry got[32];
  bool ok = r >= 0;
  if (ok && sc.sent) {
    ok = StitchCollect(device, sc.step, k, got, redo_mask, &error);
  } else {
    // another scan path ran (or the call failed: the neighbours must still get this step's record)
    Carry leaving[32];
    for (int j = 0; j < k; ++j) {
      leaving[j].cur = (ok ? out[j].cur : own_begin) + base_offset;
      leaving[j].tail = (ok && out[j].tail != ~0ull) ? out[j].tail + base_offset : ~0ull;
    }
    const bool sent = StitchExchange(device, k, leaving, sc.slab_begin, got, redo_mask, ok ? &error : nullptr, sc.step);
    ok = ok && sent;
  }
  if (!ok) { SetErr(err, err_length, error); return -1; }
  for (int j = 0; j < k; ++j) {
    arrived[j].cur = got[j].cur; arrived[j].tail = got[j].tail;
    if (carry_out) { carry_out[j].cur = out[j].cur; carry_out[j].tail = out[j].tail; }
  }
  FillStats(rs, stats);
  return 0;
} RJ_CATCH("rejit_b200_match_all_set_device_stitched", -1)

void rejit_b200_free(void* ptr) { free(ptr); }

}  // extern "C"
