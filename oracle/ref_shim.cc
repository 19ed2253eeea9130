// oracle/ref_shim.cc — TEST INFRASTRUCTURE, not product code.
//
// A thin extern "C" wrapper around the *unmodified* reference engine
// (coreperf/rejit, compiled from /root/reference by oracle/Makefile into
// oracle/_ref/librejit_ref.so).  It lets Python tests / golden-vector
// generators / bench.py's CPU baseline call the reference's public API
// (include/rejit.h:105-138) through ctypes, and toggle the three result-
// affecting run-time flags (src/flags.h:36-55; mutable because the library is
// built with -DMOD_FLAGS).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
#include <stdint.h>
#include <string.h>
#include <vector>
#include <string>
#include <thread>
#include <algorithm>
#include <omp.h>

#include <stdlib.h>

#include "rejit.h"
#include "flags.h"

extern "C" {

// flagset: 0 = reference defaults, 1 = "noreduce" (FF on, use_ff_reduce=0),
// 2 = "noff" (use_fast_forward=0; the parity configuration).
void ref_set_flagset(int flagset) {
  FLAG_use_fast_forward = (flagset != 2);
  FLAG_use_fast_forward_early = (flagset != 2);
  FLAG_use_ff_reduce = (flagset == 0);
}

void ref_set_parser_opt(int on) { FLAG_use_parser_opt = (on != 0); }

// Programs that link this library through rejit.h only (samples/jrep.cc built on the reference, the
// checker of its batching) choose the flag set with REJIT_REF_FLAGSET=0|1|2 in the environment.
__attribute__((constructor)) static void ref_flagset_from_environment() {
  if (const char* v = getenv("REJIT_REF_FLAGSET")) ref_set_flagset(atoi(v));
}

// Returns the parse status (0 ok, -1 ParserError); on error copies the
// reference's status string into msg.
int ref_parse_status(const char* pattern, char* msg, size_t msglen) {
  rejit::Regej re(pattern);
  if (re.status() != rejit::RejitSuccess && msg && msglen) {
    strncpy(msg, rejit::rejit_status_string, msglen - 1);
    msg[msglen - 1] = 0;
  }
  return (int)re.status();
}

// MatchAll: writes up to cap (begin,end) offset pairs, returns the number of
// matches found (may exceed cap), or -1 on parse error.
int64_t ref_match_all(const char* pattern, const char* text, size_t n,
                      uint64_t* out_pairs, size_t cap) {
  rejit::Regej re(pattern);
  if (re.status() != rejit::RejitSuccess) return -1;
  std::vector<rejit::Match> m;
  re.MatchAll(text, n, &m);
  size_t k = std::min(cap, m.size());
  for (size_t i = 0; i < k; i++) {
    out_pairs[2 * i] = (uint64_t)(m[i].begin - text);
    out_pairs[2 * i + 1] = (uint64_t)(m[i].end - text);
  }
  return (int64_t)m.size();
}

// MatchFirst: returns 1/0 (found / not), -1 on parse error.
int ref_match_first(const char* pattern, const char* text, size_t n,
                    uint64_t* out_pair) {
  rejit::Regej re(pattern);
  if (re.status() != rejit::RejitSuccess) return -1;
  rejit::Match m;
  m.begin = m.end = text;
  bool r = re.MatchFirst(text, n, &m);
  if (r && out_pair) {
    out_pair[0] = (uint64_t)(m.begin - text);
    out_pair[1] = (uint64_t)(m.end - text);
  }
  return r ? 1 : 0;
}

int ref_match_full(const char* pattern, const char* text, size_t n) {
  rejit::Regej re(pattern);
  if (re.status() != rejit::RejitSuccess) return -1;
  return re.MatchFull(text, n) ? 1 : 0;
}

int ref_match_anywhere(const char* pattern, const char* text, size_t n) {
  rejit::Regej re(pattern);
  if (re.status() != rejit::RejitSuccess) return -1;
  return re.MatchAnywhere(text, n) ? 1 : 0;
}

// ---- timing helpers for the CPU baseline -------------------------------
// A compiled-once handle so that the timed region contains only MatchAll
// ("best" speed in the reference's terms, tools/benchmarks/engines/rejit/
// engine.cc:69-106; the vector is cleared between iterations, which the
// reference harness forgets to do — SURVEY.md B14).
struct RefHandle {
  rejit::Regej* re;
  // One privately compiled matcher per worker thread: threads that share ONE
  // compiled Regej slow each other down severely (measured here: 8 threads on
  // one Regej run 6x SLOWER than one thread, 8 processes scale linearly), so
  // the all-cores baseline gives every thread its own copy of the JIT'd code.
  std::vector<rejit::Regej*> per_thread;
  std::string pattern;
};

void* ref_compile(const char* pattern) {
  rejit::Regej* re = new rejit::Regej(pattern);
  if (re->status() != rejit::RejitSuccess || !re->Compile(rejit::kMatchAll)) {
    delete re;
    return NULL;
  }
  RefHandle* h = new RefHandle;
  h->re = re;
  h->pattern = pattern;
  return h;
}

void ref_free(void* handle) {
  RefHandle* h = (RefHandle*)handle;
  if (!h) return;
  delete h->re;
  for (size_t i = 0; i < h->per_thread.size(); i++) delete h->per_thread[i];
  delete h;
}

// One single-threaded MatchAll call; returns the match count.
int64_t ref_run_match_all(void* handle, const char* text, size_t n) {
  RefHandle* h = (RefHandle*)handle;
  std::vector<rejit::Match> m;
  h->re->MatchAll(text, n, &m);
  return (int64_t)m.size();
}

// One call on a slab of `n` bytes; returns how many matches BEGIN in the first
// `own` bytes (the rest of the slab is overlap with the next worker's slab).
int64_t ref_match_all_handle(void* handle, const char* text, size_t n, size_t own) {
  RefHandle* h = (RefHandle*)handle;
  std::vector<rejit::Match> m;
  h->re->MatchAll(text, n, &m);
  int64_t c = 0;
  for (size_t i = 0; i < m.size(); i++)
    if ((size_t)(m[i].begin - text) < own) c++;
  return c;
}

// All-host-threads variant: the text is cut into `threads`
// contiguous slabs, each extended to the right by `overlap` bytes; every
// thread runs the (re-entrant, src/x64/codegen-x64.cc:99-207 keeps all state
// on its own stack) compiled MatchAll on its slab and counts the matches that
// BEGIN inside the slab proper.  Returns the summed count.  This is a timing
// harness: stitching of chains across slab edges is not attempted, so the
// count can differ from the single-thread count for overlapping matches.
int64_t ref_run_match_all_mt(void* handle, const char* text, size_t n,
                             int threads, size_t overlap) {
  RefHandle* h = (RefHandle*)handle;
  if (threads < 1) threads = 1;
  // more slabs than threads so that slabs with many matches do not straggle;
  // OpenMP keeps its worker threads alive between calls (no spawn cost in the
  // timed region after the first call)
  while ((int)h->per_thread.size() < threads) {     // untimed on the warm-up pass
    rejit::Regej* r = new rejit::Regej(h->pattern.c_str());
    r->Compile(rejit::kMatchAll);
    h->per_thread.push_back(r);
  }
  int slabs = threads * 4;
  if ((size_t)slabs > n / 4096 + 1) slabs = (int)(n / 4096 + 1);
  size_t slab = (n + slabs - 1) / slabs;
  int64_t total = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(+ : total)
  for (int t = 0; t < slabs; t++) {
    size_t b = std::min(n, (size_t)t * slab);
    size_t e = std::min(n, b + slab);
    size_t ee = std::min(n, e + overlap);
    if (b >= e) continue;
    std::vector<rejit::Match> m;
    h->per_thread[omp_get_thread_num()]->MatchAll(text + b, ee - b, &m);
    int64_t c = 0;
    for (size_t i = 0; i < m.size(); i++)
      if ((size_t)(m[i].begin - text) < e) c++;
    total += c;
  }
  return total;
}

}  // extern "C"
