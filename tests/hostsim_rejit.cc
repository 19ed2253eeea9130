// tests/hostsim_rejit.cc — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// The public surface of include/rejit.h (and the handful of C-ABI entry points the
// samples touch) on top of tests/hostsim.cc: the product's own parser, lowering and
// automaton tables, with the kernels' work emulated on the CPU.  Linking a sample
// against this double instead of librejit_b200.so runs the sample's REJIT_B200 code
// paths (rejit::Text, pinned staging, MatchAllParallel, the device-side chains) where
// no GPU exists, so that the CPU tier can compare the samples' output with the golden
// fixtures end to end.  The product library itself has no such path: without a CUDA
// device it fails (tests/test_host_frontend.py::test_no_cpu_fallback_without_device).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/rejit.h"
#include "../include/rejit_b200.h"

extern "C" {
int64_t hostsim_match_all(const char* pattern, size_t plen, int parser_opt, const uint8_t* text, uint64_t n, int strategy,
                          uint64_t* out_pairs, uint64_t cap, char* describe, size_t dlen);
int64_t hostsim_match_all_slabs(const char* pattern, size_t plen, const uint8_t* text, uint64_t n, int slabs,
                                uint64_t* out_pairs, uint64_t cap);
int64_t hostsim_replace_all(const char* pattern, size_t plen, const uint8_t* text, uint64_t n, const uint8_t* with,
                            uint32_t with_len, uint8_t* out, uint64_t out_cap, uint64_t* out_len);
int hostsim_match_full(const char* pattern, size_t plen, const uint8_t* text, uint64_t n);

// the C-ABI entry points samples call directly
int rejit_b200_device_count(void) { return 8; }
void* rejit_b200_pinned_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void rejit_b200_pinned_free(void* p) { free(p); }
}

namespace rejit {

namespace {
char g_status[200] = "";

[[noreturn]] void Fatal(const char* what, long long code) {
  fprintf(stderr, "hostsim_rejit: %s failed (%lld)\n", what, code);
  abort();
}

// A Text produced "on the device" owns its bytes; one made from a host buffer borrows them.
struct TextBody {
  std::string owned;
};

size_t Collect(const char* pattern, const char* text, size_t n, int slabs, std::vector<Match>* out) {
  const uint8_t* t = reinterpret_cast<const uint8_t*>(text ? text : "");
  std::vector<uint64_t> pairs(2 * (n + 2));
  const int64_t k = slabs > 1 ? hostsim_match_all_slabs(pattern, strlen(pattern), t, n, slabs, pairs.data(), n + 2)
                              : hostsim_match_all(pattern, strlen(pattern), 1, t, n, -1, pairs.data(), n + 2, nullptr, 0);
  if (k < 0) Fatal("MatchAll", k);
  if (out)
    for (int64_t i = 0; i < k; ++i) out->push_back(Match{text + pairs[2 * i], text + pairs[2 * i + 1]});
  return static_cast<size_t>(k);
}
}  // namespace

char* const rejit_status_string = g_status;

namespace internal { class RegexpInfo {}; }

Regej::Regej(const char* regexp) : regexp_(regexp), rinfo_(nullptr), status_(RejitSuccess) {
  char message[sizeof g_status] = "";
  uint64_t none[2];
  if (hostsim_match_all(regexp, strlen(regexp), 1, reinterpret_cast<const uint8_t*>(""), 0, -1, none, 1, message, sizeof message) == -1) {
    snprintf(g_status, sizeof g_status, "%s", message);
    status_ = ParserError;
  }
}
Regej::Regej(const string& regexp) : Regej(regexp.c_str()) {}
Regej::~Regej() {}
bool Regej::Compile(MatchType) { return status_ == RejitSuccess; }

bool Regej::MatchFull(const string& text) { return MatchFull(text.c_str(), text.size()); }
bool Regej::MatchFull(const char* text, size_t n) {
  return status_ == RejitSuccess && hostsim_match_full(regexp_, strlen(regexp_), reinterpret_cast<const uint8_t*>(text), n) == 1;
}
bool Regej::MatchAnywhere(const string& text) { return MatchAnywhere(text.c_str(), text.size()); }
bool Regej::MatchAnywhere(const char* text, size_t n) { return status_ == RejitSuccess && Collect(regexp_, text, n, 1, nullptr) > 0; }
bool Regej::MatchFirst(const string& text, Match* m) { return MatchFirst(text.c_str(), text.size(), m); }
bool Regej::MatchFirst(const char* text, size_t n, Match* m) {
  if (status_ != RejitSuccess) return false;
  std::vector<Match> all;
  if (Collect(regexp_, text, n, 1, &all) == 0) return false;
  if (m) *m = all[0];
  return true;
}
size_t Regej::MatchAll(const string& text, std::vector<Match>* out) { return MatchAll(text.c_str(), text.size(), out); }
size_t Regej::MatchAll(const char* text, size_t n, std::vector<Match>* out) {
  if (status_ != RejitSuccess) return 0;
  const size_t k = Collect(regexp_, text, n, 1, out);
  return out ? out->size() : k;
}
size_t Regej::MatchAll(const Text& text, std::vector<Match>* out) { return MatchAll(text.data(), text.size(), out); }
size_t Regej::MatchAllCount(const string& text) { return MatchAllCount(text.c_str(), text.size()); }
size_t Regej::MatchAllCount(const char* text, size_t n) { return status_ == RejitSuccess ? Collect(regexp_, text, n, 1, nullptr) : 0; }
size_t Regej::MatchAllCount(const Text& text) {
  const TextBody* body = static_cast<const TextBody*>(text.handle_);
  return MatchAllCount(body ? body->owned.data() : text.data(), text.size());
}
size_t Regej::MatchAllParallel(const char* text, size_t n, std::vector<Match>* out, int n_gpus) {
  if (status_ != RejitSuccess) return 0;
  const size_t k = Collect(regexp_, text, n, n_gpus, out);
  return out ? out->size() : k;
}
size_t Regej::MatchAllSet(const std::vector<Regej*>& patterns, const char* text, size_t n, std::vector<std::vector<Match> >* out) {
  size_t total = 0;
  if (out) out->resize(patterns.size());
  for (size_t j = 0; j < patterns.size(); ++j) total += Collect(patterns[j]->regexp_, text, n, 1, out ? &(*out)[j] : nullptr);
  return total;
}
size_t Regej::MatchAllCountSet(const std::vector<Regej*>& patterns, const Text& text, std::vector<size_t>* counts) {
  size_t total = 0;
  if (counts) counts->clear();
  for (Regej* re : patterns) {
    const size_t k = re->MatchAllCount(text);
    if (counts) counts->push_back(k);
    total += k;
  }
  return total;
}

bool Regej::ReplaceFirst(string& text, const string& with) {
  Match m;
  if (!MatchFirst(text, &m)) return false;
  Replace(m, text, with);
  return true;
}
size_t Regej::ReplaceAll(string& text, const string& with) {
  if (status_ != RejitSuccess) return 0;
  std::string out(text.size() + (text.size() + 2) * with.size() + 64, '\0');
  uint64_t len = 0;
  const int64_t k = hostsim_replace_all(regexp_, strlen(regexp_), reinterpret_cast<const uint8_t*>(text.data()), text.size(),
                                        reinterpret_cast<const uint8_t*>(with.data()), static_cast<uint32_t>(with.size()),
                                        reinterpret_cast<uint8_t*>(&out[0]), out.size(), &len);
  if (k < 0) Fatal("ReplaceAll", k);
  out.resize(len);
  text.swap(out);
  return static_cast<size_t>(k);
}

Text::Text(const char* text, size_t size, int) : text_(text), size_(size), handle_(nullptr) {}
Text::Text(void* handle, size_t size) : text_(nullptr), size_(size), handle_(handle) {}
Text::~Text() { delete static_cast<TextBody*>(handle_); }
Text* Text::ReplaceAll(Regej& re, const string& with, size_t* n_matches) const {
  const TextBody* body = static_cast<const TextBody*>(handle_);
  TextBody* next = new TextBody;
  next->owned.assign(body ? body->owned.data() : text_, size_);
  const size_t k = re.ReplaceAll(next->owned, with);
  if (n_matches) *n_matches = k;
  return new Text(next, next->owned.size());
}
Text* Text::ReplaceAllSet(const vector<Regej*>& patterns, const vector<string>& withs, vector<size_t>* n_matches) const {
  // the test double applies the calls one after the other (what the device's table must reproduce)
  if (patterns.empty() || patterns.size() != withs.size()) return nullptr;
  const TextBody* body = static_cast<const TextBody*>(handle_);
  TextBody* next = new TextBody;
  next->owned.assign(body ? body->owned.data() : text_, size_);
  if (n_matches) n_matches->clear();
  for (size_t i = 0; i < patterns.size(); ++i) {
    const size_t k = patterns[i]->ReplaceAll(next->owned, withs[i]);
    if (n_matches) n_matches->push_back(k);
  }
  return new Text(next, next->owned.size());
}
string Text::Download() const {
  const TextBody* body = static_cast<const TextBody*>(handle_);
  return body ? body->owned : string(text_, size_);
}

bool MatchFull(const char* re, const string& t) { return Regej(re).MatchFull(t); }
bool MatchFull(const char* re, const char* t, size_t n) { return Regej(re).MatchFull(t, n); }
bool MatchAnywhere(const char* re, const string& t) { return Regej(re).MatchAnywhere(t); }
bool MatchAnywhere(const char* re, const char* t, size_t n) { return Regej(re).MatchAnywhere(t, n); }
bool MatchFirst(const char* re, const string& t, Match* m) { return Regej(re).MatchFirst(t, m); }
bool MatchFirst(const char* re, const char* t, size_t n, Match* m) { return Regej(re).MatchFirst(t, n, m); }
size_t MatchAll(const char* re, const string& t, std::vector<Match>* out) { return Regej(re).MatchAll(t, out); }
size_t MatchAll(const char* re, const char* t, size_t n, std::vector<Match>* out) { return Regej(re).MatchAll(t, n, out); }
size_t MatchAllCount(const char* re, const string& t) { return Regej(re).MatchAllCount(t); }
size_t MatchAllCount(const char* re, const char* t, size_t n) { return Regej(re).MatchAllCount(t, n); }
size_t MatchAllParallel(const char* re, const char* t, size_t n, std::vector<Match>* out, int g) { return Regej(re).MatchAllParallel(t, n, out, g); }

void Replace(Match to_replace, string& text, const string& with) {
  std::vector<Match> one(1, to_replace);
  Replace(&one, text, with);
}
void Replace(std::vector<Match>* to_replace, string& text, const string& with) {
  string rebuilt;
  const char* at = text.c_str();
  for (const Match& m : *to_replace) {
    rebuilt.append(at, static_cast<size_t>(m.begin - at));
    rebuilt.append(with);
    at = m.end;
  }
  rebuilt.append(at, static_cast<size_t>(text.c_str() + text.size() - at));
  text.swap(rebuilt);
}
bool ReplaceFirst(const char* re, string& text, const string& with) { return Regej(re).ReplaceFirst(text, with); }
size_t ReplaceAll(const char* re, string& text, const string& with) { return Regej(re).ReplaceAll(text, with); }

}  // namespace rejit
