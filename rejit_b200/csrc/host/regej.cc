// The rejit::Regej facade over the C ABI (include/rejit_b200.h).
//
// Mirrors the behaviour of the reference's API glue
// (/root/reference/src/rejit.cc:29-267): the constructor parses (ERE only) and
// records the status; matchers build lazily; MatchAll APPENDS to the caller's
// vector and returns its size; MatchAllCount/ReplaceAll are built on MatchAll;
// Replace rebuilds the string in one pass.  Internal failures (no CUDA device,
// CUDA errors) are fatal, like the reference's rejit_fatal
// (/root/reference/src/checks.cc:19-26): there is no CPU matcher to fall back to.
#include "../../../include/rejit.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../../include/rejit_b200.h"

namespace rejit {

namespace {
char g_status[200] = "";
}
char* const rejit_status_string = g_status;

namespace internal {
class RegexpInfo {
 public:
  rejit_b200_ir* ir = nullptr;
  rejit_b200_program* program = nullptr;
  ~RegexpInfo() {
    if (program) rejit_b200_program_free(program);
    if (ir) rejit_b200_ir_free(ir);
  }
};
}  // namespace internal

namespace {

[[noreturn]] void Fatal(const char* what, const char* detail) {
  fprintf(stderr, "rejit_b200 fatal: %s: %s\n", what, detail);
  abort();
}

Status ParseInto(const char* regexp, internal::RegexpInfo* info) {
  char err[sizeof g_status];
  err[0] = 0;
  int rc = rejit_b200_parse(regexp, regexp ? strlen(regexp) : 0, /*parser_opt=*/1, &info->ir, err, sizeof err);
  if (rc != 0) {
    snprintf(g_status, sizeof g_status, "%s", err);
    return ParserError;
  }
  return RejitSuccess;
}

}  // namespace

Regej::Regej(const char* regexp) : regexp_(regexp), rinfo_(new internal::RegexpInfo()) {
  status_ = ParseInto(regexp_, rinfo_);
}

Regej::Regej(const string& regexp) : regexp_(regexp.c_str()), rinfo_(new internal::RegexpInfo()) {
  status_ = ParseInto(regexp_, rinfo_);
}

Regej::~Regej() { delete rinfo_; }

bool Regej::Compile(MatchType) {
  if (status_ != RejitSuccess) return false;
  if (rinfo_->program) return true;
  char err[256];
  err[0] = 0;
  rinfo_->program = rejit_b200_compile(rinfo_->ir, err, sizeof err);
  if (!rinfo_->program) {
    snprintf(g_status, sizeof g_status, "%s", err);
    return false;
  }
  return true;
}

bool Regej::MatchFull(const string& text) { return MatchFull(text.c_str(), text.size()); }
bool Regej::MatchFull(const char* text, size_t text_size) {
  if (!Compile(kMatchFull)) return false;
  char err[256];
  int r = rejit_b200_match_full(rinfo_->program, text, text_size, err, sizeof err);
  if (r < 0) Fatal("MatchFull", err);
  return r == 1;
}

bool Regej::MatchAnywhere(const string& text) { return MatchAnywhere(text.c_str(), text.size()); }
bool Regej::MatchAnywhere(const char* text, size_t text_size) {
  if (!Compile(kMatchAnywhere)) return false;
  char err[256];
  int r = rejit_b200_match_anywhere(rinfo_->program, text, text_size, err, sizeof err);
  if (r < 0) Fatal("MatchAnywhere", err);
  return r == 1;
}

bool Regej::MatchFirst(const string& text, Match* match) { return MatchFirst(text.c_str(), text.size(), match); }
bool Regej::MatchFirst(const char* text, size_t text_size, Match* match) {
  if (!Compile(kMatchFirst)) return false;
  char err[256];
  uint64_t pair[2] = {0, 0};
  int r = rejit_b200_match_first(rinfo_->program, text, text_size, pair, err, sizeof err);
  if (r < 0) Fatal("MatchFirst", err);
  if (r == 1 && match) {
    match->begin = text + pair[0];
    match->end = text + pair[1];
  }
  return r == 1;
}

size_t Regej::MatchAll(const string& text, vector<Match>* matches) {
  return MatchAll(text.c_str(), text.size(), matches);
}
size_t Regej::MatchAll(const char* text, size_t text_size, vector<Match>* matches) {
  if (!Compile(kMatchAll)) return 0;
  char err[256];
  uint64_t* pairs = nullptr;
  int64_t n = rejit_b200_match_all_alloc(rinfo_->program, text, text_size, &pairs, nullptr, err, sizeof err);
  if (n < 0) Fatal("MatchAll", err);
  if (matches) {
    matches->reserve(matches->size() + static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) matches->push_back(Match{text + pairs[2 * i], text + pairs[2 * i + 1]});
  }
  rejit_b200_free(pairs);
  return matches ? matches->size() : static_cast<size_t>(n);
}

Text::Text(const char* text, size_t size, int device) : text_(text), size_(size), handle_(nullptr) {
  char err[256];
  err[0] = 0;
  handle_ = rejit_b200_text_upload(device, text, size, err, sizeof err);
  if (!handle_) Fatal("Text", err);
}

Text::Text(void* handle, size_t size) : text_(nullptr), size_(size), handle_(handle) {}

Text* Text::ReplaceAll(Regej& re, const string& with, size_t* n_matches) const {
  if (!re.Compile(kMatchAll)) return nullptr;
  char err[256];
  err[0] = 0;
  int64_t n = 0;
  rejit_b200_text* out = rejit_b200_replace_all_text(re.rinfo_->program, static_cast<const rejit_b200_text*>(handle_),
                                                     with.data(), with.size(), &n, nullptr, err, sizeof err);
  if (!out) Fatal("Text::ReplaceAll", err);
  if (n_matches) *n_matches = static_cast<size_t>(n);
  return new Text(out, rejit_b200_text_length(out));
}

Text* Text::ReplaceAllSet(const vector<Regej*>& patterns, const vector<string>& withs, vector<size_t>* n_matches) const {
  if (patterns.empty() || patterns.size() != withs.size()) return nullptr;
  vector<rejit_b200_program*> progs;
  vector<const char*> ws;
  vector<size_t> lens;
  for (size_t i = 0; i < patterns.size(); ++i) {
    if (!patterns[i] || !patterns[i]->Compile(kMatchAll)) return nullptr;
    progs.push_back(patterns[i]->rinfo_->program);
    ws.push_back(withs[i].data());
    lens.push_back(withs[i].size());
  }
  vector<int64_t> counts(patterns.size(), 0);
  char err[256];
  err[0] = 0;
  rejit_b200_text* out = rejit_b200_replace_all_set_text(progs.data(), static_cast<int>(progs.size()),
                                                         static_cast<const rejit_b200_text*>(handle_), ws.data(), lens.data(),
                                                         counts.data(), nullptr, err, sizeof err);
  if (!out) Fatal("Text::ReplaceAllSet", err);
  if (n_matches) n_matches->assign(counts.begin(), counts.end());
  return new Text(out, rejit_b200_text_length(out));
}

string Text::Download() const {
  string out(size_, '\0');
  char err[256];
  err[0] = 0;
  if (rejit_b200_text_download(static_cast<const rejit_b200_text*>(handle_), size_ ? &out[0] : nullptr, size_, err, sizeof err) != 0)
    Fatal("Text::Download", err);
  return out;
}

Text::~Text() {
  if (handle_) rejit_b200_text_free(static_cast<rejit_b200_text*>(handle_));
}

size_t Regej::MatchAll(const Text& text, vector<Match>* matches) {
  if (!Compile(kMatchAll)) return 0;
  char err[256];
  uint64_t* pairs = nullptr;
  int64_t n = rejit_b200_match_all_text(rinfo_->program, static_cast<const rejit_b200_text*>(text.handle_), &pairs, nullptr,
                                        err, sizeof err);
  if (n < 0) Fatal("MatchAll", err);
  if (matches) {
    matches->reserve(matches->size() + static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) matches->push_back(Match{text.text_ + pairs[2 * i], text.text_ + pairs[2 * i + 1]});
  }
  rejit_b200_free(pairs);
  return matches ? matches->size() : static_cast<size_t>(n);
}

size_t Regej::MatchAllCount(const Text& text) {
  if (!Compile(kMatchAll)) return 0;
  char err[256];
  int64_t n = rejit_b200_match_all_text(rinfo_->program, static_cast<const rejit_b200_text*>(text.handle_), nullptr, nullptr,
                                        err, sizeof err);
  if (n < 0) Fatal("MatchAllCount", err);
  return static_cast<size_t>(n);
}

size_t Regej::MatchAllCountSet(const std::vector<Regej*>& patterns, const Text& text, std::vector<size_t>* counts) {
  std::vector<rejit_b200_program*> progs;
  for (Regej* r : patterns) {
    if (!r->Compile(kMatchAll)) return 0;
    progs.push_back(r->rinfo_->program);
  }
  if (progs.empty()) return 0;
  char err[256];
  err[0] = 0;
  rejit_b200_set* set = rejit_b200_set_create(progs.data(), static_cast<int>(progs.size()));
  if (!set) Fatal("MatchAllCountSet", "cannot create the pattern set");
  std::vector<int64_t> found(progs.size(), 0);
  if (rejit_b200_match_all_set_text(set, static_cast<const rejit_b200_text*>(text.handle_), found.data(), nullptr, nullptr, err,
                                    sizeof err) != 0)
    Fatal("MatchAllCountSet", err);
  rejit_b200_set_free(set);
  size_t total = 0;
  if (counts) counts->assign(found.begin(), found.end());
  for (int64_t c : found) total += static_cast<size_t>(c);
  return total;
}

size_t Regej::MatchAllParallel(const char* text, size_t text_size, vector<Match>* matches, int n_gpus) {
  if (!Compile(kMatchAll)) return 0;
  char err[256];
  uint64_t* pairs = nullptr;
  int64_t n = rejit_b200_match_all_multi_gpu(rinfo_->program, text, text_size, n_gpus, &pairs, nullptr, err, sizeof err);
  if (n < 0) Fatal("MatchAllParallel", err);
  if (matches) {
    matches->reserve(matches->size() + static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) matches->push_back(Match{text + pairs[2 * i], text + pairs[2 * i + 1]});
  }
  rejit_b200_free(pairs);
  return matches ? matches->size() : static_cast<size_t>(n);
}

size_t Regej::MatchAllSet(const std::vector<Regej*>& patterns, const char* text, size_t text_size,
                          std::vector<std::vector<Match> >* matches) {
  std::vector<rejit_b200_program*> progs;
  for (Regej* r : patterns) {
    if (!r->Compile(kMatchAll)) return 0;
    progs.push_back(r->rinfo_->program);
  }
  if (progs.empty()) return 0;
  char err[256];
  err[0] = 0;
  rejit_b200_set* set = rejit_b200_set_create(progs.data(), static_cast<int>(progs.size()));
  rejit_b200_text* dtext = rejit_b200_text_upload(0, text, text_size, err, sizeof err);
  if (!set || !dtext) Fatal("MatchAllSet", err);
  std::vector<int64_t> counts(progs.size(), 0);
  std::vector<uint64_t*> pairs(progs.size(), nullptr);
  if (rejit_b200_match_all_set_text(set, dtext, counts.data(), pairs.data(), nullptr, err, sizeof err) != 0)
    Fatal("MatchAllSet", err);
  size_t total = 0;
  if (matches) matches->resize(progs.size());
  for (size_t j = 0; j < progs.size(); ++j) {
    if (matches) {
      std::vector<Match>& out = (*matches)[j];
      out.reserve(out.size() + static_cast<size_t>(counts[j]));
      for (int64_t i = 0; i < counts[j]; ++i) out.push_back(Match{text + pairs[j][2 * i], text + pairs[j][2 * i + 1]});
    }
    total += static_cast<size_t>(counts[j]);
    rejit_b200_free(pairs[j]);
  }
  rejit_b200_text_free(dtext);
  rejit_b200_set_free(set);
  return total;
}

size_t Regej::MatchAllCount(const string& text) { return MatchAllCount(text.c_str(), text.size()); }
size_t Regej::MatchAllCount(const char* text, size_t text_size) {
  vector<Match> found;
  return MatchAll(text, text_size, &found);
}

bool Regej::ReplaceFirst(string& text, const string& with) {
  Match m;
  if (!MatchFirst(text, &m)) return false;
  Replace(m, text, with);
  return true;
}

size_t Regej::ReplaceAll(string& text, const string& with) {
  // the rebuild (src/rejit.cc:97-112) runs on the device too
  if (!Compile(kMatchAll)) return 0;
  char err[256];
  char* rebuilt = nullptr;
  size_t len = 0;
  int64_t n = rejit_b200_replace_all(rinfo_->program, text.data(), text.size(), with.data(), with.size(), &rebuilt, &len,
                                     nullptr, err, sizeof err);
  if (n < 0) Fatal("ReplaceAll", err);
  text.assign(rebuilt, len);
  rejit_b200_free(rebuilt);
  return static_cast<size_t>(n);
}

// ---- free helpers ------------------------------------------------------------
bool MatchFull(const char* regexp, const string& text) { return MatchFull(regexp, text.c_str(), text.size()); }
bool MatchFull(const char* regexp, const char* text, size_t n) { Regej re(regexp); return re.MatchFull(text, n); }
bool MatchAnywhere(const char* regexp, const string& text) { return MatchAnywhere(regexp, text.c_str(), text.size()); }
bool MatchAnywhere(const char* regexp, const char* text, size_t n) { Regej re(regexp); return re.MatchAnywhere(text, n); }
bool MatchFirst(const char* regexp, const string& text, Match* m) { return MatchFirst(regexp, text.c_str(), text.size(), m); }
bool MatchFirst(const char* regexp, const char* text, size_t n, Match* m) { Regej re(regexp); return re.MatchFirst(text, n, m); }
size_t MatchAll(const char* regexp, const string& text, vector<Match>* out) { return MatchAll(regexp, text.c_str(), text.size(), out); }
size_t MatchAll(const char* regexp, const char* text, size_t n, vector<Match>* out) { Regej re(regexp); return re.MatchAll(text, n, out); }
size_t MatchAllCount(const char* regexp, const string& text) { return MatchAllCount(regexp, text.c_str(), text.size()); }
size_t MatchAllCount(const char* regexp, const char* text, size_t n) { Regej re(regexp); return re.MatchAllCount(text, n); }
size_t MatchAllParallel(const char* regexp, const char* text, size_t n, vector<Match>* out, int n_gpus) {
  Regej re(regexp);
  return re.MatchAllParallel(text, n, out, n_gpus);
}

void Replace(Match to_replace, string& text, const string& with) {
  vector<Match> one(1, to_replace);
  Replace(&one, text, with);
}

void Replace(vector<Match>* to_replace, string& text, const string& with) {
  string rebuilt;
  rebuilt.reserve(text.size() + text.size() / 16);
  const char* base = text.c_str();
  const char* at = base;
  for (const Match& m : *to_replace) {
    rebuilt.append(at, static_cast<size_t>(m.begin - at));
    rebuilt.append(with);
    at = m.end;
  }
  rebuilt.append(at, static_cast<size_t>(base + text.size() - at));
  text.swap(rebuilt);
}

bool ReplaceFirst(const char* regexp, string& text, const string& with) { Regej re(regexp); return re.ReplaceFirst(text, with); }
size_t ReplaceAll(const char* regexp, string& text, const string& with) { Regej re(regexp); return re.ReplaceAll(text, with); }

}  // namespace rejit
