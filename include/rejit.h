// rejit.h — public C++ interface of rejit_b200.
//
// Source-compatible with the interface of coreperf/rejit
// (/root/reference/include/rejit.h:41-138): the same names, argument meaning
// and error behaviour, so that a program written against rejit (the regexdna
// and jrep samples, the benchmark engine) compiles and links against this
// library unchanged.  Behind it the pattern is lowered ahead of time to
// tables for hand-written sm_100a CUDA kernels instead of being JIT-compiled
// to x64 (see include/rejit_b200.h for the C boundary and DESIGN.md).
//
// Additions (a program can test for them with `#ifdef REJIT_B200`):
// Regej::MatchAllParallel and the free MatchAllParallel helper, Regej::MatchAllSet,
// and rejit::Text, a text that is copied to the device once and matched many times.
//
// Written from scratch for this project; only the declarations' shapes are
// shared with the reference, because they ARE the compatibility contract.
#ifndef REJIT_H_
#define REJIT_H_

#include <stddef.h>

#include <string>
#include <vector>

#define REJIT_B200 1

using namespace std;   // the reference header exports std:: names to its users

namespace rejit {

// Half-open byte range [begin, end) inside the searched text.
// A zero-length match has begin == end (e.g. "^$" on "" yields one match whose
// two pointers both address the terminator).
struct Match {
  const char* begin;
  const char* end;
};

enum MatchType { kMatchFull, kMatchAnywhere, kMatchFirst, kMatchAll, kNMatchTypes };

// RejitSuccess or a (negative) parser error; the message of the most recent
// error is kept in rejit_status_string (a process-wide 200-byte buffer).
enum Status { RejitSuccess = 0, ParserError = -1 };
extern char* const rejit_status_string;

namespace internal { class RegexpInfo; }

class Regej;

// A text resident in device memory (new in rejit_b200): the constructor copies
// `size` bytes to the device once; every Regej::MatchAll(const Text&, ...) then
// costs a scan and the copy of its match list, no upload.  Matches point into
// the caller's host buffer, which must stay alive and unchanged.
// ReplaceAll produces a NEW text on the device (the rebuild of Replace runs there
// too), so substitutions chain without the text ever returning to the host
// (regex-dna: strip, nine counts, eleven substitutions on one upload).  Such a
// text has no host buffer: data() is NULL, use MatchAllCount / MatchAllCountSet
// / ReplaceAll / Download on it.
class Text {
 public:
  Text(const char* text, size_t size, int device = 0);
  ~Text();
  const char* data() const { return text_; }
  size_t size() const { return size_; }
  // Every match of `re` replaced by `with`; the caller owns the result.
  Text* ReplaceAll(Regej& re, const string& with, size_t* n_matches = NULL) const;
  // ReplaceAll(*patterns[0], withs[0]), then ReplaceAll(*patterns[1], withs[1]), ... — the same text as chaining
  // the calls; patterns that each match one byte (regex-dna's IUB codes) are applied in ONE pass over the text.
  Text* ReplaceAllSet(const std::vector<Regej*>& patterns, const std::vector<string>& withs,
                      std::vector<size_t>* n_matches = NULL) const;
  string Download() const;

 private:
  friend class Regej;
  Text(void* handle, size_t size);
  Text(const Text&);
  Text& operator=(const Text&);
  const char* text_;
  size_t size_;
  void* handle_;
};

class Regej {
 public:
  explicit Regej(const char* regexp);
  explicit Regej(const string& regexp);
  ~Regej();

  Status status() const { return status_; }

  // True iff the whole text is one match.
  bool MatchFull(const string& text);
  bool MatchFull(const char* text, size_t text_size);
  // True iff some match exists.
  bool MatchAnywhere(const string& text);
  bool MatchAnywhere(const char* text, size_t text_size);
  // Left-most longest match.
  bool MatchFirst(const string& text, Match* match);
  bool MatchFirst(const char* text, size_t text_size, Match* match);
  // All left-most longest, non-overlapping matches, APPENDED to *matches;
  // returns matches->size().
  size_t MatchAll(const string& text, std::vector<struct Match>* matches);
  size_t MatchAll(const char* text, size_t text_size, std::vector<struct Match>* matches);
  size_t MatchAllCount(const string& text);
  size_t MatchAllCount(const char* text, size_t text_size);
  // MatchAll over a text that is already on the device (new in rejit_b200).
  size_t MatchAll(const Text& text, std::vector<struct Match>* matches);
  size_t MatchAllCount(const Text& text);
  // (*counts)[i] = patterns[i]->MatchAllCount(text); sets of fixed-length
  // alternations are fused into one scan of the text.  Returns the total.
  static size_t MatchAllCountSet(const std::vector<Regej*>& patterns, const Text& text, std::vector<size_t>* counts);
  // Same result as MatchAll, with the text sharded by contiguous slab over
  // n_gpus devices of this machine (new in rejit_b200).
  size_t MatchAllParallel(const char* text, size_t text_size, std::vector<struct Match>* matches,
                          int n_gpus);

  bool ReplaceFirst(string& text, const string& with);
  size_t ReplaceAll(string& text, const string& with);

  // Several patterns over the same text (new in rejit_b200): the text is copied
  // to the device once; fixed-length alternation sets (regex-dna's variants) are
  // fused into a single scan.  (*matches)[i] receives what
  // patterns[i]->MatchAll(text, text_size, ...) would append.  Returns the total.
  static size_t MatchAllSet(const std::vector<Regej*>& patterns, const char* text, size_t text_size,
                            std::vector<std::vector<struct Match> >* matches);

  // Builds the matcher eagerly (it is otherwise built on first use).
  bool Compile(MatchType match_type);

 private:
  friend class Text;
  char const* const regexp_;
  internal::RegexpInfo* rinfo_;
  Status status_;
};

// One-shot helpers; each call parses and lowers the pattern again.
bool MatchFull(const char* regexp, const string& text);
bool MatchFull(const char* regexp, const char* text, size_t text_size);
bool MatchAnywhere(const char* regexp, const string& text);
bool MatchAnywhere(const char* regexp, const char* text, size_t text_size);
bool MatchFirst(const char* regexp, const string& text, Match* match);
bool MatchFirst(const char* regexp, const char* text, size_t text_size, Match* match);
size_t MatchAll(const char* regexp, const string& text, std::vector<struct Match>* matches);
size_t MatchAll(const char* regexp, const char* text, size_t text_size,
                std::vector<struct Match>* matches);
size_t MatchAllCount(const char* regexp, const string& text);
size_t MatchAllCount(const char* regexp, const char* text, size_t text_size);
size_t MatchAllParallel(const char* regexp, const char* text, size_t text_size,
                        std::vector<struct Match>* matches, int n_gpus);

void Replace(Match to_replace, string& text, const string& with);
void Replace(vector<Match>* to_replace, string& text, const string& with);
bool ReplaceFirst(const char* regexp, string& text, const string& with);
size_t ReplaceAll(const char* regexp, string& text, const string& with);

}  // namespace rejit

#endif  // REJIT_H_
