// jrep — a grep-like front end on rejit, written for a
// device-side matcher: files are staged into ONE host blob per batch (a per-file
// offset table beside it) and every batch costs one upload and one pass for the
// pattern plus one for the line index "^", instead of two MatchAll calls per
// file.  The walk plans a batch from the sizes it already has, N threads fill it
// (-j N), and a matcher thread per device (--gpus) uploads, scans and formats it
// while the next batches are being planned and filled; output is written in
// batch order.  Output, options and exit codes follow the reference's jrep
// (/root/reference/sample/jrep.cc: options :84-126, per-file printing :261-405,
// path handling :518-545), which this file restates and does not copy.
//
// Compiles against include/rejit.h of this repository (REJIT_B200 defined: pinned
// staging, one rejit::Text upload per batch, one device per matcher) and,
// unchanged, against the reference's header and library (host-pointer MatchAll
// calls).  tests/test_samples.py checks both builds: the second against the
// reference's own jrep and golden output, the first against the same golden
// output on a test double of the library (CPU tier) and on librejit_b200.so
// (GPU tier).
//
// Exactness of batching.  Files are independent texts in the reference.  In the
// blob they are separated by one '\n', which gives every file the same line
// context at both ends as a text of its own ("^" holds after '\n' and at offset
// 0, "$" before '\n' and at the end).  A match that swallows a separator would
// not exist in the reference, and it may hide matches that do (also the empty
// match at the first byte of a file it merely ends at): every file such a match
// touches or abuts is scanned again on its own bytes (rare: the pattern has to
// match across "last bytes of a file, '\n'").
#include <errno.h>
#include <fcntl.h>
#include <ftw.h>
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "rejit.h"
#ifdef REJIT_B200
#include "rejit_b200.h"
#endif

namespace {

// JREP_TRACE=1: seconds spent staging (reading files into the blob), matching and printing, on stderr at exit.
struct Trace {
  bool on = getenv("JREP_TRACE") != nullptr;
  double stage = 0, match = 0, print = 0;
  size_t bytes = 0, batches = 0, files = 0, reruns = 0;
  static double Now() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return static_cast<double>(t.tv_sec) + 1e-9 * static_cast<double>(t.tv_nsec);
  }
  ~Trace() {
    if (on)
      fprintf(stderr, "[jrep] %zu files, %zu bytes, %zu batches, %zu files matched again alone; stage %.3f s, match %.3f s, "
              "print %.3f s\n", files, bytes, batches, reruns, stage, match, print);
  }
};
Trace g_trace;

struct Options {
  const char* pattern = nullptr;
  std::vector<const char*> paths;
  bool with_filename = false;
  bool line_number = false;
  bool colour = false;
  int recursive = 0;             // 0 no, 1 do not follow symlinks, 2 follow them
  // threads that read a batch's files into the blob (0: the walking thread reads).  The reference's default is 0;
  // with a device-side matcher staging IS the bottleneck (1.6 GB/s with one thread, 12 GB/s with eight, page
  // cache hot), so this library's build stages in parallel unless told otherwise.
#ifdef REJIT_B200
  unsigned jobs = DefaultJobs();
#else
  unsigned jobs = 0;
#endif
  static unsigned DefaultJobs() {
    const unsigned hw = std::thread::hardware_concurrency();
    return hw > 1 ? std::min(16u, hw - 1u) : 1u;
  }
  unsigned nopenfd = 1024;
  unsigned before = 0;
  unsigned after = 0;
  // bytes staged per matcher call; 0: one call per file (the reference's granularity).  16 MiB: the pipeline
  // fills and drains by one batch, so small batches finish sooner (measured: 0.09-0.13 s at 8-16 MiB against
  // 0.14 s at 64 MiB and 0.22 s at 1 GiB for a 512 MB tree); a batch stays far below the B200's L2, so the
  // line-index pass and the per-file re-runs find it there; and 16 MiB still gives every SM ~100 KB to scan.
  size_t batch_bytes = size_t(16) << 20;
  int gpus = 1;                  // devices: batches go to them in rotation (independent files: replicas)
  bool shard = false;            // instead: every batch is one text cut into slabs over the devices (MatchAllParallel)
};

struct FileSpan {
  std::string path;
  size_t begin;   // offset of the first byte in the blob
  size_t size;
  int error;      // errno of a failed open by a staging thread (-j)
};

// ---- staging ------------------------------------------------------------------
// One grow-only host buffer; pinned when the library offers it, so that the
// upload of a batch is one DMA transfer.
class Blob {
 public:
  ~Blob() { Release(); }
  char* data() { return data_; }
  size_t used() const { return used_; }
  void Clear() { used_ = 0; }
  char* Extend(size_t bytes) {
    if (used_ + bytes > capacity_) Grow(used_ + bytes);
    char* at = data_ + used_;
    used_ += bytes;
    return at;
  }
  void Shrink(size_t bytes) { used_ -= bytes; }

 private:
  void Grow(size_t need) {
    size_t cap = std::max(need, 2 * capacity_);
    cap = std::max(cap, size_t(1) << 20);
    char* fresh = Allocate(cap);
    if (!fresh) {
      fprintf(stderr, "jrep: cannot allocate %zu bytes\n", cap);
      exit(ENOMEM);
    }
    if (used_) memcpy(fresh, data_, used_);
    bool was_pinned = pinned_;
    pinned_ = fresh_pinned_;
    std::swap(fresh, data_);
    Free(fresh, was_pinned);
    capacity_ = cap;
  }
  char* Allocate(size_t bytes) {
    fresh_pinned_ = false;
#ifdef REJIT_B200
    if (void* p = rejit_b200_pinned_alloc(bytes)) {
      fresh_pinned_ = true;
      return static_cast<char*>(p);
    }
#endif
    return static_cast<char*>(malloc(bytes));
  }
  static void Free(char* p, bool pinned) {
    if (!p) return;
#ifdef REJIT_B200
    if (pinned) {
      rejit_b200_pinned_free(p);
      return;
    }
#endif
    (void)pinned;
    free(p);
  }
  void Release() {
    Free(data_, pinned_);
    data_ = nullptr;
  }
  char* data_ = nullptr;
  size_t capacity_ = 0, used_ = 0;
  bool pinned_ = false, fresh_pinned_ = false;
};

// ---- printing (one file) ---------------------------------------------------------
class Printer {
 public:
  Printer(const Options& o, std::string* out) : o_(o), out_(*out) {}

  // `lines` = the "^" matches of the file followed by one (end, end) sentinel;
  // `found` = the pattern's matches, at least one.  The walk below reproduces the
  // reference's output byte for byte, including what it prints for matches that
  // span lines (a line may be printed twice) and for a match at the very end of a
  // file (a head without a line).
  void File(const std::string& name, const rejit::Match* found, size_t n_found,
            const std::vector<rejit::Match>& lines) {
    const size_t n_lines = lines.size();
    const char* const file_end = lines.back().begin;
    size_t il = 0, im = 0;
    while (il < n_lines && im < n_found) {
      while (il < n_lines && lines[il].begin <= found[im].begin) ++il;
      --il;                                            // the line the match begins on

      if (o_.before) {
        out_ += "--\n";
        for (size_t i = il > o_.before ? il - o_.before : 0; i < il; ++i) {
          Head(name, i + 1, '-');
          Bytes(lines[i].begin, lines[i + 1].begin);
        }
      }

      Head(name, il + 1, ':');
      const char* at = lines[il].begin;
      if (at == file_end) {                            // nothing after the last line start
        ++im;
        continue;
      }
      while (im < n_found && found[im].begin < lines[il + 1].begin) {
        Bytes(at, found[im].begin);
        if (o_.colour) out_ += "\x1B[31m";
        Bytes(found[im].begin, found[im].end);
        if (o_.colour) out_ += "\x1B[0m";
        at = found[im].end;
        ++im;
      }
      size_t tail = il;                                // first line start at or after the last match's end
      while (lines[tail].begin < found[im - 1].end) ++tail;
      Bytes(found[im - 1].end, lines[tail].end);

      if (o_.after) {
        const size_t stop = std::min(tail + size_t(o_.after), n_lines - 1);
        size_t i = tail;
        for (; i < stop; ++i) {
          Head(name, i + 1, '-');
          Bytes(lines[i].begin, lines[i + 1].begin);
        }
        if (i == n_lines - 1) out_ += "\n";
        out_ += "--\n";
      }
    }
  }

 private:
  void Head(const std::string& name, size_t line, char separator) {
    if (o_.with_filename) {
      out_ += name;
      out_ += separator;
    }
    if (o_.line_number) {
      char number[24];
      int n = snprintf(number, sizeof number, "%d", static_cast<int>(line));
      out_.append(number, n);
      out_ += separator;
    }
  }
  void Bytes(const char* from, const char* to) {
    if (to > from) out_.append(from, static_cast<size_t>(to - from));
  }
  const Options& o_;
  std::string& out_;
};

// ---- batches -------------------------------------------------------------------------
// One more than there are matchers: while each matcher (one per device) works on a batch (upload, scans,
// formatting), the walk plans and the staging threads fill the next one.  Output is written in batch order.
struct Batch {
  Blob blob;
  std::vector<FileSpan> files;
  size_t planned = 0;            // -j: bytes of the batch being planned (files + separators)
  bool gaps = false;             // a file shrank or vanished while it was staged: no batch-wide scan
  bool staged = false;           // handed to a matcher, not yet printed
  size_t seq = 0;                // position in the run: batches are printed in this order
  std::string output;
};

class Jrep {
 public:
  Jrep(const Options& o) : o_(o), re_(o.pattern), sol_("^"), n_matchers_(o.shard ? 1 : std::max(1, o.gpus)) {
    for (size_t i = 0; i < n_matchers_ + 1; ++i) batches_.emplace_back(new Batch);
    cur_ = batches_[0].get();
  }
  ~Jrep() { Finish(); }

  bool Ready() {
    if (re_.status() != rejit::RejitSuccess) {
      fprintf(stderr, "jrep: %s\n", rejit::rejit_status_string);
      return false;
    }
    re_.Compile(rejit::kMatchAll);
    sol_.Compile(rejit::kMatchAll);
    return true;
  }

  // Stages one file.  Returns 0 or the errno of the failed open (the reference
  // stops the whole run there: sample/jrep.cc:269-274, 540-541).  With -j N only
  // the place in the blob is reserved here; N threads fill the batch in Run().
  int Add(const char* path, const struct stat& known) {
    if (pending_error_) return pending_error_;
    int fd = -1;
    struct stat st;
    size_t size;
    if (o_.jobs > 0) {                                 // the walk has the size already: no system call here
      size = static_cast<size_t>(known.st_size);
      if (size == 0) return access(path, R_OK) == 0 ? 0 : errno;
    } else {
      fd = open(path, O_RDONLY);
      if (fd < 0) return errno;
      size = fstat(fd, &st) == 0 ? static_cast<size_t>(st.st_size) : 0;
      if (size == 0) {
        close(fd);
        return 0;
      }
    }
    const size_t staged = o_.jobs > 0 ? cur_->planned : cur_->blob.used();
    if (!cur_->files.empty() && (o_.batch_bytes == 0 || staged + size + 1 > o_.batch_bytes)) {
      Run();
      if (pending_error_) {
        if (fd >= 0) close(fd);
        return pending_error_;
      }
    }
    if (o_.jobs > 0) {                                 // only a place in the batch; Stage() allocates once and fills it
      if (!cur_->files.empty()) ++cur_->planned;
      cur_->files.push_back(FileSpan{path, cur_->planned, size, 0});
      cur_->planned += size;
      return 0;
    }
    if (!cur_->files.empty()) *cur_->blob.Extend(1) = '\n';      // the separator
    const size_t begin = cur_->blob.used();
    char* at = cur_->blob.Extend(size);
    const double t0 = Trace::Now();
    const size_t got = ReadFile(fd, at, size);
    close(fd);
    g_trace.stage += Trace::Now() - t0;
    cur_->blob.Shrink(size - got);                          // the file shrank while we read it
    if (got == 0) {
      if (!cur_->files.empty()) cur_->blob.Shrink(1);
      return 0;
    }
    cur_->files.push_back(FileSpan{path, begin, got, 0});
    return 0;
  }

  // The errno that ended the run, if a staging thread could not open a file.
  int error() const { return pending_error_; }

  // Stages the batch that was planned and hands it to a matcher: inline without -j, else to the matcher
  // thread of its device, so that uploading, scanning and formatting batch k overlap planning and staging
  // batch k + 1 (and, with several devices, the work on batches k - 1, k - 2, ...).
  void Run() {
    Batch& b = *cur_;
    if (b.files.empty()) return;
    b.gaps = false;
    const double t0 = Trace::Now();
    if (o_.jobs > 0 && !Stage(&b.gaps)) return;
    g_trace.stage += Trace::Now() - t0;
    if (o_.jobs == 0) {
      Process(b, 0);
      Write(&b);
      return;
    }
    {
      std::unique_lock<std::mutex> lk(mu_);
      if (matchers_.empty())
        for (size_t d = 0; d < n_matchers_; ++d) matchers_.push_back(std::thread(&Jrep::MatcherLoop, this, d));
      b.seq = submitted_++;
      b.staged = true;
      cur_ = batches_[submitted_ % batches_.size()].get();
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !cur_->staged; });
  }

  // Flushes the last batch and waits for the matchers.
  void Finish() {
    Run();
    if (matchers_.empty()) return;
    {
      std::unique_lock<std::mutex> lk(mu_);
      done_ = true;
    }
    cv_.notify_all();
    for (std::thread& t : matchers_) t.join();
    matchers_.clear();
  }

 private:
  // Matcher d takes batches d, d + n, d + 2n, ... (n matchers), each on device d, and writes a batch's
  // output when all batches before it have been written.
  void MatcherLoop(size_t d) {
    for (size_t k = d;; k += n_matchers_) {
      Batch& b = *batches_[k % batches_.size()];
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return (b.staged && b.seq == k) || (done_ && submitted_ <= k); });
        if (!(b.staged && b.seq == k)) return;
      }
      Process(b, static_cast<int>(d));
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return printed_ == k; });
      }
      Write(&b);
      {
        std::unique_lock<std::mutex> lk(mu_);
        ++printed_;
        b.staged = false;
      }
      cv_.notify_all();
    }
  }

  static void Write(Batch* b) {                        // one writer at a time: inline, or the matcher whose turn it is
    if (!b->output.empty()) fwrite(b->output.data(), 1, b->output.size(), stdout);
    fflush(stdout);
    b->output.clear();
  }

  // Scans one staged batch on `device`, formats its output, and empties it.
  void Process(Batch& batch, int device) {
    Printer printer(o_, &batch.output);
    std::vector<FileSpan>& files = batch.files;
    const bool gaps = batch.gaps;
    const char* text = batch.blob.data();
    const size_t length = batch.blob.used();
    std::vector<rejit::Match> found, lines;
    double t1 = Trace::Now(), t0;
    const size_t n_files = files.size();
    size_t reruns = 0;
    double match_s = 0;
    (void)device;
#ifdef REJIT_B200
    std::unique_ptr<rejit::Text> resident;             // the only upload of the batch
    if (!gaps) {
      resident.reset(new rejit::Text(text, length, device));
      if (o_.shard && o_.gpus > 1) re_.MatchAllParallel(text, length, &found, o_.gpus);
      else re_.MatchAll(*resident, &found);
    }
#else
    if (!gaps) re_.MatchAll(text, length, &found);
#endif

    // Which file does each match begin in (a file owns [begin, begin + size], its
    // separator's position included), and which files does a match that swallows
    // a separator touch.
    std::vector<char> alone(files.size(), gaps ? 1 : 0);   // a batch with holes: every file on its own bytes
    std::vector<size_t> first(files.size() + 1, 0);   // found[first[f] .. first[f+1]) begin in file f
    size_t f = 0;
    for (size_t i = 0; i < found.size(); ++i) {
      const size_t b = static_cast<size_t>(found[i].begin - text), e = static_cast<size_t>(found[i].end - text);
      while (f + 1 < files.size() && b > files[f].begin + files[f].size) first[++f] = i;
      if (e > files[f].begin + files[f].size)                         // a file the match only abuts counts too:
        for (size_t g = f; g < files.size() && files[g].begin <= e; ++g) alone[g] = 1;   // its empty match there is lost
    }
    while (f < files.size()) first[++f] = found.size();

    // The line index: one pass over the resident batch when much of it has
    // matches, else per matching file (the reference indexes only those).
    size_t hit_files = 0, hit_bytes = 0;
    for (f = 0; f < files.size(); ++f)
      if (!alone[f] && first[f] != first[f + 1]) ++hit_files, hit_bytes += files[f].size;
    const bool whole = !gaps && (hit_files > 64 || hit_bytes * 8 > length);
    if (whole && hit_files) {
#ifdef REJIT_B200
      if (o_.shard && o_.gpus > 1) sol_.MatchAllParallel(text, length, &lines, o_.gpus);
      else sol_.MatchAll(*resident, &lines);
#else
      sol_.MatchAll(text, length, &lines);
#endif
    }

    t0 = Trace::Now();
    match_s += t0 - t1;
    double again = 0;
    size_t line_at = 0;
    std::vector<rejit::Match> file_lines, file_found;
    for (f = 0; f < files.size(); ++f) {
      const char* begin = text + files[f].begin;
      const char* end = begin + files[f].size;
      const rejit::Match* mine = found.data() + first[f];
      size_t n_mine = first[f + 1] - first[f];
      file_lines.clear();
      const double m0 = Trace::Now();
      if (alone[f]) {                                  // as a text of its own
        file_found.clear();
        re_.MatchAll(begin, files[f].size, &file_found);
        mine = file_found.data();
        n_mine = file_found.size();
        ++reruns;
      }
      if (n_mine == 0) {
        again += Trace::Now() - m0;
        continue;
      }
      if (alone[f] || !whole) {
        sol_.MatchAll(begin, files[f].size, &file_lines);
        again += Trace::Now() - m0;
      } else {
        while (line_at < lines.size() && lines[line_at].begin < begin) ++line_at;
        for (; line_at < lines.size() && lines[line_at].begin <= end; ++line_at) file_lines.push_back(lines[line_at]);
      }
      file_lines.push_back(rejit::Match{end, end});    // lets the last line be printed (sample/jrep.cc:294-296)
      printer.File(files[f].path, mine, n_mine, file_lines);
    }
    {
      std::unique_lock<std::mutex> lk(mu_);            // several matchers report here
      g_trace.bytes += length;
      g_trace.files += n_files;
      ++g_trace.batches;
      g_trace.reruns += reruns;
      g_trace.match += match_s + again;
      g_trace.print += Trace::Now() - t0 - again;
    }
    files.clear();
    batch.blob.Clear();
    batch.planned = 0;
  }

  static size_t ReadFile(int fd, char* at, size_t size) {
    size_t got = 0;
    while (got < size) {
      ssize_t r = read(fd, at + got, size - got);
      if (r < 0 && errno == EINTR) continue;
      if (r <= 0) break;
      got += static_cast<size_t>(r);
    }
    return got;
  }

  // -j N: N threads read the batch's files into their reserved places (the blob does not move while
  // they run).  A file that cannot be opened ends the run there, as in Add(): the batch is cut before
  // it.  A file that shrank since stat() leaves a hole, filled with separators.  False: nothing to scan.
  bool Stage(bool* gaps) {
    cur_->blob.Clear();
    cur_->blob.Extend(cur_->planned);
    cur_->planned = 0;
    std::atomic<size_t> next(0);
    char* const base = cur_->blob.data();
    for (size_t i = 1; i < cur_->files.size(); ++i) base[cur_->files[i].begin - 1] = '\n';      // the separators
    auto work = [&]() {
      for (size_t i; (i = next.fetch_add(1)) < cur_->files.size();) {
        FileSpan& f = cur_->files[i];
        int fd = open(f.path.c_str(), O_RDONLY);
        if (fd < 0) {
          f.error = errno ? errno : EIO;
          continue;
        }
        const size_t got = ReadFile(fd, base + f.begin, f.size);
        close(fd);
        if (got < f.size) memset(base + f.begin + got, '\n', f.size - got);
        f.size = got | (got < f.size ? kShort : 0);
      }
    };
    std::vector<std::thread> pool;
    const size_t n_threads = std::min<size_t>(o_.jobs, cur_->files.size());
    for (size_t t = 1; t < n_threads; ++t) pool.push_back(std::thread(work));
    work();
    for (std::thread& t : pool) t.join();
    size_t keep = cur_->files.size();
    for (size_t i = 0; i < cur_->files.size(); ++i) {
      if (cur_->files[i].error) {
        pending_error_ = cur_->files[i].error;
        keep = i;
        break;
      }
      if (cur_->files[i].size & kShort) {
        cur_->files[i].size &= ~kShort;
        *gaps = true;
      }
    }
    if (keep < cur_->files.size()) {
      cur_->files.resize(keep);
      cur_->blob.Clear();
      if (keep) cur_->blob.Extend(cur_->files[keep - 1].begin + cur_->files[keep - 1].size);
    }
    // files that turned out empty own nothing (the reference skips them, sample/jrep.cc:277-279)
    size_t w = 0;
    for (size_t i = 0; i < cur_->files.size(); ++i)
      if (cur_->files[i].size) cur_->files[w++] = cur_->files[i]; else *gaps = true;
    cur_->files.resize(w);
    if (cur_->files.empty()) {
      cur_->blob.Clear();
      return false;
    }
    return true;
  }

  static const size_t kShort = size_t(1) << (sizeof(size_t) * 8 - 1);
  int pending_error_ = 0;
  const Options& o_;
  rejit::Regej re_, sol_;
  const size_t n_matchers_;
  std::vector<std::unique_ptr<Batch> > batches_;       // slot of batch seq: seq % (n_matchers_ + 1)
  Batch* cur_;                                         // the batch being planned
  std::vector<std::thread> matchers_;
  std::mutex mu_;
  std::condition_variable cv_;
  size_t submitted_ = 0, printed_ = 0;
  bool done_ = false;
};

Jrep* g_jrep = nullptr;   // nftw has no user pointer

int Visit(const char* path, const struct stat* st, int type, struct FTW*) {
  return type == FTW_F ? g_jrep->Add(path, *st) : 0;
}

void Usage(FILE* to, const char* self) {
  fprintf(to,
          "Usage: %s [OPTION...] regexp file...\n"
          "grep-like search powered by rejit (Extended Regular Expression syntax;\n"
          "patterns may span lines, e.g. \"a\\nb\").\n\n"
          "  -H, --with-filename           print the file name with output lines\n"
          "  -n, --line-number             print the line number with output lines\n"
          "  -r, --recursive               search directories, do not follow symbolic links\n"
          "  -R, --dereference-recursive   search directories, follow symbolic links\n"
          "  -c, --color_output            highlight matches in red\n"
          "  -A, --after-context[=N]       print N lines of context after every match\n"
          "  -B, --before-context[=N]      print N lines of context before every match\n"
          "  -C, --context[=N]             both\n"
          "  -j, --jobs[=N]                N threads stage (read) the files of a batch (0: none); matching is per batch\n"
          "  -k, --nopenfd[=N]             directories nftw() may hold open (default 1024)\n"
          "      --batch-bytes=N           bytes staged per matcher call (default 16 MiB; 0 = one call per file)\n"
          "      --gpus=N                  use N devices: batches go to them in rotation (this library only)\n"
          "      --shard                   with --gpus: cut every batch into N slabs instead, one per device\n",
          self);
}

unsigned Number(const char* s) {
  while (*s == ' ') ++s;
  return static_cast<unsigned>(atoi(s));
}

bool ParseArguments(int argc, char** argv, Options* o) {
  static const struct option kLong[] = {
      {"with-filename", no_argument, nullptr, 'H'},       {"line-number", no_argument, nullptr, 'n'},
      {"recursive", no_argument, nullptr, 'r'},           {"dereference-recursive", no_argument, nullptr, 'R'},
      {"color_output", no_argument, nullptr, 'c'},        {"jobs", optional_argument, nullptr, 'j'},
      {"nopenfd", optional_argument, nullptr, 'k'},       {"after-context", optional_argument, nullptr, 'A'},
      {"before-context", optional_argument, nullptr, 'B'}, {"context", optional_argument, nullptr, 'C'},
      {"batch-bytes", required_argument, nullptr, 1000},  {"gpus", required_argument, nullptr, 1001},
      {"shard", no_argument, nullptr, 1002},
      {"help", no_argument, nullptr, '?'},                {nullptr, 0, nullptr, 0}};
  int c;
  while ((c = getopt_long(argc, argv, "HnrRcj::k::A::B::C::", kLong, nullptr)) != -1) {
    switch (c) {
      case 'H': o->with_filename = true; break;
      case 'n': o->line_number = true; break;
      case 'r': o->recursive = 1; break;
      case 'R': o->recursive = 2; break;
      case 'c': o->colour = true; break;
      case 'j': o->jobs = optarg ? Number(optarg) : Options::DefaultJobs(); break;
      case 'k': if (optarg) o->nopenfd = Number(optarg); break;
      case 'A': if (optarg) o->after = Number(optarg); break;
      case 'B': if (optarg) o->before = Number(optarg); break;
      case 'C': if (optarg) o->before = o->after = Number(optarg); break;
      case 1000: o->batch_bytes = static_cast<size_t>(strtoull(optarg, nullptr, 10)); break;
      case 1001: o->gpus = std::max(1, atoi(optarg)); break;
      case 1002: o->shard = true; break;
      default: return false;
    }
  }
  if (argc - optind < 2) return false;
  o->pattern = argv[optind];
  for (int i = optind + 1; i < argc; ++i) o->paths.push_back(argv[i]);
  return true;
}

}  // namespace

int main(int argc, char** argv) {
  Options options;
  if (!ParseArguments(argc, argv, &options)) {
    Usage(stderr, argv[0]);
    return 64;
  }
  if (options.pattern[0] == 0) return 0;
#ifdef REJIT_B200
  const int devices = rejit_b200_device_count();
  if (devices > 0 && options.gpus > devices) options.gpus = devices;
#endif

  Jrep jrep(options);
  if (!jrep.Ready()) return 2;
  g_jrep = &jrep;

  int rc = 0;
  for (const char* path : options.paths) {
    struct stat st;
    rc = stat(path, &st);
    if (rc) {
      fprintf(stderr, "jrep: %s: %s\n", path, strerror(errno));
      continue;
    }
    if (S_ISDIR(st.st_mode)) {
      if (!options.recursive) {
        fprintf(stderr, "jrep: %s: Is a directory.\n", path);
        continue;
      }
      rc = nftw(path, Visit, static_cast<int>(options.nopenfd), options.recursive == 2 ? 0 : FTW_PHYS);
    } else if (S_ISREG(st.st_mode)) {
      rc = jrep.Add(path, st);
    } else {
      rc = 0;
    }
    if (rc != 0) break;            // an unreadable file ends the run; what was staged before it is still printed
  }
  jrep.Finish();
  return rc ? rc : jrep.error();
}
