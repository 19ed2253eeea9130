// bench_engine — the benchmark engine of the reference's harness for THIS library
// (SURVEY.md §8f rank 4): same command line and same table on stdout as
// /root/reference/tools/benchmarks/engines/rejit/engine.cc + ../bench_engine.cc
// (options :31-46, text :182-207, table :211-241, speed :244-249), so that
// tools/benchmarks/run.py can drive it as one more engine and plot it next to
// the others.  Written against rejit.h only; restated, not copied.
//
//   bench_engine <regexp> [--size=a,b,..] [--iterations=N] [--low_char=c] [--high_char=c]
//                [--file=path] [--run_worst_case=0|1] [--resident=0|1]
//
// Prints bytes/s per text size: "worse" (the pattern built again for every run),
// "amortised" (one build for all runs, build included) and "best" (build
// excluded).  The text is the reference's: low + rand() % (high - low) per byte
// with the C library's default seed, the last byte of the largest size left 0.
//
// Differences, both deliberate: the match vector is emptied between runs (the
// reference lets it grow), and --resident=1 (this library
// only) uploads each text once, outside the timed loops, and times
// Regej::MatchAll(const Text&): the scan without the PCIe copy.
#include <fcntl.h>
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include "rejit.h"

namespace {

struct Arguments {
  const char* regexp = nullptr;
  const char* file = nullptr;
  std::vector<size_t> sizes;
  unsigned iterations = 1000;
  char low = 'a', high = 'z';
  bool worst_case = true;
  bool resident = false;
};

struct Row {
  size_t text_size;
  double worse, amortised, best;
};

[[noreturn]] void Die(const char* message) {
  printf("ERROR: %s\nExiting.\n", message);
  exit(1);
}

bool OnOff(const char* s, bool* out) {
  if (!strcmp(s, "on") || !strcmp(s, "true") || !strcmp(s, "1")) return *out = true, true;
  if (!strcmp(s, "off") || !strcmp(s, "false") || !strcmp(s, "0")) return *out = false, true;
  return false;
}

void ParseSizes(const char* s, std::vector<size_t>* sizes) {
  while (*s) {
    char* end;
    unsigned long long v = strtoull(s, &end, 10);
    if (end == s) break;
    sizes->push_back(static_cast<size_t>(v));
    s = *end == ',' ? end + 1 : end;
  }
  if (sizes->empty() || *s) Die("Invalid sizes arguments.");
}

void Parse(int argc, char** argv, Arguments* a) {
  static const struct option kLong[] = {{"file", optional_argument, nullptr, 'f'},
                                        {"size", optional_argument, nullptr, 's'},
                                        {"iterations", optional_argument, nullptr, 'i'},
                                        {"low_char", optional_argument, nullptr, 'l'},
                                        {"high_char", optional_argument, nullptr, 'h'},
                                        {"run_worst_case", optional_argument, nullptr, 1000},
                                        {"resident", optional_argument, nullptr, 1001},
                                        {nullptr, 0, nullptr, 0}};
  int c;
  while ((c = getopt_long(argc, argv, "f::s::i::l::h::", kLong, nullptr)) != -1) {
    switch (c) {
      case 'f': a->file = optarg; break;
      case 's': if (optarg) ParseSizes(optarg, &a->sizes); break;
      case 'i':
        if (optarg && (a->iterations = static_cast<unsigned>(strtoul(optarg, nullptr, 10))) == 0)
          Die("The number of iterations to run must be greater than 0.");
        break;
      case 'l': if (optarg) a->low = optarg[0]; break;
      case 'h': if (optarg) a->high = optarg[0]; break;
      case 1000:
        if (!optarg || !OnOff(optarg, &a->worst_case))
          Die("Invalid value for option 'run_worst_case'. Expected one of (on|true|1|off|false|0).\n");
        break;
      case 1001:
        if (!optarg || !OnOff(optarg, &a->resident)) Die("Invalid value for option 'resident'.");
        break;
      default:
        fprintf(stderr, "Usage: %s [OPTION...] regexp\n", argv[0]);
        exit(64);
    }
  }
  if (argc - optind != 1) {
    fprintf(stderr, "Usage: %s [OPTION...] regexp\n", argv[0]);
    exit(64);
  }
  a->regexp = argv[optind];
  if (a->regexp[0] == 0) Die("Cannot test an empty regular expression.");
  if (a->sizes.empty()) a->sizes.push_back(65536);
  std::sort(a->sizes.begin(), a->sizes.end());
}

void PrepareText(const Arguments& a, std::string* text) {
  const size_t n = a.sizes.back();
  if (a.file) {                                       // the file (without its last byte), repeated to n bytes
    std::string body;
    FILE* f = fopen(a.file, "rb");
    if (!f) Die("Cannot open the source file.");
    char chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) body.append(chunk, got);
    fclose(f);
    if (body.size() < 2) Die("The source file is too small.");
    body.resize(body.size() - 1);
    while (text->size() < n) text->append(body, 0, std::min(body.size(), n - text->size()));
    return;
  }
  text->resize(n);
  if (a.high <= a.low) Die("high_char must be greater than low_char.");
  for (size_t i = 0; i + 1 < n; ++i) (*text)[i] = static_cast<char>(a.low + rand() % (a.high - a.low));
}

double Speed(const timeval& t0, const timeval& t1, size_t bytes, unsigned runs) {
  const double usec = static_cast<double>(t1.tv_usec - t0.tv_usec) + static_cast<double>(t1.tv_sec - t0.tv_sec) * 1e6;
  return static_cast<double>(bytes) / usec * 1e6 * static_cast<double>(runs);
}

void Print(const std::vector<Row>& rows, bool worst_case) {
  int width = static_cast<int>(strlen("text_size"));
  for (const Row& r : rows) width = std::max(width, snprintf(nullptr, 0, "%zu", r.text_size));
  printf("%*s", width, "text_size");
  if (worst_case) printf("%16s", "worse");
  printf("%16s%16s\n", "amortised", "best");
  for (const Row& r : rows) {
    printf("%*zu", width, r.text_size);
    if (worst_case) printf("%16g", r.worse);
    printf("%16g%16g\n", r.amortised, r.best);
  }
}

}  // namespace

int main(int argc, char** argv) {
  Arguments a;
  Parse(argc, argv, &a);

  rejit::Regej probe(a.regexp);
  if (probe.status()) {
    printf("%s\n", rejit::rejit_status_string);
    Die("Invalid regular expression.\n");
  }
#ifndef REJIT_B200
  if (a.resident) Die("--resident needs a library with device-resident texts.");
#endif

  std::string text;
  PrepareText(a, &text);

  std::vector<Row> rows;
  for (size_t size : a.sizes) {
    Row row = {size, 0, 0, 0};
    timeval t0, t1, t2;
    std::vector<rejit::Match> matches;
#ifdef REJIT_B200
    rejit::Text* on_device = a.resident ? new rejit::Text(text.c_str(), size) : nullptr;
#endif
    auto run = [&](rejit::Regej& re) {
      matches.clear();
#ifdef REJIT_B200
      if (on_device) {
        re.MatchAll(*on_device, &matches);
        return;
      }
#endif
      re.MatchAll(text.c_str(), size, &matches);
    };

    if (a.worst_case) {
      gettimeofday(&t0, nullptr);
      for (unsigned i = 0; i < a.iterations; ++i) {
        rejit::Regej re(a.regexp);
        run(re);
      }
      gettimeofday(&t1, nullptr);
      row.worse = Speed(t0, t1, size, a.iterations);
    }

    gettimeofday(&t0, nullptr);
    rejit::Regej re(a.regexp);
    re.Compile(rejit::kMatchAll);
    gettimeofday(&t1, nullptr);
    for (unsigned i = 0; i < a.iterations; ++i) run(re);
    gettimeofday(&t2, nullptr);
    row.amortised = Speed(t0, t2, size, a.iterations);
    row.best = Speed(t1, t2, size, a.iterations);
    rows.push_back(row);
#ifdef REJIT_B200
    delete on_device;
#endif
  }
  Print(rows, a.worst_case);
  return 0;
}
