// e2e_dropin — the regex-dna counting step exactly as a program written against the reference's rejit.h does it
// (/root/reference/sample/regexdna.cc:52-67): nine Regej objects, nine calls of the UNMODIFIED signature
//     size_t Regej::MatchAll(const char* text, size_t text_size, std::vector<Match>* matches)
// on a std::string (pageable memory).  Every call uploads the text, scans it and copies its match list back.
// Timed inside the program (the library's first-use initialisation is warmed up first); prints one JSON line.
//   usage: e2e_dropin <file with the sequence> <repetitions>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <memory>
#include <string>
#include <vector>

#include "rejit.h"

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s <sequence file> <repetitions>\n", argv[0]); return 2; }
  std::ifstream in(argv[1], std::ios::binary);
  std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  const int reps = atoi(argv[2]) > 0 ? atoi(argv[2]) : 3;
  static const char* const kVariants[] = {
      "agggtaaa|tttaccct",         "[cgt]gggtaaa|tttaccc[acg]", "a[act]ggtaaa|tttacc[agt]t",
      "ag[act]gtaaa|tttac[agt]ct", "agg[act]taaa|ttta[agt]cct", "aggg[acg]aaa|ttt[cgt]ccct",
      "agggt[cgt]aa|tt[acg]accct", "agggta[cgt]a|t[acg]taccct", "agggtaa[cgt]|[acg]ttaccct"};
  std::vector<std::unique_ptr<rejit::Regej> > res;
  for (const char* v : kVariants) res.emplace_back(new rejit::Regej(v));
  size_t total = 0;
  auto step = [&]() {
    total = 0;
    for (auto& re : res) {
      std::vector<rejit::Match> matches;
      total += re->MatchAll(text.data(), text.size(), &matches);
    }
  };
  step();
  step();                                               // warm-up: contexts, staging buffers, capacities
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) step();
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
  printf("{\"value\": %.3f, \"unit\": \"GB/s\", \"ms_per_step\": %.3f, \"matches\": %zu, \"h2d_bytes_per_step\": %zu, "
         "\"uploaded_gbs\": %.2f, \"api\": \"rejit::Regej::MatchAll(const char*, size_t, vector<Match>*) x 9, std::string (pageable)\"}\n",
         text.size() / s / 1e9, s * 1e3, total, 9 * text.size(), 9 * text.size() / s / 1e9);
  return 0;
}
