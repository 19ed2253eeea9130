// regex-dna (the Benchmarks Game task that BASELINE.json's configs[1] and [4] are
// cut from), written against include/rejit.h only: it compiles unchanged against
// the reference library or against librejit_b200.so.  Reads a FASTA file on
// stdin, strips headers and newlines, counts the nine variants, applies the
// eleven IUB substitutions and prints the three lengths.
//
// Same task as /root/reference/sample/regexdna.cc (which this file does not
// copy); the use of rejit::Regej::MatchAllCount / ReplaceAll is the point.
#include <cstdio>
#include <iostream>
#include <iterator>
#include <string>

#include "rejit.h"

int main() {
  std::string text((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
  const size_t read_length = text.size();

  rejit::ReplaceAll(">.*\n|\n", text, "");
  const size_t stripped_length = text.size();

  static const char* const kVariants[] = {
      "agggtaaa|tttaccct",         "[cgt]gggtaaa|tttaccc[acg]", "a[act]ggtaaa|tttacc[agt]t",
      "ag[act]gtaaa|tttac[agt]ct", "agg[act]taaa|ttta[agt]cct", "aggg[acg]aaa|ttt[cgt]ccct",
      "agggt[cgt]aa|tt[acg]accct", "agggta[cgt]a|t[acg]taccct", "agggtaa[cgt]|[acg]ttaccct"};
  for (const char* v : kVariants) printf("%s %zu\n", v, rejit::MatchAllCount(v, text));

  static const char* const kIub[][2] = {{"B", "(c|g|t)"}, {"D", "(a|g|t)"},   {"H", "(a|c|t)"}, {"K", "(g|t)"},
                                        {"M", "(a|c)"},   {"N", "(a|c|g|t)"}, {"R", "(a|g)"},   {"S", "(c|g)"},
                                        {"V", "(a|c|g)"}, {"W", "(a|t)"},     {"Y", "(c|t)"}};
  for (const auto& s : kIub) rejit::ReplaceAll(s[0], text, s[1]);

  printf("\n%zu\n%zu\n%zu\n", read_length, stripped_length, text.size());
  return 0;
}
