// regex-dna with the text kept on the device (BASELINE.json configs[4] on one GPU; SURVEY.md §8f ranks
// 1 and 2): the FASTA file is uploaded ONCE; the header / newline strip, the nine counts (fused into one
// scan of the sequence) and the eleven IUB substitutions all run on device-resident texts, and only the
// counts and lengths come back.  Same output as samples/regexdna.cc (and as the reference's
// sample/regexdna.cc:49-91, whose twelve ReplaceAll calls each rebuild the string on the host).
// Uses the rejit_b200 additions of include/rejit.h (rejit::Text, Regej::MatchAllCountSet, Text::ReplaceAllSet).
#include <cstdio>
#include <iostream>
#include <iterator>
#include <memory>
#include <string>
#include <vector>

#include "rejit.h"

#ifndef REJIT_B200
#error "this sample needs the device-resident texts of rejit_b200 (samples/regexdna.cc is the portable one)"
#endif

int main() {
  std::string file((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
  rejit::Text raw(file.data(), file.size());

  rejit::Regej strip(">.*\n|\n");
  std::unique_ptr<rejit::Text> sequence(raw.ReplaceAll(strip, ""));

  static const char* const kVariants[] = {
      "agggtaaa|tttaccct",         "[cgt]gggtaaa|tttaccc[acg]", "a[act]ggtaaa|tttacc[agt]t",
      "ag[act]gtaaa|tttac[agt]ct", "agg[act]taaa|ttta[agt]cct", "aggg[acg]aaa|ttt[cgt]ccct",
      "agggt[cgt]aa|tt[acg]accct", "agggta[cgt]a|t[acg]taccct", "agggtaa[cgt]|[acg]ttaccct"};
  std::vector<std::unique_ptr<rejit::Regej> > owned;
  std::vector<rejit::Regej*> variants;
  for (const char* v : kVariants) {
    owned.emplace_back(new rejit::Regej(v));
    variants.push_back(owned.back().get());
  }
  std::vector<size_t> counts;
  rejit::Regej::MatchAllCountSet(variants, *sequence, &counts);
  for (size_t i = 0; i < counts.size(); ++i) printf("%s %zu\n", kVariants[i], counts[i]);

  static const char* const kIub[][2] = {{"B", "(c|g|t)"}, {"D", "(a|g|t)"},   {"H", "(a|c|t)"}, {"K", "(g|t)"},
                                        {"M", "(a|c)"},   {"N", "(a|c|g|t)"}, {"R", "(a|g)"},   {"S", "(c|g)"},
                                        {"V", "(a|c|g)"}, {"W", "(a|t)"},     {"Y", "(c|t)"}};
  // the eleven substitutions, in the reference's order: every pattern is one byte and no replacement holds a later
  // pattern's byte, so Text::ReplaceAllSet applies them in one pass over the sequence (a byte -> string table)
  std::vector<std::unique_ptr<rejit::Regej> > codes;
  std::vector<rejit::Regej*> code_ptrs;
  std::vector<std::string> withs;
  for (const auto& s : kIub) {
    codes.emplace_back(new rejit::Regej(s[0]));
    code_ptrs.push_back(codes.back().get());
    withs.push_back(s[1]);
  }
  std::unique_ptr<rejit::Text> substituted(sequence->ReplaceAllSet(code_ptrs, withs));

  printf("\n%zu\n%zu\n%zu\n", file.size(), sequence->size(), substituted->size());
  return 0;
}
