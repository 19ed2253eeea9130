// threads — concurrent MatchAll calls on ONE compiled Regej per pattern, the way the
// reference's jrep uses the library (/root/reference/sample/jrep.cc:461-493, 508-511:
// Compile up front, then worker threads call MatchAll on the shared object).  Every
// thread matches its own texts and one shared text, many times; the results must equal
// the ones a single thread computed first.  Prints "ok <calls>" or the first difference.
// Written against rejit.h only.
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "rejit.h"

namespace {

std::string MakeText(unsigned seed, size_t n) {
  std::string t(n, 'a');
  unsigned x = seed * 2654435761u + 1u;
  for (size_t i = 0; i < n; ++i) {
    x = x * 1664525u + 1013904223u;
    unsigned r = (x >> 24) & 31u;
    t[i] = r == 0 ? '\n' : r < 4 ? ';' : r < 6 ? '}' : static_cast<char>('a' + r % 8);
  }
  return t;
}

typedef std::vector<std::pair<size_t, size_t> > Offsets;

Offsets Run(rejit::Regej& re, const std::string& text) {
  std::vector<rejit::Match> m;
  re.MatchAll(text.data(), text.size(), &m);
  Offsets out;
  for (const rejit::Match& x : m) out.push_back(std::make_pair(size_t(x.begin - text.data()), size_t(x.end - text.data())));
  return out;
}

}  // namespace

int main(int argc, char** argv) {
  const int n_threads = argc > 1 ? atoi(argv[1]) : 8;
  const int rounds = argc > 2 ? atoi(argv[2]) : 12;
  const char* const kPatterns[] = {";\n}", "^", "ab|ba", "[a-c]h;", "}\n*[a-d]", "a.*h$"};
  const int n_patterns = sizeof kPatterns / sizeof kPatterns[0];
  std::vector<rejit::Regej*> res;
  for (const char* p : kPatterns) {
    res.push_back(new rejit::Regej(p));
    if (!res.back()->Compile(rejit::kMatchAll)) {
      printf("cannot compile %s\n", p);
      return 2;
    }
  }
  std::vector<std::string> texts;
  texts.push_back(MakeText(99, 300000));                                  // shared by all threads
  for (int t = 0; t < n_threads; ++t) texts.push_back(MakeText(t, 1000 + 37000 * (t % 5)));
  std::vector<std::vector<Offsets> > expected(texts.size(), std::vector<Offsets>(n_patterns));
  for (size_t i = 0; i < texts.size(); ++i)
    for (int p = 0; p < n_patterns; ++p) expected[i][p] = Run(*res[p], texts[i]);

  std::atomic<long> calls(0);
  std::atomic<int> bad(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t)
    pool.push_back(std::thread([&, t]() {
      for (int r = 0; r < rounds && !bad; ++r)
        for (int p = 0; p < n_patterns; ++p) {
          const int q = (p + t) % n_patterns;                             // threads meet on different patterns
          for (size_t i : {size_t(0), size_t(t + 1)}) {
            if (Run(*res[q], texts[i]) != expected[i][q] && !bad.exchange(1))
              printf("thread %d round %d: pattern %d on text %zu differs\n", t, r, q, i);
            ++calls;
          }
        }
    }));
  for (std::thread& th : pool) th.join();
  for (rejit::Regej* re : res) delete re;
  if (bad) return 1;
  printf("ok %ld\n", calls.load());
  return 0;
}
