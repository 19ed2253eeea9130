// tests/hostsim.cc — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Runs the product's host front end (parser, lowering, automaton tables) and the
// host+device shared logic of rejit_b200/csrc/cuda/device_program.h (NfaRun,
// ChainTake) on the CPU, emulating what each kernel of engine.cu does with the
// tables: the literal scan, the sub-stream DFA scan with its warm-up, the
// needle-window verification, the generic per-start run, and both resolve
// paths (sequential chain and restart-point segments).  It exists so that the
// CPU-only test tier (`pytest -m "not gpu"`, run where no GPU exists) can
// compare the tables and the selection logic with the oracle; the kernels
// themselves are only exercised by the GPU tier.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../rejit_b200/csrc/cuda/device_program.h"
#include "../rejit_b200/csrc/host/automaton.h"
#include "../rejit_b200/csrc/host/ir.h"

using namespace rejit_b200;

namespace {

struct Compiled {
  CompiledAutomaton ca;
  FlatTables ft;
  NfaTables nfa;
};

bool CompilePattern(const char* pattern, size_t plen, int parser_opt, Compiled* c, std::string* error) {
  ParseOptions opt;
  opt.parser_opt = parser_opt != 0;
  NodePtr root = ParseERE(pattern, plen, opt, error);
  if (!root || !WithinBudget(root.get(), error)) return false;
  LoweredRegexp lr = Lower(root.get());
  if (!BuildAutomaton(lr, &c->ca, error)) return false;
  FlattenTables(c->ca, &c->ft);
  NfaTables& t = c->nfa;
  t.n_pos = c->ft.n_pos;
  t.words = c->ft.words;
  t.has_anchor = c->ca.nfa.has_anchor ? 1 : 0;
  t.byte_mask = c->ft.byte_mask.data();
  t.first = c->ft.first.data();
  t.follow = c->ft.follow.data();
  t.accept = c->ft.accept.data();
  t.chain = c->ft.chain.data();
  t.start_ok = c->ft.start_ok.data();
  for (int i = 0; i < 4; ++i) t.accept_empty[i] = c->ft.accept_empty[i];
  return true;
}

typedef std::vector<std::pair<uint64_t, uint64_t>> Cands;

void ScanLiteral(const std::vector<uint8_t>& needle, const uint8_t* text, uint64_t n, Cands* out) {
  uint64_t m = needle.size();
  for (uint64_t p = 0; p + m <= n; ++p)
    if (memcmp(text + p, needle.data(), m) == 0) out->push_back({p, p + m});
}

// Emulates k_dfa_tma: every 272-byte lane stream is walked as two chains
// (144 + 128 bytes), each entered 16 bytes early from the start state, two
// bytes per step through the pair table (bit 31 = accept between the bytes).
void ScanDfaStreams(const Compiled& c, const uint8_t* text, uint64_t n, uint32_t, Cands* out) {
  const ScanDfa& d = c.ca.dfa;
  const uint32_t L = (uint32_t)d.match_len;
  const uint32_t C = (uint32_t)d.n_classes;
  const uint32_t acc = (uint32_t)(d.first_accept * d.n_classes);
  const uint64_t kStream = 272;
  for (uint64_t a0 = 0; a0 < n; a0 += kStream) {
    const uint64_t seg[3] = {a0, a0 + 144, a0 + kStream};
    for (int h = 0; h < 2; ++h) {
      uint64_t a = seg[h], b = std::min<uint64_t>(n, seg[h + 1]);
      if (a >= n) break;
      uint64_t p = a >= 16 ? a - 16 : 0;
      // the kernel steps in byte PAIRS from a 16-byte aligned start; emulate the
      // pair table and cross-check it against the single-byte table
      uint32_t state = 0;          // state id
      while (p < b) {
        uint32_t c1 = c.ft.dfa_class[text[p]];
        uint32_t c2 = (p + 1 < n) ? c.ft.dfa_class[text[p + 1]] : 0;
        uint32_t one = c.ft.dfa_next[state * C + c1];                 // pre-multiplied
        uint32_t ent = c.ft.dfa_pair[(state * C + c1) * C + c2];
        bool mid_acc = one >= acc;
        if (mid_acc != ((ent & 0x80000000u) != 0)) { out->push_back({~0ull, ~0ull}); return; }
        if (mid_acc) {
          uint64_t e = p + 1;
          if (e > a && e <= b && e >= L) out->push_back({e - L, e});
        }
        uint32_t two = c.ft.dfa_next[one + c2];
        if ((two / C) != (ent & 0x7FFFFFFFu) && p + 1 < n) { out->push_back({~0ull, ~0ull}); return; }
        if (two >= acc) {
          uint64_t e = p + 2;
          if (e > a && e <= b && e >= L) out->push_back({e - L, e});
        }
        state = two / C;
        p += 2;
      }
    }
  }
}

// Emulates k_window_verify: windows clipped against the previous hit's window.
void ScanWindow(const Compiled& c, const uint8_t* text, uint64_t n, Cands* out) {
  Cands hits;
  ScanLiteral(c.ca.literal, text, n, &hits);
  uint32_t lo = c.ca.window_lo, hi = c.ca.window_hi;
  for (size_t i = 0; i < hits.size(); ++i) {
    uint64_t h = hits[i].first;
    if (h < lo) continue;
    uint64_t s_max = h - lo, s_min = h >= hi ? h - hi : 0;
    if (i > 0) {
      uint64_t prev = hits[i - 1].first;
      if (prev >= lo && prev - lo + 1 > s_min) s_min = prev - lo + 1;
    }
    for (uint64_t s = s_min; s <= s_max && s < n; ++s) {
      int ctx = c.nfa.has_anchor ? ContextAt(text, n, s) : 0;
      if (!c.nfa.start_ok[ctx * 256 + text[s]]) continue;
      uint64_t e = NfaRunAny(c.nfa, text, n, s);
      if (e != kNoMatch) out->push_back({s, e});
    }
  }
}

void ScanGeneric(const Compiled& c, const uint8_t* text, uint64_t n, Cands* out) {
  for (uint64_t s = 0; s <= n; ++s) {
    int ctx = c.nfa.has_anchor ? ContextAt(text, n, s) : 0;
    bool ok = c.nfa.accept_empty[ctx] || (s < n && c.nfa.start_ok[ctx * 256 + text[s]]);
    if (!ok) continue;
    uint64_t e = NfaRunAny(c.nfa, text, n, s);
    if (e != kNoMatch) out->push_back({s, e});
  }
}

// k_resolve_small's chain
void ResolveSequential(Cands cands, ChainState st, Cands* out, ChainState* final_state) {
  std::sort(cands.begin(), cands.end());
  uint64_t prev = kNoMatch;
  for (auto& c : cands) {
    if (c.first == prev) continue;
    prev = c.first;
    if (ChainTake(&st, c.first, c.second)) out->push_back(c);
  }
  if (final_state) *final_state = st;
}

// the large path: exclusive max scan, restart points, per-segment chains
void ResolveSegments(Cands cands, ChainState carry, Cands* out) {
  std::sort(cands.begin(), cands.end());
  size_t m = cands.size();
  std::vector<uint64_t> reach(m);
  uint64_t run = carry.cur;
  for (size_t i = 0; i < m; ++i) { reach[i] = run; run = std::max(run, cands[i].second); }
  auto restart = [&](size_t i) {
    return reach[i] < cands[i].first || (reach[i] == cands[i].first && cands[i].second > cands[i].first);
  };
  std::vector<char> take(m, 0);
  for (size_t i = 0; i < m; ++i) {
    bool head = (i == 0) || (cands[i].first != cands[i - 1].first && restart(i));
    if (!head) continue;
    ChainState st = (i == 0) ? carry : ChainState{0, kNoMatch};
    uint64_t prev = kNoMatch;
    for (size_t j = i; j < m; ++j) {
      if (j > i && cands[j].first != cands[j - 1].first && restart(j)) break;
      if (cands[j].first == prev) { take[j] = 0; continue; }
      prev = cands[j].first;
      take[j] = ChainTake(&st, cands[j].first, cands[j].second);
    }
  }
  for (size_t i = 0; i < m; ++i) if (take[i]) out->push_back(cands[i]);
}

// clusters separated where no earlier candidate reaches the next begin (strictly),
// each replayed with the reference's thread labels (FaithfulSegment)
void ResolveFaithful(const Compiled& c, const uint8_t* text, uint64_t n, Cands cands, Cands* out) {
  std::sort(cands.begin(), cands.end());
  size_t m = cands.size();
  if (!m) return;
  std::vector<uint64_t> b(m), e(m), fin(m);
  std::vector<uint32_t> take(m, 0);
  for (size_t i = 0; i < m; ++i) { b[i] = cands[i].first; e[i] = cands[i].second; }
  int P = std::max(c.nfa.n_pos, 1), W = c.nfa.words;
  std::vector<uint64_t> lab(P), nlab(P);
  std::vector<uint32_t> act(W), nact(W), blocked(W);
  FaithfulScratch sc{lab.data(), nlab.data(), act.data(), nact.data(), blocked.data()};
  size_t i = 0;
  while (i < m) {
    uint64_t reach = e[i];
    size_t j = i + 1;
    while (j < m && !(reach < b[j])) { reach = std::max(reach, e[j]); ++j; }
    FaithfulSegment(c.nfa, text, n, b.data(), e.data(), i, j, sc, take.data(), fin.data());
    i = j;
  }
  for (size_t k = 0; k < m; ++k) if (take[k]) out->push_back({b[k], fin[k]});
}

}  // namespace

extern "C" {

// strategy: -1 = the one the compiler chose, 0..3 = force (3 always legal).
// Returns the match count; -1 parse/compile error; -2 the two resolve paths
// disagree; -3 forced strategy not applicable.
int64_t hostsim_match_all(const char* pattern, size_t plen, int parser_opt, const uint8_t* text, uint64_t n,
                          int strategy, uint64_t* out_pairs, uint64_t cap, char* describe, size_t dlen) {
  Compiled c;
  std::string error;
  if (!CompilePattern(pattern, plen, parser_opt, &c, &error)) {
    if (describe && dlen) snprintf(describe, dlen, "%s", error.c_str());
    return -1;
  }
  if (describe && dlen) snprintf(describe, dlen, "%s", c.ca.describe.c_str());
  int chosen = (int)c.ca.strategy;
  int use = strategy < 0 ? chosen : strategy;
  if (use != 3 && use != chosen) return -3;
  Cands cands;
  switch (use) {
    case 0: ScanLiteral(c.ca.literal, text, n, &cands); break;
    case 1: ScanDfaStreams(c, text, n, 256, &cands); break;
    case 2: ScanWindow(c, text, n, &cands); break;
    default: ScanGeneric(c, text, n, &cands);
  }
  Cands a, b;
  ChainState st{0, kNoMatch};
  if (c.ca.reentrant) {
    ResolveFaithful(c, text, n, cands, &a);
  } else {
    ResolveSequential(cands, st, &a, nullptr);
    ResolveSegments(cands, st, &b);
    if (a != b) return -2;
    // the label replay must agree with the chain wherever the chain is exact
    Cands f;
    ResolveFaithful(c, text, n, cands, &f);
    if (f != a) return -4;
  }
  for (size_t i = 0; i < a.size() && i < cap; ++i) {
    out_pairs[2 * i] = a[i].first;
    out_pairs[2 * i + 1] = a[i].second;
  }
  return (int64_t)a.size();
}

// Slab-sharded run with the carry protocol of MatchAllHostMultiGpu (engine.cu):
// every slab first resolved with carry = (slab start, none), then re-resolved
// when the chain arriving from the left differs.
int64_t hostsim_match_all_slabs(const char* pattern, size_t plen, const uint8_t* text, uint64_t n, int slabs,
                                uint64_t* out_pairs, uint64_t cap) {
  Compiled c;
  std::string error;
  if (!CompilePattern(pattern, plen, 1, &c, &error)) return -1;
  Cands all;
  ScanGeneric(c, text, n, &all);
  std::vector<Cands> part(slabs);
  std::vector<ChainState> carry_out(slabs);
  auto bound = [&](int i) { return (n / slabs) * (uint64_t)i; };
  auto run = [&](int i, ChainState in) {
    uint64_t lo = bound(i), hi = (i + 1 == slabs) ? n + 1 : bound(i + 1);
    Cands mine;
    for (auto& x : all) if (x.first >= lo && x.first < hi) mine.push_back(x);
    part[i].clear();
    ResolveSequential(mine, in, &part[i], &carry_out[i]);
  };
  for (int i = 0; i < slabs; ++i) run(i, ChainState{bound(i), kNoMatch});
  ChainState running = carry_out[0];
  for (int i = 1; i < slabs; ++i) {
    uint64_t lo = bound(i);
    if (running.cur > lo || running.tail == lo) {
      ChainState in = running;
      if (in.cur < lo) in.cur = lo;
      run(i, in);
    }
    running = carry_out[i];
  }
  uint64_t k = 0;
  for (auto& p : part)
    for (auto& x : p) {
      if (k < cap) { out_pairs[2 * k] = x.first; out_pairs[2 * k + 1] = x.second; }
      ++k;
    }
  return (int64_t)k;
}

// One rank of the one-process-per-GPU sharding: resolve the starts in [lo, hi)
// (the last slab also owns the offset n) with the given carry (global offsets).
int64_t hostsim_slab_run(const char* pattern, size_t plen, const uint8_t* text, uint64_t n, uint64_t lo,
                         uint64_t hi, int last, uint64_t carry_cur, uint64_t carry_tail, uint64_t* out_cur,
                         uint64_t* out_tail) {
  Compiled c;
  std::string error;
  if (!CompilePattern(pattern, plen, 1, &c, &error)) return -1;
  Cands all, mine, res;
  ScanGeneric(c, text, n, &all);
  uint64_t end = last ? n + 1 : hi;
  for (auto& x : all) if (x.first >= lo && x.first < end) mine.push_back(x);
  ChainState fin;
  ResolveSequential(mine, ChainState{carry_cur, carry_tail}, &res, &fin);
  *out_cur = fin.cur;
  *out_tail = fin.tail;
  return (int64_t)res.size();
}

// Fused pattern set: emulates k_set_tma (same chain geometry as k_dfa_tma, the
// union DFA's pair table) and resolves pattern `which`.  -1: parse error,
// -5: the set cannot be fused.
int64_t hostsim_set_match_all(const char* joined, size_t len, char sep, const uint8_t* text, uint64_t n, int which,
                              uint64_t* out_pairs, uint64_t cap) {
  std::vector<std::string> pats;
  std::string cur;
  for (size_t i = 0; i < len; ++i) { if (joined[i] == sep) { pats.push_back(cur); cur.clear(); } else cur.push_back(joined[i]); }
  pats.push_back(cur);
  std::vector<Compiled> comp(pats.size());
  std::vector<const CompiledAutomaton*> members;
  std::string error;
  for (size_t j = 0; j < pats.size(); ++j) {
    if (!CompilePattern(pats[j].data(), pats[j].size(), 1, &comp[j], &error)) return -1;
    members.push_back(&comp[j].ca);
  }
  SetDfa d;
  if (!BuildSetDfa(members, &d)) return -5;
  const uint32_t C = (uint32_t)d.n_classes;
  const uint32_t RW = (1u << d.row_shift) / 4;
  const uint32_t acc_row = (uint32_t)d.first_accept << d.row_shift;
  std::vector<Cands> per(pats.size());
  auto report = [&](uint32_t state, uint64_t e, uint64_t a, uint64_t b) {
    uint32_t m = d.accept_mask[state];
    for (int j = 0; m; ++j, m >>= 1)
      if ((m & 1u) && e > a && e <= b && e >= d.match_len[j]) per[j].push_back({e - d.match_len[j], e});
  };
  const uint64_t kStream = 272;
  for (uint64_t a0 = 0; a0 < n; a0 += kStream) {
    const uint64_t seg[3] = {a0, a0 + 144, a0 + kStream};
    for (int h = 0; h < 2; ++h) {
      uint64_t a = seg[h], b = std::min<uint64_t>(n, seg[h + 1]);
      if (a >= n) break;
      uint64_t p = a >= 16 ? a - 16 : 0;
      uint32_t row = 0;                       // state << row_shift
      while (p < b) {
        uint32_t state = row >> d.row_shift;            // a row: a state or the shadow of one
        if ((int)state >= d.n_rows) return -6;
        uint32_t c1 = d.byte_class[text[p]];
        uint32_t c2 = (p + 1 < n) ? d.byte_class[text[p + 1]] : 0;
        uint32_t ent = d.t2[state * RW + c1 * C + c2];
        uint32_t mid = d.t1[state * C + c1] / C;
        if (mid != d.next[(size_t)d.row_state[state] * C + c1]) return -6;
        if ((int)mid >= d.first_accept) report(mid, p + 1, a, b);
        row = ent;
        if (p + 1 < n) {
          uint32_t fin = d.next[(size_t)mid * C + c2];
          uint32_t r2 = row >> d.row_shift;
          if ((int)r2 >= d.n_rows || d.row_state[r2] != fin) return -6;
          // an accept in between must be visible in the row address
          if ((int)mid >= d.first_accept && row < acc_row) return -6;
          if ((int)r2 >= d.n_states && !((int)mid >= d.first_accept && (int)fin < d.first_accept)) return -6;
          if (row >= acc_row) report(r2, p + 2, a, b);      // shadow rows carry an empty mask
        }
        p += 2;
      }
    }
  }
  Cands res;
  ResolveSequential(per[which], ChainState{0, kNoMatch}, &res, nullptr);
  for (size_t i = 0; i < res.size() && i < cap; ++i) { out_pairs[2 * i] = res[i].first; out_pairs[2 * i + 1] = res[i].second; }
  return (int64_t)res.size();
}

// ReplaceAll: the kernel's placement functions (device_program.h) driven tile by
// tile, thread by thread, over the matches of the sequential resolve.
int64_t hostsim_replace_all(const char* pattern, size_t plen, const uint8_t* text, uint64_t n, const uint8_t* with,
                            uint32_t w, uint8_t* out, uint64_t cap, uint64_t* out_len) {
  Compiled c;
  std::string error;
  if (!CompilePattern(pattern, plen, 1, &c, &error)) return -1;
  Cands cands, res;
  ScanGeneric(c, text, n, &cands);
  if (c.ca.reentrant) ResolveFaithful(c, text, n, cands, &res);
  else ResolveSequential(cands, ChainState{0, kNoMatch}, &res, nullptr);
  const uint64_t m = res.size();
  std::vector<uint64_t> pairs(2 * m + 2), removed(m + 1, 0);
  for (uint64_t i = 0; i < m; ++i) { pairs[2 * i] = res[i].first; pairs[2 * i + 1] = res[i].second; }
  for (uint64_t i = 0; i < m; ++i) removed[i + 1] = removed[i] + (res[i].second - res[i].first);
  const uint64_t len = n - removed[m] + m * w;
  *out_len = len;
  if (len > cap) return -2;
  std::vector<uint8_t> canary(len + 64, 0xEE);
  const uint64_t n_tiles = n / kReplaceTile + 1;
  std::vector<uint16_t> s_b(kReplaceTile + 2), s_e(kReplaceTile + 2), s_r(kReplaceTile + 2);
  for (uint64_t tile = 0; tile < n_tiles; ++tile) {
    const uint64_t tile_lo = tile * kReplaceTile;
    const bool last = tile + 1 == n_tiles;
    const uint64_t tile_hi = last ? n : tile_lo + kReplaceTile;
    ReplaceTileHead h;
    h.m0 = ReplaceLowerBound(pairs.data(), m, tile_lo);
    h.m1 = last ? m : ReplaceLowerBound(pairs.data(), m, tile_hi);
    ReplaceHead(pairs.data(), removed.data(), m, tile_lo, &h);
    const uint32_t cnt = (uint32_t)(h.m1 - h.m0);
    if (cnt > kReplaceTile + 1) return -7;
    for (uint32_t i = 0; i < cnt; ++i) {
      const uint64_t b = pairs[2 * (h.m0 + i)], e = pairs[2 * (h.m0 + i) + 1];
      s_b[i] = (uint16_t)(b - tile_lo);
      s_e[i] = (uint16_t)((e < tile_hi ? e : tile_hi) - tile_lo);
      s_r[i] = (uint16_t)(removed[h.m0 + i] - h.r0);
    }
    for (uint32_t t = 0; t < 256; ++t)
      ReplacePlace(t, text, tile_lo, tile_hi, last, h, cnt, s_b.data(), s_e.data(), s_r.data(), with, w, canary.data());
  }
  for (uint64_t i = len; i < len + 64; ++i) if (canary[i] != 0xEE) return -8;     // wrote past the end
  memcpy(out, canary.data(), len);
  return (int64_t)m;
}

int hostsim_match_full(const char* pattern, size_t plen, const uint8_t* text, uint64_t n) {
  Compiled c;
  std::string error;
  if (!CompilePattern(pattern, plen, 1, &c, &error)) return -1;
  return NfaRunAny(c.nfa, text, n, 0, true) == n ? 1 : 0;
}

}  // extern "C"
