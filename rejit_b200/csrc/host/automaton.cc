// See automaton.h.  Semantics restated from the reference's per-byte matching
// loop (the transitions of /root/reference/src/x64/codegen-x64.cc:653-677 and
// :735-933 and the control edges of :366-398, :680-732):
//   literal edge .... n exact bytes, must fit before the end of the text
//   '.' ............. any byte except \n and \r            (:851-873)
//   bracket ......... any listed byte, or lo <= c <= hi compared as SIGNED
//                     chars; inverted when negated; may match \n   (:876-933)
//   ^ / $ ........... pass when the previous / current byte is \n or \r or
//                     the offset is the start / end of the text   (:686-732)
//   epsilon ......... always
#include "automaton.h"

#include <algorithm>
#include <deque>
#include <map>
#include <sstream>

namespace rejit_b200 {
namespace {

constexpr int kMaxPositions = 4096;
constexpr int kMaxDfaStates = 2048;
constexpr uint32_t kMaxWindow = 2048;

void ClassSet(std::array<uint32_t, 8>* c, int b) { (*c)[b >> 5] |= 1u << (b & 31); }

std::array<uint32_t, 8> EdgeClass(const Edge& e, size_t byte_index) {
  std::array<uint32_t, 8> c{};
  if (e.kind == kEdgeLiteral) {
    ClassSet(&c, e.bytes[byte_index]);
  } else if (e.kind == kEdgeAnyChar) {
    for (int b = 0; b < 256; ++b) if (b != '\n' && b != '\r') ClassSet(&c, b);
  } else {  // kEdgeCharSet
    for (int b = 0; b < 256; ++b) {
      bool hit = false;
      for (uint8_t s : e.singles) hit |= (s == b);
      signed char sb = static_cast<signed char>(b);
      for (const ByteRange& r : e.ranges)
        hit |= (sb >= static_cast<signed char>(r.lo) && sb <= static_cast<signed char>(r.hi));
      if (hit != e.negated) ClassSet(&c, b);
    }
  }
  return c;
}

int SingleByte(const std::array<uint32_t, 8>& c) {   // -1 unless exactly one byte
  int found = -1;
  for (int b = 0; b < 256; ++b)
    if ((c[b >> 5] >> (b & 31)) & 1u) {
      if (found >= 0) return -1;
      found = b;
    }
  return found;
}

struct Graph {           // position graph under the most permissive context
  int n;
  std::vector<std::vector<int>> succ;
  std::vector<char> start, accept;
};

Graph MakeGraph(const PositionNfa& a) {
  Graph g;
  g.n = a.n_pos;
  g.succ.resize(g.n);
  g.start.assign(g.n, 0);
  g.accept.assign(g.n, 0);
  for (int k = 0; k < g.n; ++k) {
    g.start[k] = a.first[3].test(k);
    g.accept[k] = a.accept[3].test(k);
    for (int j = 0; j < g.n; ++j) if (a.follow[3][k].test(j)) g.succ[k].push_back(j);
  }
  return g;
}

// Reachability helpers (optionally with one position removed).
std::vector<char> ForwardReach(const Graph& g, int removed) {
  std::vector<char> seen(g.n, 0);
  std::deque<int> q;
  for (int k = 0; k < g.n; ++k) if (g.start[k] && k != removed) { seen[k] = 1; q.push_back(k); }
  while (!q.empty()) {
    int k = q.front(); q.pop_front();
    for (int j : g.succ[k]) if (j != removed && !seen[j]) { seen[j] = 1; q.push_back(j); }
  }
  return seen;
}

std::vector<char> BackwardReach(const Graph& g) {
  std::vector<std::vector<int>> pred(g.n);
  for (int k = 0; k < g.n; ++k) for (int j : g.succ[k]) pred[j].push_back(k);
  std::vector<char> seen(g.n, 0);
  std::deque<int> q;
  for (int k = 0; k < g.n; ++k) if (g.accept[k]) { seen[k] = 1; q.push_back(k); }
  while (!q.empty()) {
    int k = q.front(); q.pop_front();
    for (int j : pred[k]) if (!seen[j]) { seen[j] = 1; q.push_back(j); }
  }
  return seen;
}

// Longest distance (in positions) from the start set to each node of `live`
// nodes restricted to `mask`; returns false when a cycle is found.
bool LongestFromStart(const Graph& g, const std::vector<char>& mask,
                      std::vector<uint64_t>* longest) {
  std::vector<int> indeg(g.n, 0);
  for (int k = 0; k < g.n; ++k) if (mask[k]) for (int j : g.succ[k]) if (mask[j]) ++indeg[j];
  std::vector<int> order;
  std::deque<int> q;
  for (int k = 0; k < g.n; ++k) if (mask[k] && indeg[k] == 0) q.push_back(k);
  while (!q.empty()) {
    int k = q.front(); q.pop_front();
    order.push_back(k);
    for (int j : g.succ[k]) if (mask[j] && --indeg[j] == 0) q.push_back(j);
  }
  int total = 0;
  for (int k = 0; k < g.n; ++k) total += mask[k] ? 1 : 0;
  if (static_cast<int>(order.size()) != total) return false;      // cycle
  longest->assign(g.n, 0);
  std::vector<char> has(g.n, 0);
  for (int k : order) {
    if (g.start[k]) { has[k] = 1; (*longest)[k] = std::max<uint64_t>((*longest)[k], 1); }
    if (!has[k]) continue;
    for (int j : g.succ[k]) if (mask[j]) {
      has[j] = 1;
      (*longest)[j] = std::max((*longest)[j], (*longest)[k] + 1);
    }
  }
  return true;
}

std::vector<uint64_t> ShortestFromStart(const Graph& g) {
  std::vector<uint64_t> d(g.n, kInfLen);
  std::deque<int> q;
  for (int k = 0; k < g.n; ++k) if (g.start[k]) { d[k] = 1; q.push_back(k); }
  while (!q.empty()) {
    int k = q.front(); q.pop_front();
    for (int j : g.succ[k]) if (d[j] == kInfLen) { d[j] = d[k] + 1; q.push_back(j); }
  }
  return d;
}

void Lengths(PositionNfa* a, const Graph& g) {
  std::vector<char> fwd = ForwardReach(g, -1), bwd = BackwardReach(g);
  std::vector<char> useful(g.n, 0);
  bool any_useful = false;
  for (int k = 0; k < g.n; ++k) { useful[k] = fwd[k] && bwd[k]; any_useful |= useful[k]; }
  std::vector<uint64_t> sh = ShortestFromStart(g);
  uint64_t mn = a->accept_empty[3] ? 0 : kInfLen;
  for (int k = 0; k < g.n; ++k) if (g.accept[k] && sh[k] != kInfLen) mn = std::min(mn, sh[k]);
  uint64_t mx = 0;
  if (any_useful) {
    std::vector<uint64_t> lg;
    if (!LongestFromStart(g, useful, &lg)) mx = kInfLen;
    else for (int k = 0; k < g.n; ++k) if (useful[k] && g.accept[k]) mx = std::max(mx, lg[k]);
  }
  a->min_len = (mn == kInfLen) ? 0 : mn;
  a->max_len = mx;
}

bool BuildDfa(const PositionNfa& a, ScanDfa* d) {
  // byte classes: bytes with identical acceptance columns are equivalent
  std::map<std::vector<uint32_t>, int> col_to_class;
  for (int b = 0; b < 256; ++b) {
    const std::vector<uint32_t>& col = a.byte_mask[b].w;
    auto it = col_to_class.find(col);
    if (it == col_to_class.end()) it = col_to_class.emplace(col, static_cast<int>(col_to_class.size())).first;
    d->byte_class[b] = static_cast<uint8_t>(it->second);
  }
  d->n_classes = static_cast<int>(col_to_class.size());
  std::vector<int> rep(d->n_classes, -1);
  for (int b = 0; b < 256; ++b) if (rep[d->byte_class[b]] < 0) rep[d->byte_class[b]] = b;

  std::map<BitSet, int> ids;
  std::vector<BitSet> sets;
  std::vector<std::vector<int>> trans;
  BitSet empty(a.n_pos);
  ids[empty] = 0;
  sets.push_back(empty);
  for (size_t s = 0; s < sets.size(); ++s) {
    trans.emplace_back(d->n_classes, 0);
    BitSet reach = a.first[0];
    for (int k = 0; k < a.n_pos; ++k) if (sets[s].test(k)) reach.or_with(a.follow[0][k]);
    for (int c = 0; c < d->n_classes; ++c) {
      BitSet nx = reach;
      nx.and_with(a.byte_mask[rep[c]]);
      auto it = ids.find(nx);
      if (it == ids.end()) {
        if (static_cast<int>(sets.size()) >= kMaxDfaStates) return false;
        it = ids.emplace(nx, static_cast<int>(sets.size())).first;
        sets.push_back(nx);
      }
      trans[s][c] = it->second;
    }
  }
  // renumber: non-accepting first (state 0 = start stays 0), accepting last
  int n = static_cast<int>(sets.size());
  std::vector<int> renum(n, -1);
  int next_id = 0;
  auto accepting = [&](int s) { BitSet t = sets[s]; t.and_with(a.accept[0]); return t.any(); };
  for (int s = 0; s < n; ++s) if (!accepting(s)) renum[s] = next_id++;
  d->first_accept = next_id;
  for (int s = 0; s < n; ++s) if (accepting(s)) renum[s] = next_id++;
  if (renum[0] != 0) return false;                     // start state accepting: not a scan DFA
  d->n_states = n;
  d->next.assign(static_cast<size_t>(n) * d->n_classes, 0);
  for (int s = 0; s < n; ++s)
    for (int c = 0; c < d->n_classes; ++c)
      d->next[static_cast<size_t>(renum[s]) * d->n_classes + c] = static_cast<uint16_t>(renum[trans[s][c]]);
  d->match_len = a.min_len;
  // rows are stored pre-multiplied by n_classes in 16-bit entries
  if (static_cast<size_t>(n) * d->n_classes > 65535) return false;
  return true;
}

// Follows the unique path first -> ... -> accept when the automaton is a plain
// byte string; returns false otherwise.
bool WholeLiteral(const PositionNfa& a, std::vector<uint8_t>* out) {
  if (a.has_anchor || a.accept_empty[3] || a.n_pos == 0) return false;
  int cur = -1, count = 0;
  for (int k = 0; k < a.n_pos; ++k) if (a.first[3].test(k)) { cur = k; ++count; }
  if (count != 1) return false;
  std::vector<char> used(a.n_pos, 0);
  out->clear();
  for (;;) {
    if (used[cur]) return false;
    used[cur] = 1;
    int b = SingleByte(a.cls[cur]);
    if (b < 0) return false;
    out->push_back(static_cast<uint8_t>(b));
    int nxt = -1, n = 0;
    for (int j = 0; j < a.n_pos; ++j) if (a.follow[3][cur].test(j)) { nxt = j; ++n; }
    bool acc = a.accept[3].test(cur);
    if (acc && n == 0) break;
    if (acc || n != 1) return false;
    cur = nxt;
  }
  if (static_cast<int>(out->size()) != a.n_pos) return false;
  return true;
}

// Finds a byte string every match must contain, together with the range of
// distances from the match start to the string's first byte.
bool RequiredLiteral(const PositionNfa& a, const Graph& g, std::vector<uint8_t>* lit,
                     uint32_t* lo, uint32_t* hi) {
  if (a.accept_empty[3] || a.n_pos == 0 || a.n_pos > 1024) return false;
  std::vector<char> required(g.n, 0);
  for (int k = 0; k < g.n; ++k) {
    std::vector<char> seen = ForwardReach(g, k);
    bool reaches = false;
    for (int j = 0; j < g.n; ++j) if (seen[j] && g.accept[j]) reaches = true;
    required[k] = !reaches;
  }
  std::vector<std::vector<int>> pred(g.n);
  for (int k = 0; k < g.n; ++k) for (int j : g.succ[k]) pred[j].push_back(k);
  // maximal runs r1 -> r2 -> ... of required single-byte positions where each
  // link is the only way out of r_i and the only way into r_{i+1}
  std::vector<int> best;
  for (int k = 0; k < g.n; ++k) {
    if (!required[k] || SingleByte(a.cls[k]) < 0) continue;
    bool is_head = true;
    if (pred[k].size() == 1 && !g.start[k]) {
      int p = pred[k][0];
      if (required[p] && SingleByte(a.cls[p]) >= 0 && g.succ[p].size() == 1 && !g.accept[p]) is_head = false;
    }
    if (!is_head) continue;
    std::vector<int> run{k};
    int cur = k;
    while (g.succ[cur].size() == 1 && !g.accept[cur]) {
      int nx = g.succ[cur][0];
      if (!required[nx] || SingleByte(a.cls[nx]) < 0 || pred[nx].size() != 1 || g.start[nx]) break;
      if (std::find(run.begin(), run.end(), nx) != run.end()) break;
      run.push_back(nx);
      cur = nx;
    }
    if (run.size() > best.size()) best = run;
  }
  if (best.size() < 3) return false;
  int head = best[0];
  // distances from the match start to `head`: restrict to nodes that reach head
  // without passing through it
  std::vector<char> to_head(g.n, 0);
  {
    std::deque<int> q;
    to_head[head] = 1;
    q.push_back(head);
    while (!q.empty()) {
      int k = q.front(); q.pop_front();
      for (int j : pred[k]) if (!to_head[j]) { to_head[j] = 1; q.push_back(j); }
    }
  }
  std::vector<char> fwd = ForwardReach(g, -1);
  std::vector<char> mask(g.n, 0);
  for (int k = 0; k < g.n; ++k) mask[k] = fwd[k] && to_head[k];
  // a cycle through head itself (head reachable from head) makes the prefix unbounded
  std::vector<uint64_t> lg;
  if (!LongestFromStart(g, mask, &lg)) return false;
  std::vector<uint64_t> sh = ShortestFromStart(g);
  if (sh[head] == kInfLen || lg[head] == 0) return false;
  uint64_t dmin = sh[head] - 1, dmax = lg[head] - 1;     // bytes before the needle
  if (dmax - dmin + 1 > kMaxWindow || dmax > 0xFFFFFFF0ull) return false;
  lit->clear();
  for (int k : best) lit->push_back(static_cast<uint8_t>(SingleByte(a.cls[k])));
  *lo = static_cast<uint32_t>(dmin);
  *hi = static_cast<uint32_t>(dmax);
  return true;
}

}  // namespace

bool BuildAutomaton(const LoweredRegexp& lr, CompiledAutomaton* out, std::string* error) {
  PositionNfa& a = out->nfa;
  // ---- positions and extended states ---------------------------------
  int n_ext = lr.n_states;
  std::vector<int> src, dst;
  {
    // the cap is checked on the edge sizes first: nothing is allocated for a pattern that is too large
    uint64_t want = 0;
    for (const Edge& e : lr.matching) want += (e.kind == kEdgeLiteral) ? e.bytes.size() : 1;
    if (want > (uint64_t)kMaxPositions) {
      if (error) *error = "regular expression too large for the sm_100a engine (more than 4096 byte positions)";
      return false;
    }
  }
  for (const Edge& e : lr.matching) {
    size_t steps = (e.kind == kEdgeLiteral) ? e.bytes.size() : 1;
    int from = e.entry;
    for (size_t i = 0; i < steps; ++i) {
      int to = (i + 1 == steps) ? e.exit : n_ext++;
      src.push_back(from);
      dst.push_back(to);
      a.cls.push_back(EdgeClass(e, i));
      from = to;
    }
  }
  a.n_pos = static_cast<int>(src.size());
  if (a.n_pos > kMaxPositions) {
    if (error) *error = "regular expression too large for the sm_100a engine (more than 4096 byte positions)";
    return false;
  }
  a.words = std::max(1, (a.n_pos + 31) / 32);
  const int nbits = a.words * 32;
  for (const Edge& e : lr.control) a.has_anchor |= (e.kind != kEdgeEpsilon);

  a.byte_mask.assign(256, BitSet(nbits));
  for (int k = 0; k < a.n_pos; ++k)
    for (int b = 0; b < 256; ++b)
      if ((a.cls[k][b >> 5] >> (b & 31)) & 1u) a.byte_mask[b].set(k);

  std::vector<std::vector<int>> pos_from(n_ext);
  for (int k = 0; k < a.n_pos; ++k) pos_from[src[k]].push_back(k);

  for (int ctx = 0; ctx < kCtxCount; ++ctx) {
    const bool sol = ctx & 1, eol = ctx & 2;
    // closure over enabled control edges, per real state
    std::vector<std::vector<int>> ctl(lr.n_states);
    for (const Edge& e : lr.control) {
      bool on = e.kind == kEdgeEpsilon || (e.kind == kEdgeLineStart && sol) || (e.kind == kEdgeLineEnd && eol);
      if (on) ctl[e.entry].push_back(e.exit);
    }
    auto closure = [&](int s) {
      std::vector<int> res;
      if (s >= lr.n_states) { res.push_back(s); return res; }
      std::vector<char> seen(lr.n_states, 0);
      std::deque<int> q{ s };
      seen[s] = 1;
      while (!q.empty()) {
        int u = q.front(); q.pop_front();
        res.push_back(u);
        for (int v : ctl[u]) if (!seen[v]) { seen[v] = 1; q.push_back(v); }
      }
      return res;
    };
    auto expand = [&](int s, BitSet* positions, bool* reaches_exit) {
      for (int u : closure(s)) {
        if (u == lr.exit_state) *reaches_exit = true;
        for (int k : pos_from[u]) positions->set(k);
      }
    };
    a.first[ctx] = BitSet(nbits);
    a.accept[ctx] = BitSet(nbits);
    a.follow[ctx].assign(a.n_pos, BitSet(nbits));
    bool e0 = false;
    expand(lr.entry_state, &a.first[ctx], &e0);
    a.accept_empty[ctx] = e0;
    for (int k = 0; k < a.n_pos; ++k) {
      bool acc = false;
      expand(dst[k], &a.follow[ctx][k], &acc);
      if (acc) a.accept[ctx].set(k);
    }
  }
  a.chain = BitSet(nbits);
  for (int k = 0; k + 1 < a.n_pos; ++k) {
    bool only_next = true;
    for (int ctx = 0; ctx < kCtxCount && only_next; ++ctx) {
      BitSet want(nbits);
      want.set(k + 1);
      only_next = (a.follow[ctx][k] == want);
    }
    if (only_next) a.chain.set(k);
  }
  Graph g = MakeGraph(a);
  Lengths(&a, g);

  for (int ctx = 0; ctx < kCtxCount; ++ctx)
    for (int b = 0; b < 256; ++b) {
      BitSet t = a.first[ctx];
      t.and_with(a.byte_mask[b]);
      out->start_ok[ctx][b] = t.any() ? 1 : 0;
    }

  out->reentrant = false;
  for (int ctx = 0; ctx < kCtxCount; ++ctx)
    for (int k = 0; k < a.n_pos; ++k) {
      BitSet t = a.follow[ctx][k];
      t.and_with(a.first[ctx]);
      if (t.any()) out->reentrant = true;
    }

  // ---- strategy ---------------------------------------------------------
  std::ostringstream ds;
  out->strategy = ScanStrategy::Generic;
  uint32_t lo = 0, hi = 0;
  std::vector<uint8_t> lit;
  if (WholeLiteral(a, &lit)) {
    out->strategy = ScanStrategy::Literal;
    out->literal = lit;
    ds << "literal scan, " << lit.size() << " bytes";
  } else if (RequiredLiteral(a, g, &lit, &lo, &hi) && lit.size() >= 4) {
    out->strategy = ScanStrategy::LiteralWindow;
    out->literal = lit;
    out->window_lo = lo;
    out->window_hi = hi;
    ds << "required literal (" << lit.size() << " bytes) + verify starts in [hit-" << hi << ", hit-" << lo << "]";
  } else if (!a.has_anchor && a.max_len != kInfLen && a.min_len == a.max_len && a.min_len >= 1 &&
             a.min_len <= 4096 && BuildDfa(a, &out->dfa)) {
    out->strategy = ScanStrategy::DfaFixed;
    ds << "fixed-length DFA scan, " << out->dfa.n_states << " states x " << out->dfa.n_classes
       << " classes, match length " << a.min_len;
  } else {
    ds << "start filter + per-start NFA";
  }
  ds << "; " << a.n_pos << " positions, len [" << a.min_len << ","
     << (a.max_len == kInfLen ? std::string("inf") : std::to_string(a.max_len)) << "]"
     << (a.has_anchor ? ", anchors" : "") << (out->reentrant ? ", reentrant start" : "");
  out->describe = ds.str();
  return true;
}

// See SetDfa::Kmer.  Replaces nothing in the reference: it is the set scan's
// table for texts over a tiny alphabet (regex-dna, sample/regexdna.cc:52-62).
static void BuildKmerIndex(SetDfa* out) {
  SetDfa::Kmer& km = out->kmer;
  km = SetDfa::Kmer();
  if (out->max_len > 8) return;
  const int C = out->n_classes;
  // live bytes: some state moves somewhere else than where a byte no member knows leads
  std::vector<int> live;
  std::vector<char> class_live(C, 0);
  for (int c = 0; c < C; ++c)
    for (int s = 0; s < out->n_states && !class_live[c]; ++s)
      if (out->next[static_cast<size_t>(s) * C + c] != 0) class_live[c] = 1;
  int dead_byte = -1;
  for (int b = 0; b < 256; ++b) {
    if (class_live[out->byte_class[b]]) live.push_back(b); else if (dead_byte < 0) dead_byte = b;
  }
  if (live.empty() || live.size() > 4 || dead_byte < 0) return;
  int shift = -1;
  for (int s = 0; s <= 6 && shift < 0; ++s) {
    uint32_t seen = 0;
    bool distinct = true;
    for (int b : live) { uint32_t c = (b >> s) & 3; if (seen >> c & 1) distinct = false; seen |= 1u << c; }
    if (distinct) shift = s;
  }
  if (shift < 0) return;
  km.shift = static_cast<uint32_t>(shift);
  int canon[4] = {-1, -1, -1, -1};
  for (int b : live) canon[(b >> shift) & 3] = b;
  for (int c = 0; c < 4; ++c)
    if (canon[c] >= 0) { km.canon |= static_cast<uint32_t>(canon[c]) << (8 * c); km.canon_ok |= 1u << c; }
  for (int v = 0; v <= 8; ++v)
    for (int j = 0; j < out->n_patterns; ++j) if (out->match_len[j] <= static_cast<uint32_t>(v)) km.len_le[v] |= 1u << j;
  km.mask16.assign(65536, 0);
  for (uint32_t x = 0; x < 65536; ++x) {
    int st = 0;
    for (int i = 0; i < 8; ++i) {
      int code = (x >> (2 * i)) & 3;
      int b = canon[code] >= 0 ? canon[code] : dead_byte;
      st = out->next[static_cast<size_t>(st) * C + out->byte_class[b]];
    }
    km.mask16[x] = out->accept_mask[st];
  }
  const int idx_bits = 2 * (7 + kKmerR), word_bits = idx_bits - 5;
  km.bitmap.assign(static_cast<size_t>(1) << word_bits, 0);
  for (uint32_t x = 0; x < (1u << idx_bits); ++x) {
    uint32_t any = 0;
    for (int k = 0; k < kKmerR; ++k) any |= km.mask16[(x >> (2 * k)) & 0xFFFF];
    if (any) km.bitmap[x & ((1u << word_bits) - 1)] |= 1u << (31 - (x >> word_bits));
  }
  km.ok = true;
}

bool BuildSetDfa(const std::vector<const CompiledAutomaton*>& members, SetDfa* out) {
  const int k = static_cast<int>(members.size());
  if (k < 2 || k > 32) return false;
  int total = 0;
  std::vector<int> base(k);
  for (int j = 0; j < k; ++j) {
    const CompiledAutomaton* m = members[j];
    // any anchor-free pattern all of whose matches have the same length (1..17)
    if (m->nfa.has_anchor || m->reentrant || m->nfa.n_pos == 0 || m->nfa.max_len == kInfLen ||
        m->nfa.min_len != m->nfa.max_len || m->nfa.min_len < 1 || m->nfa.min_len > 17) return false;
    base[j] = total;
    total += m->nfa.n_pos;
  }
  if (total > 2048) return false;
  const int nbits = ((total + 31) / 32) * 32;
  auto lift = [&](const BitSet& src, int j) {
    BitSet r(nbits);
    for (int q = 0; q < members[j]->nfa.n_pos; ++q) if (src.test(q)) r.set(base[j] + q);
    return r;
  };
  BitSet first(nbits);
  std::vector<BitSet> follow(total, BitSet(nbits)), byte_mask(256, BitSet(nbits)), accept(k, BitSet(nbits));
  for (int j = 0; j < k; ++j) {
    const PositionNfa& a = members[j]->nfa;
    first.or_with(lift(a.first[0], j));
    accept[j] = lift(a.accept[0], j);
    for (int q = 0; q < a.n_pos; ++q) follow[base[j] + q] = lift(a.follow[0][q], j);
    for (int b = 0; b < 256; ++b) byte_mask[b].or_with(lift(a.byte_mask[b], j));
  }
  std::map<std::vector<uint32_t>, int> col_to_class;
  for (int b = 0; b < 256; ++b) {
    auto it = col_to_class.find(byte_mask[b].w);
    if (it == col_to_class.end()) it = col_to_class.emplace(byte_mask[b].w, static_cast<int>(col_to_class.size())).first;
    out->byte_class[b] = static_cast<uint8_t>(it->second);
  }
  const int C = static_cast<int>(col_to_class.size());
  std::vector<int> rep(C, -1);
  for (int b = 0; b < 256; ++b) if (rep[out->byte_class[b]] < 0) rep[out->byte_class[b]] = b;
  std::map<BitSet, int> ids;
  std::vector<BitSet> sets;
  std::vector<std::vector<int>> trans;
  BitSet empty(nbits);
  ids[empty] = 0;
  sets.push_back(empty);
  const size_t kMaxStates = 4096;
  for (size_t s = 0; s < sets.size(); ++s) {
    trans.emplace_back(C, 0);
    BitSet reach = first;
    for (int q = 0; q < total; ++q) if (sets[s].test(q)) reach.or_with(follow[q]);
    for (int c = 0; c < C; ++c) {
      BitSet nx = reach;
      nx.and_with(byte_mask[rep[c]]);
      auto it = ids.find(nx);
      if (it == ids.end()) {
        if (sets.size() >= kMaxStates) return false;
        it = ids.emplace(nx, static_cast<int>(sets.size())).first;
        sets.push_back(nx);
      }
      trans[s][c] = it->second;
    }
  }
  const int S = static_cast<int>(sets.size());
  // budget: 16-bit premultiplied rows and a two-byte table of at most 96 KB
  if (static_cast<size_t>(S) * C > 65535) return false;
  auto mask_of = [&](int s) {
    uint32_t m = 0;
    for (int j = 0; j < k; ++j) { BitSet t = sets[s]; t.and_with(accept[j]); if (t.any()) m |= 1u << j; }
    return m;
  };
  std::vector<int> renum(S, -1);
  int next_id = 0;
  for (int s = 0; s < S; ++s) if (!mask_of(s)) renum[s] = next_id++;
  out->first_accept = next_id;
  for (int s = 0; s < S; ++s) if (mask_of(s)) renum[s] = next_id++;
  if (renum[0] != 0) return false;
  out->n_patterns = k;
  out->n_states = S;
  out->n_classes = C;
  out->next.assign(static_cast<size_t>(S) * C, 0);
  out->accept_mask.assign(S, 0);
  for (int s = 0; s < S; ++s) {
    out->accept_mask[renum[s]] = mask_of(s);
    for (int c = 0; c < C; ++c) out->next[static_cast<size_t>(renum[s]) * C + c] = static_cast<uint16_t>(renum[trans[s][c]]);
  }
  out->match_len.clear();
  out->max_len = 0;
  for (int j = 0; j < k; ++j) {
    out->match_len.push_back(static_cast<uint32_t>(members[j]->nfa.min_len));
    out->max_len = std::max<uint32_t>(out->max_len, out->match_len.back());
  }
  if (out->max_len > 17) return false;          // the kernel warms every chain up on 16 bytes
  // the kernel's class tables hold class*C*4 and class*4 in one byte each
  if ((C - 1) * C * 4 > 255) return false;
  // shadow rows (see automaton.h)
  std::vector<int> shadow_of_state(S, -1);
  out->row_state.resize(S);
  for (int s = 0; s < S; ++s) out->row_state[s] = static_cast<uint16_t>(s);
  for (int s = 0; s < S; ++s)
    for (int c1 = 0; c1 < C; ++c1) {
      int mid = out->next[static_cast<size_t>(s) * C + c1];
      if (mid < out->first_accept) continue;
      for (int c2 = 0; c2 < C; ++c2) {
        int fin = out->next[static_cast<size_t>(mid) * C + c2];
        if (fin < out->first_accept && shadow_of_state[fin] < 0) {
          shadow_of_state[fin] = static_cast<int>(out->row_state.size());
          out->row_state.push_back(static_cast<uint16_t>(fin));
        }
      }
    }
  const int R = static_cast<int>(out->row_state.size());
  out->n_rows = R;
  out->accept_mask.resize(R, 0);
  out->t1.assign(static_cast<size_t>(R) * C, 0);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c)
      out->t1[static_cast<size_t>(r) * C + c] = static_cast<uint16_t>(out->next[static_cast<size_t>(out->row_state[r]) * C + c] * C);
  // pair-table rows are padded to a power of two so that the row of an entry is a shift
  int row_words = 1;
  while (row_words < C * C) row_words <<= 1;
  out->row_shift = 2;
  while ((1 << out->row_shift) < row_words * 4) ++out->row_shift;
  if (static_cast<size_t>(R) * row_words * 4 > 128 * 1024) return false;
  out->t2.assign(static_cast<size_t>(R) * row_words, 0);
  for (int r = 0; r < R; ++r) {
    const int s = out->row_state[r];
    for (int c1 = 0; c1 < C; ++c1) {
      int mid = out->next[static_cast<size_t>(s) * C + c1];
      for (int c2 = 0; c2 < C; ++c2) {
        int fin = out->next[static_cast<size_t>(mid) * C + c2];
        int row = (mid >= out->first_accept && fin < out->first_accept) ? shadow_of_state[fin] : fin;
        out->t2[static_cast<size_t>(r) * row_words + c1 * C + c2] = static_cast<uint32_t>(row) << out->row_shift;
      }
    }
  }
  BuildKmerIndex(out);
  return true;
}

void FlattenTables(const CompiledAutomaton& ca, FlatTables* out) {
  const PositionNfa& a = ca.nfa;
  const int W = a.words;
  out->n_pos = a.n_pos;
  out->words = W;
  out->byte_mask.assign(256 * static_cast<size_t>(W), 0);
  out->first.assign(4 * static_cast<size_t>(W), 0);
  out->accept.assign(4 * static_cast<size_t>(W), 0);
  out->chain.assign(W, 0);
  out->follow.assign(static_cast<size_t>(4) * std::max(a.n_pos, 1) * W, 0);
  out->start_ok.assign(4 * 256, 0);
  for (int b = 0; b < 256; ++b)
    for (int i = 0; i < W; ++i) out->byte_mask[static_cast<size_t>(b) * W + i] = a.byte_mask[b].w[i];
  for (int c = 0; c < 4; ++c) {
    for (int i = 0; i < W; ++i) {
      out->first[c * W + i] = a.first[c].w[i];
      out->accept[c * W + i] = a.accept[c].w[i];
    }
    for (int k = 0; k < a.n_pos; ++k)
      for (int i = 0; i < W; ++i)
        out->follow[(static_cast<size_t>(c) * a.n_pos + k) * W + i] = a.follow[c][k].w[i];
    out->accept_empty[c] = a.accept_empty[c] ? 1 : 0;
    for (int b = 0; b < 256; ++b) out->start_ok[c * 256 + b] = ca.start_ok[c][b];
  }
  for (int i = 0; i < W; ++i) out->chain[i] = a.chain.w[i];
  if (ca.strategy == ScanStrategy::DfaFixed) {
    const ScanDfa& d = ca.dfa;
    out->dfa_next.resize(d.next.size());
    for (size_t i = 0; i < d.next.size(); ++i)
      out->dfa_next[i] = static_cast<uint16_t>(d.next[i] * d.n_classes);
    out->dfa_class.assign(d.byte_class.begin(), d.byte_class.end());
    const int C = d.n_classes;
    out->dfa_pair.assign(static_cast<size_t>(d.n_states) * C * C, 0);
    for (int s = 0; s < d.n_states; ++s)
      for (int c1 = 0; c1 < C; ++c1) {
        int mid = d.next[static_cast<size_t>(s) * C + c1];
        for (int c2 = 0; c2 < C; ++c2) {
          uint32_t fin = d.next[static_cast<size_t>(mid) * C + c2];
          if (mid >= d.first_accept) fin |= 0x80000000u;
          out->dfa_pair[(static_cast<size_t>(s) * C + c1) * C + c2] = fin;
        }
      }
  }
}

}  // namespace rejit_b200
