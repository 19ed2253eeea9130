// State numbering and flattening of the regular-expression tree into the NFA
// edge lists that cross the C ABI.
//
// Restates, with the same resulting state ids and edge order:
//   RegexpIndexer ... /root/reference/src/codegen.cc:91-150
//   RegexpLister .... /root/reference/src/codegen.cc:155-324
// including the behaviours that are visible in match results
// (SURVEY.md Appendix B): a bounded repetition whose upper bound is 1 receives
// a "repeat" epsilon (x? behaves as x*, B4); unrolled copies of a negated
// bracket lose the negation (Bracket::DeepCopy, src/regexp.cc:102-108).
#include "ir.h"

#include <algorithm>
#include <sstream>

namespace rejit_b200 {
namespace {

void SetEntry(Node* n, int s) {
  n->entry = s;
  if (n->kind == NodeKind::Sequence) SetEntry(n->kids.front().get(), s);
  else if (n->kind == NodeKind::Choice) for (auto& k : n->kids) SetEntry(k.get(), s);
}

void SetExit(Node* n, int s) {
  n->exit = s;
  if (n->kind == NodeKind::Sequence) SetExit(n->kids.back().get(), s);
  else if (n->kind == NodeKind::Choice) for (auto& k : n->kids) SetExit(k.get(), s);
}

// Walks a (sub)tree handing out state ids; `cursor` is the state the next
// element starts from and `high` the highest id handed out so far.
struct Numberer {
  int cursor, high;
  void Walk(Node* n) {
    switch (n->kind) {
      case NodeKind::Choice: {
        int from = cursor;
        for (auto& k : n->kids) { Walk(k.get()); --high; }
        ++high;
        SetEntry(n, from);
        SetExit(n, high);
        cursor = n->exit;
        break;
      }
      case NodeKind::Sequence: {
        int from = cursor;
        for (auto& k : n->kids) Walk(k.get());
        SetEntry(n, from);
        SetExit(n, high);
        cursor = n->exit;
        break;
      }
      default:      // physical nodes and Repeat: one fresh exit state
        n->entry = cursor;
        n->exit = ++high;
        cursor = n->exit;
    }
  }
};

NodePtr CloneForUnroll(const Node* n) {
  NodePtr c(new Node(n->kind));
  c->bytes = n->bytes;
  c->singles = n->singles;
  c->ranges = n->ranges;
  c->negated = false;                 // the reference's copy drops the flag
  c->rep_min = n->rep_min;
  c->rep_max = n->rep_max;
  for (auto& k : n->kids) c->kids.push_back(CloneForUnroll(k.get()));
  return c;
}

struct Flattener {
  LoweredRegexp* out;
  int high;                                   // highest state id in use
  std::vector<NodePtr> arena;                 // unrolled copies live here

  void Eps(int a, int b) {
    Edge e;
    e.kind = kEdgeEpsilon; e.entry = a; e.exit = b;
    out->control.push_back(e);
  }

  void Emit(const Node* n) {
    Edge e;
    e.entry = n->entry; e.exit = n->exit;
    switch (n->kind) {
      case NodeKind::Literal: e.kind = kEdgeLiteral; e.bytes = n->bytes; break;
      case NodeKind::AnyChar: e.kind = kEdgeAnyChar; break;
      case NodeKind::CharSet:
        e.kind = kEdgeCharSet; e.negated = n->negated; e.singles = n->singles; e.ranges = n->ranges;
        break;
      case NodeKind::LineStart: e.kind = kEdgeLineStart; break;
      case NodeKind::LineEnd: e.kind = kEdgeLineEnd; break;
      default: return;
    }
    (n->is_control() ? out->control : out->matching).push_back(e);
  }

  void Walk(Node* n) {
    if (n->kind == NodeKind::Sequence || n->kind == NodeKind::Choice) {
      for (auto& k : n->kids) Walk(k.get());
    } else if (n->kind == NodeKind::Repeat) {
      Unroll(n);
    } else {
      Emit(n);
    }
  }

  void Unroll(Node* rep) {
    Node* unit = rep->kids[0].get();
    const uint32_t lo = rep->rep_min, hi = rep->rep_max;
    const bool bounded = hi != kUnbounded;
    if (lo == 0 && hi == 0) { Eps(rep->entry, rep->exit); return; }

    const bool chain = lo > 1 || (hi > 1 && bounded);
    Node* body = unit;
    Node* last_copy = unit;
    std::vector<Node*> copies{unit};
    if (chain) {
      uint32_t n = bounded ? hi : lo;
      NodePtr seq(new Node(NodeKind::Sequence));
      // the sequence borrows `unit`; keep ownership in the Repeat node and
      // record raw pointers for the walk below.
      for (uint32_t i = 1; i < n; ++i) {
        arena.push_back(CloneForUnroll(unit));
        copies.push_back(arena.back().get());
      }
      last_copy = copies.back();
      body = nullptr;
    }

    int body_entry = rep->entry, body_exit = rep->exit;
    if (!bounded) {
      body_exit = -1;
      if (lo <= 1) body_entry = ++high;
    }
    // number the body (a virtual Sequence over `copies` when chained)
    Numberer nb{body_entry, high};
    for (Node* c : copies) nb.Walk(c);
    SetEntry(copies.front(), body_entry);
    int inner_exit = nb.high;              // exit of the virtual sequence
    if (copies.size() == 1) inner_exit = copies[0]->exit;
    if (body_exit != -1) { SetExit(copies.back(), body_exit); inner_exit = body_exit; }
    high = nb.high;
    (void)body;

    for (Node* c : copies) Walk(c);

    if (lo == 0) Eps(rep->entry, rep->exit);
    if (bounded && hi > 1) {
      uint32_t from = lo > 1 ? lo : 1;
      for (size_t i = from - 1; i + 1 < copies.size(); ++i) Eps(copies[i]->exit, rep->exit);
    } else {
      if (lo <= 1) Eps(rep->entry, body_entry);
      Eps(inner_exit, rep->exit);
      Eps(last_copy->exit, last_copy->entry);
    }
  }
};

// What Lower() would produce for the subtree, computed without producing it (saturating): byte positions (exactly
// what BuildAutomaton counts) and tree nodes after unrolling (every one becomes an edge or a few epsilons).
struct Footprint { uint64_t positions, nodes; };
constexpr uint64_t kFootprintSat = 1ull << 40;

Footprint Measure(const Node* n) {
  Footprint f{0, 1};
  switch (n->kind) {
    case NodeKind::Literal: f.positions = n->bytes.size(); break;
    case NodeKind::AnyChar:
    case NodeKind::CharSet: f.positions = 1; break;
    case NodeKind::Sequence:
    case NodeKind::Choice:
      for (auto& k : n->kids) {
        const Footprint g = Measure(k.get());
        f.positions = std::min(kFootprintSat, f.positions + g.positions);
        f.nodes = std::min(kFootprintSat, f.nodes + g.nodes);
      }
      break;
    case NodeKind::Repeat: {
      // Unroll(): {0,0} is one epsilon; otherwise `hi` copies when bounded, else max(lo, 1)
      if (n->rep_min == 0 && n->rep_max == 0) break;
      const uint64_t copies = n->rep_max != kUnbounded ? std::max<uint64_t>(n->rep_max, 1) : std::max<uint64_t>(n->rep_min, 1);
      const Footprint g = Measure(n->kids[0].get());
      f.positions = g.positions > kFootprintSat / copies ? kFootprintSat : g.positions * copies;
      f.nodes = g.nodes + 1 > kFootprintSat / copies ? kFootprintSat : (g.nodes + 1) * copies;
      break;
    }
    default: break;              // anchors: a control edge, no position
  }
  return f;
}

}  // namespace

bool WithinBudget(const Node* root, std::string* error) {
  const Footprint f = Measure(root);
  if (f.positions <= kMaxPatternPositions && f.nodes <= kMaxPatternNodes) return true;
  if (error) *error = "regular expression too large for the sm_100a engine (more than 4096 byte positions after unrolling its repetitions)";
  return false;
}

LoweredRegexp Lower(Node* root) {
  LoweredRegexp lr;
  Numberer nb{0, 0};
  nb.Walk(root);
  SetEntry(root, 0);
  lr.entry_state = 0;
  lr.exit_state = nb.cursor;
  Flattener fl{&lr, nb.high, {}};
  fl.Walk(root);
  lr.n_states = fl.high + 1;
  return lr;
}

std::string DumpLowered(const LoweredRegexp& lr) {
  std::ostringstream os;
  os << "states " << lr.n_states << " entry " << lr.entry_state << " exit " << lr.exit_state << "\n";
  auto one = [&](const Edge& e) {
    switch (e.kind) {
      case kEdgeLiteral:
        os << "MultipleChar {" << e.entry << "," << e.exit << "} ";
        for (uint8_t b : e.bytes) { char h[4]; snprintf(h, sizeof h, "%02x", b); os << h; }
        break;
      case kEdgeAnyChar: os << "Period {" << e.entry << "," << e.exit << "}"; break;
      case kEdgeCharSet: os << "Bracket {" << e.entry << "," << e.exit << "} " << (e.negated ? "neg" : "pos"); break;
      case kEdgeLineStart: os << "StartOfLine {" << e.entry << "," << e.exit << "}"; break;
      case kEdgeLineEnd: os << "EndOfLine {" << e.entry << "," << e.exit << "}"; break;
      default: os << "Epsilon {" << e.entry << "," << e.exit << "}";
    }
    os << "\n";
  };
  os << "control\n";
  for (auto& e : lr.control) one(e);
  os << "matching\n";
  for (auto& e : lr.matching) one(e);
  return os.str();
}

}  // namespace rejit_b200
