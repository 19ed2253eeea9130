// rejit_b200 host front end: regular-expression tree and the lowered NFA
// ("indexed" physical regexps) that is handed across the C ABI
// (include/rejit_b200.h: rejit_b200_ir) to the sm_100a engine.
//
// The SHAPE of this IR — node kinds, the <=64-character literal nodes, the
// entry/exit state numbering, the matching/control edge lists — restates the
// reference's Regexp IR so that both engines lower a pattern to the same NFA:
//   node kinds ............ /root/reference/src/regexp.h:27-68, 115-471
//   RegexpInfo lists ...... /root/reference/src/regexp.h:538-636
// The code is written from scratch for this project.
#ifndef REJIT_B200_HOST_IR_H_
#define REJIT_B200_HOST_IR_H_

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace rejit_b200 {

constexpr uint32_t kUnbounded = 0xFFFFFFFFu;   // "no upper repetition bound"
constexpr unsigned kMaxLiteralNode = 64;       // literal nodes hold at most 64 bytes
// Budgets enforced BEFORE anything is expanded (a hostile pattern must fail with an error, not exhaust memory or
// the stack): nesting depth of the tree (recursive walks, destructors), byte positions after unrolling (the
// engine's own cap, automaton.cc), and open parentheses on the parser stack.
constexpr uint32_t kMaxTreeDepth = 200;
constexpr uint64_t kMaxPatternPositions = 4096;
constexpr uint64_t kMaxPatternNodes = 1u << 18;   // tree nodes after unrolling (epsilon-only repetitions have no positions)

enum class NodeKind : uint8_t {
  Literal,        // a run of bytes matched verbatim ("MultipleChar")
  AnyChar,        // '.' : any byte except \n and \r
  CharSet,        // [...] : singles + signed ranges, optionally negated
  LineStart,      // ^
  LineEnd,        // $
  Sequence,       // concatenation
  Choice,         // alternation
  Repeat,         // {min,max}
  OpenParen,      // parser-only marker
  Bar             // parser-only marker
};

struct ByteRange { uint8_t lo, hi; };

struct Node {
  NodeKind kind;
  int entry = -1, exit = -1;                  // NFA state ids once indexed
  std::vector<uint8_t> bytes;                 // Literal
  bool negated = false;                       // CharSet
  std::vector<uint8_t> singles;               // CharSet
  std::vector<ByteRange> ranges;              // CharSet
  std::vector<std::unique_ptr<Node>> kids;    // Sequence / Choice; Repeat has exactly one
  uint32_t rep_min = 0, rep_max = 0;          // Repeat
  uint32_t depth = 1;                         // height of the subtree (parser-maintained, bounded by kMaxTreeDepth)

  explicit Node(NodeKind k) : kind(k) {}
  bool is_marker() const { return kind == NodeKind::OpenParen || kind == NodeKind::Bar; }
  bool is_physical() const { return kind <= NodeKind::LineEnd; }
  bool is_control() const { return kind == NodeKind::LineStart || kind == NodeKind::LineEnd; }
};
using NodePtr = std::unique_ptr<Node>;

// Edge kinds of the lowered NFA.  Values are part of the C ABI
// (include/rejit_b200.h: REJIT_B200_EDGE_*).
enum EdgeKind : int32_t {
  kEdgeLiteral = 0, kEdgeAnyChar = 1, kEdgeCharSet = 2,
  kEdgeLineStart = 3, kEdgeLineEnd = 4, kEdgeEpsilon = 5
};

struct Edge {
  int32_t kind;
  int32_t entry, exit;
  std::vector<uint8_t> bytes;        // Literal
  bool negated = false;              // CharSet
  std::vector<uint8_t> singles;
  std::vector<ByteRange> ranges;
};

// Result of lowering one pattern: what the reference keeps in RegexpInfo after
// RegexpIndexer + RegexpLister have run.
struct LoweredRegexp {
  int n_states = 0;
  int entry_state = 0;
  int exit_state = 0;
  std::vector<Edge> matching;        // byte-consuming edges, listing order
  std::vector<Edge> control;         // ^, $ and epsilon edges, listing order
};

struct ParseOptions {
  bool parser_opt = true;            // the reference's --use_parser_opt flag
};

// Parses an ERE; on failure returns nullptr and fills *error with a message in
// the reference's format ("Error parsing at index N\n<re>\n<spaces>^ \n<msg>").
NodePtr ParseERE(const char* pattern, size_t len, const ParseOptions& opt,
                 std::string* error);

// Size of the NFA that Lower() would produce, checked before it is produced: false (and *error, a parse-error
// style message) when unrolling the repetitions would exceed kMaxPatternPositions byte positions.
bool WithinBudget(const Node* root, std::string* error);

// Assigns NFA state numbers and flattens the tree into edge lists.  Call WithinBudget first.
LoweredRegexp Lower(Node* root);

// Human-readable dump used by the IR-parity tests.
std::string DumpLowered(const LoweredRegexp& lr);

}  // namespace rejit_b200

#endif  // REJIT_B200_HOST_IR_H_
