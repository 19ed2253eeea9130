// Ahead-of-time lowering of the indexed NFA (rejit_b200_ir) to the tables the
// sm_100a kernels consume.  This replaces the reference's x64 code generator
// (/root/reference/src/x64/codegen-x64.cc) — nothing here emits machine code.
//
// The "position automaton" has one position per byte-consuming step of the
// NFA (each byte of a literal edge, each '.', each bracket).  All epsilon /
// ^ / $ edges are folded into context-dependent closure sets, where the
// context of a text offset p is two bits:
//     sol(p) = p==0 || text[p-1] in {\n,\r}      (codegen-x64.cc:686-732)
//     eol(p) = p==N || text[p]   in {\n,\r}
// so that one NFA step at offset p is   A' = Follow[ctx(p)](A) & B[text[p]].
#ifndef REJIT_B200_HOST_AUTOMATON_H_
#define REJIT_B200_HOST_AUTOMATON_H_

#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "ir.h"

namespace rejit_b200 {

struct BitSet {
  std::vector<uint32_t> w;
  BitSet() {}
  explicit BitSet(int nbits) : w((nbits + 31) / 32, 0u) {}
  void set(int i) { w[i >> 5] |= 1u << (i & 31); }
  bool test(int i) const { return (w[i >> 5] >> (i & 31)) & 1u; }
  bool any() const { for (uint32_t x : w) if (x) return true; return false; }
  void or_with(const BitSet& o) { for (size_t i = 0; i < w.size(); ++i) w[i] |= o.w[i]; }
  void and_with(const BitSet& o) { for (size_t i = 0; i < w.size(); ++i) w[i] &= o.w[i]; }
  bool operator==(const BitSet& o) const { return w == o.w; }
  bool operator<(const BitSet& o) const { return w < o.w; }
};

constexpr int kCtxCount = 4;                 // ctx = sol | (eol << 1)
constexpr uint64_t kInfLen = ~0ull;

struct PositionNfa {
  int n_pos = 0;
  int words = 0;                             // 32-bit words per position set
  bool has_anchor = false;                   // any ^ or $ edge
  std::vector<std::array<uint32_t, 8>> cls;  // 256-bit byte class per position
  std::vector<BitSet> byte_mask;             // [256] positions accepting byte c
  BitSet first[kCtxCount];
  std::vector<BitSet> follow[kCtxCount];     // [n_pos]
  BitSet accept[kCtxCount];                  // positions after which exit is live
  bool accept_empty[kCtxCount] = {false, false, false, false};
  BitSet chain;                              // k with follow[*][k] == {k+1}
  uint64_t min_len = 0, max_len = 0;         // bytes; max_len may be kInfLen
};

// Table-driven DFA over byte classes for anchor-free fixed-length patterns
// (unanchored search: state = set of live positions).
struct ScanDfa {
  int n_states = 0;
  int n_classes = 0;
  int first_accept = 0;                      // states >= first_accept accept
  std::array<uint8_t, 256> byte_class{};
  std::vector<uint16_t> next;                // [n_states * n_classes]
  uint64_t match_len = 0;
};

enum class ScanStrategy : int32_t {
  Literal = 0,        // whole pattern is one byte string
  DfaFixed = 1,       // fixed-length, anchor-free: exact DFA scan
  LiteralWindow = 2,  // required literal + bounded prefix: scan, then verify windows
  Generic = 3         // start filter + per-start NFA run over every offset
};

struct CompiledAutomaton {
  PositionNfa nfa;
  ScanStrategy strategy = ScanStrategy::Generic;
  std::vector<uint8_t> literal;              // Literal / LiteralWindow needle
  uint32_t window_lo = 0, window_hi = 0;     // LiteralWindow: start in [hit-hi, hit-lo]
  ScanDfa dfa;                               // DfaFixed
  std::array<uint8_t, 256> start_ok[kCtxCount];  // Generic: byte can begin a match
  // A running thread can re-enter positions of the start set
  // (first[ctx] & follow[ctx][k] != 0): selection must replay the reference's
  // thread labels (device_program.h: FaithfulSegment) instead of chaining E(s).
  bool reentrant = false;
  std::string describe;                      // one-line human summary
};

// Builds everything from the lowered regexp.  Returns false (with *error) for
// patterns beyond the engine's static limits.
bool BuildAutomaton(const LoweredRegexp& lr, CompiledAutomaton* out, std::string* error);

// One DFA for a SET of anchor-free fixed-length patterns (regex-dna's nine
// variants are counted over the same text): the subset construction runs over
// the union of the patterns' position automata, so one pass over the text
// advances all patterns at once; an accepting state carries the bitmask of the
// patterns that end there.  ("fused multi-pattern".)
constexpr int kKmerR = 3;               // ends answered by one bitmap lookup (2: 32 KB bitmap, 3: 128 KB)

struct SetDfa {
  int n_patterns = 0;
  int n_states = 0, n_classes = 0, first_accept = 0;
  std::array<uint8_t, 256> byte_class{};
  std::vector<uint16_t> next;            // [n_states * n_classes] -> state id
  std::vector<uint32_t> accept_mask;     // [n_rows] bit j: pattern j ends here (0 for shadow rows)
  std::vector<uint32_t> match_len;       // [n_patterns]
  uint32_t max_len = 0;
  // flat device layouts.  Rows = the states, followed by SHADOW rows: a copy of a
  // non-accepting state's row, entered when the state in between two bytes
  // accepted, so that "some accept happened" is visible in the row address alone
  // (rows >= first_accept) and the scan needs no flag bit in the entries.
  int n_rows = 0;
  std::vector<uint16_t> row_state;       // [n_rows] the state a row stands for
  std::vector<uint16_t> t1;              // [n_rows*C]   next state * C
  std::vector<uint32_t> t2;              // [n_rows][2^row_shift / 4]: (row after two bytes) << row_shift
  int row_shift = 0;                     // log2 of the padded row size in bytes
  // k-mer index (k_set_kmer): when every member is at most 8 bytes long and the
  // bytes any pattern position accepts ("live" bytes) are at most four values
  // that (b >> shift) & 3 tells apart, "does a member end at e" depends only on
  // the 2-bit codes of the eight bytes before e: no state, no dependent chain.
  //   mask16[x]  x = eight codes, oldest in the low bits; bit j: member j accepts
  //              the last match_len[j] letters of the (canonical) 8-mer
  //   bitmap[w]  7 + R codes x (R = kKmerR consecutive ends at once), B = 2 (7 + R) - 5:
  //              bit 31 - (x >> B) of word x & (2^B - 1) is set iff one of the R
  //              8-mers (x >> 2k) & 0xFFFF, k < R, is accepted by some member
  // A byte whose code aliases a live byte is weeded out by the exact check on a hit.
  struct Kmer {
    bool ok = false;
    uint32_t shift = 0;
    uint32_t canon = 0;                  // byte c: the live byte with code c
    uint32_t canon_ok = 0;               // bit c: code c has a live byte
    uint32_t len_le[9] = {0};            // bit j: match_len[j] <= v
    std::vector<uint32_t> bitmap;        // [2^(2 (7 + R) - 5)]
    std::vector<uint32_t> mask16;        // [65536]
  } kmer;
};
// Returns false when the set cannot be fused (a member is not a fixed-length
// anchor-free DFA pattern, or the tables exceed the kernel's budget).
bool BuildSetDfa(const std::vector<const CompiledAutomaton*>& members, SetDfa* out);

// The tables in the flat layouts the kernels index (device_program.h:
// NfaTables; engine.cu: DfaTables).  The engine uploads these vectors verbatim.
struct FlatTables {
  int n_pos = 0, words = 1;
  std::vector<uint32_t> byte_mask;   // [256][W]
  std::vector<uint32_t> first;       // [4][W]
  std::vector<uint32_t> follow;      // [4][max(n_pos,1)][W]
  std::vector<uint32_t> accept;      // [4][W]
  std::vector<uint32_t> chain;       // [W]
  std::vector<uint8_t> start_ok;     // [4][256]
  uint8_t accept_empty[4] = {0, 0, 0, 0};
  std::vector<uint16_t> dfa_next;    // [states*classes], entries pre-multiplied by classes
  std::vector<uint8_t> dfa_class;    // [256]
  // two-byte steps: [state][c1 * classes + c2] = state after both bytes, bit 31
  // set when the state after the FIRST byte is accepting (a match ends between
  // the two bytes)
  std::vector<uint32_t> dfa_pair;
};
void FlattenTables(const CompiledAutomaton& ca, FlatTables* out);

}  // namespace rejit_b200

#endif  // REJIT_B200_HOST_AUTOMATON_H_
