// ERE parser for the rejit_b200 host front end.
//
// Accepts exactly the dialect the reference accepts and builds the same tree,
// including the behaviours that change match results:
//   * literals coalesce into nodes of at most 64 bytes; a byte is kept out of
//     the preceding literal only when the NEXT pattern byte is '*' or '{'
//     (so "ab+c" means "(ab)+c")  — /root/reference/src/parser.cc:467-495,
//     /root/reference/src/parser.h:100-104
//   * escapes: \( \) \{ \} \[ \] \| \* \+ \^ \$ \\ literal; \d \D \s \S \n \t
//     \xHH (hex LETTERS decode as 0..5) — parser.cc:24-37, 53-117
//   * ad-hoc bracket expressions — parser.cc:428-464
//   * {m,n} on a literal with m>1 is rewritten literal^m + {0,n-m} when
//     parser_opt is on — parser.cc:372-418
//   * alternation branches are stored in reverse source order and one-branch
//     groups vanish when parser_opt is on — parser.cc:574-610
//   * an unmatched ')' is a literal — parser.cc:510-525
// Patterns on which the reference aborts or invokes undefined behaviour
// (empty pattern, empty alternative, unbalanced '(', leading repetition
// operator, unterminated '[', stray ']') are rejected with a parse error.
#include "ir.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace rejit_b200 {
namespace {

struct Fail { size_t index; std::string msg; };

class EreParser {
 public:
  EreParser(const char* re, size_t len, const ParseOptions& opt)
      : re_(reinterpret_cast<const uint8_t*>(re)), len_(len), opt_(opt) {}

  NodePtr Run() {
    if (len_ == 0) throw Fail{0, "empty regular expression\n"};
    while (pos_ < len_) {
      size_t step = 1;
      uint8_t c = re_[pos_];
      switch (c) {
        case '\\': step = Escape(); break;
        case '{': step = Braces(pos_); break;
        case '.': Push(NodeKind::AnyChar); break;
        case '*': PushRepeat(0, kUnbounded); break;
        case '+': PushRepeat(1, kUnbounded); break;
        case '?': PushRepeat(0, 1); break;
        case '^': Push(NodeKind::LineStart); break;
        case '$': Push(NodeKind::LineEnd); break;
        case '(':
          if (++open_parens_ > kMaxTreeDepth) throw Fail{pos_, "regular expression nested too deeply\n"};
          Push(NodeKind::OpenParen);
          break;
        case ')': CloseParen(); break;
        case '|': Concatenate(); Push(NodeKind::Bar); break;
        case '[': step = Brackets(pos_); break;
        case ']': throw Fail{pos_, Unexpected(c)};
        default: PushByteAt(pos_);
      }
      pos_ += step;
    }
    Alternate();
    if (stack_.size() != 1 || stack_[0]->is_marker()) {
      int open = 0;
      for (auto& n : stack_) open += n->kind == NodeKind::OpenParen;
      char buf[64];
      snprintf(buf, sizeof buf, "Missing %d right-parenthis ')'.\n", open);
      throw Fail{pos_, buf};
    }
    return std::move(stack_[0]);
  }

 private:
  uint8_t At(size_t i) const { return i < len_ ? re_[i] : 0; }
  static std::string Unexpected(uint8_t c) {
    char buf[40];
    snprintf(buf, sizeof buf, "unexpected character %c\n", c);
    return buf;
  }
  Node* Top() { return stack_.empty() ? nullptr : stack_.back().get(); }
  Node* Push(NodeKind k) {
    stack_.emplace_back(new Node(k));
    return stack_.back().get();
  }
  NodePtr Pop() {
    NodePtr n = std::move(stack_.back());
    stack_.pop_back();
    return n;
  }
  NodePtr PopOperand() {
    if (stack_.empty() || Top()->is_marker()) throw Fail{pos_, "nothing to repeat\n"};
    return Pop();
  }

  // Appends one byte to the literal on top of the stack when allowed.
  void PushByte(uint8_t c, bool may_append) {
    Node* t = Top();
    if (may_append && t && t->kind == NodeKind::Literal && t->bytes.size() < kMaxLiteralNode) {
      t->bytes.push_back(c);
      return;
    }
    Push(NodeKind::Literal)->bytes.push_back(c);
  }
  void PushByteAt(size_t i) {
    uint8_t next = At(i + 1);
    PushByte(re_[i], !(next == '*' || next == '{'));
  }
  // A node that wraps others is one level higher than the highest of them.
  void Adopt(Node* parent, NodePtr kid) {
    if (kid->depth + 1 > parent->depth) parent->depth = kid->depth + 1;
    if (parent->depth > kMaxTreeDepth) throw Fail{pos_, "regular expression nested too deeply\n"};
    parent->kids.push_back(std::move(kid));
  }
  void PushRepeat(uint32_t lo, uint32_t hi) {
    NodePtr sub = PopOperand();
    Node* r = Push(NodeKind::Repeat);
    r->rep_min = lo;
    r->rep_max = hi;
    Adopt(r, std::move(sub));
  }

  static int HexQuirk(uint8_t c) {      // letters map to 0..5, as in the reference
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'A' && c <= 'F') return c - 'A';
    if (c >= 'a' && c <= 'f') return c - 'a';
    return -1;
  }

  size_t Escape() {
    uint8_t e = At(pos_ + 1);
    switch (e) {
      case '(': case ')': case '{': case '}': case '[': case ']': case '|':
      case '*': case '+': case '^': case '$': case '\\':
        PushByteAt(pos_ + 1);
        return 2;
      case 'd': case 'D': {
        Node* b = Push(NodeKind::CharSet);
        b->ranges.push_back({'0', '9'});
        b->negated = (e == 'D');
        return 2;
      }
      case 's': case 'S': {
        Node* b = Push(NodeKind::CharSet);
        b->singles = {' ', '\t'};
        b->negated = (e == 'S');
        return 2;
      }
      case 'n': PushByte('\n', true); return 2;
      case 't': PushByte('\t', true); return 2;
      case 'x': {
        int hi = HexQuirk(At(pos_ + 2)), lo = HexQuirk(At(pos_ + 3));
        if (hi < 0 || lo < 0) throw Fail{pos_ + 2, "expected: <two hexadecimal digits>\n"};
        PushByte(static_cast<uint8_t>((hi << 4) | lo), true);
        return 4;
      }
      default:
        throw Fail{pos_ + 1, Unexpected(e)};
    }
  }

  uint32_t Number(size_t* i) {
    size_t j = *i;
    unsigned long long v = 0;
    while (At(j) >= '0' && At(j) <= '9') {
      v = v * 10 + (At(j) - '0');
      if (v > 0xFFFFFFFFull) v = 0xFFFFFFFFull;
      ++j;
    }
    if (j == *i) throw Fail{j, "expected: <base 10 integer>\n"};
    *i = j;
    return static_cast<uint32_t>(v);
  }
  void Need(size_t i, char c) {
    if (At(i) != static_cast<uint8_t>(c)) {
      char buf[32];
      snprintf(buf, sizeof buf, "expected: %c\n", c);
      throw Fail{i, buf};
    }
  }

  size_t Braces(size_t open) {
    size_t i = open + 1;
    uint32_t lo, hi;
    if (At(i) == ',') {
      lo = 0;
      ++i;
      hi = Number(&i);
      Need(i, '}');
      ++i;
    } else {
      lo = Number(&i);
      if (At(i) == ',') {
        ++i;
        if (At(i) == '}') {
          hi = kUnbounded;
          ++i;
        } else {
          hi = Number(&i);
          Need(i, '}');
          ++i;
        }
      } else {
        Need(i, '}');
        ++i;
        hi = lo;
      }
    }
    if (lo > hi) {
      char buf[80];
      snprintf(buf, sizeof buf, "Invalid repetition bounds: %u > %u\n", lo, hi);
      throw Fail{i - 1, buf};
    }
    NodePtr operand = PopOperand();
    if (opt_.parser_opt && operand->kind == NodeKind::Literal && lo > 1) {
      // literal{lo,hi}  ->  literal^lo  literal{0,hi-lo}
      const std::vector<uint8_t> unit = operand->bytes;
      if (static_cast<uint64_t>(lo) * unit.size() > kMaxPatternPositions)
        throw Fail{i - 1, "regular expression too large (more than 4096 byte positions)\n"};
      std::vector<NodePtr> parts;
      NodePtr run(new Node(NodeKind::Literal));
      run->bytes = unit;
      for (uint32_t k = 1; k < lo; ++k) {
        if (run->bytes.size() + unit.size() > kMaxLiteralNode) {
          parts.push_back(std::move(run));
          run.reset(new Node(NodeKind::Literal));
        }
        run->bytes.insert(run->bytes.end(), unit.begin(), unit.end());
      }
      bool as_sequence = (unit.size() * static_cast<size_t>(lo) > kMaxLiteralNode) || lo != hi;
      if (!as_sequence) {
        stack_.push_back(std::move(run));
      } else {
        parts.push_back(std::move(run));
        if (lo != hi) {
          NodePtr tail(new Node(NodeKind::Repeat));
          tail->rep_min = 0;
          tail->rep_max = (hi == kUnbounded) ? kUnbounded : hi - lo;
          NodePtr u(new Node(NodeKind::Literal));
          u->bytes = unit;
          Adopt(tail.get(), std::move(u));
          parts.push_back(std::move(tail));
        }
        Node* seq = Push(NodeKind::Sequence);
        for (auto& part : parts) Adopt(seq, std::move(part));
      }
    } else {
      Node* r = Push(NodeKind::Repeat);
      r->rep_min = lo;
      r->rep_max = hi;
      Adopt(r, std::move(operand));
    }
    return i - open;
  }

  size_t Brackets(size_t open) {
    size_t i = open + 1;
    NodePtr set(new Node(NodeKind::CharSet));
    if (At(i) == '^') { set->negated = true; ++i; }
    if (At(i) == '-') { set->singles.push_back('-'); ++i; }
    for (;;) {
      if (i >= len_) throw Fail{i, "expected: ]\n"};
      if (At(i) == ']') { ++i; break; }
      if (At(i + 1) == ']') {
        set->singles.push_back(At(i));
        i += 1;
      } else if (At(i + 2) == ']') {
        if (i + 1 >= len_) throw Fail{i + 1, "expected: ]\n"};
        set->singles.push_back(At(i));
        set->singles.push_back(At(i + 1));
        i += 2;
      } else if (At(i + 1) == '-') {
        if (i + 2 >= len_) throw Fail{i + 2, "expected: ]\n"};
        set->ranges.push_back({At(i), At(i + 2)});
        i += 3;
      } else {
        set->singles.push_back(At(i));
        i += 1;
      }
    }
    stack_.push_back(std::move(set));
    return i - open;
  }

  void CloseParen() {
    if (open_parens_ == 0) {  // stray ')' is an ordinary byte
      PushByteAt(pos_);
      return;
    }
    --open_parens_;
    Alternate();
    NodePtr inner = Pop();
    if (inner->is_marker() || stack_.empty() || Top()->kind != NodeKind::OpenParen)
      throw Fail{pos_, "empty group\n"};
    stack_.pop_back();
    stack_.push_back(std::move(inner));
  }

  // Folds everything above the nearest marker into one Sequence node.
  void Concatenate() {
    if (stack_.empty()) throw Fail{pos_, "empty alternative\n"};
    size_t i = stack_.size() - 1;
    while (i > 0 && !stack_[i]->is_marker()) --i;
    size_t first = stack_[i]->is_marker() ? i + 1 : i;
    size_t n = stack_.size() - first;
    if (n == 0) throw Fail{pos_, "empty alternative\n"};
    if (n == 1) return;
    NodePtr seq(new Node(NodeKind::Sequence));
    for (size_t k = first; k < stack_.size(); ++k) Adopt(seq.get(), std::move(stack_[k]));
    stack_.resize(first);
    stack_.push_back(std::move(seq));
  }

  // Folds "a | b | c" above the nearest '(' into one Choice (branches reversed).
  void Alternate() {
    Concatenate();
    size_t last = stack_.size() - 1;
    if (opt_.parser_opt &&
        (stack_[last]->kind == NodeKind::OpenParen ||
         (last >= 1 && stack_[last - 1]->kind == NodeKind::OpenParen) || last == 0))
      return;
    NodePtr alt(new Node(NodeKind::Choice));
    size_t i = stack_.size();
    while (i > 0 && stack_[i - 1]->kind != NodeKind::OpenParen) {
      --i;
      if (!stack_[i]->is_marker()) Adopt(alt.get(), std::move(stack_[i]));
    }
    stack_.resize(i);
    stack_.push_back(std::move(alt));
  }

  const uint8_t* re_;
  size_t len_;
  ParseOptions opt_;
  size_t pos_ = 0;
  uint32_t open_parens_ = 0;          // OpenParen markers on the stack
  std::vector<NodePtr> stack_;
};

}  // namespace

NodePtr ParseERE(const char* pattern, size_t len, const ParseOptions& opt, std::string* error) {
  try {
    EreParser p(pattern, len, opt);
    return p.Run();
  } catch (const Fail& f) {
    if (error) {
      // Same layout as the reference's Parser::ParseError (parser.cc:652-665).
      *error = "Error parsing at index " + std::to_string(f.index) + "\n" +
               std::string(pattern, len) + "\n" + std::string(f.index, ' ') + "^ \n" + f.msg;
    }
    return nullptr;
  }
}

}  // namespace rejit_b200
