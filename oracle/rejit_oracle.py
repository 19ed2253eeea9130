"""oracle/rejit_oracle.py — TEST INFRASTRUCTURE (the parity checker), never shipped.

A CPU restatement of the reference engine's *semantics* for the MatchAll hot
path (and its MatchFirst / MatchFull / MatchAnywhere siblings), independent of
the product code under rejit_b200/:

  * front end (this file, pure Python; patterns are tens of bytes):
      ERE parser      follows /root/reference/src/parser.cc:40-195, 317-649
      state indexer   follows /root/reference/src/codegen.cc:91-150
      lister          follows /root/reference/src/codegen.cc:155-324
      control-list topological sort  follows /root/reference/src/regexp.cc:286-352
  * back end (oracle/nfa_sim.c, plain C, called through ctypes): a sequential
    simulation of the code the reference JIT emits with fast-forward disabled,
    /root/reference/src/x64/codegen-x64.cc:535-677 (per-byte loop), :401-522
    (match bookkeeping), :951-1097 (state ring), plus the C++ call-back
    /root/reference/src/codegen.cc:36-86 (MatchAllAppendFilter).

Parity pin: this oracle is checked against the real reference compiled from
/root/reference (oracle/_ref/librejit_ref.so, flag set "noff" =
use_fast_forward=0, the only configuration in which the reference passes
282/282 of its own tests — SURVEY.md §8c) by tests/test_oracle.py: on every
check of tools/tests/test.cc (committed as tests/golden/ref_test_table.json),
on the committed MatchAll offset fixtures (tests/golden/matchall_offsets.json),
on the lowered-IR dumps (tests/golden/ir_dumps.json) and — when _ref is
present — on randomized differential runs.

Where the reference itself is wrong in that configuration the oracle states what
it does: B19 (a match lost after an abutting match of a re-entrant pattern) is
reproduced, because it is a property of the matching semantics; B18 (use after
free in the parser for literal{m}) and B20 (literal nodes longer than 16 bytes
compared on the last repeated compare's flags only) are NOT: the oracle expands
and compares literals exactly, and B20 is restated behind a switch
(Oracle(long_literal_defect=True), nfa_sim.c mc_equal) only so that
tests/test_oracle.py can show that this compare is the whole difference.
DESIGN.md section 2 has the evidence.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

INF = 0xFFFFFFFF          # kMaxUInt, src/globals.h
MAX_NODE_LEN = 64         # kMaxNodeLength, src/regexp.h:107

# edge kinds handed to nfa_sim.c
K_MC, K_PERIOD, K_BRACKET, K_SOL, K_EOL, K_EPS = range(6)


class ParserError(Exception):
    def __init__(self, index: int, msg: str):
        super().__init__(f"Error parsing at index {index}: {msg}")
        self.index = index
        self.msg = msg


# --------------------------------------------------------------------------
# Tree nodes (src/regexp.h:115-471).  `entry`/`exit` are NFA state ids.
# --------------------------------------------------------------------------
@dataclass
class Node:
    kind: str                      # mc dot br sol eol eps cat alt rep lpar bar
    entry: int = -1
    exit: int = -1
    chars: bytearray = field(default_factory=bytearray)      # mc
    neg: bool = False                                        # br
    singles: List[int] = field(default_factory=list)         # br
    ranges: List[Tuple[int, int]] = field(default_factory=list)  # br (lo, hi)
    subs: List["Node"] = field(default_factory=list)         # cat / alt
    sub: Optional["Node"] = None                             # rep
    rmin: int = 0
    rmax: int = 0

    def is_marker(self) -> bool:
        return self.kind in ("lpar", "bar")

    def is_control(self) -> bool:
        return self.kind in ("sol", "eol", "eps")

    # SetEntryState / SetExitState, src/regexp.cc:152-203
    def set_entry(self, s: int) -> None:
        self.entry = s
        if self.kind == "cat":
            self.subs[0].set_entry(s)          # .at(0): raises on empty concat
        elif self.kind == "alt":
            for r in self.subs:
                r.set_entry(s)

    def set_exit(self, s: int) -> None:
        self.exit = s
        if self.kind == "cat":
            self.subs[-1].set_exit(s)
        elif self.kind == "alt":
            for r in self.subs:
                r.set_exit(s)

    # DeepCopy, src/regexp.cc:62-66,102-108,131-135,157-161,243-245.
    # NB Bracket::DeepCopy does not copy flags_ (the non_matching bit is lost
    # on unrolled copies) — reproduced on purpose.
    def deep_copy(self) -> "Node":
        k = self.kind
        if k == "mc":
            return Node("mc", chars=bytearray(self.chars))
        if k == "br":
            return Node("br", neg=False, singles=list(self.singles),
                        ranges=list(self.ranges))
        if k in ("dot", "sol", "eol"):
            return Node(k)
        if k in ("cat", "alt"):
            return Node(k, subs=[r.deep_copy() for r in self.subs])
        if k == "rep":
            return Node("rep", sub=self.sub.deep_copy(), rmin=self.rmin,
                        rmax=self.rmax)
        raise AssertionError(k)


# --------------------------------------------------------------------------
# Parser (ERE only, as exposed by Regej — src/rejit.cc:127-137)
# --------------------------------------------------------------------------
def _hex_code(c: int) -> int:
    # hex_code_from_char, src/parser.cc:24-37: letters decode as 0..5 (bug kept)
    if 0x30 <= c <= 0x39:
        return c - 0x30
    if 0x41 <= c <= 0x46:
        return c - 0x41
    if 0x61 <= c <= 0x66:
        return c - 0x61
    raise ParserError(0, "bad hex digit")


class _Parser:
    def __init__(self, pattern: bytes, parser_opt: bool = True):
        self.re = pattern
        self.opt = parser_opt
        self.stack: List[Node] = []
        self.index = 0

    def at(self, i: int) -> int:
        return self.re[i] if i < len(self.re) else 0

    def tos(self) -> Optional[Node]:
        return self.stack[-1] if self.stack else None

    def pop(self, where: int) -> Node:
        if not self.stack:
            # the reference pops an empty vector here (undefined behaviour)
            raise ParserError(where, "nothing to repeat")
        return self.stack.pop()

    # PushChar, src/parser.cc:467-495
    def push_char(self, c: int, append: bool = True) -> None:
        t = self.tos()
        if append and t is not None and t.kind == "mc" and len(t.chars) < MAX_NODE_LEN:
            t.chars.append(c)
            return
        self.stack.append(Node("mc", chars=bytearray([c])))

    def push_char_at(self, i: int) -> None:
        la = self.at(i + 1) if self.at(i) != 0 else 0
        retro = la in (0x2A, 0x7B)            # '*' '{' — IsRetroactiveChar, parser.h:100-104
        self.push_char(self.at(i), not retro)

    def parse(self) -> Node:
        re = self.re
        if len(re) == 0:
            raise ParserError(0, "empty regular expression")
        while self.index < len(re):
            c = re[self.index]
            la = self.at(self.index + 1)
            adv = 1
            if c == 0x5C:                      # backslash, parser.cc:53-117
                adv = 2
                if la in b"(){}[]|*+^$\\" and la != 0:
                    self.push_char_at(self.index + 1)
                elif la in (0x64, 0x44):       # \d \D
                    self.stack.append(Node("br", neg=(la == 0x44), ranges=[(0x30, 0x39)]))
                elif la == 0x6E:               # \n
                    self.push_char(0x0A)
                elif la in (0x73, 0x53):       # \s \S
                    self.stack.append(Node("br", neg=(la == 0x53), singles=[0x20, 0x09]))
                elif la == 0x74:               # \t
                    self.push_char(0x09)
                elif la == 0x78:               # \xHH
                    adv = 4
                    try:
                        v = ((_hex_code(self.at(self.index + 2)) << 4) |
                             _hex_code(self.at(self.index + 3))) & 0xFF
                    except ParserError:
                        raise ParserError(self.index + 2, "bad hex escape")
                    self.push_char(v)
                else:
                    raise ParserError(self.index + 1, "unexpected character")
            elif c == 0x7B:                    # '{'
                adv = self.parse_curly(self.index)
            elif c == 0x2E:                    # '.'
                self.stack.append(Node("dot"))
            elif c == 0x2A:                    # '*'  parser.cc:613-617
                self.stack.append(Node("rep", sub=self.pop_operand(), rmin=0, rmax=INF))
            elif c == 0x2B:                    # '+'
                self.stack.append(Node("rep", sub=self.pop_operand(), rmin=1, rmax=INF))
            elif c == 0x3F:                    # '?'
                self.stack.append(Node("rep", sub=self.pop_operand(), rmin=0, rmax=1))
            elif c == 0x5E:
                self.stack.append(Node("sol"))
            elif c == 0x24:
                self.stack.append(Node("eol"))
            elif c == 0x28:
                self.stack.append(Node("lpar"))
            elif c == 0x29:
                self.do_right_paren()
            elif c == 0x7C:
                self.do_concatenation()
                self.stack.append(Node("bar"))
            elif c == 0x5B:
                adv = self.parse_brackets(self.index)
            elif c == 0x5D:
                raise ParserError(self.index, "unexpected character")   # UNREACHABLE() in the reference
            else:
                self.push_char_at(self.index)
            self.index += adv
        # DoFinish, parser.cc:634-649
        self.do_alternation()
        if len(self.stack) != 1:
            raise ParserError(self.index, "Missing right-parenthesis ')'")
        root = self.stack[0]
        if root.is_marker():
            raise ParserError(self.index, "Missing right-parenthesis ')'")
        return root

    def pop_operand(self) -> Node:
        r = self.pop(self.index)
        if r.is_marker():
            # the reference would wrap a parser marker in a Repetition and
            # crash later; reject instead.
            raise ParserError(self.index, "nothing to repeat")
        return r

    def parse_uint(self, i: int) -> Tuple[int, int]:
        j = i
        while 0x30 <= self.at(j) <= 0x39:
            j += 1
        if j == i:
            raise ParserError(i, "expected: <base 10 integer>")
        return int(self.re[i:j]) & 0xFFFFFFFF, j

    # ParseCurlyBrackets, parser.cc:317-425
    def parse_curly(self, lcb: int) -> int:
        c = lcb + 1
        if self.at(c) == 0x2C:                 # {,n}
            rmin = 0
            rmax, c = self.parse_uint(c + 1)
            if self.at(c) != 0x7D:
                raise ParserError(c, "expected: }")
            c += 1
        else:
            rmin, c = self.parse_uint(c)
            if self.at(c) == 0x2C:
                c += 1
                if self.at(c) == 0x7D:
                    rmax = INF
                    c += 1
                else:
                    rmax, c = self.parse_uint(c)
                    if self.at(c) != 0x7D:
                        raise ParserError(c, "expected: }")
                    c += 1
            else:
                if self.at(c) != 0x7D:
                    raise ParserError(c, "expected: }")
                c += 1
                rmax = rmin
        if rmin > rmax:
            raise ParserError(c - 1, f"Invalid repetition bounds: {rmin} > {rmax}")
        re = self.pop_operand()
        if self.opt and re.kind == "mc" and rmin > 1:
            # a{min,max} -> a^min a{0,max-min}, parser.cc:372-418
            base = bytes(re.chars)
            pieces: List[Node] = []
            cur = Node("mc", chars=bytearray(base))
            use_concat = (len(base) * rmin > MAX_NODE_LEN) or (rmin != rmax)
            for _ in range(rmin - 1):
                if len(cur.chars) + len(base) > MAX_NODE_LEN:
                    pieces.append(cur)
                    cur = Node("mc")
                cur.chars.extend(base)
            if use_concat:
                pieces.append(cur)
                if rmin != rmax:
                    pieces.append(Node("rep", sub=Node("mc", chars=bytearray(base)), rmin=0,
                                       rmax=INF if rmax == INF else rmax - rmin))
                self.stack.append(Node("cat", subs=pieces))
            else:
                self.stack.append(cur)
        else:
            self.stack.append(Node("rep", sub=re, rmin=rmin, rmax=rmax))
        return c - lcb

    # ParseBrackets, parser.cc:428-464 (ad hoc: no escapes, no classes)
    def parse_brackets(self, lb: int) -> int:
        c = lb + 1
        br = Node("br")
        if self.at(c) == 0x5E:
            br.neg = True
            c += 1
        if self.at(c) == 0x2D:
            br.singles.append(0x2D)
            c += 1
        while True:
            if self.at(c) == 0:
                raise ParserError(c, "expected: ]")      # reference reads past the NUL
            if self.at(c) == 0x5D:
                c += 1
                break
            if self.at(c + 1) == 0x5D:
                br.singles.append(self.at(c))
                c += 1
            elif self.at(c + 2) == 0x5D:
                if self.at(c + 1) == 0:
                    raise ParserError(c + 1, "expected: ]")
                br.singles.append(self.at(c))
                br.singles.append(self.at(c + 1))
                c += 2
            elif self.at(c + 1) == 0x2D:
                if self.at(c + 2) == 0:
                    raise ParserError(c + 2, "expected: ]")
                br.ranges.append((self.at(c), self.at(c + 2)))
                c += 3
            else:
                br.singles.append(self.at(c))
                c += 1
        self.stack.append(br)
        return c - lb

    # DoRightParenthesis, parser.cc:510-525
    def do_right_paren(self) -> None:
        if not any(r.kind == "lpar" for r in self.stack):
            self.push_char_at(self.index)          # unmatched ')' is a literal
            return
        self.do_alternation()
        inner = self.stack.pop()
        if not self.stack or self.stack[-1].kind != "lpar":
            raise ParserError(self.index, "empty group")
        self.stack.pop()
        if inner.is_marker():
            raise ParserError(self.index, "empty group")
        self.stack.append(inner)

    # DoConcatenation, parser.cc:542-571
    def do_concatenation(self) -> None:
        if not self.stack:
            raise ParserError(self.index, "empty alternative")
        i = len(self.stack) - 1
        while i > 0 and not self.stack[i].is_marker():
            i -= 1
        first = i + 1 if self.stack[i].is_marker() else i
        n = len(self.stack) - first
        if n == 0:
            # the reference builds an empty Concatenation here and later throws
            # std::out_of_range from SetEntryState(.at(0)); reject up front.
            raise ParserError(self.index, "empty alternative")
        if n != 1:
            cat = Node("cat", subs=self.stack[first:])
            del self.stack[first:]
            self.stack.append(cat)

    # DoAlternation, parser.cc:574-610 (branches pushed in REVERSE order)
    def do_alternation(self) -> None:
        self.do_concatenation()
        last = len(self.stack) - 1
        if self.opt and (self.stack[last].kind == "lpar" or
                         (last - 1 >= 0 and self.stack[last - 1].kind == "lpar") or
                         last == 0):
            return
        alt = Node("alt")
        i = last
        while i >= 0 and self.stack[i].kind != "lpar":
            if not self.stack[i].is_marker():
                alt.subs.append(self.stack[i])
            i -= 1
        first = 0 if i < 0 else i + 1
        del self.stack[first:]
        self.stack.append(alt)


# --------------------------------------------------------------------------
# Indexer + Lister
# --------------------------------------------------------------------------
class _Indexer:
    """RegexpIndexer, src/codegen.cc:91-150."""

    def __init__(self, entry_state: int = 0, last_state: int = 0):
        self.entry_state = entry_state
        self.last_state = last_state

    def visit(self, r: Node) -> None:
        if r.kind == "alt":
            orig = self.entry_state
            for s in r.subs:
                self.visit(s)
                self.last_state -= 1
            self.last_state += 1
            r.set_entry(orig)
            r.set_exit(self.last_state)
            self.entry_state = r.exit
        elif r.kind == "cat":
            orig = self.entry_state
            for s in r.subs:
                self.visit(s)
            r.set_entry(orig)
            r.set_exit(self.last_state)
            self.entry_state = r.exit
        else:                                   # physical regexps and Repetition
            r.set_entry(self.entry_state)
            self.last_state += 1
            r.set_exit(self.last_state)
            self.entry_state = r.exit


@dataclass
class Edge:
    kind: int
    entry: int
    exit: int
    node: Optional[Node] = None

    def label(self) -> str:
        n = self.node
        if self.kind == K_MC:
            return "MC[" + bytes(n.chars).decode("latin-1") + "]"
        if self.kind == K_BRACKET:
            s = bytes(n.singles).decode("latin-1")
            r = ",".join(f"{chr(a)}-{chr(b)}" for a, b in n.ranges)
            return ("NBR[" if n.neg else "BR[") + s + "|" + r + "]"
        return ["MC", "PERIOD", "BRACKET", "SOL", "EOL", "EPS"][self.kind]


@dataclass
class LoweredRegexp:
    """What RegexpInfo holds after Codegen::Compile, src/regexp.h:538-636."""
    n_states: int
    entry_state: int
    exit_state: int
    matching: List[Edge]
    control: List[Edge]           # in PROCESSING order (after SortTopoligcal)
    control_unsorted: List[Edge]  # in listing order (what --print_re_list shows)
    topo_sorted: bool


class _Lister:
    """RegexpLister, src/codegen.cc:155-324."""

    def __init__(self, last_state: int):
        self.last_state = last_state            # rinfo->last_state()
        self.matching: List[Edge] = []
        self.control: List[Edge] = []

    def list(self, r: Node) -> None:
        kind = {"mc": K_MC, "dot": K_PERIOD, "br": K_BRACKET, "sol": K_SOL,
                "eol": K_EOL, "eps": K_EPS}[r.kind]
        e = Edge(kind, r.entry, r.exit, r)
        (self.control if r.is_control() else self.matching).append(e)

    def eps(self, a: int, b: int) -> None:
        self.control.append(Edge(K_EPS, a, b, None))

    def visit(self, r: Node) -> None:
        if r.kind in ("alt", "cat"):
            for s in r.subs:
                self.visit(s)
        elif r.kind == "rep":
            self.visit_rep(r)
        else:
            self.list(r)

    def visit_rep(self, rep: Node) -> None:
        base, rmin, rmax = rep.sub, rep.rmin, rep.rmax
        limited = rmax != INF
        if rmin == 0 and rmax == 0:
            self.eps(rep.entry, rep.exit)
            return
        needs_concat = rmin > 1 or (rmax > 1 and limited)
        if not needs_concat:
            inside = base
        else:
            n_rep = rmax if limited else rmin
            inside = Node("cat", subs=[base] + [base.deep_copy() for _ in range(n_rep - 1)])
        inside_entry, inside_exit = rep.entry, rep.exit
        if not limited:
            inside_exit = -1
            if rmin <= 1:
                self.last_state += 1
                inside_entry = self.last_state
        ix = _Indexer(inside_entry, self.last_state)
        ix.visit(inside)                        # IndexSub, codegen.cc:98-105
        inside.set_entry(inside_entry)
        if inside_exit != -1:
            inside.set_exit(inside_exit)
        self.last_state = ix.last_state
        self.visit(inside)
        if rmin == 0:
            self.eps(rep.entry, rep.exit)
        if limited and rmax > 1:
            lo = max(1, rmin)
            for it in inside.subs[lo - 1:-1]:
                self.eps(it.exit, rep.exit)
        else:
            if rmin <= 1:
                self.eps(rep.entry, inside.entry)
            self.eps(inside.exit, rep.exit)
            last = inside.subs[-1] if needs_concat else inside
            self.eps(last.exit, last.entry)


def _sort_topological(ctrl: List[Edge]) -> Tuple[List[Edge], bool]:
    """SortTopoligcal, src/regexp.cc:286-352 (multimaps iterate by key, equal
    keys in insertion order)."""
    n = len(ctrl)
    if n <= 1:
        return list(ctrl), True
    entries = sorted(range(n), key=lambda i: ctrl[i].entry)     # stable
    exits = {}
    for i in range(n):
        exits.setdefault(ctrl[i].exit, []).append(i)
    sorted_states: List[int] = []
    for i in entries:
        s = ctrl[i].entry
        if s not in exits and s not in sorted_states:
            sorted_states.append(s)
    if len(sorted_states) == n:
        return list(ctrl), True
    exits = {k: list(v) for k, v in exits.items()}
    out: List[int] = []
    while sorted_states:
        cur = sorted_states.pop()
        for i in [j for j in entries if ctrl[j].entry == cur]:
            out.append(i)
            ex = ctrl[i].exit
            if ex in exits and i in exits[ex]:
                exits[ex].remove(i)
                if not exits[ex]:
                    del exits[ex]
            if ex not in exits:
                sorted_states.append(ex)
    if len(out) == n:
        return [ctrl[i] for i in out], True
    return list(ctrl), False


def lower(pattern, parser_opt: bool = True) -> LoweredRegexp:
    """Parser → Indexer → Lister, the arch-independent half of
    Codegen::Compile (src/codegen.cc:591-656)."""
    if isinstance(pattern, str):
        pattern = pattern.encode("latin-1")
    root = _Parser(pattern, parser_opt).parse()
    ix = _Indexer(0, 0)
    ix.visit(root)
    root.set_entry(0)
    exit_state = ix.entry_state
    lister = _Lister(ix.last_state)
    lister.visit(root)
    ctrl_sorted, ok = _sort_topological(lister.control)
    return LoweredRegexp(lister.last_state + 1, 0, exit_state, lister.matching,
                         ctrl_sorted, list(lister.control), ok)


# --------------------------------------------------------------------------
# Back end: oracle/nfa_sim.c through ctypes
# --------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _sim():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libnfa_sim.so")
        src = os.path.join(_HERE, "nfa_sim.c")
        if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-o", so, src])
        _lib = ctypes.CDLL(so)
        _lib.nfa_sim_run.restype = ctypes.c_int64
        _lib.nfa_sim_run.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,       # mode, n_states, entry, exit
            ctypes.POINTER(ctypes.c_int32), ctypes.c_int,                 # edges (6 ints each), n_match
            ctypes.c_int, ctypes.c_int,                                   # n_ctrl, topo_sorted
            ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t,            # payload, text, n
            ctypes.POINTER(ctypes.c_uint64), ctypes.c_size_t]             # out pairs, cap
    return _lib


def _pack(lr: LoweredRegexp):
    payload = bytearray()
    rows: List[int] = []
    for e in list(lr.matching) + list(lr.control):
        off, ln, flags = len(payload), 0, 0
        if e.kind == K_MC:
            payload += bytes(e.node.chars)
            ln = len(e.node.chars)
        elif e.kind == K_BRACKET:
            n = e.node
            blob = bytes([len(n.singles) & 0xFF, len(n.singles) >> 8]) + bytes(n.singles)
            blob += bytes([len(n.ranges) & 0xFF, len(n.ranges) >> 8])
            for lo, hi in n.ranges:
                blob += bytes([lo, hi])
            payload += blob
            ln = len(blob)
            flags = 1 if n.neg else 0
        rows += [e.kind, e.entry, e.exit, off, ln, flags]
    arr = (ctypes.c_int32 * max(1, len(rows)))(*rows)
    return arr, bytes(payload)


MODE_ALL, MODE_FIRST, MODE_FULL, MODE_ANYWHERE = 0, 1, 2, 3
MODE_LONG_LITERAL_DEFECT = 0x100      # nfa_sim.c: literals compared the way the reference's emitted code does (B20)


class Oracle:
    """Compiled-once handle, mirrors rejit::Regej (include/rejit.h:105-138)."""

    def __init__(self, pattern, parser_opt: bool = True, long_literal_defect: bool = False):
        """long_literal_defect=True reproduces the reference's compare of literals longer than 16 bytes
        (nfa_sim.c mc_equal; defect B20): used only to pin this oracle against the compiled reference on
        such patterns.  The parity oracle (default) compares literals exactly."""
        self.lowered = lower(pattern, parser_opt)
        self._edges, self._payload = _pack(self.lowered)
        self._mode_bits = MODE_LONG_LITERAL_DEFECT if long_literal_defect else 0

    @property
    def longest_literal(self) -> int:
        """Bytes of the longest MultipleChar node (literals coalesce up to 64 bytes, src/regexp.h:107)."""
        return max([len(e.node.chars) for e in self.lowered.matching if e.kind == K_MC] or [0])

    def _run(self, mode: int, text: bytes, cap: int):
        lr = self.lowered
        out = (ctypes.c_uint64 * (2 * max(1, cap)))()
        r = _sim().nfa_sim_run(mode | self._mode_bits, lr.n_states, lr.entry_state, lr.exit_state,
                               self._edges, len(lr.matching), len(lr.control),
                               1 if lr.topo_sorted else 0, self._payload, text,
                               len(text), out, cap)
        return r, out

    def match_all(self, text: bytes) -> List[Tuple[int, int]]:
        cap = 1024
        while True:
            r, out = self._run(MODE_ALL, text, cap)
            if r <= cap:
                return [(out[2 * i], out[2 * i + 1]) for i in range(r)]
            cap = int(r)

    def match_all_count(self, text: bytes) -> int:
        r, _ = self._run(MODE_ALL, text, 0)
        return int(r)

    def match_first(self, text: bytes) -> Optional[Tuple[int, int]]:
        r, out = self._run(MODE_FIRST, text, 1)
        return (out[0], out[1]) if r else None

    def match_full(self, text: bytes) -> bool:
        r, _ = self._run(MODE_FULL, text, 0)
        return bool(r)

    def match_anywhere(self, text: bytes) -> bool:
        r, _ = self._run(MODE_ANYWHERE, text, 0)
        return bool(r)


def match_all(pattern, text: bytes, parser_opt: bool = True):
    return Oracle(pattern, parser_opt).match_all(text)
