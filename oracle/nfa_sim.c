/* oracle/nfa_sim.c — TEST INFRASTRUCTURE (parity checker back end), never shipped.
 *
 * Sequential C restatement of the machine code the reference JIT emits when
 * fast-forward is disabled (the parity configuration).  One
 * function, four modes (MatchAll / MatchFirst / MatchFull / MatchAnywhere).
 *
 * Follows, step for step:
 *   per-byte loop ........ /root/reference/src/x64/codegen-x64.cc:535-640
 *                          (Codegen::GenerateMatchDirection, kForward)
 *   time-flow checks ..... codegen-x64.cc:247-315 (CheckTimeFlow, !fast_forward_)
 *   control regexps ...... codegen-x64.cc:366-398, 680-732 (eps, ^, $)
 *   match bookkeeping .... codegen-x64.cc:401-466 (CheckMatch), :469-522 (RegisterMatch)
 *   transitions .......... codegen-x64.cc:653-677, 735-933 (MC / '.' / bracket)
 *   state ring ........... codegen-x64.cc:951-1017 (SetState: the OLDER start
 *                          wins), :1042-1097 (ClearTime, ClearStates)
 *   match list filter .... /root/reference/src/codegen.cc:36-86 (MatchAllAppendFilter)
 *
 * The ring stores, per (time slot, NFA state), the offset+1 of the text
 * position where the thread now in that state started (0 = no thread); the
 * reference stores raw pointers, for which 0 is likewise "empty".
 *
 * The front end (parser / indexer / lister) lives in oracle/rejit_oracle.py and
 * passes the lowered regexp as flat arrays.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { K_MC = 0, K_PERIOD, K_BRACKET, K_SOL, K_EOL, K_EPS };
enum { MODE_ALL = 0, MODE_FIRST, MODE_FULL, MODE_ANYWHERE };
/* OR-ed into `mode`: compare literals the way the reference's emitted code does, defect included (below). */
#define MODE_LONG_LITERAL_DEFECT 0x100

#define RING_TIMES 66 /* 1 + kMaxNodeLength (src/codegen.cc:617) + spare */

typedef struct {
  int32_t kind, entry, exit, off, len, flags;
} edge_t;

typedef struct {
  uint64_t *pairs;
  size_t cap;
  size_t count;
  /* the reference keeps whole matches in a std::vector; we keep everything in
   * a growable shadow vector so that erase-from-the-back works past `cap`. */
  uint64_t *vec;
  size_t vcap;
} matches_t;

static void vec_push(matches_t *m, uint64_t b, uint64_t e) {
  if (m->count == m->vcap) {
    m->vcap = m->vcap ? 2 * m->vcap : 256;
    m->vec = (uint64_t *)realloc(m->vec, 2 * m->vcap * sizeof(uint64_t));
  }
  m->vec[2 * m->count] = b;
  m->vec[2 * m->count + 1] = e;
  m->count++;
}

/* MultipleChar against the text window w (both `len` bytes).
 *
 * exact = 1: byte equality, what the node means.  This is what the parity
 * oracle uses and what the product implements.
 *
 * exact = 0: what the code emitted by MatchMultipleChar
 * (/root/reference/src/x64/codegen-x64.cc:757-837) accepts.  For len > 8 it
 * pre-checks the first 8 bytes, then runs a repeated quadword compare over
 * len / 8 quadwords and, without looking at its flags, a repeated byte
 * compare over len % 8 bytes FROM WHERE THE QUADWORD COMPARE STOPPED; only
 * the last compare's flags are tested (:823-833).  So for len > 16 with
 * len % 8 != 0 a window that differs from the literal in quadword q >= 1 is
 * accepted whenever the len % 8 bytes after that quadword are equal
 * (defect B20; tests/test_oracle.py pins this model
 * against the compiled reference).  Lengths <= 16 and multiples of 8 are
 * compared correctly. */
static int mc_equal(const unsigned char *w, const unsigned char *lit, int len, int exact) {
  if (exact || len <= 16 || len % 8 == 0) return memcmp(w, lit, (size_t)len) == 0;
  if (memcmp(w, lit, 8) != 0) return 0;
  int at = 8 * (len / 8);
  for (int q = 0; q < len / 8; q++)
    if (memcmp(w + 8 * q, lit + 8 * q, 8) != 0) { at = 8 * (q + 1); break; }
  return memcmp(w + at, lit + at, (size_t)(len % 8)) == 0;
}

/* MatchAllAppend(filter=true), src/codegen.cc:36-76 */
static void append_filter(matches_t *m, uint64_t b, uint64_t e) {
  while (m->count && m->vec[2 * (m->count - 1)] >= b) m->count--;
  if (b == e && m->count && m->vec[2 * (m->count - 1) + 1] == b) return;
  vec_push(m, b, e);
}

static int bracket_hit(const unsigned char *pl, const edge_t *ed, unsigned char ch) {
  const unsigned char *p = pl + ed->off;
  int ns = p[0] | (p[1] << 8);
  p += 2;
  int hit = 0;
  for (int i = 0; i < ns; i++)
    if (p[i] == ch) hit = 1;
  p += ns;
  int nr = p[0] | (p[1] << 8);
  p += 2;
  /* signed-char range compare: cmpb + setcc(greater_equal / less_equal),
   * codegen-x64.cc:898-908 */
  signed char c = (signed char)ch;
  for (int i = 0; i < nr; i++)
    if (c >= (signed char)p[2 * i] && c <= (signed char)p[2 * i + 1]) hit = 1;
  return (ed->flags & 1) ? !hit : hit;
}

int64_t nfa_sim_run(int mode, int n_states, int entry_state, int exit_state,
                    const int32_t *edges_raw, int n_match, int n_ctrl,
                    int topo_sorted, const char *payload_c, const char *text_c,
                    size_t n, uint64_t *out_pairs, size_t cap) {
  const edge_t *medges = (const edge_t *)edges_raw;
  const edge_t *cedges = medges + n_match;
  const unsigned char *pl = (const unsigned char *)payload_c;
  const unsigned char *text = (const unsigned char *)text_c;
  const int exact_literals = !(mode & MODE_LONG_LITERAL_DEFECT);
  mode &= 0xFF;

  uint64_t *ring = (uint64_t *)calloc((size_t)RING_TIMES * n_states, sizeof(uint64_t));
  unsigned char summary[RING_TIMES]; /* time summary bits, codegen-x64.cc:210-245 */
  memset(summary, 0, sizeof(summary));
  int base = 0; /* ring_index: slot holding "time 0" */
  matches_t M = {out_pairs, cap, 0, NULL, 0};
  uint64_t forward_match = 0, backward_match = 0; /* offset+1; 0 = none */
  int64_t result = 0;
  int have_first = 0;
  uint64_t first_b = 0, first_e = 0;

#define SLOT(t) (ring + (size_t)(((base + (t)) % RING_TIMES)) * n_states)
  /* SetState: update only when the source thread is strictly older,
   * codegen-x64.cc:951-987 (dec + unsigned compare makes 0 the "youngest"). */
#define SET_STATE(t, tgt, srcv)                                  \
  do {                                                           \
    uint64_t *_c = &SLOT(t)[tgt];                                \
    if ((uint64_t)((srcv)-1) < (uint64_t)(*_c - 1)) {            \
      *_c = (srcv);                                              \
      summary[t] = 1;                                            \
    }                                                            \
  } while (0)

  if (mode == MODE_FULL) { /* codegen-x64.cc:162-165 */
    SLOT(0)[entry_state] = 0 + 1;
    summary[0] = 1;
  }

  size_t p = 0;
  for (;;) {
    uint64_t *cur = SLOT(0);
    /* ---- CheckTimeFlow, !fast_forward_ ---- */
    int flowing = 0;
    for (int t = 0; t < RING_TIMES; t++) flowing |= summary[t];
    if (mode == MODE_FULL) {
      if (!flowing) { result = 0; goto done; }
    } else if (forward_match) {
      if (mode == MODE_ANYWHERE) { result = 1; goto done; }
      if (mode == MODE_FIRST) {
        if (!flowing) goto limit;
      } else { /* MODE_ALL: RegisterMatch */
        append_filter(&M, backward_match - 1, forward_match - 1);
        if (forward_match - 1 == n) goto done; /* codegen-x64.cc:506-508 */
        forward_match = backward_match = 0;
      }
    }
    /* ---- new thread at this position, codegen-x64.cc:544-554 ---- */
    if (mode != MODE_FULL) {
      cur[entry_state] = p + 1;
      summary[0] = 1;
    }
    /* ---- HandleControlRegexps ---- */
    {
      int rounds = topo_sorted ? 1 : (n_ctrl > 0 ? n_ctrl : 1);
      for (int r = 0; r < rounds; r++)
        for (int i = 0; i < n_ctrl; i++) {
          const edge_t *ed = &cedges[i];
          uint64_t sv = cur[ed->entry];
          int ok;
          if (ed->kind == K_EPS) ok = 1;
          else if (ed->kind == K_SOL)
            ok = (p == 0) || text[p - 1] == '\n' || text[p - 1] == '\r';
          else /* K_EOL */
            ok = (p == n) || text[p] == '\n' || text[p] == '\r';
          if (ok) SET_STATE(0, ed->exit, sv);
        }
    }
    /* ---- CheckMatch (forward) ---- */
    if (mode != MODE_FULL && cur[exit_state]) {
      if (mode == MODE_ANYWHERE) { result = 1; goto done; }
      forward_match = p + 1;
      backward_match = cur[exit_state];
      /* ClearStates(begin[, end]), codegen-x64.cc:1075-1097 */
      for (size_t i = 0; i < (size_t)RING_TIMES * n_states; i++) {
        uint64_t v = ring[i];
        if (backward_match >= v) continue;
        if (mode == MODE_ALL && forward_match <= v) continue;
        ring[i] = 0;
      }
    }
    if (p == n) goto limit;
    /* ---- GenerateTransitions ---- */
    for (int i = 0; i < n_match; i++) {
      const edge_t *ed = &medges[i];
      uint64_t sv = cur[ed->entry];
      if (!sv) continue;
      unsigned char ch = text[p];
      if (ed->kind == K_MC) {
        if (p + (size_t)ed->len <= n && mc_equal(text + p, pl + ed->off, ed->len, exact_literals))
          SET_STATE(ed->len, ed->exit, sv);
      } else if (ed->kind == K_PERIOD) {
        if (ch != '\n' && ch != '\r') SET_STATE(1, ed->exit, sv);
      } else { /* K_BRACKET */
        if (bracket_hit(pl, ed, ch)) SET_STATE(1, ed->exit, sv);
      }
    }
    /* ---- ClearTime(0); advance; FlowTime ---- */
    memset(cur, 0, sizeof(uint64_t) * n_states);
    base = (base + 1) % RING_TIMES;
    memmove(summary, summary + 1, RING_TIMES - 1);
    summary[RING_TIMES - 1] = 0;
    p++;
  }

limit:
  if (mode == MODE_FULL) {
    result = SLOT(0)[exit_state] != 0; /* codegen-x64.cc:586-590 */
  } else if (forward_match) {
    if (mode == MODE_FIRST) {
      have_first = 1;
      first_b = backward_match - 1;
      first_e = forward_match - 1;
      result = 1;
    } else { /* MODE_ALL */
      append_filter(&M, backward_match - 1, forward_match - 1);
    }
  }

done:
  if (mode == MODE_ALL) {
    result = (int64_t)M.count;
    size_t k = M.count < cap ? M.count : cap;
    if (k) memcpy(out_pairs, M.vec, 2 * k * sizeof(uint64_t));
  } else if (mode == MODE_FIRST && have_first && cap >= 1) {
    out_pairs[0] = first_b;
    out_pairs[1] = first_e;
  }
  free(M.vec);
  free(ring);
  return result;
#undef SLOT
#undef SET_STATE
}
