/* include/rejit_b200.h — the C ABI of the B200-native matching engine.
 *
 * This is the drop-in boundary for rejit's MatchAll hot path.  In the
 * reference, `Regej::Compile` turns the lowered regexp (RegexpInfo,
 * /root/reference/src/regexp.h:538-636) into four JIT'd x64 functions
 *     bool     MatchFull    (const char*, size_t)
 *     bool     MatchAnywhere(const char*, size_t)
 *     bool     MatchFirst   (const char*, size_t, Match*)
 *     unsigned MatchAll     (const char*, size_t, std::vector<Match>*)
 * (/root/reference/src/regexp.h:533-536) that `Regej::Match*` call through
 * function pointers (/root/reference/src/rejit.cc:154,167,180,193).  The entry
 * points below replace exactly that pair — "compile the lowered regexp" and
 * "call the compiled matcher" — with an ahead-of-time lowering to tables for
 * hand-written sm_100a kernels.  Plain pointers and sizes only; offsets (not
 * pointers) cross the boundary, the C++ facade converts them back to
 * rejit::Match {begin,end}.  INTEGRATION.md shows the reference-side binding.
 *
 * There is no CPU fallback: every matching entry point fails with an error
 * when no CUDA device (or the CUDA part of the library) is available.
 *
 * Error convention: entry points that take (err, err_length) return -1 (or
 * NULL) and write a NUL-terminated message; a NULL program / set / text handle
 * is such an error, not a crash; no C++ exception leaves the library (an
 * allocation failure inside is reported the same way).
 */
#ifndef REJIT_B200_H_
#define REJIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the lowered regexp: a flat restatement of RegexpInfo ---------------
 * One rejit_b200_edge per "physical regexp" of re_matching_list_ /
 * re_control_list_ (src/regexp.h:586-600); `kind` follows the order of
 * LIST_PHYSICAL_REGEXP_TYPES (src/regexp.h:27-41).                          */
enum {
  REJIT_B200_EDGE_MULTIPLE_CHAR = 0, /* payload: the bytes                    */
  REJIT_B200_EDGE_PERIOD = 1,
  REJIT_B200_EDGE_BRACKET = 2,       /* payload: u16 n_single, singles, u16 n_range, (lo,hi)*; flags bit0 = non_matching */
  REJIT_B200_EDGE_START_OF_LINE = 3,
  REJIT_B200_EDGE_END_OF_LINE = 4,
  REJIT_B200_EDGE_EPSILON = 5
};

typedef struct rejit_b200_edge {
  int32_t kind;
  int32_t entry_state;
  int32_t exit_state;
  int32_t payload_offset;
  int32_t payload_length;
  int32_t flags;
} rejit_b200_edge;

typedef struct rejit_b200_ir {
  int32_t n_states;      /* RegexpInfo::last_state() + 1                      */
  int32_t entry_state;   /* RegexpInfo::entry_state()                         */
  int32_t exit_state;    /* RegexpInfo::exit_state()                          */
  int32_t n_matching;    /* edges[0 .. n_matching) = re_matching_list_        */
  int32_t n_control;     /* edges[n_matching .. +n_control) = re_control_list_ */
  const rejit_b200_edge* edges;
  const uint8_t* payload;
  size_t payload_length;
} rejit_b200_ir;

typedef struct rejit_b200_program rejit_b200_program;

/* Chain state carried across a slab boundary when one text is scanned in
 * pieces (multi-GPU sharding): `cur` = smallest offset at which the next match
 * may begin, `tail` = end of the last selected non-empty match (UINT64_MAX if
 * none).                                                                     */
typedef struct rejit_b200_carry {
  uint64_t cur;
  uint64_t tail;
} rejit_b200_carry;

typedef struct rejit_b200_stats {
  float scan_ms;        /* text-scanning kernel(s), CUDA events on the engine stream */
  float total_ms;       /* whole device pipeline                                  */
  uint32_t launches;    /* kernels launched by the call                           */
  uint32_t reruns;
  uint64_t candidates;
  uint64_t matches;
  int32_t strategy;     /* 0 literal, 1 fixed-length DFA, 2 literal+window, 3 generic */
  int32_t large_path;
} rejit_b200_stats;

/* ---- front end (host only; mirrors Parser + RegexpIndexer + RegexpLister) -- */
/* Parses an ERE and lowers it.  Returns 0 (RejitSuccess) or -1 (ParserError,
 * message — reference format — copied to err).  The IR must be released with
 * rejit_b200_ir_free.  parser_opt mirrors the reference's --use_parser_opt.    */
int rejit_b200_parse(const char* pattern, size_t pattern_length, int parser_opt,
                     rejit_b200_ir** out_ir, char* err, size_t err_length);
void rejit_b200_ir_free(rejit_b200_ir* ir);
/* Text dump of the lowered IR (tests diff it against the reference's
 * --print_re_list output).  Returns bytes needed (excluding NUL).             */
size_t rejit_b200_ir_dump(const rejit_b200_ir* ir, char* buffer, size_t buffer_length);

/* ---- compile: replaces Codegen::Compile (src/codegen.cc:591-656) ----------- */
/* Needs no GPU: builds the automaton tables; device copies are made lazily.    */
rejit_b200_program* rejit_b200_compile(const rejit_b200_ir* ir, char* err, size_t err_length);
void rejit_b200_program_free(rejit_b200_program* program);
/* One-line description of the chosen scan strategy.                            */
const char* rejit_b200_program_describe(const rejit_b200_program* program);
/* 1 when the pattern's text may be cut into slabs (the *_slab entry points, a
 * non-trivial carry_in, MatchAllParallel on several devices), 0 when it is
 * "re-entrant": a running thread can re-enter the pattern's start (e.g.
 * `.{,4}t`), which makes the reference drop a match that begins where the
 * previous one ended; reproducing that needs the thread labels of the whole
 * run, which a (cur, tail) carry cannot resume.  The slab entry points return
 * an error for such a pattern; rejit_b200_match_all_multi_gpu matches it on
 * one device and sets stats->large_path |= 2 to say so.                        */
int rejit_b200_program_is_shardable(const rejit_b200_program* program);

/* ---- the compiled matchers: replace the four JIT'd functions -------------- */
/* Host text in, results out.  MatchAll writes up to `capacity` (begin,end)
 * byte-offset pairs and returns the number of matches found (may exceed
 * capacity), or -1 on error (no device, CUDA failure; message in err).        */
int64_t rejit_b200_match_all(rejit_b200_program* program, const char* text, size_t text_length,
                             uint64_t* out_pairs, size_t capacity, char* err, size_t err_length);
/* As above but returns a malloc'ed array of pairs (free with rejit_b200_free). */
int64_t rejit_b200_match_all_alloc(rejit_b200_program* program, const char* text, size_t text_length,
                                   uint64_t** out_pairs, rejit_b200_stats* stats,
                                   char* err, size_t err_length);
/* 1 = found (out_pair filled), 0 = not found, -1 = error.                      */
int rejit_b200_match_first(rejit_b200_program* program, const char* text, size_t text_length,
                           uint64_t out_pair[2], char* err, size_t err_length);
int rejit_b200_match_full(rejit_b200_program* program, const char* text, size_t text_length,
                          char* err, size_t err_length);
int rejit_b200_match_anywhere(rejit_b200_program* program, const char* text, size_t text_length,
                              char* err, size_t err_length);
/* MatchAllParallel (new; named by BASELINE.json, absent from the reference):
 * the text is cut into n_gpus contiguous slabs, one per device.               */
int64_t rejit_b200_match_all_multi_gpu(rejit_b200_program* program, const char* text, size_t text_length,
                                       int n_gpus, uint64_t** out_pairs, rejit_b200_stats* stats,
                                       char* err, size_t err_length);

/* ---- device-resident text (benchmarks, pipelines that keep text in HBM) ---- */
int rejit_b200_device_count(void);
void* rejit_b200_device_alloc(int device, size_t bytes);          /* 16-byte aligned, padded */
void rejit_b200_device_free(int device, void* ptr);
void* rejit_b200_pinned_alloc(size_t bytes);
void rejit_b200_pinned_free(void* ptr);
int rejit_b200_copy_to_device(int device, void* dst, const void* src, size_t bytes);
int rejit_b200_copy_from_device(int device, void* dst, const void* src, size_t bytes);
void rejit_b200_flush_l2(int device);
/* d_text: device pointer (16-byte aligned) to text_length bytes; d_out_pairs:
 * device buffer for `capacity` pairs (may be NULL with capacity 0 to count).
 * carry_in / carry_out may be NULL (whole text in one call).
 * PADDING: the kernels read whole 16-byte groups, and the 16 bytes after the
 * group that holds the last text byte: the allocation behind d_text must be
 * readable up to ((text_length + 15) & ~15) + 16 bytes (their contents do not
 * matter).  rejit_b200_device_alloc and rejit_b200_text_upload add 64 bytes of
 * slack; a foreign pointer that ends at an allocation boundary must be padded
 * by the caller.  This holds for every entry point that takes a device text.   */
int64_t rejit_b200_match_all_device(rejit_b200_program* program, int device, const void* d_text,
                                    size_t text_length, uint64_t* d_out_pairs, size_t capacity,
                                    const rejit_b200_carry* carry_in, rejit_b200_carry* carry_out,
                                    rejit_b200_stats* stats, char* err, size_t err_length);

/* Upload-once text: one host-to-device copy, then any number of patterns are
 * matched against the resident copy (regex-dna counts nine patterns over one
 * sequence); each match call copies only its match list back.                 */
typedef struct rejit_b200_text rejit_b200_text;
rejit_b200_text* rejit_b200_text_upload(int device, const char* text, size_t text_length,
                                        char* err, size_t err_length);
/* The same from bytes that are already in device memory (a device-to-device copy
 * into a padded buffer the text owns): a pipeline that produces its input on the
 * device chains into ReplaceAll / set calls without a host round trip.            */
rejit_b200_text* rejit_b200_text_from_device(int device, const void* d_text, size_t text_length,
                                             char* err, size_t err_length);
void rejit_b200_text_free(rejit_b200_text* text);
int64_t rejit_b200_match_all_text(rejit_b200_program* program, const rejit_b200_text* text,
                                  uint64_t** out_pairs, rejit_b200_stats* stats,
                                  char* err, size_t err_length);

/* Regej::ReplaceAll (reference include/rejit.h:131, src/rejit.cc:221-226 and
 * Replace, src/rejit.cc:97-112) with the rebuild done on the device (SURVEY.md
 * §8f rank 2): every match is replaced by with[0..with_length).  Returns the
 * number of matches replaced; *out is malloc'ed (rejit_b200_free).              */
int64_t rejit_b200_replace_all(rejit_b200_program* program, const char* text, size_t text_length,
                               const char* with, size_t with_length, char** out, size_t* out_length,
                               rejit_b200_stats* stats, char* err, size_t err_length);
/* The same on an uploaded text; the result stays on the device as a new text, so
 * that substitutions chain without host round trips (regex-dna: the header /
 * newline strip followed by the eleven IUB substitutions, sample/regexdna.cc:49,69-85). */
rejit_b200_text* rejit_b200_replace_all_text(rejit_b200_program* program, const rejit_b200_text* text,
                                             const char* with, size_t with_length, int64_t* n_matches,
                                             rejit_b200_stats* stats, char* err, size_t err_length);
/* `count` ReplaceAll calls applied one after the other — programs[0] with
 * withs[0] first — as the eleven IUB substitutions of regex-dna are
 * (sample/regexdna.cc:69-85).  When every pattern matches exactly one byte and
 * no replacement holds a byte that a LATER pattern matches, the calls collapse
 * into one byte -> string table: one counting pass, one prefix sum, one writing
 * pass over the text instead of `count` scan + rebuild passes; any other set is
 * simply run call by call.  The result is identical either way.
 * n_matches[i] (may be NULL) = matches replaced by programs[i].                 */
rejit_b200_text* rejit_b200_replace_all_set_text(rejit_b200_program* const* programs, int count,
                                                 const rejit_b200_text* text, const char* const* withs,
                                                 const size_t* with_lengths, int64_t* n_matches,
                                                 rejit_b200_stats* stats, char* err, size_t err_length);
size_t rejit_b200_text_length(const rejit_b200_text* text);
/* Where the text lives on its device (readable for 64 bytes past its length); valid until the text is freed. */
const void* rejit_b200_text_device_ptr(const rejit_b200_text* text);
int rejit_b200_text_download(const rejit_b200_text* text, char* dst, size_t capacity, char* err, size_t err_length);

/* Pattern sets: several compiled patterns matched
 * against the same text.  When every member is a fixed-length, anchor-free
 * alternation (regex-dna's nine variants) they are fused into ONE automaton and
 * the text is scanned once for all of them; other sets are run member by member.
 * Results are identical to calling rejit_b200_match_all per member.
 * out_counts[j] = matches of member j; out_pairs (may be NULL) receives one
 * malloc'ed (begin,end) array per member (free each with rejit_b200_free).      */
typedef struct rejit_b200_set rejit_b200_set;
rejit_b200_set* rejit_b200_set_create(rejit_b200_program* const* programs, int count);
void rejit_b200_set_free(rejit_b200_set* set);
const char* rejit_b200_set_describe(const rejit_b200_set* set);
/* Inspection (tests): the k-mer index of a fused set whose members are at most 8
 * bytes long over at most four live byte values (the table k_set_kmer scans
 * with).  Returns 1 and fills info[0..13] = {shift, canon, canon_ok, n_members,
 * len_le[0..8], R}, bitmap[2^(2 (7 + R) - 5)] (at most 32768 words), mask16[65536]
 * (any of them may be NULL); 0 when the set has no such index.  R = ends answered
 * by one bitmap lookup.                                                           */
int rejit_b200_set_kmer_tables(const rejit_b200_set* set, uint32_t* info, uint32_t* bitmap, uint32_t* mask16);
int rejit_b200_match_all_set_text(rejit_b200_set* set, const rejit_b200_text* text, int64_t* out_counts,
                                  uint64_t** out_pairs, rejit_b200_stats* stats, char* err, size_t err_length);
int rejit_b200_match_all_set_device(rejit_b200_set* set, int device, const void* d_text, size_t text_length,
                                    int64_t* out_counts, rejit_b200_stats* stats, char* err, size_t err_length);

/* Slab variant of the set call (one-process-per-GPU sharding): ownership range,
 * base offset and one carry per member, as for rejit_b200_match_all_device_slab. */
int rejit_b200_match_all_set_device_slab(rejit_b200_set* set, int device, const void* d_text, size_t text_length,
                                         uint64_t own_begin, uint64_t own_end, uint64_t base_offset,
                                         const rejit_b200_carry* carry_in, rejit_b200_carry* carry_out,
                                         int64_t* out_counts, rejit_b200_stats* stats, char* err, size_t err_length);

/* Slab variant for one-process-per-GPU sharding: only matches that BEGIN in
 * [own_begin, own_end) of the buffer are reported (own_end > text_length means
 * "to the end, including the empty match at text_length"); the buffer should
 * extend past own_end by the pattern's longest match (right halo) and, for
 * patterns with ^, start one byte before own_begin.  base_offset is added to
 * every reported offset; the carries are in buffer coordinates.               */
int64_t rejit_b200_match_all_device_slab(rejit_b200_program* program, int device, const void* d_text,
                                         size_t text_length, uint64_t own_begin, uint64_t own_end,
                                         uint64_t base_offset, uint64_t* d_out_pairs, size_t capacity,
                                         const rejit_b200_carry* carry_in, rejit_b200_carry* carry_out,
                                         rejit_b200_stats* stats, char* err, size_t err_length);

/* ---- device-side stitch (one process per GPU) ----------------
 * The "tiny allgather to stitch boundary matches" as peer-to-peer stores over
 * NVLink: every rank owns an inbox in device memory; its CUDA IPC handle (64
 * bytes) goes to the neighbouring ranks by whatever channel the job has (an
 * all-gather of the handles at start-up), and rejit_b200_stitch_connect maps
 * the neighbours' inboxes (NULL at the ends of the chain).  After a slab call
 * rejit_b200_stitch_exchange sends the chain states that leave the slab
 * (`leaving[j]`, the slab call's carry_out in GLOBAL offsets) into the right
 * neighbour's device memory, waits for the states arriving from the left
 * (`arrived[j]`) and sets bit j of *redo_mask when pattern j's arriving chain
 * reaches into this slab (first owned start = slab_begin, global): only then
 * must the slab call be repeated with carry_in = arrived.  One warp on the
 * engine's stream; no host-to-host hop, no collective.  Every rank must call it
 * the same number of times.  Returns 0, or -1 (a neighbour did not answer).     */
int rejit_b200_stitch_open(int device, int rank, int world, void* handle_out /* 64 bytes */, char* err, size_t err_length);
int rejit_b200_stitch_connect(int device, const void* left_handle, const void* right_handle, char* err, size_t err_length);
void rejit_b200_stitch_close(int device);
int rejit_b200_stitch_exchange(int device, int count, const rejit_b200_carry* leaving, uint64_t slab_begin,
                               rejit_b200_carry* arrived, uint32_t* redo_mask, char* err, size_t err_length);

/* The slab call of a pattern set and the exchange in ONE step: as
 * rejit_b200_match_all_set_device_slab with "nothing arrives from the left",
 * followed by rejit_b200_stitch_exchange — but when the fused k-mer scan runs, the
 * CTA that reports the result sends and receives the chain states itself, inside
 * the scan kernel (no second launch, no host hop in between; the step's device time
 * includes the neighbour's answer).  carry_out: buffer coordinates, as for the slab
 * call; arrived / redo_mask: as for rejit_b200_stitch_exchange.                       */
int rejit_b200_match_all_set_device_stitched(rejit_b200_set* set, int device, const void* d_text, size_t text_length,
                                             uint64_t own_begin, uint64_t own_end, uint64_t base_offset,
                                             rejit_b200_carry* carry_out, rejit_b200_carry* arrived, uint32_t* redo_mask,
                                             int64_t* out_counts, rejit_b200_stats* stats, char* err, size_t err_length);

void rejit_b200_free(void* ptr);

#ifdef __cplusplus
}
#endif

#endif /* REJIT_B200_H_ */
