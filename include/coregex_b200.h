/* coregex_b200.h — C ABI of the B200 bulk-scan engine (libcoregex_b200.so).
 *
 * The reference (coregx/coregex) is a pure-Go package with no FFI boundary of its own; the
 * boundary this library replaces is the internal seam every public Regex method funnels into
 * (SURVEY.md §8b).  Each entry point below names the reference interface it stands in for; the
 * cgo binding a coregex maintainer would add is shown in INTEGRATION.md and go/coregex/.
 *
 * Conventions (modelled on the reference's Go->asm seam, e.g. simd/memchr_amd64.go:26-35,
 * prefilter/teddy_ssse3_amd64.go:38): pointer + length in, scalars / caller-owned buffers out,
 * no callbacks, haystacks are borrowed read-only and never retained.  Only cgx_compile returns
 * an error text (the reference's search methods cannot fail either); search calls return
 * CGX_OK or a negative status for resource/driver problems.  There is NO CPU fallback: without
 * a CUDA device every search call returns CGX_ERR_NO_DEVICE.
 *
 * Thread-safety: a cgx_regex may be shared by threads; each call serialises on the regex's own
 * device scratch (the reference pools per-goroutine SearchState, meta/engine.go:258-296).
 */
#ifndef COREGEX_B200_H
#define COREGEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgx_regex cgx_regex;

enum {
  CGX_OK = 0,
  CGX_ERR_SYNTAX = -1,      /* pattern rejected; errbuf holds the Go-formatted message        */
  CGX_ERR_UNSUPPORTED = -2, /* valid pattern outside the GPU engines' scope; errbuf says why  */
  CGX_ERR_NO_DEVICE = -3,   /* no usable CUDA device / driver                                  */
  CGX_ERR_CUDA = -4,        /* a CUDA call failed; cgx_last_error() has the text               */
  CGX_ERR_ARGS = -5,        /* bad arguments (null pointer, misaligned device pointer, ...)    */
  CGX_ERR_NOMEM = -6,
  CGX_ERR_CONFIG = -7       /* cgx_config out of range; errbuf holds "regexp: invalid config: ..." */
};

/* ---- compile -------------------------------------------------------------------------------
 * replaces coregex.Compile / meta.Compile (reference regex.go:110, meta/compile.go:40-60).
 * On failure *out is NULL and errbuf receives "error parsing regexp: ..." exactly as
 * syntax.Error formats it (reference meta/compile.go:775-784), or an "unsupported: ..." text. */
int cgx_compile(const char* pattern, size_t pattern_len, cgx_regex** out, char* errbuf, size_t errcap);
void cgx_free(cgx_regex* re);

/* replaces coregex.CompileWithConfig / meta.CompileWithConfig (reference regex.go:198,
 * meta/compile.go:62) with meta.Config (meta/config.go:31-113), field for field.
 * cgx_default_config = meta.DefaultConfig (config.go:101-112); cgx_config_validate = Config.Validate
 * (config.go:132-170, called from meta/compile.go:53): same ranges, same message ("regexp: invalid config: <Field>: <Message>",
 * config.go:179-181) in errbuf, status CGX_ERR_CONFIG.  What the fields do here: enable_dfa,
 * enable_prefilter and min_literal_len steer strategy selection exactly where the reference reads
 * them (meta/strategy.go:515,963,976,1265,1392,1447; meta/compile.go:466) — cgx_strategy reports the
 * result, and the digit-prefilter and multi-literal engines are only used when the reference would
 * use theirs.  max_dfa_states, determinization_limit, max_literals, max_recursion_depth and
 * enable_ascii_optimization size the reference's lazy-DFA cache and NFA compiler; they are
 * validated and have no effect on results (the GPU tables are built eagerly, host-side).          */
typedef struct cgx_config {
  int enable_dfa;                /* EnableDFA               default 1     */
  int enable_prefilter;          /* EnablePrefilter         default 1     */
  uint32_t max_dfa_states;       /* MaxDFAStates            default 10000 */
  int determinization_limit;     /* DeterminizationLimit    default 1000  */
  int min_literal_len;           /* MinLiteralLen           default 1     */
  int max_literals;              /* MaxLiterals             default 256   */
  int max_recursion_depth;       /* MaxRecursionDepth       default 100   */
  int enable_ascii_optimization; /* EnableASCIIOptimization default 1     */
} cgx_config;
void cgx_default_config(cgx_config* cfg);
int cgx_config_validate(const cgx_config* cfg, char* errbuf, size_t errcap);
int cgx_compile_cfg(const char* pattern, size_t pattern_len, const cgx_config* cfg /* NULL = default */,
                    cgx_regex** out, char* errbuf, size_t errcap);

/* replaces Regex.Longest / meta.Engine.SetLongest (reference regex.go:464, meta/engine.go:250):
 * leftmost-longest (POSIX) instead of leftmost-first matching for all later searches.  Like the
 * reference's, this call must not run concurrently with searches.  Patterns whose engine cannot
 * switch return CGX_ERR_UNSUPPORTED (cgx_last_error says why) and keep leftmost-first.           */
int cgx_set_longest(cgx_regex* re, int longest);

/* reference meta.Engine.Strategy() (meta/engine.go:191): the strategy name the reference would
 * pick for this pattern, e.g. "UseDigitPrefilter".                                            */
const char* cgx_strategy(const cgx_regex* re);
/* which GPU engine runs it: "dfa-runstart", "dfa-byteset", "dfa-lut" (+"+flat": bit-parallel
 * start filter in front of the DFA walk; +"+bitstream": flat deterministic pattern, starts AND
 * ends bit-parallel, scan_bits.cu), "line-dfa", "teddy", "fat-teddy"                            */
const char* cgx_engine(const cgx_regex* re);
/* the record delimiter of this pattern: a byte no match can contain ('\n' whenever the pattern
 * allows it).  The scan treats the haystack as records separated by it; callers that split a
 * corpus into shards (cgx_scan_shard_device) must cut right after this byte.                   */
int cgx_delimiter(const cgx_regex* re);
/* reference meta.Engine.NumCaptures() (meta/engine.go:232): groups including group 0          */
int cgx_num_captures(const cgx_regex* re);
/* reference meta.Engine.SubexpNames() (regex.go:575): name of group i ("" for group 0, unnamed groups
 * and i out of range); the pointer lives as long as the regex                                    */
const char* cgx_subexp_name(const cgx_regex* re, int i);
const char* cgx_last_error(void);

/* ---- host-buffer searches (haystack in host memory; H2D/D2H inside the call) ----------------
 * cgx_is_match            replaces meta.Engine.IsMatch (meta/ismatch.go:27) = Regex.Match
 * cgx_find_all_index      replaces meta.Engine.FindAllIndicesStreaming (meta/findall.go:155)
 *                         = Regex.FindAllIndex / AppendAllIndex (regex.go:695,748).
 *                         limit<0: all matches, limit==0: none (regex.go:696-698).
 *                         Writes up to cap_pairs (start,end) int64 pairs in match order and stores
 *                         the TOTAL number of matches in *count (two-call sizing: call with
 *                         cap_pairs=0 to size, or grow and retry when *count > cap_pairs).
 * cgx_count               replaces meta.Engine.Count (meta/findall.go:297)
 * cgx_find_all_submatch_index  replaces meta.Engine.FindAllSubmatch (meta/findall.go:390)
 *                         = Regex.FindAllSubmatchIndex (regex.go:1423); stride is
 *                         2*cgx_num_captures ints per match, unmatched groups are -1,-1.     */
int cgx_is_match(cgx_regex* re, const uint8_t* haystack, size_t len, int* matched);
int cgx_find_all_index(cgx_regex* re, const uint8_t* haystack, size_t len, int64_t limit,
                       int64_t* out_pairs, size_t cap_pairs, size_t* count);
int cgx_count(cgx_regex* re, const uint8_t* haystack, size_t len, int64_t limit, size_t* count);
int cgx_find_all_submatch_index(cgx_regex* re, const uint8_t* haystack, size_t len, int64_t limit,
                                int64_t* out, size_t cap_matches, size_t* count);

/* ---- device-resident batch entry (what a GPU-side caller wants; no host copies) --------------
 * d_haystack: device pointer, 16-byte aligned.  base_offset is added to every reported offset
 * (shard base when a corpus is split across GPUs).  d_out_pairs: device int64 pairs, 16-byte
 * aligned, capacity cap_pairs (may be 0/NULL for mode COUNT / ISMATCH).  d_result: device
 * uint64[2] = {total matches, is-match flag}, written when the scan completes on `stream`
 * (a cudaStream_t passed as void*; NULL = default stream).  The call only enqueues work.      */
enum { CGX_MODE_FINDALL = 0, CGX_MODE_COUNT = 1, CGX_MODE_ISMATCH = 2 };
int cgx_scan_device(cgx_regex* re, const uint8_t* d_haystack, size_t len, int64_t base_offset,
                    int mode, int64_t* d_out_pairs, size_t cap_pairs, uint64_t* d_result,
                    void* stream);
/* One shard of a larger logical haystack (corpus split across GPUs, or pieces of a pipelined
 * host copy).  Shards are cut right after a record delimiter (cgx_delimiter, normally '\n'): base_offset > 0 promises
 * that the byte before d_haystack[0] is a delimiter, bytes_after > 0 that the shard ends with
 * one.  base_offset == 0 marks the true start of text (\A, non-multiline ^), bytes_after == 0
 * the true end (\z, $), and bytes_after also feeds the multi-literal engine's end-of-haystack
 * verify regime (reference prefilter/teddy.go:415-428), so that shard results concatenate to
 * exactly the whole-haystack result.  cgx_scan_device is the bytes_after == 0 case.            */
int cgx_scan_shard_device(cgx_regex* re, const uint8_t* d_haystack, size_t len, int64_t base_offset,
                          int64_t bytes_after, int mode, int64_t* d_out_pairs, size_t cap_pairs,
                          uint64_t* d_result, void* stream);
/* Batch of independent records — what the reference does with one FindAllIndex call per haystack
 * (regex.go:695) — in one scan.  Record r is d_haystack[d_rec_off[r] .. d_rec_off[r+1]): nrec + 1
 * ascending device offsets, d_rec_off[0] = 0, d_rec_off[nrec] = len.  Every record but the last
 * must END with the record delimiter (cgx_delimiter, normally '\n'), so that no match spans two
 * records, and the pattern must be free of anchors and look-around (else CGX_ERR_UNSUPPORTED): then
 * the batch result is exactly the per-record results, concatenated.  Pairs are written in global
 * order with offsets relative to d_haystack (+ base_offset); d_rec_prefix[r] (nrec + 1 entries) =
 * number of written pairs that start before record r, so record r owns pairs
 * [d_rec_prefix[r], d_rec_prefix[r+1]) and record-relative offsets are pair - d_rec_off[r].
 * d_result: device uint64[3] = {total matches, is-match flag, number of inner record boundaries
 * that are NOT preceded by the delimiter (the caller must see 0)}.                              */
int cgx_scan_records_device(cgx_regex* re, const uint8_t* d_haystack, size_t len,
                            const uint64_t* d_rec_off, size_t nrec, int64_t base_offset,
                            int64_t* d_out_pairs, size_t cap_pairs, uint64_t* d_rec_prefix,
                            uint64_t* d_result, void* stream);
/* submatch variant: d_out receives stride int64 per match */
int cgx_scan_submatch_device(cgx_regex* re, const uint8_t* d_haystack, size_t len,
                             int64_t base_offset, int64_t* d_out, size_t cap_matches,
                             uint64_t* d_result, void* stream);
/* ... of a shard that bytes_after more bytes of the logical haystack follow (as cgx_scan_shard_device):
 * `\z`, the empty record behind a trailing delimiter and the reference's treatment of a match AT the
 * end of the haystack (groups unset, nfa/pikevm.go:2201-2206) apply to the last shard only        */
int cgx_scan_submatch_shard_device(cgx_regex* re, const uint8_t* d_haystack, size_t len,
                                   int64_t base_offset, int64_t bytes_after, int64_t* d_out,
                                   size_t cap_matches, uint64_t* d_result, void* stream);

/* ---- compact offset wire format (multi-GPU offset gather, SURVEY.md §8e / BASELINE config 5) ----
 * The reference returns [][2]int (16 B per match, regex.go:710-723); between GPUs a shard's sorted
 * pairs travel as 6 B per match: the low 32 bits of the shard-relative start + a 16-bit length,
 * behind a table of the first match index of every 4 GiB segment of the shard.  One message =
 * seg_first[nseg] u64 | lo[count] u32 | len[count] u16, cgx_wire_bytes(count, nseg) bytes with
 * nseg = cgx_wire_segments(shard_len).  d_bad (device uint64) counts matches that do not fit
 * (longer than 65535 bytes): the caller then sends plain int64 pairs instead.  Both calls only
 * enqueue work on `stream`; d_wire 8-byte aligned, pair buffers 16-byte aligned.               */
size_t cgx_wire_bytes(size_t count, int nseg);
int cgx_wire_segments(size_t shard_len);
int cgx_pack_offsets_device(const int64_t* d_pairs, size_t count, int64_t shard_base, size_t shard_len,
                            uint8_t* d_wire, uint64_t* d_bad, void* stream);
int cgx_unpack_offsets_device(const uint8_t* d_wire, size_t count, int64_t shard_base, size_t shard_len,
                              int64_t* d_pairs_out, void* stream);

/* number of kernels launched by this regex since creation (bench.py reports it) */
uint64_t cgx_launch_count(const cgx_regex* re);

/* ---- synthetic corpora (bench/test utility; deterministic from seed) --------------------------
 * kind 0: access-log lines   kind 1: text with planted literals (C3/C5)
 * kind 2: 80-byte e-mail lines (C4).  Fills d_out[0..len) on the device; every block of
 * `block` bytes ends with '\n' so shards can be generated independently.                        */
int cgx_synth_device(int kind, uint64_t seed, uint64_t first_block, uint8_t* d_out, size_t len,
                     const uint8_t* d_literals, const int32_t* d_lit_offsets, int nlit, void* stream);
int cgx_synth_host(int kind, uint64_t seed, uint64_t first_block, uint8_t* out, size_t len,
                   const uint8_t* literals, const int32_t* lit_offsets, int nlit);

/* ---- the byte-search family of the reference's prefilter layer, device resident ------------------
 * "The first position whose byte is in a set" — reference simd/memchr_amd64.go:67 Memchr, :114
 * Memchr2, :159 Memchr3, simd/memchr_digit_amd64.go:17 MemchrDigit (:34 MemchrDigitAt = the same on
 * the tail of the buffer), simd/memchr_class_amd64.go:35 MemchrWord, :58 MemchrNotWord, :76
 * MemchrInTable, :90 MemchrNotInTable all are cgx_memchr_table_device with the right 256-entry
 * table (non-zero = in the set); :202 MemchrPair and simd/memmem.go:53 Memmem have their own entry.
 * d_h: device haystack, 16-byte aligned.  d_result: device int64[2]; [0] receives the index or -1
 * (the reference's return value), [1] is scratch.  The scan stops shortly after the first hit.
 * Needles of cgx_memmem_device are at most 256 bytes long (CGX_ERR_UNSUPPORTED beyond).           */
int cgx_memchr_table_device(const uint8_t* d_h, size_t n, const uint8_t* table256, int64_t* d_result, void* stream);
/* the same from position `at` on (absolute index back): MemchrDigitAt, memchr_digit_amd64.go:34 */
int cgx_memchr_table_at_device(const uint8_t* d_h, size_t n, size_t at, const uint8_t* table256, int64_t* d_result,
                               void* stream);
int cgx_memchr_pair_device(const uint8_t* d_h, size_t n, uint8_t byte1, uint8_t byte2, int64_t offset,
                           int64_t* d_result, void* stream);
int cgx_memmem_device(const uint8_t* d_h, size_t n, const uint8_t* needle, size_t m, int64_t* d_result, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COREGEX_B200_H */
