/*
 * lc_b200.h -- C-ABI of the B200-native log-parsing engine (libloongcollector_b200.so).
 *
 * POD only: plain pointers and sizes, int return codes, no exceptions, no C++/torch types.
 * These are the entry points a LoongCollector build binds to replace the CPU arithmetic inside
 *   Processor::Process(PipelineEventGroup&)      core/collection_pipeline/plugin/interface/Processor.h:27-37
 * for the four native processors on the hot path; each function cites the reference code whose
 * RESULT it reproduces bit-exactly.  The reference-side binding is shown in INTEGRATION.md.
 *
 * Conventions
 *  - "base" is the contiguous SourceBuffer allocation holding the group's bytes
 *    (core/common/memory/SourceBuffer.h:98-131,156-181); every offset is u32 relative to base.
 *  - Host-pointer entry points (no suffix) validate the event table against base_len (LC_ERR_INVALID_ARG), copy
 *    base + the event table to the GPU, run the kernels and copy the results back before returning.
 *    *_dev entry points take DEVICE pointers (arena already resident in HBM) and queue their kernels on the
 *    engine's current stream -- lc_engine_stream(), replaceable with lc_engine_set_stream().  The regex / delimiter
 *    *_dev calls never wait for the device (results are valid once the stream has drained: lc_engine_sync or the
 *    caller's own event); the split / multiline *_dev calls return a count and therefore synchronise.  *_dev
 *    callers guarantee ev_off[i] + ev_len[i] <= base_len (device tables are not re-read on the host).
 *  - One lc_engine per (GPU, host thread): mirrors the reference's per-thread regex copies
 *    (ProcessorParseRegexNative.cpp:64-67,255-257).  An engine is not thread-safe; regexes are immutable
 *    after compilation and may be shared between engines.
 *  - There is NO CPU fallback: if no CUDA device is usable every compute call fails with LC_ERR_CUDA.
 */
#ifndef LC_B200_H
#define LC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LC_OK 0
#define LC_ERR_INVALID_ARG 1
#define LC_ERR_CUDA 2          /* no device / CUDA runtime failure (message in lc_last_error) */
#define LC_ERR_REGEX_INVALID 3 /* pattern does not parse (reference: Init returns false, ParamExtractor.cpp:199-209) */
#define LC_ERR_REGEX_UNSUPPORTED 4 /* valid for boost but outside the automaton subset (back-refs, look-behind, multi-byte look-ahead...) */
#define LC_ERR_CAPACITY 5      /* caller-provided output capacity too small; *n_out holds the needed count */
#define LC_ERR_TOO_LARGE 6     /* buffer >= 4 GiB or >= 2^30 lines in one call */

/* per-event status of lc_regex_parse (ProcessorParseRegexNative.cpp:186-253) */
#define LC_REGEX_OK 0
#define LC_REGEX_NOMATCH 1       /* regex_match false            -> out_failed++ (:225) */
#define LC_REGEX_KEYS_MISMATCH 2 /* what.size() <= keys.size()   -> fail, out_failed NOT incremented (:227-244) */

/* per-event status of lc_delim_parse (ProcessorParseDelimiterNative.cpp:206-364) */
#define LC_DELIM_OK 0
#define LC_DELIM_PARSE_FAIL 1 /* FSM error (DelimiterModeFsmParser.cpp:260-294) or SplitString false */
#define LC_DELIM_BLANK 2      /* empty / all-blank value: out_failed++, event left untouched (:220-242) */
#define LC_DELIM_COLUMNS 3    /* column-count rule failed (:285) */

/* flag bits of lc_multiline_split output events */
#define LC_ML_IS_LAST 1u /* isLastLog flag the reference passes to CreateNewEvent (rawSize rule, :329-332) */
#define LC_ML_MATCHED 2u /* emitted as a matched record (matched_events++), else an unmatched single line */

typedef struct lc_engine lc_engine_t;
typedef struct lc_regex lc_regex_t;

/* ---- library / engine ---------------------------------------------------------------------- */
const char* lc_version(void);
/* Thread-local message of the last failing call on this thread. */
const char* lc_last_error(void);
/* Number of visible CUDA devices (0 if none / runtime unusable). */
int lc_device_count(void);
/* Creates an engine bound to CUDA device `device` with its own stream, workspace and pinned staging. */
int lc_engine_create(int device, lc_engine_t** out);
void lc_engine_destroy(lc_engine_t* e);
/* Blocks until all work queued on the engine's stream is complete. */
int lc_engine_sync(lc_engine_t* e);
/* CUDA stream (cudaStream_t) the engine queues its work on, for callers that enqueue their own work. */
void* lc_engine_stream(lc_engine_t* e);
/* Makes the engine queue on the caller's stream (cudaStream_t; NULL = back to the engine's own).  Drains the
 * previous stream first: the engine's workspace is re-used from call to call and ordered by the stream. */
int lc_engine_set_stream(lc_engine_t* e, void* stream);
/* Number of kernel launches issued by this engine so far (bench.py's gpu_launches). */
uint64_t lc_engine_launch_count(const lc_engine_t* e);
/* Pinned host memory helpers (a SourceBuffer arena allocated here is DMA-able without staging). */
void* lc_host_alloc(size_t bytes);
void lc_host_free(void* p);

/* ---- regex compilation (host only, no GPU needed) -------------------------------------------
 * Replaces boost::regex(pattern) at ProcessorParseRegexNative.cpp:66 and
 * ProcessorSplitMultilineLogStringNative.cpp:72-78.  *out is set (and must be freed) whenever the
 * return code is LC_OK or LC_ERR_REGEX_UNSUPPORTED/INVALID so that lc_regex_error() can be read. */
int lc_regex_compile(const char* pattern, size_t len, lc_regex_t** out);
void lc_regex_free(lc_regex_t* r);
const char* lc_regex_error(const lc_regex_t* r);
uint32_t lc_regex_ngroups(const lc_regex_t* r);
/* info[0..7] = mode (0 forward-only, 1 two-pass), byte classes, NFA walker states, context kinds,
 * reverse-DFA states, prefix-DFA states, table bytes, NFA instructions. */
void lc_regex_info(const lc_regex_t* r, uint32_t info[8]);

/* ---- a1: ProcessorSplitLogStringNative::ProcessEvent (inner/ProcessorSplitLogStringNative.cpp:101-174)
 * Cuts buf[0,len) on split_char.  Piece k = (out_off[k], out_len[k]); empty pieces kept; a trailing
 * split char yields no extra empty piece.  *n_out = number of pieces (even when > cap => LC_ERR_CAPACITY). */
int lc_split_lines(lc_engine_t* e, const uint8_t* buf, uint64_t len, uint8_t split_char, uint32_t* out_off,
                   uint32_t* out_len, uint64_t cap, uint64_t* n_out);
int lc_split_lines_dev(lc_engine_t* e, const uint8_t* d_buf, uint64_t len, uint8_t split_char, uint32_t* d_out_off,
                       uint32_t* d_out_len, uint64_t cap, uint64_t* n_out /* host */);

/* ---- a3: ProcessorParseRegexNative::RegexLogLineParser (ProcessorParseRegexNative.cpp:186-253)
 * For each event i (value = base[ev_off[i], +ev_len[i])): boost::regex_match over the whole value
 * (StringTools.cpp:183-211).  status[i] as LC_REGEX_*; on LC_REGEX_OK row i of cap_off/cap_len
 * ([n][ngroups], offsets relative to base) holds what[g+1] = (begin, length); groups that did not
 * participate report (end of value, 0).  Rows of failed events are zero-filled. */
int lc_regex_parse(lc_engine_t* e, const lc_regex_t* re, const uint8_t* base, uint64_t base_len,
                   const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n, uint32_t nkeys, uint8_t* status,
                   uint32_t* cap_off, uint32_t* cap_len);
int lc_regex_parse_dev(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                       const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint32_t nkeys,
                       uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len);

/* Batched event groups -- Processor::Process(std::vector<PipelineEventGroup>&) (Processor.h:31): the arenas of many
 * groups (<= 512 KB each, LogFileReader.cpp:97) go up back to back into ONE packed device arena and are parsed by one
 * launch sequence, instead of one launch + six copies + one sync per group.  Span k = the bytes
 * [span_ptr[k], + span_len[k]) of one group's SourceBuffer chunk (DMA-able without staging when it was allocated
 * with lc_host_alloc); it lands at [span_dst[k], + span_len[k]) of the packed arena (16-byte aligned, ascending,
 * non-overlapping, below packed_len).  Events [span_first_ev[k], span_first_ev[k+1]) belong to span k; ev_off[] and
 * the returned cap_off[] are relative to the PACKED arena (host address = span_ptr[k] + (off - span_dst[k])). */
int lc_regex_parse_packed(lc_engine_t* e, const lc_regex_t* re, uint64_t nspans, const uint8_t* const* span_ptr,
                          const uint32_t* span_len, const uint32_t* span_dst, const uint64_t* span_first_ev,
                          uint64_t packed_len, const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n,
                          uint32_t nkeys, uint8_t* status, uint32_t* cap_off, uint32_t* cap_len);

/* Same, handing the spans back as they finish: on_done(ctx, first_span, span_count) is called on the calling thread, in
 * span order, once the result rows of those spans' events have landed in status / cap_off / cap_len -- later spans are
 * still being uploaded and parsed meanwhile, so the caller's per-event epilogue overlaps the GPU pipeline. */
typedef void (*lc_spans_done_fn)(void* ctx, uint64_t first_span, uint64_t span_count);
int lc_regex_parse_packed_cb(lc_engine_t* e, const lc_regex_t* re, uint64_t nspans, const uint8_t* const* span_ptr,
                             const uint32_t* span_len, const uint32_t* span_dst, const uint64_t* span_first_ev,
                             uint64_t packed_len, const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n,
                             uint32_t nkeys, uint8_t* status, uint32_t* cap_off, uint32_t* cap_len,
                             lc_spans_done_fn on_done, void* ctx);

/* Same, with the event table read in place from a strided table: event i = (d_ev_off[i * ev_stride],
 * d_ev_len[i * ev_stride]).  Lets one processor's output feed the next without a gather -- e.g. column k of
 * lc_delim_parse_dev's [n][max_fields] tables (d_f_off + k, d_f_len + k, stride max_fields) is the event table of
 * the regex that parses that column (the delimiter -> regex chain of a pipeline, BASELINE config C4). */
int lc_regex_parse_strided_dev(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                               const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint32_t ev_stride, uint64_t n,
                               uint32_t nkeys, uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len);

/* Several patterns evaluated in ONE grid (BASELINE config C5 "multi-pattern"; north_star: "evaluate every log line
 * in the batch in one grid").  The reference runs one ProcessorParseRegexNative per pattern, each holding its own
 * boost::regex (ProcessorParseRegexNative.cpp:64-67), and a line parsed by pattern p sees exactly RegexLogLineParser
 * of that instance (:186-253).  Here all automata are resident in shared memory together and every line is tried on
 * the patterns in array order until one matches (== regex_match of `(?:p0)|(?:p1)|...`, groups numbered per
 * pattern); with sel != NULL line i is tried on pattern sel[i] only (LC_MULTI_ANY = on all, in order).
 * which[i] = index of the pattern that matched or LC_MULTI_NONE; status[i] as LC_REGEX_* for that pattern's nkeys;
 * rows of cap_off / cap_len are [n][row_pitch] (row_pitch >= the largest group count), columns beyond the matching
 * pattern's groups and rows of unmatched lines are zero.  1..8 patterns. */
#define LC_MULTI_NONE 0xFFu
#define LC_MULTI_ANY 0xFFu
int lc_regex_parse_multi(lc_engine_t* e, const lc_regex_t* const* res, uint32_t npat, const uint32_t* nkeys,
                         const uint8_t* base, uint64_t base_len, const uint32_t* ev_off, const uint32_t* ev_len,
                         uint64_t n, const uint8_t* sel, uint8_t* which, uint8_t* status, uint32_t row_pitch,
                         uint32_t* cap_off, uint32_t* cap_len);
int lc_regex_parse_multi_dev(lc_engine_t* e, const lc_regex_t* const* res, uint32_t npat, const uint32_t* nkeys,
                             const uint8_t* d_base, uint64_t base_len, const uint32_t* d_ev_off,
                             const uint32_t* d_ev_len, uint64_t n, const uint8_t* d_sel, uint8_t* d_which,
                             uint8_t* d_status, uint32_t row_pitch, uint32_t* d_cap_off, uint32_t* d_cap_len);

/* Boolean whole-value match == BoostRegexMatch(buf, size, reg, exception) without captures
 * (core/common/StringTools.cpp:213-236), the arithmetic of ProcessorFilterNative::IsMatched
 * (core/plugin/processor/ProcessorFilterNative.cpp:258-275): out_match[i] = 1 iff regex_match holds.
 * Only the reverse pass of the automaton runs. */
int lc_regex_match(lc_engine_t* e, const lc_regex_t* re, const uint8_t* base, uint64_t base_len,
                   const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n, uint8_t* out_match);
int lc_regex_match_dev(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                       const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint8_t* d_out_match);

/* Anchored prefix probe == BoostRegexSearch / regex_search(match_continuous) (StringTools.cpp:263-288),
 * one boolean per event. */
int lc_regex_prefix_match(lc_engine_t* e, const lc_regex_t* re, const uint8_t* base, uint64_t base_len,
                          const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n, uint8_t* out_match);

/* ---- a2: ProcessorSplitMultilineLogStringNative::ProcessEvent (inner/...Multiline...cpp:127-393)
 * One source value buf[0,len).  start/cont/end: compiled pattern or NULL (pattern string empty, :68-70).
 * Output event k = (out_off[k], out_len[k], out_flags[k] = LC_ML_*) in reference emission order.
 * counters[0..2] += matched_events, input_lines, unmatched_lines (:82-84,106-107). */
int lc_multiline_split(lc_engine_t* e, const uint8_t* buf, uint64_t len, const lc_regex_t* start,
                       const lc_regex_t* cont, const lc_regex_t* end, int discard_unmatched, uint32_t* out_off,
                       uint32_t* out_len, uint8_t* out_flags, uint64_t cap, uint64_t* n_out, uint64_t counters[3]);
int lc_multiline_split_dev(lc_engine_t* e, const uint8_t* d_buf, uint64_t len, const lc_regex_t* start,
                           const lc_regex_t* cont, const lc_regex_t* end, int discard_unmatched, uint32_t* d_out_off,
                           uint32_t* d_out_len, uint8_t* d_out_flags, uint64_t cap, uint64_t* n_out /* host */,
                           uint64_t counters[3] /* host */);

/* ---- f3 (next row): LogFileReader::RemoveLastIncompleteLog (core/file_server/reader/LogFileReader.cpp:1997-2064) for
 * raw text (RawTextParser::GetLastLine, :2186-2204): how many leading bytes of a freshly read chunk form complete logs.
 * start / end: MultilineOptions::GetStartPatternReg / GetEndPatternReg or NULL (multiline mode iff one is set).
 * *keep_bytes = the function's return value ("the number of bytes left, including \n"); *rollback_line_feeds =
 * rollbackLineFeedCount (left untouched when allow_rollback == 0 or len == 0, like the reference).  The walk back
 * over the chunk's lines becomes "the last line whose probe flag is set" over the split pass's line table, so a raw
 * file chunk can go to the GPU before the reader knows where its last complete record ends. */
int lc_remove_last_incomplete_log(lc_engine_t* e, const uint8_t* buf, uint64_t len, const lc_regex_t* start,
                                  const lc_regex_t* end, int allow_rollback, uint64_t* keep_bytes,
                                  int32_t* rollback_line_feeds);
int lc_remove_last_incomplete_log_dev(lc_engine_t* e, const uint8_t* d_buf, uint64_t len, const lc_regex_t* start,
                                      const lc_regex_t* end, int allow_rollback, uint64_t* keep_bytes /* host */,
                                      int32_t* rollback_line_feeds /* host */);

/* ---- a4: ProcessorParseDelimiterNative::ProcessEvent (ProcessorParseDelimiterNative.cpp:206-409)
 *          + DelimiterModeFsmParser::ParseDelimiterLine (core/parser/DelimiterModeFsmParser.cpp:260-294)
 * Per event: trim (:226-238), then the quote FSM (sep_len == 1 && quote != sep[0]) or the multi-char
 * SplitString (:366-409).  status[i] as LC_DELIM_*; nfields[i] = parsed column count; rows of
 * f_off/f_len/f_dq are [n][max_fields]: raw span of column j and, for the FSM path, the number of doubled
 * quotes inside it (the un-escaped value has f_len - f_dq bytes; the host shim materialises it in the
 * arena exactly as AddFieldWithUnQuote does, :83-113).  Columns beyond max_fields are counted, not stored. */
int lc_delim_parse(lc_engine_t* e, const uint8_t* base, uint64_t base_len, const uint32_t* ev_off,
                   const uint32_t* ev_len, uint64_t n, const uint8_t* sep, uint32_t sep_len, uint8_t quote,
                   uint32_t nkeys, int extend, int allow_short, uint32_t max_fields, uint8_t* status,
                   uint32_t* nfields, uint32_t* f_off, uint32_t* f_len, uint32_t* f_dq);
int lc_delim_parse_dev(lc_engine_t* e, const uint8_t* d_base, uint64_t base_len, const uint32_t* d_ev_off,
                       const uint32_t* d_ev_len, uint64_t n, const uint8_t* sep, uint32_t sep_len, uint8_t quote,
                       uint32_t nkeys, int extend, int allow_short, uint32_t max_fields, uint8_t* d_status,
                       uint32_t* d_nfields, uint32_t* d_f_off, uint32_t* d_f_len, uint32_t* d_f_dq);

/* Same, with a column tap for processor chains (ProcessorParseDelimiterNative -> ProcessorParseRegexNative on one
 * column, BASELINE config C4): the (offset, length) of column tap_col of every line are ALSO written to the dense
 * tables d_tap_off / d_tap_len -- exactly the event table lc_regex_parse_dev needs for that column, so the next stage
 * reads 8 bytes per line instead of striding through the [n][max_fields] tables (rows of failed lines are (0, 0)). */
int lc_delim_parse_tap_dev(lc_engine_t* e, const uint8_t* d_base, uint64_t base_len, const uint32_t* d_ev_off,
                           const uint32_t* d_ev_len, uint64_t n, const uint8_t* sep, uint32_t sep_len, uint8_t quote,
                           uint32_t nkeys, int extend, int allow_short, uint32_t max_fields, uint8_t* d_status,
                           uint32_t* d_nfields, uint32_t* d_f_off, uint32_t* d_f_len, uint32_t* d_f_dq,
                           uint32_t tap_col, uint32_t* d_tap_off, uint32_t* d_tap_len);

/* The same chain with HOST buffers (pinned memory recommended: lc_host_alloc): the arena is uploaded ONCE, in chunks of
 * whole events; per chunk the delimiter stage runs with the column tap and the regex stage on the tapped column, and both
 * stages' tables travel back while later chunks are still being uploaded (three streams, as in lc_regex_parse).
 * Outputs: the delimiter tables of lc_delim_parse ([n], [n][max_fields]) and the regex tables of lc_regex_parse for
 * column `column` ([n], [n][regex_nkeys groups]; a line whose delimiter stage failed or that has no such column is
 * parsed as the empty value).  Replaces: ProcessorParseDelimiterNative::Process followed by
 * ProcessorParseRegexNative::Process on one of its keys (core/plugin/processor/ProcessorParseDelimiterNative.cpp:206-364,
 * ProcessorParseRegexNative.cpp:132-168) -- a pipeline's `processors` list, collection_pipeline/CollectionPipeline.cpp. */
int lc_delim_regex_chain(lc_engine_t* e, const uint8_t* base, uint64_t base_len, const uint32_t* ev_off,
                         const uint32_t* ev_len, uint64_t n, const uint8_t* sep, uint32_t sep_len, uint8_t quote,
                         uint32_t nkeys, int extend, int allow_short, uint32_t max_fields, uint8_t* status,
                         uint32_t* nfields, uint32_t* f_off, uint32_t* f_len, uint32_t* f_dq, uint32_t column,
                         const lc_regex_t* re, uint32_t regex_nkeys, uint8_t* re_status, uint32_t* cap_off,
                         uint32_t* cap_len);

/* ---- f4 (next row): SLSEventGroupSerializer::Serialize for LOG events
 *          (core/collection_pipeline/serializer/SLSSerializer.cpp:254-269,377-395 over the writer of
 *           core/protobuf/sls/LogGroupSerializer.cpp:33-143,232-262)
 * Emits the `Logs` fields (field 1 of sls_logs::LogGroup), concatenated in event order.  Event i owns the contents
 * entries [ent_begin[i], ent_begin[i+1]) (m = ent_begin[n] entries in all); entry k is the key
 * base[ent_koff[k], +ent_klen[k]) and the value base[ent_voff[k], +ent_vlen[k]).  Events without entries are skipped
 * (LogEvent::Empty); Time below 2^28 is raised to 2^28 (the reference keeps the varint at 5 bytes);
 * ev_time_ns may be NULL, and ev_time_ns[i] == LC_SLS_NO_NS means "no Time_ns field" (nanoseconds disabled or not
 * set).  *out_len receives the total size; if it exceeds out_cap nothing is written and LC_ERR_CAPACITY is returned.
 * The group-level fields (Topic, Source, MachineUUID, LogTags) are a few bytes appended by the caller. */
#define LC_SLS_NO_NS 0xFFFFFFFFu
int lc_sls_serialize_logs(lc_engine_t* e, const uint8_t* base, uint64_t base_len, uint64_t n, const uint32_t* ev_time,
                          const uint32_t* ev_time_ns, const uint64_t* ent_begin, const uint32_t* ent_koff,
                          const uint32_t* ent_klen, const uint32_t* ent_voff, const uint32_t* ent_vlen, uint8_t* out,
                          uint64_t out_cap, uint64_t* out_len);

/* Device-fed variant for the regex -> serialise hand-over: the `Logs` fields of the events a ProcessorParseRegexNative
 * leaves behind, written straight from its DEVICE result tables (status / cap_off / cap_len of lc_regex_parse_dev,
 * rows of row_pitch entries) and the constant key strings -- no host-side (key, value) span lists, the bytes never
 * leave the GPU between parsing and the wire format.  Event i with LC_REGEX_OK carries keys[k] -> capture k for
 * k < nkeys in key order (AddLog per capture, ProcessorParseRegexNative.cpp:249-251; the source key is deleted,
 * :153-155); a failed event carries the one content fail_key -> the whole line when fail_key != NULL
 * (KeepingSourceWhenParseFail with RenamedSourceKey, :156-158) and is skipped otherwise (erased,
 * CommonParserOptions.cpp:99-117).  keys must be distinct.  d_ev_time_ns may be NULL; LC_SLS_NO_NS per event = no
 * Time_ns.  d_out receives the bytes on the device; *out_len (host) their count; LC_ERR_CAPACITY if > out_cap. */
int lc_sls_serialize_parsed_dev(lc_engine_t* e, const uint8_t* d_base, uint64_t base_len, const uint32_t* d_ev_off,
                                const uint32_t* d_ev_len, const uint8_t* d_status, const uint32_t* d_cap_off,
                                const uint32_t* d_cap_len, uint32_t row_pitch, uint64_t n, const char* const* keys,
                                const uint32_t* key_lens, uint32_t nkeys, const char* fail_key, uint32_t fail_key_len,
                                const uint32_t* d_ev_time, const uint32_t* d_ev_time_ns, uint8_t* d_out,
                                uint64_t out_cap, uint64_t* out_len);

#ifdef __cplusplus
}
#endif
#endif /* LC_B200_H */
