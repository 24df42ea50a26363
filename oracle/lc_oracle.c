/*
 * lc_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the observable results of LoongCollector's four
 * native log-parsing processors on FLAT buffers.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load
 * this file's shared object; the product library (loongcollector_b200/) never
 * links, loads or calls it.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference checkout, /root/reference in the build container):
 *   split     core/plugin/processor/inner/ProcessorSplitLogStringNative.cpp:101-174
 *   multiline core/plugin/processor/inner/ProcessorSplitMultilineLogStringNative.cpp:127-393
 *   regex     core/plugin/processor/ProcessorParseRegexNative.cpp:186-253,
 *             core/common/StringTools.cpp:183-211 (regex_match), :263-288 (regex_search|match_continuous)
 *   delimiter core/plugin/processor/ProcessorParseDelimiterNative.cpp:206-409,
 *             core/parser/DelimiterModeFsmParser.cpp:49-113,134-154,172-186,201-223,260-294
 *
 * Regex arithmetic: the reference delegates to Boost.Regex 1.68 (un-vendored,
 * not installable here).  This oracle delegates to PCRE2 10.42 (libpcre2-8.so.0,
 * loaded with dlopen because the image ships no pcre2.h) configured for the
 * boost defaults that matter on this path: '.' matches
 * '\n' (DOTALL), '^'/'$' are line anchors (MULTILINE), bytes not UTF, full match =
 * ANCHORED|ENDANCHORED, match_continuous = ANCHORED.  Unset groups are reported
 * boost-style as (end_of_input, 0).
 *
 * PARITY PINNING: pinned against every golden vector the reference's own unit
 * tests and docs hold for this path (tests/golden/, extracted by
 * tests/golden/extract_reference_vectors.py).  Patterns outside those vectors
 * are pinned only by agreement of independent Perl-semantics engines
 * (PCRE2 here, Python `re` in tests) -- "parity unpinned at the boost boundary".
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ PCRE2 */
typedef struct pcre2_real_code_8 pcre2_code;
typedef struct pcre2_real_match_data_8 pcre2_match_data;
#define ORC_PCRE2_ANCHORED 0x80000000u
#define ORC_PCRE2_ENDANCHORED 0x20000000u
#define ORC_PCRE2_DOTALL 0x00000020u
#define ORC_PCRE2_MULTILINE 0x00000400u
#define ORC_PCRE2_INFO_CAPTURECOUNT 4
#define ORC_PCRE2_UNSET (~(size_t)0)

static struct {
    void* lib;
    pcre2_code* (*compile)(const uint8_t*, size_t, uint32_t, int*, size_t*, void*);
    void (*code_free)(pcre2_code*);
    pcre2_match_data* (*md_create)(const pcre2_code*, void*);
    void (*md_free)(pcre2_match_data*);
    int (*match)(const pcre2_code*, const uint8_t*, size_t, size_t, uint32_t, pcre2_match_data*, void*);
    size_t* (*ovector)(pcre2_match_data*);
    int (*info)(const pcre2_code*, uint32_t, void*);
    int (*jit_compile)(pcre2_code*, uint32_t);
} P;

static int orc_load_pcre2(void) {
    if (P.lib)
        return 0;
    const char* names[] = {"libpcre2-8.so.0", "libpcre2-8.so", "/usr/lib/x86_64-linux-gnu/libpcre2-8.so.0", 0};
    for (int i = 0; names[i] && !P.lib; ++i)
        P.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!P.lib)
        return -1;
    P.compile = dlsym(P.lib, "pcre2_compile_8");
    P.code_free = dlsym(P.lib, "pcre2_code_free_8");
    P.md_create = dlsym(P.lib, "pcre2_match_data_create_from_pattern_8");
    P.md_free = dlsym(P.lib, "pcre2_match_data_free_8");
    P.match = dlsym(P.lib, "pcre2_match_8");
    P.ovector = dlsym(P.lib, "pcre2_get_ovector_pointer_8");
    P.info = dlsym(P.lib, "pcre2_pattern_info_8");
    P.jit_compile = dlsym(P.lib, "pcre2_jit_compile_8");
    if (!P.compile || !P.code_free || !P.md_create || !P.md_free || !P.match || !P.ovector || !P.info)
        return -2;
    return 0;
}

typedef struct orc_regex {
    pcre2_code* code;
    uint32_t ngroups;
} orc_regex;

/* boost::regex(pattern) -- ProcessorParseRegexNative.cpp:66, ProcessorSplitMultilineLogStringNative.cpp:72-78.
 * jit!=0 additionally JIT-compiles (used only to report a best-case CPU number; boost has no JIT). */
orc_regex* orc_regex_compile(const char* pattern, uint64_t len, int jit) {
    if (orc_load_pcre2() != 0)
        return NULL;
    int err = 0;
    size_t erroff = 0;
    pcre2_code* c = P.compile((const uint8_t*)pattern, (size_t)len, ORC_PCRE2_DOTALL | ORC_PCRE2_MULTILINE, &err,
                              &erroff, NULL);
    if (!c)
        return NULL;
    if (jit && P.jit_compile)
        P.jit_compile(c, 1u /* PCRE2_JIT_COMPLETE */);
    orc_regex* r = (orc_regex*)calloc(1, sizeof(orc_regex));
    r->code = c;
    P.info(c, ORC_PCRE2_INFO_CAPTURECOUNT, &r->ngroups);
    return r;
}

void orc_regex_free(orc_regex* r) {
    if (!r)
        return;
    P.code_free(r->code);
    free(r);
}

uint32_t orc_regex_ngroups(const orc_regex* r) {
    return r->ngroups;
}

/* One matcher scratch per calling thread (mirrors the reference's per-thread regex copies). */
typedef struct orc_matcher {
    const orc_regex* re;
    pcre2_match_data* md;
} orc_matcher;

orc_matcher* orc_matcher_create(const orc_regex* re) {
    orc_matcher* m = (orc_matcher*)calloc(1, sizeof(orc_matcher));
    m->re = re;
    m->md = P.md_create(re->code, NULL);
    return m;
}

void orc_matcher_free(orc_matcher* m) {
    if (!m)
        return;
    P.md_free(m->md);
    free(m);
}

/* BoostRegexSearch(buf,size,reg,exc) == regex_search(..., match_continuous): StringTools.cpp:263-288. */
int orc_regex_prefix_match(orc_matcher* m, const uint8_t* buf, uint64_t len) {
    int rc = P.match(m->re->code, buf, (size_t)len, 0, ORC_PCRE2_ANCHORED, m->md, NULL);
    return rc >= 0;
}

/* BoostRegexMatch(buf,len,reg,exc,what,match_default) == regex_match: StringTools.cpp:183-211.
 * cap_off/cap_len receive ngroups entries (groups 1..ngroups), offsets relative to buf.
 * Unset groups -> (len, 0) (boost: first == second == end of input). Returns 1 on match. */
int orc_regex_full_match(orc_matcher* m, const uint8_t* buf, uint64_t len, uint32_t* cap_off, uint32_t* cap_len) {
    int rc = P.match(m->re->code, buf, (size_t)len, 0, ORC_PCRE2_ANCHORED | ORC_PCRE2_ENDANCHORED, m->md, NULL);
    if (rc < 0)
        return 0;
    size_t* ov = P.ovector(m->md);
    for (uint32_t g = 1; g <= m->re->ngroups; ++g) {
        if (cap_off) {
            /* rc == highest set group + 1; groups >= rc are unset and their ovector entries hold PCRE2_UNSET */
            if ((int)g >= rc || ov[2 * g] == ORC_PCRE2_UNSET) {
                cap_off[g - 1] = (uint32_t)len;
                cap_len[g - 1] = 0;
            } else {
                cap_off[g - 1] = (uint32_t)ov[2 * g];
                cap_len[g - 1] = (uint32_t)(ov[2 * g + 1] - ov[2 * g]);
            }
        }
    }
    return 1;
}

/* ProcessorParseRegexNative::RegexLogLineParser over a flat batch (:186-253).
 * status: 0 = ok, 1 = no match (out_failed++), 2 = matched but what.size() <= nkeys (:227, no out_failed++).
 * cap arrays are [n * ngroups]; rows of failed events are zero-filled. */
void orc_regex_parse_batch(orc_matcher* m, const uint8_t* base, const uint32_t* ev_off, const uint32_t* ev_len,
                           uint64_t n, uint32_t nkeys, uint8_t* status, uint32_t* cap_off, uint32_t* cap_len) {
    uint32_t G = m->re->ngroups;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t* co = cap_off + i * G;
        uint32_t* cl = cap_len + i * G;
        if (!orc_regex_full_match(m, base + ev_off[i], ev_len[i], co, cl)) {
            status[i] = 1;
            memset(co, 0, G * sizeof(uint32_t));
            memset(cl, 0, G * sizeof(uint32_t));
        } else if (G + 1 <= nkeys) {
            status[i] = 2;
            memset(co, 0, G * sizeof(uint32_t));
            memset(cl, 0, G * sizeof(uint32_t));
        } else {
            status[i] = 0;
            for (uint32_t g = 0; g < G; ++g)
                co[g] += ev_off[i]; /* make offsets relative to base, like the engine */
        }
    }
}

/* BoostRegexMatch(buffer, size, reg, exception) as used by ProcessorFilterNative::IsMatched
 * (core/plugin/processor/ProcessorFilterNative.cpp:258-275, core/common/StringTools.cpp:213-236): boolean per event. */
void orc_regex_match_batch(orc_matcher* m, const uint8_t* base, const uint32_t* ev_off, const uint32_t* ev_len,
                           uint64_t n, uint8_t* out) {
    for (uint64_t i = 0; i < n; ++i)
        out[i] = (uint8_t)orc_regex_full_match(m, base + ev_off[i], ev_len[i], NULL, NULL);
}

/* ------------------------------------------------------------------ split */
/* ProcessorSplitLogStringNative::ProcessEvent + GetNextLine (:127-174): pieces between split chars;
 * empty pieces kept; a trailing split char yields no extra empty piece (loop stops at begin >= size).
 * Returns the number of pieces; writes at most cap. */
uint64_t orc_split_lines(const uint8_t* buf, uint64_t len, uint8_t split_char, uint32_t* off, uint32_t* ln,
                         uint64_t cap) {
    uint64_t n = 0, begin = 0;
    while (begin < len) {
        const uint8_t* p = (const uint8_t*)memchr(buf + begin, split_char, (size_t)(len - begin));
        uint64_t end = p ? (uint64_t)(p - buf) : len;
        if (n < cap) {
            off[n] = (uint32_t)begin;
            ln[n] = (uint32_t)(end - begin);
        }
        ++n;
        begin = end + 1;
    }
    return n;
}

/* ------------------------------------------------------------------ multiline */
typedef struct orc_ml_out {
    uint32_t* off;
    uint32_t* len;
    uint8_t* flags; /* bit0 = isLastLog flag used for SetPosition (:329-332), bit1 = matched record */
    uint64_t cap, n;
    uint64_t matched_events, input_lines, unmatch_lines;
    int discard;
} orc_ml_out;

static void ml_emit(orc_ml_out* o, uint64_t off, uint64_t len, int is_last, int matched) {
    if (o->n < o->cap) {
        o->off[o->n] = (uint32_t)off;
        o->len[o->n] = (uint32_t)len;
        o->flags[o->n] = (uint8_t)((is_last ? 1 : 0) | (matched ? 2 : 0));
    }
    o->n++;
}

/* HandleUnmatchLogs (:342-380): re-split the span; single_line -> one event per line carrying the CALLER's
 * isLastLog flag; discard -> nothing; unmatch_lines += lines. */
static void ml_unmatch(orc_ml_out* o, const uint8_t* buf, uint64_t off, uint64_t len, int is_last) {
    uint64_t begin = 0;
    while (begin < len) {
        const uint8_t* p = (const uint8_t*)memchr(buf + off + begin, '\n', (size_t)(len - begin));
        uint64_t end = p ? (uint64_t)(p - (buf + off)) : len;
        o->unmatch_lines++;
        if (!o->discard)
            ml_emit(o, off + begin, end - begin, is_last, 0);
        begin = end + 1;
    }
}

/* ProcessorSplitMultilineLogStringNative::ProcessEvent (:162-308) on one source value.
 * start/cont/end: matcher or NULL (== pattern string empty, :68-70).
 * counters[0..2] += matched_events, input_lines, unmatch_lines.  Returns number of output events. */
uint64_t orc_multiline_split(const uint8_t* buf, uint64_t len, orc_matcher* start, orc_matcher* cont, orc_matcher* end,
                             int discard_unmatched, uint32_t* out_off, uint32_t* out_len, uint8_t* out_flags,
                             uint64_t cap, uint64_t* counters) {
    orc_ml_out o = {out_off, out_len, out_flags, cap, 0, 0, 0, 0, discard_unmatched};
    const int S = start != NULL, C = cont != NULL, E = end != NULL;
    int partial = 0;
    uint64_t ms = 0; /* multiStartIndex as an offset */
    if (!S && !C && E) {
        partial = 1;
        ms = 0;
    }
    uint64_t begin = 0;
    while (begin < len) {
        const uint8_t* p = (const uint8_t*)memchr(buf + begin, '\n', (size_t)(len - begin));
        uint64_t lend = p ? (uint64_t)(p - buf) : len;
        const uint8_t* line = buf + begin;
        uint64_t llen = lend - begin;
        int is_last = (lend == len); /* begin + content.size() == sourceVal.size() (:174) */
        o.input_lines++;
        if (!partial) {
            orc_matcher* probe = S ? start : cont;
            if (probe && orc_regex_prefix_match(probe, line, llen)) {
                ms = begin;
                partial = 1;
            } else if (E && !S && C && orc_regex_prefix_match(end, line, llen)) {
                ml_emit(&o, begin, llen, is_last, 1);
                ms = lend + 1;
                o.matched_events++;
            } else {
                ml_unmatch(&o, buf, begin, llen, is_last);
            }
        } else {
            if (C && orc_regex_prefix_match(cont, line, llen)) {
                begin = lend + 1;
                continue;
            }
            if (E) {
                if (C) {
                    if (orc_regex_prefix_match(end, line, llen)) {
                        ml_emit(&o, ms, lend - ms, is_last, 1);
                        o.matched_events++;
                    } else {
                        ml_unmatch(&o, buf, ms, lend - ms, is_last);
                    }
                    partial = 0;
                } else {
                    if (orc_regex_prefix_match(end, line, llen)) {
                        ml_emit(&o, ms, lend - ms, is_last, 1);
                        if (S)
                            partial = 0;
                        else
                            ms = lend + 1;
                        o.matched_events++;
                    }
                }
            } else {
                if (!C) {
                    if (orc_regex_prefix_match(start, line, llen)) {
                        ml_emit(&o, ms, begin - 1 - ms, is_last, 1);
                        ms = begin;
                        o.matched_events++;
                    }
                } else {
                    ml_emit(&o, ms, begin - 1 - ms, is_last, 1);
                    o.matched_events++;
                    if (!orc_regex_prefix_match(start, line, llen)) {
                        ml_unmatch(&o, buf, begin, llen, is_last);
                        partial = 0;
                    } else {
                        ms = begin;
                    }
                }
            }
        }
        begin = lend + 1;
    }
    if (partial && ms < len) { /* :289-308 */
        if (!E) {
            ml_emit(&o, ms, len - ms, 1, 1);
            o.matched_events++;
        } else {
            ml_unmatch(&o, buf, ms, len - ms, 1);
        }
    }
    if (counters) {
        counters[0] += o.matched_events;
        counters[1] += o.input_lines;
        counters[2] += o.unmatch_lines;
    }
    return o.n;
}

/* ------------------------------------------------------------------ last incomplete log (next row f3)
 * LogFileReader::RemoveLastIncompleteLog (core/file_server/reader/LogFileReader.cpp:1997-2064) over
 * RawTextParser::GetLastLine (:2186-2204): how many leading bytes of a freshly read chunk form complete logs, and how
 * many line feeds are rolled back.  start / end = MultilineOptions::GetStartPatternReg / GetEndPatternReg (NULL = not
 * configured; multiline mode iff one of them is set).  buf[size] must be readable (the reader's buffer is NUL
 * terminated one past the end, :2521). */
typedef struct orc_line_info {
    int32_t begin, end, rollback, full;
} orc_line_info;

static orc_line_info orc_get_last_line(const uint8_t* buf, int32_t end) { /* :2186-2204 */
    orc_line_info r = {0, 0, 0, 0};
    if (end == 0)
        return r;
    r.end = end;
    r.rollback = 1;
    r.full = 1;
    for (int32_t b = end; b > 0; --b)
        if (buf[b - 1] == '\n') {
            r.begin = b;
            return r;
        }
    r.begin = 0;
    return r;
}

int32_t orc_remove_last_incomplete_log(const uint8_t* buf, int32_t size, orc_matcher* start, orc_matcher* end,
                                       int allow_rollback, int32_t* rollback_out) {
    int32_t rollback = *rollback_out;
    if (!allow_rollback || size == 0) /* :1999-2001 (the count is left as the caller initialised it) */
        return size;
    int32_t endPs = buf[size - 1] == '\n' ? size - 1 : size;
    rollback = 0;
    int foundEnd = 0;
    if (start || end) { /* IsMultiline() */
        while (endPs >= 0) {
            orc_line_info c = orc_get_last_line(buf, endPs);
            const uint8_t* d = buf + c.begin;
            uint64_t dl = (uint64_t)(c.end - c.begin);
            if (end) {
                if (orc_regex_prefix_match(end, d, dl)) {
                    foundEnd = 1;
                    if (buf[c.end] == '\n') {
                        *rollback_out = rollback;
                        return c.end + 1;
                    }
                }
            } else if (start && orc_regex_prefix_match(start, d, dl)) {
                rollback += c.rollback;
                *rollback_out = rollback;
                return c.begin;
            }
            rollback += c.rollback;
            endPs = c.begin - 1;
        }
    }
    if (end && foundEnd) {
        *rollback_out = rollback;
        return 0;
    }
    rollback = 0;
    endPs = buf[size - 1] == '\n' ? size - 1 : size;
    orc_line_info c = orc_get_last_line(buf, endPs);
    if (c.full && buf[c.end] == '\n') {
        *rollback_out = rollback;
        return c.end + 1;
    }
    rollback += c.rollback;
    *rollback_out = rollback;
    return c.begin;
}

/* ------------------------------------------------------------------ delimiter */
enum { ST_INITIAL = 0, ST_QUOTE = 1, ST_DATA = 2, ST_DOUBLE_QUOTE = 3 };

/* Trim of ProcessorParseDelimiterNative::ProcessEvent (:219-242).  Returns 0 when the value is empty or
 * blank (caller: out_failed++, event untouched), else 1 with [*beg,*end). */
int orc_delim_trim(const uint8_t* v, uint32_t len, int32_t* beg, int32_t* end) {
    int32_t endIdx = (int32_t)len;
    if (endIdx == 0)
        return 0;
    for (int32_t i = endIdx - 1; i >= 0; --i) {
        if (v[i] == ' ' || v[i] == '\r')
            endIdx = i;
        else
            break;
    }
    int32_t begIdx = 0;
    for (int32_t i = 0; i < endIdx; ++i) {
        if (v[i] == ' ')
            begIdx = i + 1;
        else
            break;
    }
    if (begIdx >= endIdx)
        return 0;
    *beg = begIdx;
    *end = endIdx;
    return 1;
}

/* DelimiterModeFsmParser::ParseDelimiterLine(StringView,...) (:260-294) with the zero-copy handlers
 * (:49-81 separator, :134-154 quote, :172-186 data, :201-223 EOF).  Emits per field the RAW span
 * [fieldStart,fieldEnd) and the number of doubled quotes inside it (the un-escaped value has
 * len - dq bytes, AddFieldWithUnQuote :83-113).  Returns the field count, or -1 on FSM error
 * (all columns cleared).  Writes at most cap fields but counts all. */
int64_t orc_delim_fsm(const uint8_t* ch, int32_t begin, int32_t end, uint8_t sep, uint8_t quote, uint32_t* f_off,
                      uint32_t* f_len, uint32_t* f_dq, int64_t cap) {
    int state = ST_INITIAL;
    int dq = 0;
    int fs = begin, fe = begin;
    int64_t n = 0;
#define ORC_ADD_FIELD()                                                                                               \
    do {                                                                                                               \
        if (n < cap) {                                                                                                 \
            f_off[n] = (uint32_t)fs;                                                                                   \
            f_len[n] = (uint32_t)(fe - fs);                                                                            \
            f_dq[n] = (uint32_t)dq;                                                                                    \
        }                                                                                                              \
        ++n;                                                                                                           \
        dq = 0;                                                                                                        \
    } while (0)
    for (int i = begin; i < end; ++i) {
        uint8_t c = ch[i];
        if (c == sep) {
            switch (state) {
                case ST_INITIAL:
                    ORC_ADD_FIELD();
                    fs = ++fe;
                    break;
                case ST_QUOTE:
                    fe++;
                    break;
                case ST_DATA:
                    state = ST_INITIAL;
                    ORC_ADD_FIELD();
                    fs = ++fe;
                    break;
                case ST_DOUBLE_QUOTE:
                    state = ST_INITIAL;
                    dq--;
                    ORC_ADD_FIELD();
                    fe += 2;
                    fs = fe;
                    break;
            }
        } else if (c == quote) {
            switch (state) {
                case ST_INITIAL:
                    state = ST_QUOTE;
                    fs++;
                    break;
                case ST_QUOTE:
                    state = ST_DOUBLE_QUOTE;
                    dq++;
                    fe++;
                    break;
                case ST_DATA:
                    return -1;
                case ST_DOUBLE_QUOTE:
                    state = ST_QUOTE;
                    fe++;
                    break;
            }
        } else {
            switch (state) {
                case ST_INITIAL:
                    state = ST_DATA;
                    fe++;
                    break;
                case ST_QUOTE:
                case ST_DATA:
                    fe++;
                    break;
                case ST_DOUBLE_QUOTE:
                    return -1;
            }
        }
    }
    if (state == ST_DOUBLE_QUOTE)
        dq--;
    if (state == ST_QUOTE)
        return -1;
    ORC_ADD_FIELD();
#undef ORC_ADD_FIELD
    return n;
}

/* AddFieldWithUnQuote (:83-113): collapse doubled quotes of a raw span into dst; returns bytes written. */
uint32_t orc_delim_unquote(const uint8_t* ch, uint32_t off, uint32_t len, uint8_t quote, uint8_t* dst) {
    uint32_t j = 0;
    for (uint32_t i = off; i < off + len; ++i) {
        if (ch[i] == quote) {
            if (i + 1 < off + len && ch[i + 1] == quote) {
                dst[j++] = quote;
                ++i;
            }
        } else {
            dst[j++] = ch[i];
        }
    }
    return j;
}

/* ProcessorParseDelimiterNative::SplitString (:366-409): multi-char separator (or quote == separator).
 * extend != 0 <=> OverflowedFieldsTreatment == EXTEND.  Returns the field count (0 == false). */
int64_t orc_delim_split(const uint8_t* buffer, int32_t begIdx, int32_t endIdx, const uint8_t* sep, uint32_t d_size,
                        uint32_t nkeys, int extend, uint32_t* f_off, uint32_t* f_len, int64_t cap) {
    int64_t n = 0;
#define ORC_PUSH(o, l)                                                                                                \
    do {                                                                                                               \
        if (n < cap) {                                                                                                 \
            f_off[n] = (uint32_t)(o);                                                                                  \
            f_len[n] = (uint32_t)(l);                                                                                  \
        }                                                                                                              \
        ++n;                                                                                                           \
    } while (0)
    if (endIdx <= begIdx || d_size == 0 || nkeys == 0)
        return 0;
    size_t size = (size_t)(endIdx - begIdx);
    if (d_size > size) {
        ORC_PUSH(begIdx, size);
        return n;
    }
    size_t pos = (size_t)begIdx;
    size_t top = (size_t)endIdx - d_size;
    while (pos <= top) {
        const uint8_t* pch = (const uint8_t*)memmem(buffer + pos, (size_t)endIdx - pos, sep, d_size);
        size_t pos2 = pch ? (size_t)(pch - buffer) : (size_t)endIdx;
        ORC_PUSH(pos, pos2 - pos);
        if (pos2 == (size_t)endIdx)
            return n;
        pos = pos2 + d_size;
        if ((uint64_t)n >= nkeys && !extend) {
            ORC_PUSH(pos2, (size_t)endIdx - pos2);
            return n;
        }
    }
    if (pos <= (size_t)endIdx)
        ORC_PUSH(pos, (size_t)endIdx - pos);
#undef ORC_PUSH
    return n;
}

/* Flat batch form used for parity against the engine and for the CPU baseline.
 * Per event: status 0 = ok, 1 = FSM/split failure, 2 = empty/blank value (event untouched, :220-242),
 *            3 = column-count failure (:285).
 * nfields[i] = number of parsed columns (before any overflow join); field rows are [n * max_fields],
 * offsets relative to base.  mode_quote != 0 -> FSM path, else SplitString path. */
void orc_delim_parse_batch(const uint8_t* base, const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n,
                           const uint8_t* sep, uint32_t sep_len, uint8_t quote, int mode_quote, uint32_t nkeys,
                           int extend, int allow_short, uint32_t max_fields, uint8_t* status, uint32_t* nfields,
                           uint32_t* f_off, uint32_t* f_len, uint32_t* f_dq) {
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t* v = base + ev_off[i];
        uint32_t* fo = f_off + i * max_fields;
        uint32_t* fl = f_len + i * max_fields;
        uint32_t* fd = f_dq + i * max_fields;
        memset(fo, 0, max_fields * 4);
        memset(fl, 0, max_fields * 4);
        memset(fd, 0, max_fields * 4);
        nfields[i] = 0;
        int32_t b, e;
        if (!orc_delim_trim(v, ev_len[i], &b, &e)) {
            status[i] = 2;
            continue;
        }
        int64_t k;
        if (nkeys == 0) {
            status[i] = 1;
            continue;
        }
        if (mode_quote)
            k = orc_delim_fsm(v, b, e, sep[0], quote, fo, fl, fd, max_fields);
        else
            k = orc_delim_split(v, b, e, sep, sep_len, nkeys, extend, fo, fl, max_fields);
        if (k < 0 || (!mode_quote && k == 0)) {
            status[i] = 1;
            memset(fo, 0, max_fields * 4);
            memset(fl, 0, max_fields * 4);
            memset(fd, 0, max_fields * 4);
            continue;
        }
        nfields[i] = (uint32_t)k;
        uint32_t w = k < (int64_t)max_fields ? (uint32_t)k : max_fields;
        for (uint32_t j = 0; j < w; ++j)
            fo[j] += ev_off[i];
        /* FSM path, non-extend: columns >= nkeys collapse into ONE joined column (:258-275) before the count test */
        uint64_t cols = (uint64_t)k;
        if (mode_quote && !extend && cols > nkeys)
            cols = (uint64_t)nkeys + 1;
        if (cols == 0 || (!allow_short && cols < nkeys))
            status[i] = 3;
        else
            status[i] = 0;
    }
}
