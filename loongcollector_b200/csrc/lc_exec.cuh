// lc_exec.cuh -- scalar interpretation of a compiled regex blob (lc_tables.h): the semantic core
// that the sm_100a kernels execute once per log line.  Written as __host__ __device__ so that the
// exact same statements can be exercised on the CPU by the test-only emulation library under
// tests/emul/ (which validates the COMPILER's tables against the oracle without a GPU).  The product
// library only ever calls these from device code.
#pragma once
#include <stdint.h>
#include <string.h>

#include "lc_tables.h"

#if defined(__CUDACC__)
#define LC_HD __host__ __device__ __forceinline__
#else
#define LC_HD inline
#endif

#define LC_SLOT_UNSET 0xFFFFFFFFu

struct LcProgView {
    const LcRegexHeader* h;
    const uint8_t* byte_class;
    const uint8_t* class_pc;
    const uint64_t* actions;
    const uint16_t* pre_next;
    const uint8_t* pre_acc;
    const uint32_t* fwd;
    const uint32_t* fwd_eof;
    const uint16_t* rev_next;
};

LC_HD LcProgView lc_view(const void* blob) {
    const uint8_t* b = (const uint8_t*)blob;
    const LcRegexHeader* h = (const LcRegexHeader*)blob;
    LcProgView v;
    v.h = h;
    v.byte_class = b + h->off_byte_class;
    v.class_pc = b + h->off_class_pc;
    v.actions = (const uint64_t*)(b + h->off_actions);
    v.pre_next = (const uint16_t*)(b + h->off_pre_next);
    v.pre_acc = b + h->off_pre_acc;
    v.fwd = (const uint32_t*)(b + h->off_fwd);
    v.fwd_eof = (const uint32_t*)(b + h->off_fwd_eof);
    v.rev_next = (const uint16_t*)(b + h->off_rev_next);
    return v;
}

// regex_search(..., match_continuous) as a boolean (reference: core/common/StringTools.cpp:263-288).
LC_HD bool lc_prefix_match(const LcProgView& v, const uint8_t* s, uint32_t n) {
    const uint32_t nc = v.h->nclasses;
    uint32_t st = v.h->pre_start;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t e = v.pre_next[st * nc + v.byte_class[s[i]]];
        if (e == LC_PREFIX_ACCEPT)
            return true;
        if (e == LC_PREFIX_DEAD)
            return false;
        st = e;
    }
    return v.pre_acc[st] != 0;
}

LC_HD void lc_apply_action(const LcProgView& v, uint32_t act, uint32_t pos, uint32_t* slots) {
    uint64_t m = v.actions[act];
    while (m) {
#if defined(__CUDA_ARCH__)
        int s = __ffsll((long long)m) - 1;
#else
        int s = __builtin_ctzll(m);
#endif
        slots[s] = pos;
        m &= m - 1;
    }
}

// regex_match with captures, forward-only automaton (mode LC_MODE_FWD1).
// slots: 2*ngroups entries, pre-filled with LC_SLOT_UNSET by the caller.
LC_HD bool lc_full_match_fwd1(const LcProgView& v, const uint8_t* s, uint32_t n, uint32_t* slots) {
    const uint32_t cols = v.h->fwd_cols;
    const uint32_t npc = v.h->npc;
    uint32_t w = 0, pk = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t c = v.byte_class[s[i]];
        uint32_t e = v.fwd[(w * npc + pk) * cols + c];
        if (e == LC_NONE_ENTRY)
            return false;
        uint32_t a = LC_ENTRY_ACT(e);
        if (a)
            lc_apply_action(v, a, i, slots);
        w = LC_ENTRY_NEXT(e);
        if (npc > 1)
            pk = v.class_pc[c];
    }
    uint32_t e = v.fwd_eof[w * npc + pk];
    if (e == LC_NONE_ENTRY)
        return false;
    uint32_t a = LC_ENTRY_ACT(e);
    if (a)
        lc_apply_action(v, a, n, slots);
    return true;
}

// Reverse labelling pass of the two-pass matcher: lab[i] = reverse-DFA state at position i (0..n).
// Returns false early when the state dies (no suffix can reach a full match).
template <class LabT>
LC_HD bool lc_rev_label(const LcProgView& v, const uint8_t* s, uint32_t n, LabT* lab) {
    const uint32_t nc = v.h->nclasses;
    uint32_t d = v.h->rev_start;
    lab[n] = (LabT)d;
    for (uint32_t i = n; i-- > 0;) {
        d = v.rev_next[d * nc + v.byte_class[s[i]]];
        if (d == LC_REV_DEAD)
            return false;
        lab[i] = (LabT)d;
    }
    return true;
}

// Guided forward walk: at every position take the highest-priority transition that stays viable.
template <class LabT>
LC_HD bool lc_fwd_walk(const LcProgView& v, const uint8_t* s, uint32_t n, const LabT* lab, uint32_t* slots) {
    const uint32_t cols = v.h->fwd_cols;
    const uint32_t npc = v.h->npc;
    uint32_t w = 0, pk = 0;
    for (uint32_t i = 0; i <= n; ++i) {
        uint32_t e = v.fwd[(w * npc + pk) * cols + (uint32_t)lab[i]];
        if (e == LC_NONE_ENTRY)
            return false; // only possible at i == 0 (no viable start)
        uint32_t a = LC_ENTRY_ACT(e);
        if (a)
            lc_apply_action(v, a, i, slots);
        w = LC_ENTRY_NEXT(e);
        if (npc > 1 && i < n)
            pk = v.class_pc[v.byte_class[s[i]]];
    }
    return true;
}

// Capture slots -> boost sub_match (offset, length); a group that did not participate reports
// first == second == end of input, i.e. (n, 0).
LC_HD void lc_slots_to_cap(const uint32_t* slots, uint32_t g, uint32_t n, uint32_t* off, uint32_t* len) {
    uint32_t b = slots[2 * g], e = slots[2 * g + 1];
    if (b == LC_SLOT_UNSET || e == LC_SLOT_UNSET || e < b) {
        *off = n;
        *len = 0;
    } else {
        *off = b;
        *len = e - b;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Stride-2 two-pass matcher over the fast2 layout (lc_tables.h: LcFast2Header): reference (loop) formulation.
// The kernel's unrolled middle section is a specialisation of exactly these statements.
struct LcFast2View {
    const LcFast2Header* h;
    const uint8_t* cls;
    const uint8_t* t2; // byte addressed (u32 entries)
    const uint8_t* pid;
    const uint8_t* pair_l;
    const uint8_t* rev1;
    const uint32_t* f2;
    const uint32_t* fwd1;
    const uint64_t* masks;
};

LC_HD LcFast2View lc_fast2_view(const void* blob) {
    const uint8_t* b = (const uint8_t*)blob;
    const LcFast2Header* h = (const LcFast2Header*)blob;
    LcFast2View v;
    v.h = h;
    v.cls = b + h->off_cls;
    v.t2 = b + h->off_t2;
    v.pid = b + h->off_pid;
    v.pair_l = b + h->off_pair_l;
    v.rev1 = b + h->off_rev1;
    v.f2 = (const uint32_t*)(b + h->off_f2);
    v.fwd1 = (const uint32_t*)(b + h->off_fwd1);
    v.masks = (const uint64_t*)(b + h->off_masks);
    return v;
}

#define LC_SLOT16_UNSET 0xFFFFu

LC_HD void lc_fast2_apply_mask(const LcFast2View& v, uint32_t act, uint32_t pos, uint16_t* slots) {
    uint64_t m = v.masks[act];
    while (m) {
#if defined(__CUDA_ARCH__)
        int s = __ffsll((long long)m) - 1;
#else
        int s = __builtin_ctzll(m);
#endif
        slots[s] = (uint16_t)pos;
        m &= m - 1;
    }
}

// single forward step of walker w under label L at position pos; returns the next walker
LC_HD uint32_t lc_fast2_single(const LcFast2View& v, uint32_t w, uint32_t L, uint32_t pos, uint16_t* slots) {
    const uint32_t e = v.fwd1[w * v.h->nrev + L];
    const uint32_t a = LC_ENTRY_ACT(e);
    if (a)
        lc_fast2_apply_mask(v, a, pos, slots);
    const uint32_t nx = LC_ENTRY_NEXT(e);
    return nx == 0xFFFFu ? 0u : nx;
}

// slow path of a pair step whose first or second transition sets several slots
LC_HD void lc_fast2_pair_slow(const LcFast2View& v, uint32_t w, uint32_t P, uint32_t pos, uint16_t* slots) {
    const uint32_t La = v.pair_l[2 * P], Lb = v.pair_l[2 * P + 1];
    const uint32_t w1 = lc_fast2_single(v, w, La, pos, slots);
    (void)lc_fast2_single(v, w1, Lb, pos + 1, slots);
}

// s = first byte of the event; `mis` = virtual index of that byte (address & 15 on the device).
// lab[j] receives the pair id of positions (2j, 2j+1) in virtual indexing; needs (n + mis) / 2 + 1 bytes.
// slots: 2 * ngroups u16 entries preset to LC_SLOT16_UNSET.  n must be < 65535.
LC_HD bool lc_fast2_event(const LcFast2View& v, const uint8_t* s, uint32_t mis, uint32_t n, uint8_t* lab,
                          uint16_t* slots) {
    const uint32_t Q = n + mis;
    const uint32_t nrev = v.h->nrev, ncls = v.h->ncls, row_bytes = v.h->row_bytes;
    const uint32_t start = v.h->rev_start;
    // ---- reverse: labels of positions Q (== start) down to mis
    uint32_t d = start;
    uint32_t q = Q;
    if ((q & 1) && q > mis) { // byte q-1 is the first slot of the pair (q-1, q): second slot is the end position
        --q;
        d = v.rev1[d * ncls + v.cls[s[q - mis]]];
        if (!d)
            return false;
        lab[q / 2] = v.pid[d * nrev + start];
    }
    while (q >= mis + 2) { // full byte pair (q-2, q-1)
        q -= 2;
        const uint32_t addr = d * row_bytes + ((uint32_t)v.cls[s[q + 1 - mis]] * ncls + v.cls[s[q - mis]]) * 4;
        const uint32_t e = *(const uint32_t*)(v.t2 + addr);
        lab[q / 2] = (uint8_t)(e >> 16);
        d = (e & 0xFFFFu) / row_bytes;
        if (!d)
            return false;
    }
    if (q > mis) { // one byte left: it sits in the second slot of a pair whose first slot precedes the event
        --q;
        d = v.rev1[d * ncls + v.cls[s[q - mis]]];
        if (!d)
            return false;
    }
    // ---- forward; d == label of position mis
    if (v.fwd1[d] == LC_NONE_ENTRY) // START row
        return false;
    uint32_t w = 0;
    q = mis;
    if (q & 1) {
        w = lc_fast2_single(v, w, d, 0, slots);
        ++q;
    }
    const uint32_t pshift = v.h->pair_shift, f2row = v.h->f2_row;
    while (q + 1 <= Q) {
        const uint32_t P = lab[q / 2] >> pshift; // labels hold pair_id << pair_shift
        const uint32_t e = v.f2[w * f2row + P];
        const uint32_t sa = (e >> 8) & 0x7Fu, sb = (e >> 16) & 0x7Fu;
        if (e & LC_FAST2_ACT_MULTI) {
            lc_fast2_pair_slow(v, w, P, q - mis, slots);
        } else {
            if (sa)
                slots[(sa - 2) / 2] = (uint16_t)(q - mis);
            if (sb)
                slots[(sb - 2) / 2] = (uint16_t)(q + 1 - mis);
        }
        w = e & 0xFFu;
        q += 2;
    }
    if (q == Q)
        (void)lc_fast2_single(v, w, start, q - mis, slots);
    return true;
}

LC_HD void lc_slots16_to_cap(const uint16_t* slots, uint32_t g, uint32_t n, uint32_t* off, uint32_t* len) {
    uint32_t b = slots[2 * g], e = slots[2 * g + 1];
    if (b == LC_SLOT16_UNSET || e == LC_SLOT16_UNSET || e < b) {
        *off = n;
        *len = 0;
    } else {
        *off = b;
        *len = e - b;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Single-pass tagged DFA (lc_tables.h: LcTdfaHeader): reference (loop) formulation; the kernel's unrolled middle
// section is a specialisation of exactly these statements.
struct LcTdfaView {
    const LcTdfaHeader* h;
    const uint8_t* cls;
    const uint8_t* t2; // byte addressed (u32 entries)
    const uint32_t* t1;
    const uint32_t* eof;
    const uint16_t* ops;
    const uint32_t* skip;
};

LC_HD LcTdfaView lc_tdfa_view(const void* blob) {
    const uint8_t* b = (const uint8_t*)blob;
    const LcTdfaHeader* h = (const LcTdfaHeader*)blob;
    LcTdfaView v;
    v.h = h;
    v.cls = b + h->off_cls;
    v.t2 = b + h->off_t2;
    v.t1 = (const uint32_t*)(b + h->off_t1);
    v.eof = (const uint32_t*)(b + h->off_eof);
    v.ops = (const uint16_t*)(b + h->off_ops);
    v.skip = (const uint32_t*)(b + h->off_skip);
    return v;
}

// does the 16-byte chunk w[0..3] hold one of the exit bytes of skip word sk (LC_TDFA_SKIP set)?
LC_HD bool lc_tdfa_chunk_has_exit(uint32_t sk, const uint32_t w[4]) {
    const uint32_t n = (sk >> 16) & 3u;
    uint32_t hit = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t splat = ((sk >> (8 * k)) & 0xFFu) * 0x01010101u;
        for (int j = 0; j < 4; ++j) {
            const uint32_t x = w[j] ^ splat;
            hit |= (x - 0x01010101u) & ~x & 0x80808080u; // some byte of x is zero
        }
    }
    return hit != 0;
}

// RegT = uint16_t (events shorter than 65535 bytes: the shared-memory register files of the kernels) or uint32_t
// (the long-event kernel); "unset" is the all-ones value of RegT.
template <class RegT>
LC_HD void lc_tdfa_run_ops(const LcTdfaView& v, uint32_t list, uint32_t pos, RegT* regs) {
    const uint16_t* p = v.ops + list;
    const uint32_t cnt = p[0];
    for (uint32_t k = 1; k <= cnt; ++k) {
        const uint32_t op = p[k], dst = op >> 8, src = op & 0xFFu;
        regs[dst] = src == LC_TDFA_SRC_POS ? (RegT)pos : src == LC_TDFA_SRC_UNSET ? (RegT)~(RegT)0 : regs[src];
    }
}

// one byte from `state` at position pos; returns the next state (0 = dead)
template <class RegT>
LC_HD uint32_t lc_tdfa_single(const LcTdfaView& v, uint32_t state, uint32_t byte, uint32_t pos, RegT* regs) {
    const uint32_t e = v.t1[state * v.h->ncls + v.cls[byte]];
    if (e >> 16)
        lc_tdfa_run_ops(v, e >> 16, pos, regs);
    return e & 0xFFFFu;
}

// s = first byte of the event; `mis` = virtual index of that byte (address & 15 on the device; pairs are aligned
// on even virtual indices).  regs: h->nregs u16 entries, the first 2 * ngroups preset to LC_SLOT16_UNSET; on a
// match they hold the capture boundaries.  n must be < 65535 for RegT = uint16_t.
template <class RegT>
LC_HD bool lc_tdfa_event(const LcTdfaView& v, const uint8_t* s, uint32_t mis, uint32_t n, RegT* regs) {
    const uint32_t ncls = v.h->ncls, row_bytes = v.h->row_bytes;
    uint32_t st = v.h->start;
    uint32_t pos = 0;
    if ((mis & 1) && n) {
        st = lc_tdfa_single(v, st, s[0], 0, regs);
        pos = 1;
    }
    while (pos + 2 <= n) {
        // run skipping exactly as the kernels do it: at a 16-byte boundary of the line's aligned frame, with a whole
        // chunk of input left, a skippable state jumps over a chunk that holds none of its exit bytes
        if (((pos + mis) & 15u) == 0 && pos + 16 <= n && v.skip[st]) {
            uint32_t w[4];
            memcpy(w, s + pos, 16);
            if (!lc_tdfa_chunk_has_exit(v.skip[st], w)) {
                pos += 16;
                continue;
            }
        }
        const uint32_t c0 = v.cls[s[pos]], c1 = v.cls[s[pos + 1]];
        const uint32_t e = *(const uint32_t*)(v.t2 + st * row_bytes + (c0 * ncls + c1) * 4);
        if (e & LC_TDFA_SLOW) {
            const uint32_t s1 = lc_tdfa_single(v, st, s[pos], pos, regs);
            st = lc_tdfa_single(v, s1, s[pos + 1], pos + 1, regs);
        } else {
            const uint32_t sa = (e >> 16) & 0x7Fu, sb = (e >> 24) & 0x7Fu;
            if (sa)
                regs[(sa - 2) / 2] = (RegT)pos;
            if (sb)
                regs[(sb - 2) / 2] = (RegT)(pos + 1);
            st = (e & 0xFFFFu) / row_bytes;
        }
        pos += 2;
    }
    if (pos < n) {
        st = lc_tdfa_single(v, st, s[pos], pos, regs);
        ++pos;
    }
    const uint32_t fin = v.eof[st];
    if (fin == LC_NONE_ENTRY)
        return false;
    lc_tdfa_run_ops(v, fin, n, regs);
    return true;
}

// ------------------------------------------------------------------------------------------------------------
// Delimiter quote FSM (DelimiterModeFsmParser::ParseDelimiterLine, zero-copy variant,
// core/parser/DelimiterModeFsmParser.cpp:260-294) with run skipping: per 16-byte aligned chunk the separator and
// quote positions are found with byte-wise SIMD compares; only those "special" bytes go through the state
// machine, every run of ordinary bytes between them is one step (the ordinary-byte transition is idempotent:
// INITIAL -> DATA, QUOTE / DATA stay, DOUBLE_QUOTE is an error).  Exactly the per-byte machine otherwise.
//   v = line base; [begin, end) = trimmed range; push(field_start, field_len, doubled_quotes) receives every column.
// Reads the aligned 16-byte chunks that contain bytes of [begin, end) (like the kernel's other loaders).
LC_HD uint32_t lc_eq_mask16(const uint32_t w[4], uint32_t splat) {
    uint32_t m = 0;
    for (int k = 0; k < 4; ++k) {
        // bit 0 of every byte of w[k] that equals the splat byte: exact zero-byte test of x (no cross-byte borrows);
        // plain integer ops on both sides (the SIMD-video compare of the GPU is emulated and costs more)
        const uint32_t x = w[k] ^ splat;
        const uint32_t eq = (~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u) >> 7;
        m |= (((eq * 0x01020408u) >> 24) & 0xFu) << (4 * k);
    }
    return m;
}

// Resumable form: the machine's registers + one call per 16-byte aligned chunk, so that a kernel can feed the chunks
// from a staging tile in shared memory stage by stage.  All positions are offsets from the line start `v`; q* are
// positions in the line's 16-byte aligned frame (frame index = offset + mis).
struct LcDelimRun {
    int state; // 0 INITIAL 1 QUOTE 2 DATA 3 DOUBLE_QUOTE
    int dq;
    int fs, fe;
    uint32_t cur; // next unprocessed frame position
};

LC_HD void lc_delim_start(LcDelimRun& r, int32_t begin, uint32_t mis) {
    r.state = 0;
    r.dq = 0;
    r.fs = r.fe = begin;
    r.cur = (uint32_t)begin + mis;
}

// chunk = the 16 bytes at frame positions [q0, q0 + 16), restricted to [qb, qe).  Returns false on a parse error.
//
// Only separators and quotes ("specials") step the machine; the run of ordinary bytes in front of each special is one
// step.  The step itself is table-driven and branch-free apart from the push: on the GPU the 32 lanes of a warp walk
// 32 different lines, and a formulation with one code path per (state, byte kind) made the warp execute most paths for
// every special (ncu: 1-6 active lanes on the hot instructions).  Entry (state * 2 + kind), kind 0 = separator,
// 1 = quote:  bits 0-1 next state | bits 2-3 fe += | bit 4 push the column first | bit 5 dq-- before the push |
// bit 6 error | bit 7 fs++ | bit 8 dq++.  After a push: dq = 0 and fs = the new fe.
//   INITIAL  sep: push, fe+1 -> INITIAL          quote: fs++ -> QUOTE
//   QUOTE    sep: fe+1 (part of the field)       quote: dq++, fe+1 -> DOUBLE_QUOTE
//   DATA     sep: push, fe+1 -> INITIAL          quote: error
//   DQUOTE   sep: dq--, push, fe+2 -> INITIAL    quote: fe+1 -> QUOTE (it was an escaped quote)
#define LC_DELIM_T(next, dfe, push, dqdec, err, fsinc, dqinc)                                                           \
    ((uint64_t)((next) | (dfe) << 2 | (push) << 4 | (dqdec) << 5 | (err) << 6 | (fsinc) << 7 | (dqinc) << 8))
template <class Push>
LC_HD bool lc_delim_chunk(LcDelimRun& r, const uint32_t w[4], uint32_t q0, uint32_t qb, uint32_t qe, uint32_t sep_splat,
                          uint32_t quote_splat, Push& push) {
    const uint64_t T_LO = LC_DELIM_T(0, 1, 1, 0, 0, 0, 0) | LC_DELIM_T(1, 0, 0, 0, 0, 1, 0) << 16 |
                          LC_DELIM_T(1, 1, 0, 0, 0, 0, 0) << 32 | LC_DELIM_T(3, 1, 0, 0, 0, 0, 1) << 48;
    const uint64_t T_HI = LC_DELIM_T(0, 1, 1, 0, 0, 0, 0) | LC_DELIM_T(2, 0, 0, 0, 1, 0, 0) << 16 |
                          LC_DELIM_T(0, 2, 1, 1, 0, 0, 0) << 32 | LC_DELIM_T(1, 1, 0, 0, 0, 0, 0) << 48;
    const uint32_t ms = lc_eq_mask16(w, sep_splat);
    uint32_t special = ms | lc_eq_mask16(w, quote_splat);
    if (q0 < qb)
        special &= ~((1u << (qb - q0)) - 1u);
    if (qe - q0 < 16)
        special &= (1u << (qe - q0)) - 1u;
    const uint32_t chunk_end = qe - q0 < 16 ? qe : q0 + 16;
    int state = r.state, dq = r.dq, fs = r.fs, fe = r.fe;
    uint32_t cur = r.cur, bad = 0;
    while (special) {
#if defined(__CUDA_ARCH__)
        const int b = __ffs((int)special) - 1;
#else
        const int b = __builtin_ctz(special);
#endif
        special &= special - 1;
        const uint32_t next = q0 + (uint32_t)b;
        const uint32_t gap = next - cur; // run of ordinary bytes in front of the special
        bad |= (gap != 0 && state == 3) ? 1u : 0u;
        state = (gap != 0 && state == 0) ? 2 : state;
        fe += (int)gap;
        cur = next + 1;
        const uint32_t idx = (uint32_t)state * 2u + (((ms >> b) & 1u) ^ 1u); // a byte equal to both counts as separator
        const uint32_t e = (uint32_t)((idx < 4 ? T_LO : T_HI) >> ((idx & 3u) * 16u)) & 0xFFFFu;
        if (e & 0x10u) {
            dq -= (int)((e >> 5) & 1u);
            if (!bad)
                push((uint32_t)fs, (uint32_t)(fe - fs), (uint32_t)dq);
            dq = 0;
        }
        fe += (int)((e >> 2) & 3u);
        dq += (int)((e >> 8) & 1u);
        bad |= (e >> 6) & 1u;
        fs = (e & 0x10u) ? fe : fs + (int)((e >> 7) & 1u);
        state = (int)(e & 3u);
    }
    {
        const uint32_t gap = chunk_end - cur; // ordinary bytes behind the last special of the chunk
        bad |= (gap != 0 && state == 3) ? 1u : 0u;
        state = (gap != 0 && state == 0) ? 2 : state;
        fe += (int)gap;
        cur = chunk_end;
    }
    r.state = state;
    r.dq = dq;
    r.fs = fs;
    r.fe = fe;
    r.cur = cur;
    return bad == 0;
}

// end of line: closes the last column; false = unterminated quote
template <class Push>
LC_HD bool lc_delim_finish(LcDelimRun& r, Push& push) {
    if (r.state == 3)
        r.dq--;
    if (r.state == 1)
        return false;
    push((uint32_t)r.fs, (uint32_t)(r.fe - r.fs), (uint32_t)r.dq);
    return true;
}

template <class Push>
LC_HD bool lc_delim_fsm(const uint8_t* v, int32_t begin, int32_t end, uint8_t sep, uint8_t quote, Push& push) {
    const uint32_t sep_splat = sep * 0x01010101u, quote_splat = quote * 0x01010101u;
    const uint32_t mis = (uint32_t)((uintptr_t)v & 15u);
    const uint8_t* abase = v - mis;
    const uint32_t qb = (uint32_t)begin + mis, qe = (uint32_t)end + mis; // range in the aligned frame
    LcDelimRun r;
    lc_delim_start(r, begin, mis);
    for (uint32_t qc = qb >> 4; end > begin && qc <= ((qe - 1) >> 4); ++qc) {
        uint32_t w[4];
#if defined(__CUDA_ARCH__)
        const uint4 vv = __ldg(reinterpret_cast<const uint4*>(abase) + qc);
        w[0] = vv.x, w[1] = vv.y, w[2] = vv.z, w[3] = vv.w;
#else
        memcpy(w, abase + (size_t)qc * 16, 16);
#endif
        if (!lc_delim_chunk(r, w, qc * 16, qb, qe, sep_splat, quote_splat, push))
            return false;
    }
    return lc_delim_finish(r, push);
}

// ---- bit-parallel form for well-formed records ---------------------------------------------------------------------
// The machine above steps once per separator / quote, and with 32 different lines in a warp the loop over the specials
// of a chunk runs as often as the busiest lane needs (ncu: the kernel issues ~40 instructions per input byte).  For a
// record in which every quote sits where the machine accepts one, the columns follow from three masks per chunk:
//   S = separators, Q = quotes, P = inclusive prefix-XOR of Q (with the carry of the chunks before): P is 1 from an
//   opening quote up to the byte before its closing quote, so the REAL separators are S & ~P, an opening quote is
//   Q & P and a closing quote is Q & ~P.
// Well-formed (exactly the paths of the machine that do not end in an error):
//   * an opening quote is the first byte of a column (record start or behind a real separator) or directly follows a
//     closing quote (the escaped quote "" inside a quoted column: DOUBLE_QUOTE -> QUOTE);
//   * a closing quote is followed by a real separator, by another quote, or by the end of the record
//     (DOUBLE_QUOTE + ordinary byte is the machine's error, DATA + quote likewise);
//   * the record does not end inside a quoted column.
// A column with quotes is then `"` content `"`: content = [first + 1, last - 1), doubled quotes = (quotes - 2) / 2.
// Anything else returns false and the caller runs lc_delim_fsm on the record (which also decides about errors).
struct LcDelimFast {
    uint32_t inq;        // 0 / all ones: inside a quoted section where the step starts
    uint32_t prev_rs;    // the byte before the step's first byte was a real separator (or the record starts here)
    uint32_t prev_close; // ... was a closing quote
    uint32_t nq;         // quotes of the open column so far
};

LC_HD void lc_delim_fast_start(LcDelimFast& r) {
    r.inq = 0;
    r.prev_rs = 1;
    r.prev_close = 0;
    r.nq = 0;
}

LC_HD uint32_t lc_popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
LC_HD uint32_t lc_low_bits(uint32_t n) { return n >= 32 ? 0xFFFFFFFFu : (1u << n) - 1u; } // the n lowest bits

// One step over W (16 or 32) bytes of the record's 16-byte aligned frame, [q0, q0 + W).  ms / mq: masks of the bytes
// equal to the separator / quote.  mark(p, quotes) is called for every real separator: p = its offset from the line
// start, quotes = the number of quotes of the column it closes (a caller that wants columns derives them: the column
// starts behind the previous mark; with quotes it is `"` content `"`, doubled quotes = (quotes - 2) / 2).
template <int W, class Mark>
LC_HD bool lc_delim_fast_step(LcDelimFast& r, uint32_t ms, uint32_t mq, uint32_t q0, uint32_t qb, uint32_t qe,
                              uint32_t mis, Mark& mark) {
    const uint32_t b0 = q0 < qb ? qb - q0 : 0u;              // first byte of the step that belongs to the record
    const uint32_t nb = qe - q0 < (uint32_t)W ? qe - q0 : W; // bytes of the step in front of the record's end
    const uint32_t V = lc_low_bits(nb) & ~lc_low_bits(b0);   // (nb > b0: the caller only passes steps with bytes)
    const uint32_t S = ms & V, Q = mq & V;
    uint32_t P = Q;
    P ^= P << 1;
    P ^= P << 2;
    P ^= P << 4;
    P ^= P << 8;
    if (W > 16)
        P ^= P << 16;
    P = (P ^ r.inq) & lc_low_bits(W);
    const uint32_t RS = S & ~P, open = Q & P, close = Q & ~P;
    const uint32_t first = 1u << b0, last = 1u << (nb - 1);
    const bool ends = qe - q0 <= (uint32_t)W; // the record ends inside this step
    // opening quotes: behind a real separator / at the record start, or behind a closing quote
    uint32_t bad = open & ~(((RS | close) << 1) | ((r.prev_rs | r.prev_close) ? first : 0u));
    // closing quotes: in front of a real separator or a quote; the record's last byte may be one; the successor of
    // the step's last byte is checked with the next step (prev_close)
    bad |= close & (ends ? V : (V & ~last)) & ~(((RS | Q) >> 1) | (ends ? last : 0u));
    if (r.prev_close && !((RS | Q) & first))
        bad |= 1u;
    if (bad)
        return false;
    uint32_t m = RS, cb = b0, nq = r.nq;
    while (m) {
#if defined(__CUDA_ARCH__)
        const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
#else
        const uint32_t b = (uint32_t)__builtin_ctz(m);
#endif
        m &= m - 1;
        mark((uint32_t)(q0 + b - mis), nq + lc_popc32(Q & lc_low_bits(b) & ~lc_low_bits(cb)));
        nq = 0;
        cb = b + 1;
    }
    r.nq = nq + lc_popc32(Q & ~lc_low_bits(cb));
    r.inq = (P & last) ? 0xFFFFFFFFu : 0u;
    r.prev_rs = (RS & last) ? 1u : 0u;
    r.prev_close = (close & last) ? 1u : 0u;
    return true;
}

// end of the record: the mark that closes the last column; false = the record ends inside a quoted column
template <class Mark>
LC_HD bool lc_delim_fast_finish(LcDelimFast& r, int32_t end, Mark& mark) {
    if (r.inq)
        return false;
    mark((uint32_t)end, r.nq);
    return true;
}

// column [start, p) that carries `quotes` quotes -> push(first byte, length, doubled quotes)
template <class Push>
LC_HD void lc_delim_fast_column(uint32_t start, uint32_t p, uint32_t quotes, Push& push) {
    if (quotes)
        push(start + 1, p - start - 2, (quotes - 2) >> 1);
    else
        push(start, p - start, 0u);
}

// whole record, 32 bytes per step as in the kernel (CPU-tier tests).  Returns false when the record is not well-formed
// in the sense above -- columns pushed so far are then to be discarded.
template <class Push>
LC_HD bool lc_delim_fast(const uint8_t* v, int32_t begin, int32_t end, uint8_t sep, uint8_t quote, Push& push) {
    const uint32_t sep_splat = sep * 0x01010101u, quote_splat = quote * 0x01010101u;
    const uint32_t mis = (uint32_t)((uintptr_t)v & 15u);
    const uint8_t* abase = v - mis;
    const uint32_t qb = (uint32_t)begin + mis, qe = (uint32_t)end + mis;
    LcDelimFast r;
    lc_delim_fast_start(r);
    uint32_t start = (uint32_t)begin;
    auto mark = [&](uint32_t p, uint32_t quotes) {
        lc_delim_fast_column(start, p, quotes, push);
        start = p + 1;
    };
    for (uint32_t qc = (qb >> 4) & ~1u; end > begin && qc <= ((qe - 1) >> 4); qc += 2) {
        uint32_t ms = 0, mq = 0;
        for (uint32_t h = 0; h < 2; ++h) {
            if ((qc + h) * 16 >= qe || (qc + h + 1) * 16 <= qb)
                continue; // no byte of the record in this half: not read
            uint32_t w[4];
#if defined(__CUDA_ARCH__)
            const uint4 vv = __ldg(reinterpret_cast<const uint4*>(abase) + qc + h);
            w[0] = vv.x, w[1] = vv.y, w[2] = vv.z, w[3] = vv.w;
#else
            memcpy(w, abase + (size_t)(qc + h) * 16, 16);
#endif
            ms |= lc_eq_mask16(w, sep_splat) << (16 * h);
            mq |= lc_eq_mask16(w, quote_splat) << (16 * h);
        }
        if (!lc_delim_fast_step<32>(r, ms, mq, qc * 16, qb, qe, mis, mark))
            return false;
    }
    return lc_delim_fast_finish(r, end, mark);
}

// ------------------------------------------------------------------------------------------------------------
// SLS wire format of LOG events (next row): the hand-rolled protobuf writer of
// core/protobuf/sls/LogGroupSerializer.cpp:33-143,232-262 as size and emit functions over spans of the arena.
//   Log record = 0x0A varint(body) body ; body = 0x08 varint5(max(time, 2^28)) { 0x12 varint(pair) pair }* [0x25 ns:4]
//   pair       = 0x0A varint(klen) key 0x12 varint(vlen) value
LC_HD uint32_t lc_varint_size(uint32_t v) { return v < (1u << 7) ? 1 : v < (1u << 14) ? 2 : v < (1u << 21) ? 3 : v < (1u << 28) ? 4 : 5; }

LC_HD uint32_t lc_put_varint(uint8_t* p, uint32_t v) {
    uint32_t n = 0;
    while (v >= 0x80u) {
        p[n++] = (uint8_t)(v | 0x80u);
        v >>= 7;
    }
    p[n++] = (uint8_t)v;
    return n;
}

// GetLogContentSize (:232-237) without the tag + length prefix
LC_HD uint32_t lc_sls_pair_inner(uint32_t klen, uint32_t vlen) {
    return 1 + lc_varint_size(klen) + klen + 1 + lc_varint_size(vlen) + vlen;
}

// GetLogSize (:239-253): body = size inside the Logs field, return = with tag and length prefix; 0 entries -> 0, 0
LC_HD uint32_t lc_sls_log_size(const uint32_t* klen, const uint32_t* vlen, uint64_t e0, uint64_t e1, bool has_ns,
                               uint32_t* body_out) {
    if (e0 == e1) {
        *body_out = 0;
        return 0;
    }
    uint32_t body = 1 + 5 + (has_ns ? 1 + 4 : 0);
    for (uint64_t k = e0; k < e1; ++k) {
        const uint32_t in = lc_sls_pair_inner(klen[k], vlen[k]);
        body += 1 + lc_varint_size(in) + in;
    }
    *body_out = body;
    return 1 + lc_varint_size(body) + body;
}

// Writes one Log record at `out` (which has room for lc_sls_log_size bytes).  Called by all `nlanes` cooperating
// lanes with identical arguments except `lane`: lane 0 writes the tags / lengths, every lane copies its share of the
// key and value bytes.
LC_HD void lc_sls_emit_log(uint8_t* out, const uint8_t* base, uint32_t time, bool has_ns, uint32_t ns,
                           const uint32_t* koff, const uint32_t* klen, const uint32_t* voff, const uint32_t* vlen,
                           uint64_t e0, uint64_t e1, uint32_t body, uint32_t lane, uint32_t nlanes) {
    uint32_t at = 0;
    uint8_t hdr[16];
    uint32_t h = 0;
    hdr[h++] = 0x0A;
    h += lc_put_varint(hdr + h, body);
    hdr[h++] = 0x08;
    h += lc_put_varint(hdr + h, time < (1u << 28) ? (1u << 28) : time); // always 5 bytes
    if (lane == 0)
        for (uint32_t j = 0; j < h; ++j)
            out[j] = hdr[j];
    at = h;
    for (uint64_t k = e0; k < e1; ++k) {
        const uint32_t kl = klen[k], vl = vlen[k];
        h = 0;
        hdr[h++] = 0x12;
        h += lc_put_varint(hdr + h, lc_sls_pair_inner(kl, vl));
        hdr[h++] = 0x0A;
        h += lc_put_varint(hdr + h, kl);
        if (lane == 0)
            for (uint32_t j = 0; j < h; ++j)
                out[at + j] = hdr[j];
        at += h;
        for (uint32_t j = lane; j < kl; j += nlanes)
            out[at + j] = base[koff[k] + j];
        at += kl;
        h = 0;
        hdr[h++] = 0x12;
        h += lc_put_varint(hdr + h, vl);
        if (lane == 0)
            for (uint32_t j = 0; j < h; ++j)
                out[at + j] = hdr[j];
        at += h;
        for (uint32_t j = lane; j < vl; j += nlanes)
            out[at + j] = base[voff[k] + j];
        at += vl;
    }
    if (has_ns && lane == 0) {
        out[at] = 0x25;
        out[at + 1] = (uint8_t)ns;
        out[at + 2] = (uint8_t)(ns >> 8);
        out[at + 3] = (uint8_t)(ns >> 16);
        out[at + 4] = (uint8_t)(ns >> 24);
    }
}

// ---- the same writer over an abstract contents list (entry k = key bytes + value bytes), so that the contents can come
// from the regex stage's capture tables instead of host-built span lists.  En: uint32_t klen(k), vlen(k);
// const uint8_t* key(k), val(k).
template <class En>
LC_HD uint32_t lc_sls_log_size_t(const En& en, uint32_t count, bool has_ns, uint32_t* body_out) {
    if (count == 0) {
        *body_out = 0;
        return 0;
    }
    uint32_t body = 1 + 5 + (has_ns ? 1 + 4 : 0);
    for (uint32_t k = 0; k < count; ++k) {
        const uint32_t in = lc_sls_pair_inner(en.klen(k), en.vlen(k));
        body += 1 + lc_varint_size(in) + in;
    }
    *body_out = body;
    return 1 + lc_varint_size(body) + body;
}

template <class En>
LC_HD void lc_sls_emit_log_t(uint8_t* out, uint32_t time, bool has_ns, uint32_t ns, const En& en, uint32_t count,
                             uint32_t body, uint32_t lane, uint32_t nlanes) {
    uint32_t at = 0;
    uint8_t hdr[16];
    uint32_t h = 0;
    hdr[h++] = 0x0A;
    h += lc_put_varint(hdr + h, body);
    hdr[h++] = 0x08;
    h += lc_put_varint(hdr + h, time < (1u << 28) ? (1u << 28) : time); // always 5 bytes
    if (lane == 0)
        for (uint32_t j = 0; j < h; ++j)
            out[j] = hdr[j];
    at = h;
    for (uint32_t k = 0; k < count; ++k) {
        const uint32_t kl = en.klen(k), vl = en.vlen(k);
        const uint8_t* kp = en.key(k);
        const uint8_t* vp = en.val(k);
        h = 0;
        hdr[h++] = 0x12;
        h += lc_put_varint(hdr + h, lc_sls_pair_inner(kl, vl));
        hdr[h++] = 0x0A;
        h += lc_put_varint(hdr + h, kl);
        if (lane == 0)
            for (uint32_t j = 0; j < h; ++j)
                out[at + j] = hdr[j];
        at += h;
        for (uint32_t j = lane; j < kl; j += nlanes)
            out[at + j] = kp[j];
        at += kl;
        h = 0;
        hdr[h++] = 0x12;
        h += lc_put_varint(hdr + h, vl);
        if (lane == 0)
            for (uint32_t j = 0; j < h; ++j)
                out[at + j] = hdr[j];
        at += h;
        for (uint32_t j = lane; j < vl; j += nlanes)
            out[at + j] = vp[j];
        at += vl;
    }
    if (has_ns && lane == 0) {
        out[at] = 0x25;
        out[at + 1] = (uint8_t)ns;
        out[at + 2] = (uint8_t)(ns >> 8);
        out[at + 3] = (uint8_t)(ns >> 16);
        out[at + 4] = (uint8_t)(ns >> 24);
    }
}

