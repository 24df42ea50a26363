// lc_exec.cuh -- scalar interpretation of a compiled regex blob (lc_tables.h): the semantic core
// that the sm_100a kernels execute once per log line.  Written as __host__ __device__ so that the
// exact same statements can be exercised on the CPU by the test-only emulation library under
// tests/emul/ (which validates the COMPILER's tables against the oracle without a GPU).  The product
// library only ever calls these from device code.
#pragma once
#include <stdint.h>

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
