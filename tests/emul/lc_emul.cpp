// lc_emul.cpp -- TEST-ONLY host emulation of the device-side table interpreter.
//
// Builds the product's regex COMPILER (loongcollector_b200/csrc/regex_compiler.cpp) together with the
// __host__ __device__ interpreter statements of lc_exec.cuh into a CPU shared object so that the
// "not gpu" test tier can check compiler + table semantics against the oracle without a B200.
// It is NOT part of the product library, is never loaded by loongcollector_b200/, and nothing here is a
// CPU fallback: the product C-ABI has no path that reaches this file.
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../loongcollector_b200/csrc/lc_exec.cuh"
#include "../../loongcollector_b200/csrc/regex_compiler.h"

struct EmulRegex {
    lcb200::CompileResult r;
};

extern "C" {

EmulRegex* emul_compile(const char* pattern, uint64_t len) {
    EmulRegex* e = new EmulRegex;
    e->r = lcb200::compile_regex(pattern, (size_t)len);
    return e;
}
void emul_free(EmulRegex* e) { delete e; }
int emul_valid(const EmulRegex* e) { return e->r.valid; }
int emul_supported(const EmulRegex* e) { return e->r.supported; }
const char* emul_error(const EmulRegex* e) { return e->r.error.c_str(); }
uint32_t emul_ngroups(const EmulRegex* e) { return e->r.ngroups; }
// info[0..7] = mode, nclasses, nw, npc, rev_nstates, pre_nstates, blob bytes, n_insts
void emul_info(const EmulRegex* e, uint32_t* info) {
    memset(info, 0, 8 * sizeof(uint32_t));
    if (!e->r.supported)
        return;
    const LcRegexHeader* h = (const LcRegexHeader*)e->r.blob.data();
    info[0] = h->mode;
    info[1] = h->nclasses;
    info[2] = h->nw;
    info[3] = h->npc;
    info[4] = h->rev_nstates;
    info[5] = h->pre_nstates;
    info[6] = h->total_bytes;
    info[7] = e->r.n_insts;
}
int emul_prefix_match(const EmulRegex* e, const uint8_t* s, uint32_t n) {
    LcProgView v = lc_view(e->r.blob.data());
    return lc_prefix_match(v, s, n) ? 1 : 0;
}
// stride-2 layout: returns -1 when the pattern has no fast2 blob; `mis` emulates the device-side 16-byte misalignment
int emul_full_match_fast2(const EmulRegex* e, const uint8_t* s, uint32_t n, uint32_t mis, uint32_t* cap_off,
                          uint32_t* cap_len) {
    if (e->r.fast2_blob.empty())
        return -1;
    LcFast2View v = lc_fast2_view(e->r.fast2_blob.data());
    std::vector<uint8_t> lab((n + mis) / 2 + 2, 0);
    uint16_t slots[2 * LC_MAX_GROUPS];
    for (uint32_t k = 0; k < 2 * LC_MAX_GROUPS; ++k)
        slots[k] = LC_SLOT16_UNSET;
    if (!lc_fast2_event(v, s, mis, n, lab.data(), slots))
        return 0;
    for (uint32_t g = 0; g < v.h->ngroups; ++g)
        lc_slots16_to_cap(slots, g, n, cap_off + g, cap_len + g);
    return 1;
}
// single-pass tagged DFA: returns -1 when the pattern has no tdfa blob
int emul_full_match_tdfa(const EmulRegex* e, const uint8_t* s, uint32_t n, uint32_t mis, uint32_t* cap_off,
                         uint32_t* cap_len) {
    if (e->r.tdfa_blob.empty())
        return -1;
    LcTdfaView v = lc_tdfa_view(e->r.tdfa_blob.data());
    uint16_t regs[64];
    for (uint32_t k = 0; k < 64; ++k)
        regs[k] = LC_SLOT16_UNSET;
    if (!lc_tdfa_event(v, s, mis, n, regs))
        return 0;
    for (uint32_t g = 0; g < v.h->ngroups; ++g)
        lc_slots16_to_cap(regs, g, n, cap_off + g, cap_len + g);
    return 1;
}
// info[0..5] = states, classes, registers, max threads per state, has_slow, table bytes
void emul_tdfa_info(const EmulRegex* e, uint32_t* info) {
    memset(info, 0, 6 * sizeof(uint32_t));
    if (e->r.tdfa_blob.empty())
        return;
    const LcTdfaHeader* h = (const LcTdfaHeader*)e->r.tdfa_blob.data();
    info[0] = h->nstates;
    info[1] = h->ncls;
    info[2] = h->nregs;
    info[3] = h->max_threads;
    info[4] = h->has_slow;
    info[5] = h->total_bytes;
}
uint32_t emul_fast2_bytes(const EmulRegex* e) { return (uint32_t)e->r.fast2_blob.size(); }
uint32_t emul_fast_bytes(const EmulRegex* e) { return (uint32_t)e->r.fast_blob.size(); }

int emul_full_match(const EmulRegex* e, const uint8_t* s, uint32_t n, uint32_t* cap_off, uint32_t* cap_len) {
    LcProgView v = lc_view(e->r.blob.data());
    uint32_t slots[2 * LC_MAX_GROUPS];
    for (uint32_t k = 0; k < 2 * LC_MAX_GROUPS; ++k)
        slots[k] = LC_SLOT_UNSET;
    bool ok;
    if (v.h->mode == LC_MODE_FWD1) {
        ok = lc_full_match_fwd1(v, s, n, slots);
    } else {
        std::vector<uint16_t> lab(n + 1);
        ok = lc_rev_label(v, s, n, lab.data()) && lc_fwd_walk(v, s, n, lab.data(), slots);
    }
    if (!ok)
        return 0;
    for (uint32_t g = 0; g < v.h->ngroups; ++g)
        lc_slots_to_cap(slots, g, n, cap_off + g, cap_len + g);
    return 1;
}
}

extern "C" {
// copies the tdfa blob (diagnostics / table inspection in tests); returns its size
uint32_t emul_tdfa_blob(const EmulRegex* e, uint8_t* out, uint32_t cap) {
    uint32_t n = (uint32_t)e->r.tdfa_blob.size();
    if (out && cap >= n)
        memcpy(out, e->r.tdfa_blob.data(), n);
    return n;
}
}

extern "C" {
// run-skipping delimiter FSM of the kernels (lc_exec.cuh: lc_delim_fsm); same contract as the oracle's orc_delim_fsm:
// returns the column count or -1 on an FSM error; writes at most cap columns.  `line` must sit in a buffer with 16
// readable bytes before and after it (aligned 16-byte chunks are read).
int64_t emul_delim_fsm(const uint8_t* line, int32_t begin, int32_t end, uint8_t sep, uint8_t quote, uint32_t* f_off,
                       uint32_t* f_len, uint32_t* f_dq, int64_t cap) {
    int64_t n = 0;
    auto push = [&](uint32_t o, uint32_t l, uint32_t dq) {
        if (n < cap) {
            f_off[n] = o;
            f_len[n] = l;
            f_dq[n] = dq;
        }
        ++n;
    };
    return lc_delim_fsm(line, begin, end, sep, quote, push) ? n : -1;
}
}

extern "C" {
// bit-parallel form of the same machine for well-formed records (lc_exec.cuh: lc_delim_fast): column count, or -2 when
// the record has to go through lc_delim_fsm (which then also decides about errors)
int64_t emul_delim_fast(const uint8_t* line, int32_t begin, int32_t end, uint8_t sep, uint8_t quote, uint32_t* f_off,
                        uint32_t* f_len, uint32_t* f_dq, int64_t cap) {
    int64_t n = 0;
    auto push = [&](uint32_t o, uint32_t l, uint32_t dq) {
        if (n < cap) {
            f_off[n] = o;
            f_len[n] = l;
            f_dq[n] = dq;
        }
        ++n;
    };
    return lc_delim_fast(line, begin, end, sep, quote, push) ? n : -2;
}
}

extern "C" {
// SLS wire format, host build of the kernels' size / emit functions (lc_exec.cuh), one "lane".  Returns the total
// size; writes when out_cap suffices.  Same contract as lc_sls_serialize_logs.
uint64_t emul_sls_serialize_logs(const uint8_t* base, uint64_t n, const uint32_t* ev_time, const uint32_t* ev_ns,
                                 const uint64_t* ent_begin, const uint32_t* koff, const uint32_t* klen,
                                 const uint32_t* voff, const uint32_t* vlen, uint8_t* out, uint64_t out_cap) {
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t body;
        total += lc_sls_log_size(klen, vlen, ent_begin[i], ent_begin[i + 1], ev_ns && ev_ns[i] != 0xFFFFFFFFu, &body);
    }
    if (total > out_cap)
        return total;
    uint64_t at = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t body;
        const bool has_ns = ev_ns && ev_ns[i] != 0xFFFFFFFFu;
        const uint32_t sz = lc_sls_log_size(klen, vlen, ent_begin[i], ent_begin[i + 1], has_ns, &body);
        if (!sz)
            continue;
        lc_sls_emit_log(out + at, base, ev_time[i], has_ns, has_ns ? ev_ns[i] : 0u, koff, klen, voff, vlen,
                        ent_begin[i], ent_begin[i + 1], body, 0, 1);
        at += sz;
    }
    return total;
}
}
