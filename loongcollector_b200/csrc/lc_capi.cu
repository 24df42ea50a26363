// lc_capi.cu -- implementation of the C-ABI declared in include/lc_b200.h: engine (stream, grow-only
// HBM workspace, look-back descriptors), host<->device staging, and the per-processor pipelines.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/lc_b200.h"
#include "lc_kernels.cuh"
#include "lc_tables.h"
#include "regex_compiler.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                                                  \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess)                                                                                         \
            return fail(LC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                               \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256; // grow-only with slack
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess)
            cap = want;
        return e;
    }
    void release() {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

std::atomic<uint64_t> g_regex_ids{1};

// engines alive in this process: lc_regex_free() tells each of them to drop its device copies of the tables
std::mutex g_engines_mu;
std::vector<lc_engine*> g_engines;

} // namespace

struct lc_regex {
    uint64_t id;
    lcb200::CompileResult res;
    std::string pattern;
};

struct lc_engine {
    int device = 0;
    cudaStream_t stream = nullptr;     // the stream every call of this engine is queued on
    cudaStream_t own_stream = nullptr; // created with the engine; `stream` unless lc_engine_set_stream replaced it
    uint64_t launches = 0;
    int num_sms = 148;
    int smem_per_block_optin = 0;
    int smem_per_sm = 0;
    bool force_basic_regex = false;   // env LC_B200_REGEX_KERNEL=basic   (tables in global memory)
    int regex_variant = 0; // env LC_B200_REGEX_KERNEL: 0 auto, 1 "fast" (stride-1), 2 "fast2" (stride-2), 3 "generic",
                           // 4 "tdfa" (single pass, staged input), 5 "tdfa_direct" (single pass, per-lane loads),
                           // 6 "tdfa_pc" (single pass, producer warps fill the tiles)
    uint64_t scratch_hint = 0;
    int length_order = -1; // env LC_B200_LENGTH_ORDER: 1 = always order ragged batches by length, 0 = never, unset = auto
    uint32_t max_warps = 32;   // env LC_B200_MAX_WARPS (tuning knob: resident warps per block of the regex kernels)
    bool multi_split = false;  // env LC_B200_MULTI_SPLIT=1: lc_regex_parse_multi launches one pattern at a time (test knob)
    // staging / workspace (grow-only)
    DevBuf in, ev_off, ev_len, out_a, out_b, out_c, out_d, out_e;
    DevBuf lines_off, lines_len, flags, state, cnt, pos, lab_sizes, lab_off, lab, order;
    DevBuf desc;   // look-back descriptors (3 regions)
    DevBuf split_scratch; // masks + per-tile counts of the three-pass split (lck::split_scratch_bytes)
    DevBuf small;  // tickets + counters: [0..3] u32 tickets, +16: u32 n_out, +32: u64 total, +64: u64 counters[2]
    void* h_small = nullptr; // pinned mirror of `small`
    std::unordered_map<uint64_t, void*> blobs; // regex id * 4 + layout -> device blob
    std::mutex freed_mu;
    std::vector<uint64_t> freed_ids; // regexes freed since the last call (their blobs are released lazily)
    // copy pipeline of the host-pointer entry points
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev_h2d, ev_comp, ev_d2h;
};

namespace {

cudaError_t engine_blob(lc_engine* e, const lc_regex* r, const void** out, int layout = 0) {
    *out = nullptr;
    if (!r)
        return cudaSuccess;
    const std::vector<uint8_t>& src = layout == 3   ? r->res.tdfa_blob
                                      : layout == 2 ? r->res.fast2_blob
                                                    : (layout == 1 ? r->res.fast_blob : r->res.blob);
    const uint64_t key = r->id * 4 + (uint64_t)layout;
    {
        // device tables of regexes freed since the last call (lc_regex_free): released here, on the engine's own
        // thread, after the stream has drained
        std::vector<uint64_t> dead;
        {
            std::lock_guard<std::mutex> lk(e->freed_mu);
            dead.swap(e->freed_ids);
        }
        if (!dead.empty()) {
            cudaStreamSynchronize(e->stream);
            for (uint64_t id : dead)
                for (uint64_t l = 0; l < 4; ++l) {
                    auto f = e->blobs.find(id * 4 + l);
                    if (f != e->blobs.end()) {
                        cudaFree(f->second);
                        e->blobs.erase(f);
                    }
                }
        }
    }
    auto it = e->blobs.find(key);
    if (it != e->blobs.end()) {
        *out = it->second;
        return cudaSuccess;
    }
    void* d = nullptr;
    cudaError_t er = cudaMalloc(&d, src.size());
    if (er != cudaSuccess)
        return er;
    er = cudaMemcpyAsync(d, src.data(), src.size(), cudaMemcpyHostToDevice, e->stream);
    if (er != cudaSuccess) {
        cudaFree(d);
        return er;
    }
    e->blobs[key] = d;
    *out = d;
    return cudaSuccess;
}

struct Small {
    uint32_t tickets[4];
    uint32_t n_out;
    uint32_t overflow;
    uint32_t pad0[2];
    uint64_t total;
    unsigned long long bump;
    unsigned long long next_batch;
    unsigned long long total_chars; // un-truncated split-char count of the last split
    unsigned long long counters[2];
};

int bind(lc_engine* e) {
    CU_TRY(cudaSetDevice(e->device));
    return LC_OK;
}

// descriptor regions: [0] split tiles, [1] state tiles, [2] sum tiles
struct DescPlan {
    uint64_t* r[3];
};

int prep_desc(lc_engine* e, size_t n0, size_t n1, size_t n2, DescPlan& plan) {
    size_t tot = n0 + n1 + n2 + 3;
    CU_TRY(e->desc.ensure(tot * 8));
    CU_TRY(cudaMemsetAsync(e->desc.p, 0, tot * 8, e->stream));
    CU_TRY(cudaMemsetAsync(e->small.p, 0, sizeof(Small), e->stream));
    plan.r[0] = e->desc.as<uint64_t>();
    plan.r[1] = plan.r[0] + n0 + 1;
    plan.r[2] = plan.r[1] + n1 + 1;
    return LC_OK;
}

int ensure_copy_streams(lc_engine* e, int nchunks) {
    if (!e->s_h2d)
        CU_TRY(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking));
    if (!e->s_d2h)
        CU_TRY(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
    while ((int)e->ev_h2d.size() < nchunks) {
        cudaEvent_t a, b, c;
        CU_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c, cudaEventDisableTiming));
        e->ev_h2d.push_back(a);
        e->ev_comp.push_back(b);
        e->ev_d2h.push_back(c);
    }
    return LC_OK;
}

// host-pointer entry points: every event must lie inside [0, base_len) -- a bad table would make the kernels read
// (or, for the serialiser, write) outside the staged arena
int check_events(const uint32_t* off, const uint32_t* len, uint64_t n, uint64_t base_len, const char* what) {
    uint64_t hi = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t en = (uint64_t)off[i] + len[i];
        hi = en > hi ? en : hi;
    }
    if (hi > base_len)
        return fail(LC_ERR_INVALID_ARG, std::string(what) + ": event beyond base_len");
    return LC_OK;
}

// first-byte filter + empty-line verdict of pattern slot p for the fused split + probe pass (prefix DFA of the blob)
void fill_probe_slot(lck::MlConfig& cfg, int p, const lc_regex* r) {
    memset(cfg.first[p], 0, sizeof cfg.first[p]);
    if (!r)
        return;
    const uint8_t* b = r->res.blob.data();
    const LcRegexHeader* h = reinterpret_cast<const LcRegexHeader*>(b);
    const uint8_t* cls = b + h->off_byte_class;
    const uint16_t* pre_next = reinterpret_cast<const uint16_t*>(b + h->off_pre_next);
    const uint8_t* pre_acc = b + h->off_pre_acc;
    for (uint32_t c = 0; c < 256; ++c)
        if (pre_next[h->pre_start * h->nclasses + cls[c]] != LC_PREFIX_DEAD)
            cfg.first[p][c >> 5] |= 1u << (c & 31);
    if (pre_acc[h->pre_start])
        cfg.empty_flags |= 1u << p;
}

int check_regex_usable(const lc_regex* r, const char* what) {
    if (!r->res.supported)
        return fail(r->res.valid ? LC_ERR_REGEX_UNSUPPORTED : LC_ERR_REGEX_INVALID,
                    std::string(what) + ": " + r->res.error);
    return LC_OK;
}

} // namespace

extern "C" {

const char* lc_version(void) { return "loongcollector_b200 0.1.0 (sm_100a)"; }
const char* lc_last_error(void) { return g_err.c_str(); }

int lc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int lc_engine_create(int device, lc_engine_t** out) {
    if (!out)
        return fail(LC_ERR_INVALID_ARG, "lc_engine_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t er = cudaGetDeviceCount(&n);
    if (er != cudaSuccess || n == 0)
        return fail(LC_ERR_CUDA, std::string("lc_engine_create: no usable CUDA device (") +
                                     (er != cudaSuccess ? cudaGetErrorString(er) : "device count 0") +
                                     "); this engine has no CPU fallback");
    if (device < 0 || device >= n)
        return fail(LC_ERR_INVALID_ARG, "lc_engine_create: bad device index");
    lc_engine* e = new (std::nothrow) lc_engine;
    if (!e)
        return fail(LC_ERR_CUDA, "out of host memory");
    e->device = device;
    CU_TRY(cudaSetDevice(device));
    CU_TRY(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    e->stream = e->own_stream;
    CU_TRY(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, device));
    CU_TRY(cudaDeviceGetAttribute(&e->smem_per_block_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CU_TRY(cudaDeviceGetAttribute(&e->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    {
        const char* k = getenv("LC_B200_REGEX_KERNEL");
        e->force_basic_regex = k && !strcmp(k, "basic");
        const char* mw = getenv("LC_B200_MAX_WARPS");
        if (mw && atoi(mw) >= 4 && atoi(mw) <= 32)
            e->max_warps = (uint32_t)atoi(mw);
        const char* msp = getenv("LC_B200_MULTI_SPLIT");
        e->multi_split = msp && !strcmp(msp, "1");
        const char* lo = getenv("LC_B200_LENGTH_ORDER");
        e->length_order = !lo ? -1 : (!strcmp(lo, "1") ? 1 : 0);
        e->regex_variant = !k ? 0 : (!strcmp(k, "fast") ? 1 : (!strcmp(k, "fast2") ? 2 : (!strcmp(k, "generic") ? 3 : (!strcmp(k, "tdfa") ? 4 : (!strcmp(k, "tdfa_direct") ? 5 : (!strcmp(k, "tdfa_pc") ? 6 : 0))))));
    }
    CU_TRY(e->small.ensure(sizeof(Small)));
    CU_TRY(cudaMallocHost(&e->h_small, sizeof(Small)));
    {
        std::lock_guard<std::mutex> lk(g_engines_mu);
        g_engines.push_back(e);
    }
    *out = e;
    return LC_OK;
}

void lc_engine_destroy(lc_engine_t* e) {
    if (!e)
        return;
    {
        std::lock_guard<std::mutex> lk(g_engines_mu);
        g_engines.erase(std::remove(g_engines.begin(), g_engines.end(), e), g_engines.end());
    }
    cudaSetDevice(e->device);
    if (e->stream)
        cudaStreamSynchronize(e->stream);
    DevBuf* bufs[] = {&e->in, &e->ev_off, &e->ev_len, &e->out_a, &e->out_b, &e->out_c, &e->out_d, &e->out_e,
                      &e->lines_off, &e->lines_len, &e->flags, &e->state, &e->cnt, &e->pos, &e->lab_sizes,
                      &e->lab_off, &e->lab, &e->order, &e->desc, &e->small, &e->split_scratch};
    for (DevBuf* b : bufs)
        b->release();
    for (auto& kv : e->blobs)
        cudaFree(kv.second);
    if (e->h_small)
        cudaFreeHost(e->h_small);
    for (cudaEvent_t ev : e->ev_h2d)
        cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->ev_comp)
        cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->ev_d2h)
        cudaEventDestroy(ev);
    if (e->s_h2d)
        cudaStreamDestroy(e->s_h2d);
    if (e->s_d2h)
        cudaStreamDestroy(e->s_d2h);
    if (e->own_stream)
        cudaStreamDestroy(e->own_stream);
    delete e;
}

int lc_engine_sync(lc_engine_t* e) {
    if (!e)
        return fail(LC_ERR_INVALID_ARG, "engine is NULL");
    CU_TRY(cudaSetDevice(e->device));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return LC_OK;
}

void* lc_engine_stream(lc_engine_t* e) { return e ? (void*)e->stream : nullptr; }

int lc_engine_set_stream(lc_engine_t* e, void* stream) {
    if (!e)
        return fail(LC_ERR_INVALID_ARG, "engine is NULL");
    CU_TRY(cudaSetDevice(e->device));
    CU_TRY(cudaStreamSynchronize(e->stream)); // workspace reuse is ordered by the stream: drain the old one first
    e->stream = stream ? (cudaStream_t)stream : e->own_stream;
    return LC_OK;
}
uint64_t lc_engine_launch_count(const lc_engine_t* e) { return e ? e->launches : 0; }

void* lc_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        g_err = "lc_host_alloc: cudaMallocHost failed";
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void lc_host_free(void* p) {
    if (p)
        cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------------ regex
int lc_regex_compile(const char* pattern, size_t len, lc_regex_t** out) {
    if (!out || (!pattern && len))
        return fail(LC_ERR_INVALID_ARG, "lc_regex_compile: bad arguments");
    lc_regex* r = new (std::nothrow) lc_regex;
    if (!r)
        return fail(LC_ERR_CUDA, "out of host memory");
    r->id = g_regex_ids.fetch_add(1);
    r->pattern.assign(pattern ? pattern : "", len);
    r->res = lcb200::compile_regex(r->pattern.data(), r->pattern.size());
    *out = r;
    if (!r->res.valid)
        return fail(LC_ERR_REGEX_INVALID, r->res.error);
    if (!r->res.supported)
        return fail(LC_ERR_REGEX_UNSUPPORTED, r->res.error);
    return LC_OK;
}

void lc_regex_free(lc_regex_t* r) {
    if (!r)
        return;
    {
        std::lock_guard<std::mutex> lk(g_engines_mu);
        for (lc_engine* e : g_engines) {
            std::lock_guard<std::mutex> lk2(e->freed_mu);
            e->freed_ids.push_back(r->id);
        }
    }
    delete r;
}
const char* lc_regex_error(const lc_regex_t* r) { return r ? r->res.error.c_str() : ""; }
uint32_t lc_regex_ngroups(const lc_regex_t* r) { return r ? r->res.ngroups : 0; }

void lc_regex_info(const lc_regex_t* r, uint32_t info[8]) {
    memset(info, 0, 8 * sizeof(uint32_t));
    if (!r || !r->res.supported)
        return;
    const LcRegexHeader* h = reinterpret_cast<const LcRegexHeader*>(r->res.blob.data());
    info[0] = h->mode;
    info[1] = h->nclasses;
    info[2] = h->nw;
    info[3] = h->npc;
    info[4] = h->rev_nstates;
    info[5] = h->pre_nstates;
    info[6] = h->total_bytes;
    info[7] = r->res.n_insts;
}

// ------------------------------------------------------------------------------------------------ split
int lc_split_lines_dev(lc_engine_t* e, const uint8_t* d_buf, uint64_t len, uint8_t split_char, uint32_t* d_out_off,
                       uint32_t* d_out_len, uint64_t cap, uint64_t* n_out) {
    if (!e || !n_out || (len && !d_buf))
        return fail(LC_ERR_INVALID_ARG, "lc_split_lines_dev: bad arguments");
    *n_out = 0;
    if (len == 0)
        return LC_OK;
    if (len >= 0xFFFFFFF0ull)
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB per call");
    int rc = bind(e);
    if (rc)
        return rc;
    uint32_t shift = (uint32_t)((uintptr_t)d_buf & 15u);
    DescPlan plan;
    rc = prep_desc(e, lck::split_tiles(len, shift), 0, 0, plan);
    if (rc)
        return rc;
    Small* ds = e->small.as<Small>();
    uint32_t cap32 = cap > 0x3FFFFFFFull ? 0x3FFFFFFFu : (uint32_t)cap;
    CU_TRY(e->split_scratch.ensure(lck::split_scratch_bytes(len, false)));
    e->launches += lck::launch_split(d_buf, (uint32_t)len, split_char, d_out_off, d_out_len, cap32, plan.r[0],
                                     &ds->tickets[0], &ds->n_out, &ds->total_chars, e->split_scratch.as<uint64_t>(),
                                     e->stream);
    CU_TRY(cudaGetLastError());
    Small* hs = (Small*)e->h_small;
    CU_TRY(cudaMemcpyAsync(hs, ds, sizeof(Small), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    if (hs->total_chars >= (1ull << 30) - 1)
        return fail(LC_ERR_TOO_LARGE, "more than 2^30 pieces in one call");
    *n_out = hs->n_out;
    if (*n_out > cap)
        return fail(LC_ERR_CAPACITY, "lc_split_lines: output capacity too small");
    return LC_OK;
}

int lc_split_lines(lc_engine_t* e, const uint8_t* buf, uint64_t len, uint8_t split_char, uint32_t* out_off,
                   uint32_t* out_len, uint64_t cap, uint64_t* n_out) {
    if (!e || !n_out || (len && !buf))
        return fail(LC_ERR_INVALID_ARG, "lc_split_lines: bad arguments");
    *n_out = 0;
    if (len == 0)
        return LC_OK;
    if (len >= 0xFFFFFFF0ull)
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB per call");
    int rc = bind(e);
    if (rc)
        return rc;
    uint64_t dcap = cap < len ? cap : len;
    CU_TRY(e->in.ensure(len + 16));
    CU_TRY(e->out_a.ensure((dcap + 1) * 4));
    CU_TRY(e->out_b.ensure((dcap + 1) * 4));
    CU_TRY(cudaMemcpyAsync(e->in.p, buf, len, cudaMemcpyHostToDevice, e->stream));
    rc = lc_split_lines_dev(e, e->in.as<uint8_t>(), len, split_char, e->out_a.as<uint32_t>(),
                            e->out_b.as<uint32_t>(), dcap, n_out);
    if (rc)
        return rc;
    if (*n_out) {
        CU_TRY(cudaMemcpyAsync(out_off, e->out_a.p, *n_out * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaMemcpyAsync(out_len, e->out_b.p, *n_out * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
    }
    return LC_OK;
}

// ------------------------------------------------------------------------------------------------ regex parse
// span_bytes: bytes of the arena this batch's events cover (== base_len for a whole-arena call; the chunked host path
// passes the chunk's own span so that per-batch heuristics see the batch, not the arena).  ev_stride: distance in
// elements between consecutive entries of d_ev_off / d_ev_len (1 = dense tables).
static int regex_parse_dev_impl(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                                uint64_t span_bytes, const uint32_t* d_ev_off, const uint32_t* d_ev_len,
                                uint32_t ev_stride, uint64_t n, uint32_t nkeys, uint8_t* d_status, uint32_t* d_cap_off,
                                uint32_t* d_cap_len, bool bool_only);

int lc_regex_parse_dev(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                       const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint32_t nkeys,
                       uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len) {
    return regex_parse_dev_impl(e, re, d_base, base_len, base_len, d_ev_off, d_ev_len, 1, n, nkeys, d_status, d_cap_off,
                                d_cap_len, false);
}

int lc_regex_parse_strided_dev(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                               const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint32_t ev_stride, uint64_t n,
                               uint32_t nkeys, uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len) {
    if (ev_stride == 0)
        return fail(LC_ERR_INVALID_ARG, "lc_regex_parse_strided_dev: ev_stride must be >= 1");
    return regex_parse_dev_impl(e, re, d_base, base_len, base_len, d_ev_off, d_ev_len, ev_stride, n, nkeys, d_status,
                                d_cap_off, d_cap_len, false);
}

int lc_regex_match_dev(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                       const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint8_t* d_out_match) {
    return regex_parse_dev_impl(e, re, d_base, base_len, base_len, d_ev_off, d_ev_len, 1, n, 0, d_out_match, nullptr,
                                nullptr, true);
}

int lc_regex_match(lc_engine_t* e, const lc_regex_t* re, const uint8_t* base, uint64_t base_len,
                   const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n, uint8_t* out_match) {
    if (!e || !re || (n && (!ev_off || !ev_len || !out_match)) || (base_len && !base))
        return fail(LC_ERR_INVALID_ARG, "lc_regex_match: bad arguments");
    int rc = check_regex_usable(re, "lc_regex_match");
    if (rc)
        return rc;
    if (n == 0)
        return LC_OK;
    rc = check_events(ev_off, ev_len, n, base_len, "lc_regex_match");
    if (rc)
        return rc;
    rc = bind(e);
    if (rc)
        return rc;
    CU_TRY(e->in.ensure(base_len + 16));
    CU_TRY(e->ev_off.ensure(n * 4));
    CU_TRY(e->ev_len.ensure(n * 4));
    CU_TRY(e->out_a.ensure(n));
    CU_TRY(cudaMemcpyAsync(e->in.p, base, base_len, cudaMemcpyHostToDevice, e->stream));
    CU_TRY(cudaMemcpyAsync(e->ev_off.p, ev_off, n * 4, cudaMemcpyHostToDevice, e->stream));
    CU_TRY(cudaMemcpyAsync(e->ev_len.p, ev_len, n * 4, cudaMemcpyHostToDevice, e->stream));
    rc = lc_regex_match_dev(e, re, e->in.as<uint8_t>(), base_len, e->ev_off.as<uint32_t>(), e->ev_len.as<uint32_t>(), n,
                            e->out_a.as<uint8_t>());
    if (rc)
        return rc;
    CU_TRY(cudaMemcpyAsync(out_match, e->out_a.p, n, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return LC_OK;
}

// gathers a strided event table into the engine's dense scratch tables (the two-pass fall-back kernels take dense
// tables only)
static int densify_events(lc_engine_t* e, const uint32_t*& d_ev_off, const uint32_t*& d_ev_len, uint32_t ev_stride,
                          uint64_t n) {
    CU_TRY(e->lines_off.ensure(n * 4));
    CU_TRY(e->lines_len.ensure(n * 4));
    CU_TRY(cudaMemcpy2DAsync(e->lines_off.p, 4, d_ev_off, (size_t)ev_stride * 4, 4, n, cudaMemcpyDeviceToDevice,
                             e->stream));
    CU_TRY(cudaMemcpy2DAsync(e->lines_len.p, 4, d_ev_len, (size_t)ev_stride * 4, 4, n, cudaMemcpyDeviceToDevice,
                             e->stream));
    d_ev_off = e->lines_off.as<uint32_t>();
    d_ev_len = e->lines_len.as<uint32_t>();
    return LC_OK;
}

static int regex_parse_dev_impl(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                                uint64_t span_bytes, const uint32_t* d_ev_off, const uint32_t* d_ev_len,
                                uint32_t ev_stride, uint64_t n, uint32_t nkeys, uint8_t* d_status, uint32_t* d_cap_off,
                                uint32_t* d_cap_len, bool bool_only) {
    if (!e || !re)
        return fail(LC_ERR_INVALID_ARG, "lc_regex_parse_dev: bad arguments");
    int rc = check_regex_usable(re, "lc_regex_parse");
    if (rc)
        return rc;
    if (n == 0)
        return LC_OK;
    if (base_len >= 0xFFFFFFF0ull || n >= (1ull << 30))
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB and < 2^30 events per call");
    rc = bind(e);
    if (rc)
        return rc;
    const void* d_blob;
    CU_TRY(engine_blob(e, re, &d_blob));
    const LcRegexHeader* h = reinterpret_cast<const LcRegexHeader*>(re->res.blob.data());
    Small* ds = e->small.as<Small>();
    Small* hs = (Small*)e->h_small;
    const bool force_basic = e->force_basic_regex;
    const size_t smem_max = (size_t)e->smem_per_block_optin;
    // ---- single-pass tagged DFA: the preferred kernel whenever the pattern's TDFA fits shared memory.  Nothing on
    // this path waits for the device: ragged-batch ordering is decided by a device-side flag and events too long for
    // the 16-bit capture registers are redone by a follow-up kernel that exits at once when there was none.
    if (!force_basic && (e->regex_variant == 0 || e->regex_variant >= 4) && !re->res.tdfa_blob.empty()) {
        const LcTdfaHeader* th = reinterpret_cast<const LcTdfaHeader*>(re->res.tdfa_blob.data());
        const uint32_t tb = (uint32_t)re->res.tdfa_blob.size();
        const bool staged = e->regex_variant != 5 || ev_stride != 1;
        auto smem_need = [&](uint32_t warps) {
            return staged ? lck::tdfa_staged_smem_bytes(tb, th->nregs, warps * 32)
                          : lck::tdfa_smem_bytes(tb, th->nregs, warps * 32);
        };
        uint32_t warps = e->max_warps;
        while (warps > 4 && smem_need(warps) > smem_max)
            warps -= 2;
        bool usable = smem_need(warps) <= smem_max;
        if (usable && !staged && base_len >= 65535) { // A/B kernel without the long-event hand-over: host-side check
            CU_TRY(cudaMemsetAsync(ds->counters, 0, sizeof ds->counters, e->stream));
            lck::launch_len_stats(d_ev_len, n, ds->counters, e->stream);
            e->launches++;
            CU_TRY(cudaMemcpyAsync(hs->counters, ds->counters, sizeof ds->counters, cudaMemcpyDeviceToHost,
                                   e->stream));
            CU_TRY(cudaStreamSynchronize(e->stream));
            usable = hs->counters[0] < 65535;
        }
        if (usable) {
            const void* d_tblob;
            CU_TRY(engine_blob(e, re, &d_tblob, 3));
            const uint32_t threads = warps * 32;
            const uint64_t need_blocks = (n + threads - 1) / threads;
            const uint32_t grid = (uint32_t)std::min<uint64_t>(need_blocks, (uint64_t)e->num_sms);
            // Ragged batches of long lines: a warp costs its longest line, so visit the events by descending length
            // bucket.  Only built when the mean length makes the three small pre-pass kernels negligible; the scan
            // kernel leaves a flag saying whether the lengths really are ragged, which the regex kernel reads.
            // LC_B200_LENGTH_ORDER=1 forces the pre-pass, =0 disables it.
            const uint32_t* d_order = nullptr;
            const uint32_t* d_order_flag = nullptr;
            if (staged && ev_stride == 1 && n >= 4096 && e->length_order != 0 &&
                (e->length_order == 1 || span_bytes / n >= 1024)) {
                CU_TRY(e->order.ensure(n * 4 + 512));
                uint32_t* hist = e->order.as<uint32_t>() + n;
                lck::launch_length_order(d_ev_len, n, hist, e->order.as<uint32_t>(), e->stream);
                e->launches += 3;
                d_order = e->order.as<uint32_t>();
                d_order_flag = e->length_order == 1 ? nullptr : hist + 64;
            }
            CU_TRY(cudaMemsetAsync(&ds->overflow, 0, offsetof(Small, total_chars) - offsetof(Small, overflow),
                                   e->stream));
            int er;
            if (staged && e->regex_variant == 6 && !d_order &&
                lck::tdfa_pc_smem_bytes(tb, th->nregs, 1024) <= smem_max)
                er = lck::launch_regex_tdfa_pc(d_tblob, tb, th->has_slow != 0, th->nregs, d_base, d_ev_off, d_ev_len,
                                               ev_stride, n, nkeys, d_status, bool_only ? nullptr : d_cap_off,
                                               bool_only ? nullptr : d_cap_len, 1024,
                                               (uint32_t)std::min<uint64_t>((n + 895) / 896, (uint64_t)e->num_sms),
                                               &ds->next_batch, &ds->overflow, e->stream);
            else if (staged)
                er = lck::launch_regex_tdfa_staged(d_tblob, tb, th->has_slow != 0, th->nregs, d_base, d_ev_off,
                                                   d_ev_len, ev_stride, n, nkeys, d_status,
                                                   bool_only ? nullptr : d_cap_off, bool_only ? nullptr : d_cap_len,
                                                   threads, grid, &ds->next_batch, &ds->overflow, d_order,
                                                   d_order_flag, e->stream);
            else
                er = lck::launch_regex_tdfa(d_tblob, tb, th->has_slow != 0, th->nregs, d_base, d_ev_off, d_ev_len, n,
                                            nkeys, d_status, bool_only ? nullptr : d_cap_off,
                                            bool_only ? nullptr : d_cap_len, threads, grid, &ds->next_batch, nullptr,
                                            e->stream);
            e->launches++;
            if (er)
                return fail(LC_ERR_CUDA, std::string("regex kernel launch: ") + cudaGetErrorString((cudaError_t)er));
            if (staged && base_len >= 65535) {
                lck::TdfaMultiArgs ma;
                memset(&ma, 0, sizeof ma);
                ma.blob[0] = d_tblob;
                ma.blob_bytes[0] = tb;
                ma.nkeys[0] = nkeys;
                ma.npat = 1;
                lck::launch_regex_tdfa_long(ma, d_base, d_ev_off, d_ev_len, ev_stride, n, nullptr, nullptr, d_status,
                                            d_cap_off, d_cap_len, th->ngroups, &ds->overflow, bool_only, e->stream);
                e->launches++;
                CU_TRY(cudaGetLastError());
            }
            return LC_OK;
        }
    }
    if (ev_stride != 1) {
        rc = densify_events(e, d_ev_off, d_ev_len, ev_stride, n);
        if (rc)
            return rc;
    }
    if (!force_basic) {
        // ---- pick the kernel variant: stride-2 layout > stride-1 fast layout > generic shared-memory interpreter
        uint64_t mx = 0, avg = span_bytes / n + 1;
        if (h->mode == LC_MODE_TWOPASS) {
            // size the per-thread label area from the actual length distribution: cover the longest event when
            // the batch is near-uniform, else ~1.25x the mean (longer events spill to the global slab)
            CU_TRY(cudaMemsetAsync(ds->counters, 0, sizeof ds->counters, e->stream));
            lck::launch_len_stats(d_ev_len, n, ds->counters, e->stream);
            e->launches++;
            CU_TRY(cudaMemcpyAsync(hs->counters, ds->counters, sizeof ds->counters, cudaMemcpyDeviceToHost,
                                   e->stream));
            CU_TRY(cudaStreamSynchronize(e->stream));
            mx = hs->counters[0];
            avg = hs->counters[1] / n + 1;
        }
        const uint64_t cover = (mx <= avg + avg / 2 + 64) ? mx : (avg + avg / 4);
        enum { V_FAST2, V_FAST, V_GENERIC } variant = V_GENERIC;
        if (e->regex_variant == 0 || e->regex_variant == 2)
            if (!re->res.fast2_blob.empty() && mx < 65535 && re->res.fast2_blob.size() + 32768 <= smem_max)
                variant = V_FAST2;
        if (variant == V_GENERIC && (e->regex_variant == 0 || e->regex_variant == 1) && !re->res.fast_blob.empty() &&
            re->res.fast_blob.size() + 32768 <= smem_max)
            variant = V_FAST;
        const std::vector<uint8_t>& vb =
            variant == V_FAST2 ? re->res.fast2_blob : (variant == V_FAST ? re->res.fast_blob : re->res.blob);
        bool convert_status = false;
        if (bool_only && variant != V_FAST2) { // these kernels always write captures: give them scratch tables
            CU_TRY(e->out_d.ensure(n * h->ngroups * 4 + 4));
            CU_TRY(e->out_e.ensure(n * h->ngroups * 4 + 4));
            d_cap_off = e->out_d.as<uint32_t>();
            d_cap_len = e->out_e.as<uint32_t>();
            convert_status = true;
        }
        const uint32_t blob_bytes = (uint32_t)vb.size();
        if (blob_bytes + 4096 <= smem_max) {
            const void* d_vblob = d_blob;
            if (variant != V_GENERIC)
                CU_TRY(engine_blob(e, re, &d_vblob, variant == V_FAST2 ? 2 : 1));
            // labels per 32-bit word: stride-2 -> 8 bytes of input, stride-1 u8 -> 4, u16 -> 2
            const uint32_t per = variant == V_FAST2 ? 8u
                                                    : ((h->mode == LC_MODE_TWOPASS && h->rev_label_bytes == 2 &&
                                                        variant == V_GENERIC)
                                                           ? 2u
                                                           : 4u);
            uint32_t lab_words = 0, threads = 512, blocks_per_sm = 2;
            size_t slot_bytes = 0; // per warp
            if (h->mode == LC_MODE_TWOPASS) {
                lab_words = (uint32_t)((cover + 15) / per + 2); // + 15: labels are shifted by the 16 B misalignment
                if (lab_words < 8)
                    lab_words = 8;
                if (variant == V_FAST2)
                    slot_bytes = (size_t)32 * lck::fast2_slot_pitch(h->ngroups) * 2;
                else if (variant == V_FAST)
                    slot_bytes = (size_t)32 * lck::fast_slot_pitch(h->ngroups) * 4;
                size_t budget = smem_max - blob_bytes - 1024;
                uint32_t warps = (uint32_t)(budget / ((size_t)lab_words * 128 + slot_bytes));
                // Long lines: labels in shared memory would leave a handful of resident warps (measured: 20x
                // slower).  Keep >= kMinWarps warps resident and let events that do not fit keep their labels in
                // the global slab instead (L2-resident while in flight; +1 B of traffic per input byte worst case).
                // The stride-2 kernel evaluates oversized events in checkpointed blocks (labels of one block only),
                // so it can afford the full 32 warps; the others keep whole-event labels in the global slab.
                const uint32_t kMinWarps = variant == V_FAST2 ? 32 : 20;
                if (warps < kMinWarps) {
                    warps = kMinWarps;
                    size_t per_warp = budget / kMinWarps;
                    lab_words = per_warp > slot_bytes + 8 * 128 ? (uint32_t)((per_warp - slot_bytes) / 128) : 8;
                    while (warps > 4 && (size_t)warps * ((size_t)lab_words * 128 + slot_bytes) > budget)
                        --warps;
                }
                if (warps > e->max_warps)
                    warps = e->max_warps;
                threads = warps * 32;
                blocks_per_sm = 1;
                if (warps <= 16 && 2 * (blob_bytes + (size_t)warps * (lab_words * 128 + slot_bytes) + 1024) <=
                                       (size_t)e->smem_per_sm)
                    blocks_per_sm = 2;
            }
            uint64_t need_blocks = (n + threads - 1) / threads;
            uint32_t grid = (uint32_t)std::min<uint64_t>(need_blocks, (uint64_t)e->num_sms * blocks_per_sm);
            // global label slab for events longer than the shared-memory budget: start small, remember what worked
            uint64_t full = span_bytes / per + 2 * n + 1024;
            uint64_t scratch_words = std::max<uint64_t>(e->scratch_hint, std::min<uint64_t>(full, 16ull << 20));
            if (h->mode == LC_MODE_TWOPASS && (mx + 15) / per + 2 > lab_words) {
                // some events spill: provision the upper bound of what ALL events could need (no retry runs)
                uint64_t bound = (hs->counters[1] + 16 * n) / per + 2 * n + 1024;
                if (variant == V_FAST2) // checkpointed blocks: 2 B per (lab_words * 8)-byte block of a long event
                    bound = 4 * n + hs->counters[1] / 32 + 1024;
                scratch_words = std::max(scratch_words, bound);
            }
            // ragged batch: visit events in descending length-bucket order (a warp costs its longest line)
            const uint32_t* d_order = nullptr;
            // (measured on C5, Zipf 64 B-8 KB: -7 %, the scattered visiting order costs more L2 locality than the
            //  balanced warps win -- kept opt-in: LC_B200_LENGTH_ORDER=1)
            if (e->length_order == 1 && h->mode == LC_MODE_TWOPASS && mx > 2 * avg + 64 && n >= 4096) {
                CU_TRY(e->order.ensure(n * 4 + 512));
                uint32_t* hist = e->order.as<uint32_t>() + n;
                lck::launch_length_order(d_ev_len, n, hist, e->order.as<uint32_t>(), e->stream);
                e->launches += 3;
                d_order = e->order.as<uint32_t>();
            }
            for (int attempt = 0; attempt < 8; ++attempt) {
                if (h->mode == LC_MODE_TWOPASS)
                    CU_TRY(e->lab.ensure(scratch_words * 4));
                CU_TRY(cudaMemsetAsync(&ds->overflow, 0, offsetof(Small, total_chars) - offsetof(Small, overflow), e->stream));
                int er;
                if (variant == V_FAST2) {
                    const LcFast2Header* f2h = reinterpret_cast<const LcFast2Header*>(vb.data());
                    const bool multi = f2h->has_multi != 0;
                    er = lck::launch_regex_fast2(d_vblob, blob_bytes, multi, f2h->pair_shift == 2, h->ngroups, d_base,
                                                 d_ev_off, d_ev_len, n,
                                                 nkeys, d_status, d_cap_off, d_cap_len, lab_words, threads, grid,
                                                 e->lab.as<uint32_t>(), scratch_words, &ds->bump, &ds->overflow,
                                                 &ds->next_batch, d_order, e->stream);
                } else if (variant == V_FAST) {
                    const bool multi = reinterpret_cast<const LcFastHeader*>(vb.data())->reserved[0] != 0;
                    er = lck::launch_regex_twopass_fast(d_vblob, blob_bytes, multi, h->ngroups, d_base, d_ev_off,
                                                        d_ev_len, n, nkeys, d_status, d_cap_off, d_cap_len, lab_words,
                                                        threads, grid, e->lab.as<uint32_t>(), scratch_words, &ds->bump,
                                                        &ds->overflow, &ds->next_batch, d_order, e->stream);
                } else {
                    er = lck::launch_regex_parse_fast(d_vblob, blob_bytes, h->rev_label_bytes, h->ngroups, d_base,
                                                      d_ev_off, d_ev_len, n, nkeys, d_status, d_cap_off, d_cap_len,
                                                      lab_words, threads, grid, e->lab.as<uint32_t>(), scratch_words,
                                                      &ds->bump, &ds->overflow, &ds->next_batch, d_order, e->stream);
                }
                e->launches++;
                if (er)
                    return fail(LC_ERR_CUDA,
                                std::string("regex kernel launch: ") + cudaGetErrorString((cudaError_t)er));
                if (h->mode != LC_MODE_TWOPASS) {
                    if (convert_status) {
                        lck::launch_status_to_bool(d_status, n, e->stream);
                        e->launches++;
                    }
                    return LC_OK;
                }
                CU_TRY(cudaMemcpyAsync(&hs->overflow, &ds->overflow, 4, cudaMemcpyDeviceToHost, e->stream));
                CU_TRY(cudaStreamSynchronize(e->stream));
                if (!hs->overflow) {
                    e->scratch_hint = scratch_words;
                    if (convert_status) {
                        lck::launch_status_to_bool(d_status, n, e->stream);
                        e->launches++;
                    }
                    return LC_OK;
                }
                scratch_words = scratch_words < full ? std::min<uint64_t>(full, scratch_words * 4) : scratch_words * 2;
            }
            return fail(LC_ERR_CUDA, "label scratch exhausted after retries");
        }
    }
    // ---- baseline kernel (tables in global memory); kept for A/B checks and for automata beyond shared memory
    const uint64_t* d_lab_off = nullptr;
    uint16_t* d_lab = nullptr;
    if (h->mode == LC_MODE_TWOPASS) {
        DescPlan plan;
        rc = prep_desc(e, 0, 0, lck::scan_tiles(n), plan);
        if (rc)
            return rc;
        CU_TRY(e->lab_sizes.ensure(n * 4));
        CU_TRY(e->lab_off.ensure(n * 8));
        lck::launch_label_sizes(d_ev_len, n, e->lab_sizes.as<uint32_t>(), e->stream);
        lck::launch_exclusive_sum(e->lab_sizes.as<uint32_t>(), n, e->lab_off.as<uint64_t>(), &ds->total, plan.r[2],
                                  &ds->tickets[2], e->stream);
        e->launches += 2;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(&hs->total, &ds->total, 8, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        CU_TRY(e->lab.ensure(hs->total * 2 + 16));
        d_lab_off = e->lab_off.as<uint64_t>();
        d_lab = e->lab.as<uint16_t>();
    }
    if (bool_only) {
        CU_TRY(e->out_d.ensure(n * h->ngroups * 4 + 4));
        CU_TRY(e->out_e.ensure(n * h->ngroups * 4 + 4));
        d_cap_off = e->out_d.as<uint32_t>();
        d_cap_len = e->out_e.as<uint32_t>();
    }
    lck::launch_regex_parse_basic(d_blob, h->mode, h->ngroups, d_base, d_ev_off, d_ev_len, n, nkeys, d_status,
                                  d_cap_off, d_cap_len, d_lab_off, d_lab, e->stream);
    e->launches++;
    if (bool_only) {
        lck::launch_status_to_bool(d_status, n, e->stream);
        e->launches++;
    }
    CU_TRY(cudaGetLastError());
    return LC_OK;
}

} // extern "C"

// Pipelined host-pointer batch: the arena is cut into chunks of whole events; chunk c's bytes and event table go up on a
// copy stream, its kernels run on the engine stream as soon as they have landed, and its result tables travel back on
// a third stream while later chunks are still being uploaded (PCIe is full duplex).  `run(c, i0, cnt, span)` queues
// the chunk's kernels; `down(c, i0, cnt)` queues its D2H copies on s_d2h.  Event ranges are validated chunk by chunk
// right before their upload is queued, so the check overlaps the copies of the previous chunks.
template <class Run, class Down>
static int pipelined_events(lc_engine_t* e, const uint8_t* base, uint64_t base_len, const uint32_t* ev_off,
                            const uint32_t* ev_len, uint64_t n, uint64_t nchunks, const char* what, Run run, Down down) {
    int rc = ensure_copy_streams(e, (int)nchunks);
    if (rc)
        return rc;
    uint8_t* d_in = e->in.as<uint8_t>();
    uint32_t* d_off = e->ev_off.as<uint32_t>();
    uint32_t* d_len = e->ev_len.as<uint32_t>();
    // every exit path drains the copy streams: they read and write caller buffers
    auto drain = [&](int code) {
        cudaStreamSynchronize(e->s_h2d);
        cudaStreamSynchronize(e->stream);
        cudaStreamSynchronize(e->s_d2h);
        return code;
    };
    // the engine stream may still be reading the workspace on behalf of an earlier asynchronous call
    cudaEvent_t ev0 = e->ev_comp[0];
    if (cudaEventRecord(ev0, e->stream) != cudaSuccess || cudaStreamWaitEvent(e->s_h2d, ev0, 0) != cudaSuccess)
        return fail(LC_ERR_CUDA, "stream ordering failed");
    for (uint64_t c = 0; c < nchunks; ++c) {
        const uint64_t i0 = n * c / nchunks, i1 = n * (c + 1) / nchunks, cnt = i1 - i0;
        uint64_t lo = ~0ull, hi = 0;
        for (uint64_t i = i0; i < i1; ++i) {
            const uint64_t o = ev_off[i], en = o + ev_len[i];
            lo = o < lo ? o : lo;
            hi = en > hi ? en : hi;
        }
        if (hi > base_len)
            return drain(fail(LC_ERR_INVALID_ARG, std::string(what) + ": event beyond base_len"));
        lo = lo == ~0ull ? 0 : (lo & ~15ull);
#define LC_PIPE_TRY(expr)                                                                                              \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess)                                                                                         \
            return drain(fail(LC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)));                        \
    } while (0)
        if (hi > lo)
            LC_PIPE_TRY(cudaMemcpyAsync(d_in + lo, base + lo, hi - lo, cudaMemcpyHostToDevice, e->s_h2d));
        LC_PIPE_TRY(cudaMemcpyAsync(d_off + i0, ev_off + i0, cnt * 4, cudaMemcpyHostToDevice, e->s_h2d));
        LC_PIPE_TRY(cudaMemcpyAsync(d_len + i0, ev_len + i0, cnt * 4, cudaMemcpyHostToDevice, e->s_h2d));
        LC_PIPE_TRY(cudaEventRecord(e->ev_h2d[c], e->s_h2d));
        LC_PIPE_TRY(cudaStreamWaitEvent(e->stream, e->ev_h2d[c], 0));
        rc = run(c, i0, cnt, hi > lo ? hi - lo : 0);
        if (rc)
            return drain(rc);
        LC_PIPE_TRY(cudaEventRecord(e->ev_comp[c], e->stream));
        LC_PIPE_TRY(cudaStreamWaitEvent(e->s_d2h, e->ev_comp[c], 0));
        rc = down(c, i0, cnt);
        if (rc)
            return drain(rc);
#undef LC_PIPE_TRY
    }
    CU_TRY(cudaStreamSynchronize(e->s_d2h));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return LC_OK;
}

static uint64_t pipeline_chunks(uint64_t base_len, uint64_t n) {
    const uint64_t kChunkBytes = 48ull << 20;
    uint64_t nchunks = (base_len + kChunkBytes - 1) / kChunkBytes;
    if (nchunks > 64)
        nchunks = 64;
    if (nchunks < 2 || n < nchunks * 1024)
        return 1;
    return nchunks;
}

extern "C" {

int lc_regex_parse(lc_engine_t* e, const lc_regex_t* re, const uint8_t* base, uint64_t base_len,
                   const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n, uint32_t nkeys, uint8_t* status,
                   uint32_t* cap_off, uint32_t* cap_len) {
    if (!e || !re || (n && (!ev_off || !ev_len || !status)) || (base_len && !base))
        return fail(LC_ERR_INVALID_ARG, "lc_regex_parse: bad arguments");
    int rc = check_regex_usable(re, "lc_regex_parse");
    if (rc)
        return rc;
    if (n == 0)
        return LC_OK;
    const uint32_t G = re->res.ngroups;
    if (G && (!cap_off || !cap_len))
        return fail(LC_ERR_INVALID_ARG, "lc_regex_parse: bad arguments");
    if (base_len >= 0xFFFFFFF0ull || n >= (1ull << 30))
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB and < 2^30 events per call");
    rc = bind(e);
    if (rc)
        return rc;
    CU_TRY(e->in.ensure(base_len + 16));
    CU_TRY(e->ev_off.ensure(n * 4));
    CU_TRY(e->ev_len.ensure(n * 4));
    CU_TRY(e->out_a.ensure(n));
    CU_TRY(e->out_b.ensure(n * G * 4 + 4));
    CU_TRY(e->out_c.ensure(n * G * 4 + 4));
    const uint64_t nchunks = pipeline_chunks(base_len, n);
    if (nchunks == 1) {
        rc = check_events(ev_off, ev_len, n, base_len, "lc_regex_parse");
        if (rc)
            return rc;
        CU_TRY(cudaMemcpyAsync(e->in.p, base, base_len, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->ev_off.p, ev_off, n * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->ev_len.p, ev_len, n * 4, cudaMemcpyHostToDevice, e->stream));
        rc = lc_regex_parse_dev(e, re, e->in.as<uint8_t>(), base_len, e->ev_off.as<uint32_t>(),
                                e->ev_len.as<uint32_t>(), n, nkeys, e->out_a.as<uint8_t>(), e->out_b.as<uint32_t>(),
                                e->out_c.as<uint32_t>());
        if (rc) {
            cudaStreamSynchronize(e->stream);
            return rc;
        }
        CU_TRY(cudaMemcpyAsync(status, e->out_a.p, n, cudaMemcpyDeviceToHost, e->stream));
        if (G) {
            CU_TRY(cudaMemcpyAsync(cap_off, e->out_b.p, n * G * 4, cudaMemcpyDeviceToHost, e->stream));
            CU_TRY(cudaMemcpyAsync(cap_len, e->out_c.p, n * G * 4, cudaMemcpyDeviceToHost, e->stream));
        }
        CU_TRY(cudaStreamSynchronize(e->stream));
        return LC_OK;
    }
    uint8_t* d_in = e->in.as<uint8_t>();
    auto run = [&](uint64_t, uint64_t i0, uint64_t cnt, uint64_t span) {
        return regex_parse_dev_impl(e, re, d_in, base_len, span, e->ev_off.as<uint32_t>() + i0,
                                    e->ev_len.as<uint32_t>() + i0, 1, cnt, nkeys, e->out_a.as<uint8_t>() + i0,
                                    e->out_b.as<uint32_t>() + i0 * G, e->out_c.as<uint32_t>() + i0 * G, false);
    };
    auto down = [&](uint64_t, uint64_t i0, uint64_t cnt) {
        CU_TRY(cudaMemcpyAsync(status + i0, e->out_a.as<uint8_t>() + i0, cnt, cudaMemcpyDeviceToHost, e->s_d2h));
        if (G) {
            CU_TRY(cudaMemcpyAsync(cap_off + i0 * G, e->out_b.as<uint32_t>() + i0 * G, cnt * G * 4,
                                   cudaMemcpyDeviceToHost, e->s_d2h));
            CU_TRY(cudaMemcpyAsync(cap_len + i0 * G, e->out_c.as<uint32_t>() + i0 * G, cnt * G * 4,
                                   cudaMemcpyDeviceToHost, e->s_d2h));
        }
        return (int)LC_OK;
    };
    return pipelined_events(e, base, base_len, ev_off, ev_len, n, nchunks, "lc_regex_parse", run, down);
}

// Batched event groups: span k (one group's arena bytes, ideally pinned: lc_host_alloc) is uploaded to
// [span_dst[k], + span_len[k]) of ONE packed device arena; the event table addresses the packed arena.  Chunks of
// ~32 MB of whole spans flow through the same three-stream pipeline as lc_regex_parse.
int lc_regex_parse_packed(lc_engine_t* e, const lc_regex_t* re, uint64_t nspans, const uint8_t* const* span_ptr,
                          const uint32_t* span_len, const uint32_t* span_dst, const uint64_t* span_first_ev,
                          uint64_t packed_len, const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n,
                          uint32_t nkeys, uint8_t* status, uint32_t* cap_off, uint32_t* cap_len) {
    return lc_regex_parse_packed_cb(e, re, nspans, span_ptr, span_len, span_dst, span_first_ev, packed_len, ev_off, ev_len,
                                    n, nkeys, status, cap_off, cap_len, nullptr, nullptr);
}

int lc_regex_parse_packed_cb(lc_engine_t* e, const lc_regex_t* re, uint64_t nspans, const uint8_t* const* span_ptr,
                             const uint32_t* span_len, const uint32_t* span_dst, const uint64_t* span_first_ev,
                             uint64_t packed_len, const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n,
                             uint32_t nkeys, uint8_t* status, uint32_t* cap_off, uint32_t* cap_len,
                             lc_spans_done_fn on_done, void* ctx) {
    if (!e || !re || (nspans && (!span_ptr || !span_len || !span_dst || !span_first_ev)) ||
        (n && (!ev_off || !ev_len || !status)))
        return fail(LC_ERR_INVALID_ARG, "lc_regex_parse_packed: bad arguments");
    int rc = check_regex_usable(re, "lc_regex_parse_packed");
    if (rc)
        return rc;
    if (n == 0)
        return LC_OK;
    const uint32_t G = re->res.ngroups;
    if (G && (!cap_off || !cap_len))
        return fail(LC_ERR_INVALID_ARG, "lc_regex_parse_packed: bad arguments");
    if (packed_len >= 0xFFFFFFF0ull || n >= (1ull << 30))
        return fail(LC_ERR_TOO_LARGE, "packed arena must be < 4 GiB and < 2^30 events per call");
    if (nspans == 0 || span_first_ev[0] != 0 || span_first_ev[nspans] != n)
        return fail(LC_ERR_INVALID_ARG, "lc_regex_parse_packed: span_first_ev must cover [0, n]");
    rc = bind(e);
    if (rc)
        return rc;
    CU_TRY(e->in.ensure(packed_len + 16));
    CU_TRY(e->ev_off.ensure(n * 4));
    CU_TRY(e->ev_len.ensure(n * 4));
    CU_TRY(e->out_a.ensure(n));
    CU_TRY(e->out_b.ensure(n * G * 4 + 4));
    CU_TRY(e->out_c.ensure(n * G * 4 + 4));
    // chunk boundaries (in spans): ~32 MB of arena bytes each, at most 64 chunks
    std::vector<uint64_t> cut{0};
    {
        const uint64_t target = std::max<uint64_t>(32ull << 20, packed_len / 64 + 1);
        uint64_t acc = 0;
        for (uint64_t k = 0; k < nspans; ++k) {
            acc += span_len[k];
            if (acc >= target && k + 1 < nspans) {
                cut.push_back(k + 1);
                acc = 0;
            }
        }
        cut.push_back(nspans);
    }
    const uint64_t nchunks = cut.size() - 1;
    rc = ensure_copy_streams(e, (int)nchunks);
    if (rc)
        return rc;
    uint8_t* d_in = e->in.as<uint8_t>();
    uint32_t* d_off = e->ev_off.as<uint32_t>();
    uint32_t* d_len = e->ev_len.as<uint32_t>();
    auto drain = [&](int code) {
        cudaStreamSynchronize(e->s_h2d);
        cudaStreamSynchronize(e->stream);
        cudaStreamSynchronize(e->s_d2h);
        return code;
    };
#define LC_PIPE_TRY(expr)                                                                                              \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess)                                                                                         \
            return drain(fail(LC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)));                        \
    } while (0)
    LC_PIPE_TRY(cudaEventRecord(e->ev_comp[0], e->stream));
    LC_PIPE_TRY(cudaStreamWaitEvent(e->s_h2d, e->ev_comp[0], 0));
    for (uint64_t c = 0; c < nchunks; ++c) {
        const uint64_t k0 = cut[c], k1 = cut[c + 1];
        const uint64_t i0 = span_first_ev[k0], i1 = span_first_ev[k1], cnt = i1 - i0;
        uint64_t span_bytes = 0;
        for (uint64_t k = k0; k < k1; ++k) {
            const uint64_t lo = span_dst[k], hi = lo + span_len[k];
            if ((lo & 15u) || hi > packed_len || span_first_ev[k] > span_first_ev[k + 1] ||
                (k + 1 < nspans && hi > span_dst[k + 1]))
                return drain(fail(LC_ERR_INVALID_ARG, "lc_regex_parse_packed: spans must be 16-byte aligned, ordered "
                                                      "and inside the packed arena"));
            for (uint64_t i = span_first_ev[k]; i < span_first_ev[k + 1]; ++i)
                if (ev_off[i] < lo || (uint64_t)ev_off[i] + ev_len[i] > hi)
                    return drain(fail(LC_ERR_INVALID_ARG, "lc_regex_parse_packed: event outside its span"));
            if (span_len[k])
                LC_PIPE_TRY(cudaMemcpyAsync(d_in + lo, span_ptr[k], span_len[k], cudaMemcpyHostToDevice, e->s_h2d));
            span_bytes += span_len[k];
        }
        if (cnt) {
            LC_PIPE_TRY(cudaMemcpyAsync(d_off + i0, ev_off + i0, cnt * 4, cudaMemcpyHostToDevice, e->s_h2d));
            LC_PIPE_TRY(cudaMemcpyAsync(d_len + i0, ev_len + i0, cnt * 4, cudaMemcpyHostToDevice, e->s_h2d));
        }
        LC_PIPE_TRY(cudaEventRecord(e->ev_h2d[c], e->s_h2d));
        if (!cnt) {
            LC_PIPE_TRY(cudaEventRecord(e->ev_d2h[c], e->s_d2h));
            continue;
        }
        LC_PIPE_TRY(cudaStreamWaitEvent(e->stream, e->ev_h2d[c], 0));
        rc = regex_parse_dev_impl(e, re, d_in, packed_len, span_bytes, d_off + i0, d_len + i0, 1, cnt, nkeys,
                                  e->out_a.as<uint8_t>() + i0, e->out_b.as<uint32_t>() + i0 * G,
                                  e->out_c.as<uint32_t>() + i0 * G, false);
        if (rc)
            return drain(rc);
        LC_PIPE_TRY(cudaEventRecord(e->ev_comp[c], e->stream));
        LC_PIPE_TRY(cudaStreamWaitEvent(e->s_d2h, e->ev_comp[c], 0));
        LC_PIPE_TRY(cudaMemcpyAsync(status + i0, e->out_a.as<uint8_t>() + i0, cnt, cudaMemcpyDeviceToHost, e->s_d2h));
        if (G) {
            LC_PIPE_TRY(cudaMemcpyAsync(cap_off + i0 * G, e->out_b.as<uint32_t>() + i0 * G, cnt * G * 4,
                                        cudaMemcpyDeviceToHost, e->s_d2h));
            LC_PIPE_TRY(cudaMemcpyAsync(cap_len + i0 * G, e->out_c.as<uint32_t>() + i0 * G, cnt * G * 4,
                                        cudaMemcpyDeviceToHost, e->s_d2h));
        }
        LC_PIPE_TRY(cudaEventRecord(e->ev_d2h[c], e->s_d2h));
    }
    if (on_done) {
        // hand the chunks back in order as their result tables land, while later chunks are still on the GPU
        for (uint64_t c = 0; c < nchunks; ++c) {
            LC_PIPE_TRY(cudaEventSynchronize(e->ev_d2h[c]));
            on_done(ctx, cut[c], cut[c + 1] - cut[c]);
        }
    }
#undef LC_PIPE_TRY
    return drain(LC_OK);
}

// ------------------------------------------------------------------------------------------------ multi-pattern
int lc_regex_parse_multi_dev(lc_engine_t* e, const lc_regex_t* const* res, uint32_t npat, const uint32_t* nkeys,
                             const uint8_t* d_base, uint64_t base_len, const uint32_t* d_ev_off,
                             const uint32_t* d_ev_len, uint64_t n, const uint8_t* d_sel, uint8_t* d_which,
                             uint8_t* d_status, uint32_t row_pitch, uint32_t* d_cap_off, uint32_t* d_cap_len) {
    if (!e || !res || !nkeys || npat == 0 || npat > lck::LC_MULTI_MAX)
        return fail(LC_ERR_INVALID_ARG, "lc_regex_parse_multi_dev: bad arguments (1..8 patterns)");
    int rc;
    uint32_t gmax = 0, max_nregs = 0;
    bool slow = false;
    for (uint32_t p = 0; p < npat; ++p) {
        if (!res[p])
            return fail(LC_ERR_INVALID_ARG, "lc_regex_parse_multi_dev: NULL pattern");
        if ((rc = check_regex_usable(res[p], "lc_regex_parse_multi")))
            return rc;
        if (res[p]->res.tdfa_blob.empty())
            return fail(LC_ERR_REGEX_UNSUPPORTED,
                        "lc_regex_parse_multi: pattern " + std::to_string(p) +
                            " has no single-pass tables (too many states / registers); parse it with lc_regex_parse");
        const LcTdfaHeader* th = reinterpret_cast<const LcTdfaHeader*>(res[p]->res.tdfa_blob.data());
        gmax = std::max(gmax, th->ngroups);
        max_nregs = std::max(max_nregs, th->nregs);
        slow = slow || th->has_slow != 0;
    }
    if (row_pitch < gmax)
        return fail(LC_ERR_INVALID_ARG, "lc_regex_parse_multi_dev: row_pitch smaller than the largest group count");
    if (n == 0)
        return LC_OK;
    if (base_len >= 0xFFFFFFF0ull || n >= (1ull << 30))
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB and < 2^30 events per call");
    rc = bind(e);
    if (rc)
        return rc;
    Small* ds = e->small.as<Small>();
    const size_t smem_max = (size_t)e->smem_per_block_optin;
    lck::TdfaMultiArgs all;
    memset(&all, 0, sizeof all);
    all.npat = npat;
    for (uint32_t p = 0; p < npat; ++p) {
        CU_TRY(engine_blob(e, res[p], &all.blob[p], 3));
        all.blob_bytes[p] = (uint32_t)res[p]->res.tdfa_blob.size();
        all.nkeys[p] = nkeys[p];
    }
    // ragged batches: same device-side decision as the single-pattern path
    const uint32_t* d_order = nullptr;
    const uint32_t* d_order_flag = nullptr;
    if (n >= 4096 && e->length_order != 0 && (e->length_order == 1 || base_len / n >= 1024)) {
        CU_TRY(e->order.ensure(n * 4 + 512));
        uint32_t* hist = e->order.as<uint32_t>() + n;
        lck::launch_length_order(d_ev_len, n, hist, e->order.as<uint32_t>(), e->stream);
        e->launches += 3;
        d_order = e->order.as<uint32_t>();
        d_order_flag = e->length_order == 1 ? nullptr : hist + 64;
    }
    // Group consecutive patterns whose tables fit shared memory together (and below shared address 64 Ki: the pair
    // tables are addressed with 16 bits).  One group = one launch; normally everything is one group.
    uint32_t p0 = 0;
    bool first = true;
    while (p0 < npat) {
        lck::TdfaMultiArgs ga;
        memset(&ga, 0, sizeof ga);
        uint32_t warps = 0;
        uint32_t p1 = p0;
        while (p1 < npat) {
            lck::TdfaMultiArgs tryg = ga;
            tryg.blob[tryg.npat] = all.blob[p1];
            tryg.blob_bytes[tryg.npat] = all.blob_bytes[p1];
            tryg.nkeys[tryg.npat] = all.nkeys[p1];
            tryg.npat++;
            // last pair table must end below 64 Ki (LC_TDFA_REBASE_ROOM covers the window base and static shared memory)
            const bool addressable = LC_TDFA_REBASE_ROOM + lck::tdfa_multi_table_bytes(tryg) <= 65535u;
            uint32_t w = e->max_warps;
            while (w > 4 && lck::tdfa_multi_smem_bytes(tryg, max_nregs, w * 32) > smem_max)
                w -= 2;
            const bool fits = lck::tdfa_multi_smem_bytes(tryg, max_nregs, w * 32) <= smem_max && w >= 16;
            if (tryg.npat > 1 && (!(addressable && fits) || e->multi_split))
                break;
            if (tryg.npat == 1 && lck::tdfa_multi_smem_bytes(tryg, max_nregs, w * 32) > smem_max)
                return fail(LC_ERR_REGEX_UNSUPPORTED, "lc_regex_parse_multi: tables do not fit shared memory");
            ga = tryg;
            warps = w;
            ++p1;
        }
        const uint32_t threads = warps * 32;
        const uint32_t grid = (uint32_t)std::min<uint64_t>((n + threads - 1) / threads, (uint64_t)e->num_sms);
        CU_TRY(cudaMemsetAsync(&ds->next_batch, 0, sizeof ds->next_batch, e->stream));
        if (first)
            CU_TRY(cudaMemsetAsync(&ds->overflow, 0, sizeof ds->overflow, e->stream));
        int er = lck::launch_regex_tdfa_multi(ga, p0, !first, slow, max_nregs, d_base, d_ev_off, d_ev_len, n, d_sel,
                                              d_which, d_status, d_cap_off, d_cap_len, row_pitch, threads, grid,
                                              &ds->next_batch, &ds->overflow, d_order, d_order_flag, e->stream);
        e->launches++;
        if (er)
            return fail(LC_ERR_CUDA, std::string("regex kernel launch: ") + cudaGetErrorString((cudaError_t)er));
        first = false;
        p0 = p1;
    }
    if (base_len >= 65535) {
        lck::launch_regex_tdfa_long(all, d_base, d_ev_off, d_ev_len, 1, n, d_sel, d_which, d_status, d_cap_off,
                                    d_cap_len, row_pitch, &ds->overflow, false, e->stream);
        e->launches++;
        CU_TRY(cudaGetLastError());
    }
    return LC_OK;
}

int lc_regex_parse_multi(lc_engine_t* e, const lc_regex_t* const* res, uint32_t npat, const uint32_t* nkeys,
                         const uint8_t* base, uint64_t base_len, const uint32_t* ev_off, const uint32_t* ev_len,
                         uint64_t n, const uint8_t* sel, uint8_t* which, uint8_t* status, uint32_t row_pitch,
                         uint32_t* cap_off, uint32_t* cap_len) {
    if (!e || !res || !nkeys || (n && (!ev_off || !ev_len || !status || !which)) || (base_len && !base) ||
        (n && row_pitch && (!cap_off || !cap_len)))
        return fail(LC_ERR_INVALID_ARG, "lc_regex_parse_multi: bad arguments");
    if (n == 0)
        return LC_OK;
    if (base_len >= 0xFFFFFFF0ull || n >= (1ull << 30))
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB and < 2^30 events per call");
    int rc = bind(e);
    if (rc)
        return rc;
    const uint64_t G = row_pitch;
    CU_TRY(e->in.ensure(base_len + 16));
    CU_TRY(e->ev_off.ensure(n * 4));
    CU_TRY(e->ev_len.ensure(n * 4));
    CU_TRY(e->out_a.ensure(n));
    CU_TRY(e->out_b.ensure(n * G * 4 + 4));
    CU_TRY(e->out_c.ensure(n * G * 4 + 4));
    CU_TRY(e->out_d.ensure(n));
    CU_TRY(e->flags.ensure(n));
    if (sel) {
        for (uint64_t i = 0; i < n; ++i)
            if (sel[i] != 0xFFu && sel[i] >= npat)
                return fail(LC_ERR_INVALID_ARG, "lc_regex_parse_multi: selector names a pattern that does not exist");
        CU_TRY(cudaMemcpyAsync(e->flags.p, sel, n, cudaMemcpyHostToDevice, e->stream));
    }
    const uint8_t* d_sel = sel ? e->flags.as<uint8_t>() : nullptr;
    uint8_t* d_in = e->in.as<uint8_t>();
    const uint64_t nchunks = pipeline_chunks(base_len, n);
    auto run = [&](uint64_t, uint64_t i0, uint64_t cnt, uint64_t) {
        return lc_regex_parse_multi_dev(e, res, npat, nkeys, d_in, base_len, e->ev_off.as<uint32_t>() + i0,
                                        e->ev_len.as<uint32_t>() + i0, cnt, d_sel ? d_sel + i0 : nullptr,
                                        e->out_d.as<uint8_t>() + i0, e->out_a.as<uint8_t>() + i0, row_pitch,
                                        e->out_b.as<uint32_t>() + i0 * G, e->out_c.as<uint32_t>() + i0 * G);
    };
    auto down_on = [&](cudaStream_t st, uint64_t i0, uint64_t cnt) {
        CU_TRY(cudaMemcpyAsync(status + i0, e->out_a.as<uint8_t>() + i0, cnt, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(which + i0, e->out_d.as<uint8_t>() + i0, cnt, cudaMemcpyDeviceToHost, st));
        if (G) {
            CU_TRY(cudaMemcpyAsync(cap_off + i0 * G, e->out_b.as<uint32_t>() + i0 * G, cnt * G * 4,
                                   cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaMemcpyAsync(cap_len + i0 * G, e->out_c.as<uint32_t>() + i0 * G, cnt * G * 4,
                                   cudaMemcpyDeviceToHost, st));
        }
        return (int)LC_OK;
    };
    if (nchunks == 1) {
        rc = check_events(ev_off, ev_len, n, base_len, "lc_regex_parse_multi");
        if (rc)
            return rc;
        CU_TRY(cudaMemcpyAsync(e->in.p, base, base_len, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->ev_off.p, ev_off, n * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->ev_len.p, ev_len, n * 4, cudaMemcpyHostToDevice, e->stream));
        rc = run(0, 0, n, base_len);
        if (!rc)
            rc = down_on(e->stream, 0, n);
        cudaError_t er = cudaStreamSynchronize(e->stream);
        if (!rc && er != cudaSuccess)
            return fail(LC_ERR_CUDA, cudaGetErrorString(er));
        return rc;
    }
    auto down = [&](uint64_t, uint64_t i0, uint64_t cnt) { return down_on(e->s_d2h, i0, cnt); };
    return pipelined_events(e, base, base_len, ev_off, ev_len, n, nchunks, "lc_regex_parse_multi", run, down);
}

int lc_regex_prefix_match(lc_engine_t* e, const lc_regex_t* re, const uint8_t* base, uint64_t base_len,
                          const uint32_t* ev_off, const uint32_t* ev_len, uint64_t n, uint8_t* out_match) {
    if (!e || !re || (n && (!ev_off || !ev_len || !out_match)) || (base_len && !base))
        return fail(LC_ERR_INVALID_ARG, "lc_regex_prefix_match: bad arguments");
    int rc = check_regex_usable(re, "lc_regex_prefix_match");
    if (rc)
        return rc;
    if (n == 0)
        return LC_OK;
    rc = check_events(ev_off, ev_len, n, base_len, "lc_regex_prefix_match");
    if (rc)
        return rc;
    rc = bind(e);
    if (rc)
        return rc;
    const void* d_blob;
    CU_TRY(engine_blob(e, re, &d_blob));
    CU_TRY(e->in.ensure(base_len + 16));
    CU_TRY(e->ev_off.ensure(n * 4));
    CU_TRY(e->ev_len.ensure(n * 4));
    CU_TRY(e->out_a.ensure(n));
    CU_TRY(cudaMemcpyAsync(e->in.p, base, base_len, cudaMemcpyHostToDevice, e->stream));
    CU_TRY(cudaMemcpyAsync(e->ev_off.p, ev_off, n * 4, cudaMemcpyHostToDevice, e->stream));
    CU_TRY(cudaMemcpyAsync(e->ev_len.p, ev_len, n * 4, cudaMemcpyHostToDevice, e->stream));
    lck::launch_prefix_match(d_blob, e->in.as<uint8_t>(), e->ev_off.as<uint32_t>(), e->ev_len.as<uint32_t>(), n,
                             e->out_a.as<uint8_t>(), e->stream);
    e->launches++;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out_match, e->out_a.p, n, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return LC_OK;
}

// ------------------------------------------------------------------------------------------------ multiline
int lc_multiline_split_dev(lc_engine_t* e, const uint8_t* d_buf, uint64_t len, const lc_regex_t* start,
                           const lc_regex_t* cont, const lc_regex_t* end, int discard_unmatched, uint32_t* d_out_off,
                           uint32_t* d_out_len, uint8_t* d_out_flags, uint64_t cap, uint64_t* n_out,
                           uint64_t counters[3]) {
    if (!e || !n_out || (len && !d_buf))
        return fail(LC_ERR_INVALID_ARG, "lc_multiline_split_dev: bad arguments");
    *n_out = 0;
    int rc;
    for (const lc_regex_t* r : {start, cont, end})
        if (r && (rc = check_regex_usable(r, "lc_multiline_split")))
            return rc;
    if (len == 0)
        return LC_OK;
    if (len >= 0xFFFFFFF0ull)
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB per call");
    rc = bind(e);
    if (rc)
        return rc;
    lck::MlConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    CU_TRY(engine_blob(e, start, &cfg.blob_start));
    CU_TRY(engine_blob(e, cont, &cfg.blob_cont));
    CU_TRY(engine_blob(e, end, &cfg.blob_end));
    cfg.discard = discard_unmatched;
    fill_probe_slot(cfg, 0, start);
    fill_probe_slot(cfg, 1, cont);
    fill_probe_slot(cfg, 2, end);

    Small* ds = e->small.as<Small>();
    Small* hs = (Small*)e->h_small;
    uint32_t shift = (uint32_t)((uintptr_t)d_buf & 15u);
    static const bool fused = [] {
        const char* k = getenv("LC_B200_ML_FUSED"); // A/B knob: 0 = the five-kernel formulation
        return !(k && !strcmp(k, "0"));
    }();
    if (fused) {
        // Two launches, no host round trip in between: (1) split + per-line probes, (2) state scan + counts + slots +
        // emission; the second reads the line count from device memory.  The line table is sized from an estimate and
        // the call repeats with the exact size in the rare case it was too small.
        uint64_t lcap = len / 24 + 4096;
        for (;;) {
            if (lcap > len)
                lcap = len;
            if (lcap > 0x3FFFFFF0ull)
                lcap = 0x3FFFFFF0ull;
            CU_TRY(e->lines_off.ensure((lcap + 1) * 4));
            CU_TRY(e->lines_len.ensure((lcap + 1) * 4));
            CU_TRY(e->flags.ensure(lcap + 1));
            const uint32_t ftiles = lck::ml_fused_tiles(lcap);
            DescPlan plan;
            rc = prep_desc(e, lck::split_tiles(len, shift), ftiles, ftiles, plan);
            if (rc)
                return rc;
            CU_TRY(e->split_scratch.ensure(lck::split_scratch_bytes(len, true)));
            e->launches += lck::launch_split_probe(cfg, d_buf, (uint32_t)len, e->lines_off.as<uint32_t>(),
                                                   e->lines_len.as<uint32_t>(), e->flags.as<uint8_t>(), (uint32_t)lcap,
                                                   plan.r[0], &ds->tickets[0], &ds->n_out, &ds->total_chars,
                                                   e->split_scratch.as<uint64_t>(), e->stream) - 1;
            static const bool ml_lookback = [] {
                const char* v = getenv("LC_B200_ML"); // A/B knob: "lookback" = ml_fused_kernel (two chained look-backs)
                return v && !strcmp(v, "lookback");
            }();
            if (ml_lookback) {
                lck::launch_ml_fused(cfg, e->flags.as<uint8_t>(), e->lines_off.as<uint32_t>(),
                                     e->lines_len.as<uint32_t>(), &ds->n_out, (uint32_t)lcap, (uint32_t)len, d_out_off,
                                     d_out_len, d_out_flags, cap, plan.r[1], plan.r[2], &ds->tickets[1], ds->counters,
                                     &ds->total, e->stream);
                e->launches += 2;
            } else {
                CU_TRY(e->state.ensure((size_t)lck::ml_pass_scratch_bytes(lcap)));
                e->launches += 1 + lck::launch_ml_passes(cfg, e->flags.as<uint8_t>(), e->lines_off.as<uint32_t>(),
                                                         e->lines_len.as<uint32_t>(), &ds->n_out, (uint32_t)lcap,
                                                         (uint32_t)len, d_out_off, d_out_len, d_out_flags, cap,
                                                         e->state.as<uint64_t>(), ds->counters, &ds->total, e->stream);
            }
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaMemcpyAsync(hs, ds, sizeof(Small), cudaMemcpyDeviceToHost, e->stream));
            CU_TRY(cudaStreamSynchronize(e->stream));
            if (hs->total_chars >= (1ull << 30) - 2)
                return fail(LC_ERR_TOO_LARGE, "more than 2^30 lines in one call");
            if (hs->n_out <= lcap)
                break;
            lcap = hs->n_out;
        }
        *n_out = hs->total;
        if (counters) {
            counters[0] += hs->counters[0];
            counters[1] += hs->n_out;
            counters[2] += hs->counters[1];
        }
        if (*n_out > cap)
            return fail(LC_ERR_CAPACITY, "lc_multiline_split: output capacity too small");
        return LC_OK;
    }
    // 1. line table (grow the workspace until it fits; typical logs fit the first estimate)
    uint64_t lcap = len / 24 + 4096;
    uint64_t n = 0;
    for (;;) {
        if (lcap > len)
            lcap = len;
        CU_TRY(e->lines_off.ensure((lcap + 1) * 4));
        CU_TRY(e->lines_len.ensure((lcap + 1) * 4));
        DescPlan plan;
        rc = prep_desc(e, lck::split_tiles(len, shift), 0, 0, plan);
        if (rc)
            return rc;
        CU_TRY(e->split_scratch.ensure(lck::split_scratch_bytes(len, false)));
        e->launches += lck::launch_split(d_buf, (uint32_t)len, '\n', e->lines_off.as<uint32_t>(),
                                         e->lines_len.as<uint32_t>(),
                                         (uint32_t)(lcap > 0x3FFFFFFFull ? 0x3FFFFFFFull : lcap), plan.r[0],
                                         &ds->tickets[0], &ds->n_out, &ds->total_chars,
                                         e->split_scratch.as<uint64_t>(), e->stream);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(hs, ds, sizeof(Small), cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        if (hs->total_chars >= (1ull << 30) - 2)
            return fail(LC_ERR_TOO_LARGE, "more than 2^30 lines in one call");
        n = hs->n_out;
        if (n <= lcap)
            break;
        lcap = n;
    }
    if (n >= (1ull << 30) - 2)
        return fail(LC_ERR_TOO_LARGE, "more than 2^30 lines in one call");
    // 2. per-line prefix probes
    CU_TRY(e->flags.ensure(n + 1));
    lck::launch_ml_probe(cfg, d_buf, e->lines_off.as<uint32_t>(), e->lines_len.as<uint32_t>(), n,
                         e->flags.as<uint8_t>(), e->stream);
    // 3. state scan + event counts, 4. output slots
    DescPlan plan;
    rc = prep_desc(e, 0, lck::scan_tiles(n + 1), lck::scan_tiles(n + 1), plan);
    if (rc)
        return rc;
    CU_TRY(e->state.ensure((n + 1) * 4));
    CU_TRY(e->cnt.ensure((n + 1) * 4));
    CU_TRY(e->pos.ensure((n + 1) * 8));
    lck::launch_ml_state(cfg, e->flags.as<uint8_t>(), e->lines_len.as<uint32_t>(), n, e->state.as<uint32_t>(),
                         e->cnt.as<uint32_t>(), plan.r[1],
                         &ds->tickets[1], e->stream);
    lck::launch_exclusive_sum(e->cnt.as<uint32_t>(), n + 1, e->pos.as<uint64_t>(), &ds->total, plan.r[2],
                              &ds->tickets[2], e->stream);
    // 5. emission
    lck::launch_ml_emit(cfg, e->flags.as<uint8_t>(), e->lines_off.as<uint32_t>(), e->lines_len.as<uint32_t>(), n,
                        (uint32_t)len, e->state.as<uint32_t>(), e->pos.as<uint64_t>(), d_out_off, d_out_len,
                        d_out_flags, cap, ds->counters, e->stream);
    e->launches += 4;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(hs, ds, sizeof(Small), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    *n_out = hs->total;
    if (counters) {
        counters[0] += hs->counters[0];
        counters[1] += n;
        counters[2] += hs->counters[1];
    }
    if (*n_out > cap)
        return fail(LC_ERR_CAPACITY, "lc_multiline_split: output capacity too small");
    return LC_OK;
}

int lc_multiline_split(lc_engine_t* e, const uint8_t* buf, uint64_t len, const lc_regex_t* start,
                       const lc_regex_t* cont, const lc_regex_t* end, int discard_unmatched, uint32_t* out_off,
                       uint32_t* out_len, uint8_t* out_flags, uint64_t cap, uint64_t* n_out, uint64_t counters[3]) {
    if (!e || !n_out || (len && !buf))
        return fail(LC_ERR_INVALID_ARG, "lc_multiline_split: bad arguments");
    *n_out = 0;
    if (len == 0)
        return LC_OK;
    if (len >= 0xFFFFFFF0ull)
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB per call");
    int rc = bind(e);
    if (rc)
        return rc;
    uint64_t dcap = cap < len ? cap : len; // an output event covers at least one input byte or one line
    CU_TRY(e->in.ensure(len + 16));
    CU_TRY(e->out_a.ensure((dcap + 1) * 4));
    CU_TRY(e->out_b.ensure((dcap + 1) * 4));
    CU_TRY(e->out_c.ensure(dcap + 1));
    CU_TRY(cudaMemcpyAsync(e->in.p, buf, len, cudaMemcpyHostToDevice, e->stream));
    rc = lc_multiline_split_dev(e, e->in.as<uint8_t>(), len, start, cont, end, discard_unmatched,
                                e->out_a.as<uint32_t>(), e->out_b.as<uint32_t>(), e->out_c.as<uint8_t>(), dcap, n_out,
                                counters);
    if (rc)
        return rc;
    if (*n_out) {
        CU_TRY(cudaMemcpyAsync(out_off, e->out_a.p, *n_out * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaMemcpyAsync(out_len, e->out_b.p, *n_out * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaMemcpyAsync(out_flags, e->out_c.p, *n_out, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
    }
    return LC_OK;
}

// ------------------------------------------------------------------------------------------------ last incomplete log
int lc_remove_last_incomplete_log_dev(lc_engine_t* e, const uint8_t* d_buf, uint64_t len, const lc_regex_t* start,
                                      const lc_regex_t* end, int allow_rollback, uint64_t* keep_bytes,
                                      int32_t* rollback_line_feeds) {
    if (!e || !keep_bytes || !rollback_line_feeds || (len && !d_buf))
        return fail(LC_ERR_INVALID_ARG, "lc_remove_last_incomplete_log_dev: bad arguments");
    int rc;
    for (const lc_regex_t* r : {start, end})
        if (r && (rc = check_regex_usable(r, "lc_remove_last_incomplete_log")))
            return rc;
    *keep_bytes = len;
    if (!allow_rollback || len == 0) // :1999-2001: nothing to do, the caller's count stays as it is
        return LC_OK;
    if (len >= 0x7FFFFFF0ull)
        return fail(LC_ERR_TOO_LARGE, "chunk must be < 2 GiB (the reference's sizes are int32_t)");
    rc = bind(e);
    if (rc)
        return rc;
    lck::MlConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    CU_TRY(engine_blob(e, start, &cfg.blob_start));
    cfg.blob_cont = nullptr;
    CU_TRY(engine_blob(e, end, &cfg.blob_end));
    cfg.discard = 0;
    fill_probe_slot(cfg, 0, start);
    fill_probe_slot(cfg, 2, end);
    Small* ds = e->small.as<Small>();
    Small* hs = (Small*)e->h_small;
    const uint32_t shift = (uint32_t)((uintptr_t)d_buf & 15u);
    uint64_t lcap = len / 24 + 4096;
    for (;;) {
        if (lcap > len)
            lcap = len;
        CU_TRY(e->lines_off.ensure((lcap + 1) * 4));
        CU_TRY(e->lines_len.ensure((lcap + 1) * 4));
        CU_TRY(e->flags.ensure(lcap + 1));
        DescPlan plan;
        rc = prep_desc(e, lck::split_tiles(len, shift), 0, 0, plan);
        if (rc)
            return rc;
        CU_TRY(e->split_scratch.ensure(lck::split_scratch_bytes(len, true)));
        e->launches += lck::launch_split_probe(cfg, d_buf, (uint32_t)len, e->lines_off.as<uint32_t>(),
                                               e->lines_len.as<uint32_t>(), e->flags.as<uint8_t>(), (uint32_t)lcap,
                                               plan.r[0], &ds->tickets[0], &ds->n_out, &ds->total_chars,
                                               e->split_scratch.as<uint64_t>(), e->stream) - 1;
        lck::launch_last_record(e->flags.as<uint8_t>(), e->lines_off.as<uint32_t>(), e->lines_len.as<uint32_t>(),
                                &ds->n_out, (uint32_t)lcap, (uint32_t)len, start != nullptr, end != nullptr,
                                ds->counters, e->stream);
        e->launches += 2;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(hs, ds, sizeof(Small), cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        if (hs->n_out <= lcap)
            break;
        lcap = hs->n_out;
    }
    *keep_bytes = hs->counters[0];
    *rollback_line_feeds = (int32_t)hs->counters[1];
    return LC_OK;
}

int lc_remove_last_incomplete_log(lc_engine_t* e, const uint8_t* buf, uint64_t len, const lc_regex_t* start,
                                  const lc_regex_t* end, int allow_rollback, uint64_t* keep_bytes,
                                  int32_t* rollback_line_feeds) {
    if (!e || !keep_bytes || !rollback_line_feeds || (len && !buf))
        return fail(LC_ERR_INVALID_ARG, "lc_remove_last_incomplete_log: bad arguments");
    *keep_bytes = len;
    if (!allow_rollback || len == 0)
        return LC_OK;
    if (len >= 0x7FFFFFF0ull)
        return fail(LC_ERR_TOO_LARGE, "chunk must be < 2 GiB (the reference's sizes are int32_t)");
    int rc = bind(e);
    if (rc)
        return rc;
    CU_TRY(e->in.ensure(len + 16));
    CU_TRY(cudaMemcpyAsync(e->in.p, buf, len, cudaMemcpyHostToDevice, e->stream));
    return lc_remove_last_incomplete_log_dev(e, e->in.as<uint8_t>(), len, start, end, allow_rollback, keep_bytes,
                                             rollback_line_feeds);
}

// ------------------------------------------------------------------------------------------------ delimiter
int lc_delim_parse_dev(lc_engine_t* e, const uint8_t* d_base, uint64_t base_len, const uint32_t* d_ev_off,
                       const uint32_t* d_ev_len, uint64_t n, const uint8_t* sep, uint32_t sep_len, uint8_t quote,
                       uint32_t nkeys, int extend, int allow_short, uint32_t max_fields, uint8_t* d_status,
                       uint32_t* d_nfields, uint32_t* d_f_off, uint32_t* d_f_len, uint32_t* d_f_dq) {
    return lc_delim_parse_tap_dev(e, d_base, base_len, d_ev_off, d_ev_len, n, sep, sep_len, quote, nkeys, extend,
                                  allow_short, max_fields, d_status, d_nfields, d_f_off, d_f_len, d_f_dq, 0xFFFFFFFFu,
                                  nullptr, nullptr);
}

int lc_delim_parse_tap_dev(lc_engine_t* e, const uint8_t* d_base, uint64_t base_len, const uint32_t* d_ev_off,
                           const uint32_t* d_ev_len, uint64_t n, const uint8_t* sep, uint32_t sep_len, uint8_t quote,
                           uint32_t nkeys, int extend, int allow_short, uint32_t max_fields, uint8_t* d_status,
                           uint32_t* d_nfields, uint32_t* d_f_off, uint32_t* d_f_len, uint32_t* d_f_dq,
                           uint32_t tap_col, uint32_t* d_tap_off, uint32_t* d_tap_len) {
    if (!e || !sep || sep_len < 1 || sep_len > 4 || max_fields == 0 || ((d_tap_off == nullptr) != (d_tap_len == nullptr)) ||
        (d_tap_off && tap_col >= max_fields))
        return fail(LC_ERR_INVALID_ARG, "lc_delim_parse_dev: bad arguments (separator must be 1..4 bytes)");
    if (n == 0)
        return LC_OK;
    if (base_len >= 0xFFFFFFF0ull)
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB per call");
    int rc = bind(e);
    if (rc)
        return rc;
    lck::DelimConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    memcpy(cfg.sep, sep, sep_len);
    cfg.sep_len = sep_len;
    cfg.quote = quote;
    cfg.nkeys = nkeys;
    cfg.extend = extend;
    cfg.allow_short = allow_short;
    cfg.max_fields = max_fields;
    cfg.tap_col = tap_col;
    cfg.tap_off = d_tap_off;
    cfg.tap_len = d_tap_len;
    Small* ds = e->small.as<Small>();
    CU_TRY(cudaMemsetAsync(&ds->next_batch, 0, sizeof ds->next_batch, e->stream));
    lck::launch_delim(cfg, d_base, d_ev_off, d_ev_len, n, d_status, d_nfields, d_f_off, d_f_len, d_f_dq,
                      &ds->next_batch, e->stream);
    e->launches++;
    CU_TRY(cudaGetLastError());
    return LC_OK;
}

int lc_delim_regex_chain(lc_engine_t* e, const uint8_t* base, uint64_t base_len, const uint32_t* ev_off,
                         const uint32_t* ev_len, uint64_t n, const uint8_t* sep, uint32_t sep_len, uint8_t quote,
                         uint32_t nkeys, int extend, int allow_short, uint32_t max_fields, uint8_t* status,
                         uint32_t* nfields, uint32_t* f_off, uint32_t* f_len, uint32_t* f_dq, uint32_t column,
                         const lc_regex_t* re, uint32_t regex_nkeys, uint8_t* re_status, uint32_t* cap_off,
                         uint32_t* cap_len) {
    if (!e || !re || !sep || sep_len < 1 || sep_len > 4 || max_fields == 0 || column >= max_fields ||
        (n && (!ev_off || !ev_len || !status || !nfields || !f_off || !f_len || !f_dq || !re_status)) ||
        (base_len && !base))
        return fail(LC_ERR_INVALID_ARG, "lc_delim_regex_chain: bad arguments");
    int rc = check_regex_usable(re, "lc_delim_regex_chain");
    if (rc)
        return rc;
    if (n == 0)
        return LC_OK;
    const uint32_t G = re->res.ngroups;
    if (G && (!cap_off || !cap_len))
        return fail(LC_ERR_INVALID_ARG, "lc_delim_regex_chain: bad arguments");
    if (base_len >= 0xFFFFFFF0ull || n >= (1ull << 30) || n * (uint64_t)max_fields >= (1ull << 32))
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB, < 2^30 events and < 2^32 columns per call");
    rc = bind(e);
    if (rc)
        return rc;
    const uint64_t MF = max_fields, fbytes = n * MF * 4;
    CU_TRY(e->in.ensure(base_len + 16));
    CU_TRY(e->ev_off.ensure(n * 4));
    CU_TRY(e->ev_len.ensure(n * 4));
    CU_TRY(e->out_a.ensure(n));
    CU_TRY(e->out_b.ensure(n * 4));
    CU_TRY(e->out_c.ensure(fbytes));
    CU_TRY(e->out_d.ensure(fbytes));
    CU_TRY(e->out_e.ensure(fbytes));
    CU_TRY(e->lines_off.ensure(n * 4)); // the tapped column = the regex stage's event table
    CU_TRY(e->lines_len.ensure(n * 4));
    CU_TRY(e->flags.ensure(n));
    CU_TRY(e->cnt.ensure(n * G * 4 + 4));
    CU_TRY(e->pos.ensure(n * G * 4 + 4));
    uint8_t* d_in = e->in.as<uint8_t>();
    auto run = [&](uint64_t, uint64_t i0, uint64_t cnt, uint64_t span) {
        int r = lc_delim_parse_tap_dev(e, d_in, base_len, e->ev_off.as<uint32_t>() + i0, e->ev_len.as<uint32_t>() + i0,
                                       cnt, sep, sep_len, quote, nkeys, extend, allow_short, max_fields,
                                       e->out_a.as<uint8_t>() + i0, e->out_b.as<uint32_t>() + i0,
                                       e->out_c.as<uint32_t>() + i0 * MF, e->out_d.as<uint32_t>() + i0 * MF,
                                       e->out_e.as<uint32_t>() + i0 * MF, column, e->lines_off.as<uint32_t>() + i0,
                                       e->lines_len.as<uint32_t>() + i0);
        if (r)
            return r;
        return regex_parse_dev_impl(e, re, d_in, base_len, span, e->lines_off.as<uint32_t>() + i0,
                                    e->lines_len.as<uint32_t>() + i0, 1, cnt, regex_nkeys, e->flags.as<uint8_t>() + i0,
                                    e->cnt.as<uint32_t>() + i0 * G, e->pos.as<uint32_t>() + i0 * G, false);
    };
    auto down = [&](uint64_t, uint64_t i0, uint64_t cnt) {
        CU_TRY(cudaMemcpyAsync(status + i0, e->out_a.as<uint8_t>() + i0, cnt, cudaMemcpyDeviceToHost, e->s_d2h));
        CU_TRY(cudaMemcpyAsync(nfields + i0, e->out_b.as<uint32_t>() + i0, cnt * 4, cudaMemcpyDeviceToHost, e->s_d2h));
        CU_TRY(cudaMemcpyAsync(f_off + i0 * MF, e->out_c.as<uint32_t>() + i0 * MF, cnt * MF * 4, cudaMemcpyDeviceToHost,
                               e->s_d2h));
        CU_TRY(cudaMemcpyAsync(f_len + i0 * MF, e->out_d.as<uint32_t>() + i0 * MF, cnt * MF * 4, cudaMemcpyDeviceToHost,
                               e->s_d2h));
        CU_TRY(cudaMemcpyAsync(f_dq + i0 * MF, e->out_e.as<uint32_t>() + i0 * MF, cnt * MF * 4, cudaMemcpyDeviceToHost,
                               e->s_d2h));
        CU_TRY(cudaMemcpyAsync(re_status + i0, e->flags.as<uint8_t>() + i0, cnt, cudaMemcpyDeviceToHost, e->s_d2h));
        if (G) {
            CU_TRY(cudaMemcpyAsync(cap_off + i0 * G, e->cnt.as<uint32_t>() + i0 * G, cnt * G * 4, cudaMemcpyDeviceToHost,
                                   e->s_d2h));
            CU_TRY(cudaMemcpyAsync(cap_len + i0 * G, e->pos.as<uint32_t>() + i0 * G, cnt * G * 4, cudaMemcpyDeviceToHost,
                                   e->s_d2h));
        }
        return (int)LC_OK;
    };
    uint64_t nchunks = pipeline_chunks(base_len, n);
    if (nchunks > 1 && nchunks < 8 && n >= 8 * 1024)
        nchunks = 8; // the tables going back are as large as the arena going up: shorter chunks, earlier overlap
    return pipelined_events(e, base, base_len, ev_off, ev_len, n, nchunks, "lc_delim_regex_chain", run, down);
}

int lc_delim_parse(lc_engine_t* e, const uint8_t* base, uint64_t base_len, const uint32_t* ev_off,
                   const uint32_t* ev_len, uint64_t n, const uint8_t* sep, uint32_t sep_len, uint8_t quote,
                   uint32_t nkeys, int extend, int allow_short, uint32_t max_fields, uint8_t* status,
                   uint32_t* nfields, uint32_t* f_off, uint32_t* f_len, uint32_t* f_dq) {
    if (!e || (n && (!ev_off || !ev_len || !status || !nfields || !f_off || !f_len || !f_dq)) || (base_len && !base))
        return fail(LC_ERR_INVALID_ARG, "lc_delim_parse: bad arguments");
    if (n == 0)
        return LC_OK;
    int rc = check_events(ev_off, ev_len, n, base_len, "lc_delim_parse");
    if (rc)
        return rc;
    rc = bind(e);
    if (rc)
        return rc;
    size_t fbytes = (size_t)n * max_fields * 4;
    CU_TRY(e->in.ensure(base_len + 16));
    CU_TRY(e->ev_off.ensure(n * 4));
    CU_TRY(e->ev_len.ensure(n * 4));
    CU_TRY(e->out_a.ensure(n));
    CU_TRY(e->out_b.ensure(n * 4));
    CU_TRY(e->out_c.ensure(fbytes));
    CU_TRY(e->out_d.ensure(fbytes));
    CU_TRY(e->out_e.ensure(fbytes));
    CU_TRY(cudaMemcpyAsync(e->in.p, base, base_len, cudaMemcpyHostToDevice, e->stream));
    CU_TRY(cudaMemcpyAsync(e->ev_off.p, ev_off, n * 4, cudaMemcpyHostToDevice, e->stream));
    CU_TRY(cudaMemcpyAsync(e->ev_len.p, ev_len, n * 4, cudaMemcpyHostToDevice, e->stream));
    rc = lc_delim_parse_dev(e, e->in.as<uint8_t>(), base_len, e->ev_off.as<uint32_t>(), e->ev_len.as<uint32_t>(), n,
                            sep, sep_len, quote, nkeys, extend, allow_short, max_fields, e->out_a.as<uint8_t>(),
                            e->out_b.as<uint32_t>(), e->out_c.as<uint32_t>(), e->out_d.as<uint32_t>(),
                            e->out_e.as<uint32_t>());
    if (rc)
        return rc;
    CU_TRY(cudaMemcpyAsync(status, e->out_a.p, n, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaMemcpyAsync(nfields, e->out_b.p, n * 4, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaMemcpyAsync(f_off, e->out_c.p, fbytes, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaMemcpyAsync(f_len, e->out_d.p, fbytes, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaMemcpyAsync(f_dq, e->out_e.p, fbytes, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return LC_OK;
}

int lc_sls_serialize_logs(lc_engine_t* e, const uint8_t* base, uint64_t base_len, uint64_t n, const uint32_t* ev_time,
                          const uint32_t* ev_time_ns, const uint64_t* ent_begin, const uint32_t* ent_koff,
                          const uint32_t* ent_klen, const uint32_t* ent_voff, const uint32_t* ent_vlen, uint8_t* out,
                          uint64_t out_cap, uint64_t* out_len) {
    if (!e || !out_len || (n && (!ev_time || !ent_begin)) || (base_len && !base))
        return fail(LC_ERR_INVALID_ARG, "lc_sls_serialize_logs: bad arguments");
    *out_len = 0;
    if (n == 0)
        return LC_OK;
    const uint64_t m = ent_begin[n];
    if (m && (!ent_koff || !ent_klen || !ent_voff || !ent_vlen))
        return fail(LC_ERR_INVALID_ARG, "lc_sls_serialize_logs: bad arguments");
    if (base_len >= 0xFFFFFFF0ull || n >= (1ull << 30) || m >= (1ull << 31))
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB, < 2^30 events and < 2^31 contents per call");
    for (uint64_t i = 0; i < n; ++i)
        if (ent_begin[i] > ent_begin[i + 1])
            return fail(LC_ERR_INVALID_ARG, "lc_sls_serialize_logs: ent_begin must be non-decreasing");
    if (ent_begin[0] != 0)
        return fail(LC_ERR_INVALID_ARG, "lc_sls_serialize_logs: ent_begin[0] must be 0");
    int rc = check_events(ent_koff, ent_klen, m, base_len, "lc_sls_serialize_logs (keys)");
    if (!rc)
        rc = check_events(ent_voff, ent_vlen, m, base_len, "lc_sls_serialize_logs (values)");
    if (rc)
        return rc;
    rc = bind(e);
    if (rc)
        return rc;
    // workspace: in = arena, ev_off/ev_len = key spans, out_b/out_c = value spans, lines_off/lines_len = time / ns,
    // pos = ent_begin, lab_sizes / cnt = record / body sizes, lab_off = record offsets, out_d = the wire bytes
    CU_TRY(e->in.ensure(base_len + 16));
    CU_TRY(e->ev_off.ensure(m * 4 + 4));
    CU_TRY(e->ev_len.ensure(m * 4 + 4));
    CU_TRY(e->out_b.ensure(m * 4 + 4));
    CU_TRY(e->out_c.ensure(m * 4 + 4));
    CU_TRY(e->lines_off.ensure(n * 4));
    CU_TRY(e->lines_len.ensure(n * 4));
    CU_TRY(e->pos.ensure((n + 1) * 8));
    CU_TRY(e->lab_sizes.ensure(n * 4));
    CU_TRY(e->cnt.ensure(n * 4));
    CU_TRY(e->lab_off.ensure(n * 8));
    DescPlan plan;
    rc = prep_desc(e, 0, 0, lck::scan_tiles(n), plan);
    if (rc)
        return rc;
    Small* ds = e->small.as<Small>();
    Small* hs = (Small*)e->h_small;
    CU_TRY(cudaMemcpyAsync(e->in.p, base, base_len, cudaMemcpyHostToDevice, e->stream));
    if (m) {
        CU_TRY(cudaMemcpyAsync(e->ev_off.p, ent_koff, m * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->ev_len.p, ent_klen, m * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->out_b.p, ent_voff, m * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->out_c.p, ent_vlen, m * 4, cudaMemcpyHostToDevice, e->stream));
    }
    CU_TRY(cudaMemcpyAsync(e->lines_off.p, ev_time, n * 4, cudaMemcpyHostToDevice, e->stream));
    if (ev_time_ns)
        CU_TRY(cudaMemcpyAsync(e->lines_len.p, ev_time_ns, n * 4, cudaMemcpyHostToDevice, e->stream));
    CU_TRY(cudaMemcpyAsync(e->pos.p, ent_begin, (n + 1) * 8, cudaMemcpyHostToDevice, e->stream));
    const uint32_t* d_ns = ev_time_ns ? e->lines_len.as<uint32_t>() : nullptr;
    lck::launch_sls_sizes(e->pos.as<uint64_t>(), e->ev_len.as<uint32_t>(), e->out_c.as<uint32_t>(), d_ns, n,
                          e->lab_sizes.as<uint32_t>(), e->cnt.as<uint32_t>(), e->stream);
    lck::launch_exclusive_sum(e->lab_sizes.as<uint32_t>(), n, e->lab_off.as<uint64_t>(), &ds->total, plan.r[2],
                              &ds->tickets[2], e->stream);
    e->launches += 2;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(&hs->total, &ds->total, 8, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    *out_len = hs->total;
    if (hs->total > out_cap)
        return fail(LC_ERR_CAPACITY, "lc_sls_serialize_logs: output capacity too small");
    if (hs->total == 0)
        return LC_OK;
    if (!out)
        return fail(LC_ERR_INVALID_ARG, "lc_sls_serialize_logs: bad arguments");
    CU_TRY(e->out_d.ensure(hs->total));
    lck::launch_sls_emit(e->in.as<uint8_t>(), e->lines_off.as<uint32_t>(), d_ns, e->pos.as<uint64_t>(),
                         e->ev_off.as<uint32_t>(), e->ev_len.as<uint32_t>(), e->out_b.as<uint32_t>(),
                         e->out_c.as<uint32_t>(), n, e->lab_off.as<uint64_t>(), e->cnt.as<uint32_t>(),
                         e->out_d.as<uint8_t>(), e->stream);
    e->launches++;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(out, e->out_d.p, hs->total, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return LC_OK;
}

int lc_sls_serialize_parsed_dev(lc_engine_t* e, const uint8_t* d_base, uint64_t base_len, const uint32_t* d_ev_off,
                                const uint32_t* d_ev_len, const uint8_t* d_status, const uint32_t* d_cap_off,
                                const uint32_t* d_cap_len, uint32_t row_pitch, uint64_t n, const char* const* keys,
                                const uint32_t* key_lens, uint32_t nkeys, const char* fail_key, uint32_t fail_key_len,
                                const uint32_t* d_ev_time, const uint32_t* d_ev_time_ns, uint8_t* d_out,
                                uint64_t out_cap, uint64_t* out_len) {
    if (!e || !out_len || (n && (!d_base || !d_ev_off || !d_ev_len || !d_status || !d_ev_time)) ||
        (nkeys && (!keys || !key_lens || !d_cap_off || !d_cap_len)) || nkeys > row_pitch || nkeys > LC_MAX_GROUPS)
        return fail(LC_ERR_INVALID_ARG, "lc_sls_serialize_parsed_dev: bad arguments");
    *out_len = 0;
    if (n == 0)
        return LC_OK;
    if (base_len >= 0xFFFFFFF0ull || n >= (1ull << 30))
        return fail(LC_ERR_TOO_LARGE, "buffer must be < 4 GiB and < 2^30 events per call");
    for (uint32_t a = 0; a < nkeys; ++a)
        for (uint32_t b = a + 1; b < nkeys; ++b)
            if (key_lens[a] == key_lens[b] && !memcmp(keys[a], keys[b], key_lens[a]))
                return fail(LC_ERR_INVALID_ARG, "lc_sls_serialize_parsed_dev: keys must be distinct (a repeated key "
                                                "overwrites its earlier content, LogEvent.cpp:83-95)");
    int rc = bind(e);
    if (rc)
        return rc;
    // constant key strings: [keys..., fail key] back to back + offsets, staged through the engine's key buffer
    std::vector<uint8_t> kb;
    std::vector<uint32_t> at(nkeys + 2);
    for (uint32_t k = 0; k < nkeys; ++k) {
        at[k] = (uint32_t)kb.size();
        kb.insert(kb.end(), keys[k], keys[k] + key_lens[k]);
    }
    at[nkeys] = (uint32_t)kb.size();
    if (fail_key)
        kb.insert(kb.end(), fail_key, fail_key + fail_key_len);
    at[nkeys + 1] = (uint32_t)kb.size();
    const size_t at_bytes = at.size() * 4;
    CU_TRY(e->order.ensure(at_bytes + kb.size() + 16));
    CU_TRY(cudaMemcpyAsync(e->order.p, at.data(), at_bytes, cudaMemcpyHostToDevice, e->stream));
    if (!kb.empty())
        CU_TRY(cudaMemcpyAsync(e->order.as<uint8_t>() + at_bytes, kb.data(), kb.size(), cudaMemcpyHostToDevice,
                               e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream)); // (pageable sources: their bytes must be on the device before they die)
    lck::SlsParsedArgs a{d_base, d_ev_off, d_ev_len, d_status, d_cap_off, d_cap_len, row_pitch,
                         e->order.as<uint8_t>() + at_bytes, e->order.as<uint32_t>(), nkeys, fail_key ? 1u : 0u};
    CU_TRY(e->lab_sizes.ensure(n * 4));
    CU_TRY(e->cnt.ensure(n * 4));
    CU_TRY(e->lab_off.ensure(n * 8));
    DescPlan plan;
    rc = prep_desc(e, 0, 0, lck::scan_tiles(n), plan);
    if (rc)
        return rc;
    Small* ds = e->small.as<Small>();
    Small* hs = (Small*)e->h_small;
    lck::launch_sls_parsed_sizes(a, d_ev_time_ns, n, e->lab_sizes.as<uint32_t>(), e->cnt.as<uint32_t>(), e->stream);
    lck::launch_exclusive_sum(e->lab_sizes.as<uint32_t>(), n, e->lab_off.as<uint64_t>(), &ds->total, plan.r[2],
                              &ds->tickets[2], e->stream);
    e->launches += 2;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(&hs->total, &ds->total, 8, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    *out_len = hs->total;
    if (hs->total > out_cap)
        return fail(LC_ERR_CAPACITY, "lc_sls_serialize_parsed_dev: output capacity too small");
    if (hs->total == 0)
        return LC_OK;
    if (!d_out)
        return fail(LC_ERR_INVALID_ARG, "lc_sls_serialize_parsed_dev: bad arguments");
    lck::launch_sls_parsed_emit(a, d_ev_time, d_ev_time_ns, n, e->lab_off.as<uint64_t>(), e->cnt.as<uint32_t>(),
                                d_out, e->stream);
    e->launches++;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(e->stream));
    return LC_OK;
}

} // extern "C"
