// lc_capi.cu -- implementation of the C-ABI declared in include/lc_b200.h: engine (stream, grow-only
// HBM workspace, look-back descriptors), host<->device staging, and the per-processor pipelines.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
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

} // namespace

struct lc_regex {
    uint64_t id;
    lcb200::CompileResult res;
    std::string pattern;
};

struct lc_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    int num_sms = 148;
    int smem_per_block_optin = 0;
    int smem_per_sm = 0;
    bool force_basic_regex = false;   // env LC_B200_REGEX_KERNEL=basic   (tables in global memory)
    int regex_variant = 0; // env LC_B200_REGEX_KERNEL: 0 auto, 1 "fast" (stride-1), 2 "fast2" (stride-2), 3 "generic",
                           // 4 "tdfa" (single pass, staged input), 5 "tdfa_direct" (single pass, per-lane loads)
    uint64_t scratch_hint = 0;
    int length_order = -1; // env LC_B200_LENGTH_ORDER: 1 = always order ragged batches by length, 0 = never, unset = auto
    uint32_t max_warps = 32;   // env LC_B200_MAX_WARPS (tuning knob: resident warps per block of the regex kernels)
    // staging / workspace (grow-only)
    DevBuf in, ev_off, ev_len, out_a, out_b, out_c, out_d, out_e;
    DevBuf lines_off, lines_len, flags, state, cnt, pos, lab_sizes, lab_off, lab, order;
    DevBuf desc;   // look-back descriptors (3 regions)
    DevBuf small;  // tickets + counters: [0..3] u32 tickets, +16: u32 n_out, +32: u64 total, +64: u64 counters[2]
    void* h_small = nullptr; // pinned mirror of `small`
    std::unordered_map<uint64_t, void*> blobs; // regex id -> device blob
    // copy pipeline of the host-pointer entry points
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev_h2d, ev_comp;
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
    if (er != cudaSuccess)
        return er;
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
    uint64_t pad1[1];
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
        cudaEvent_t a, b;
        CU_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        e->ev_h2d.push_back(a);
        e->ev_comp.push_back(b);
    }
    return LC_OK;
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
    CU_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CU_TRY(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, device));
    CU_TRY(cudaDeviceGetAttribute(&e->smem_per_block_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CU_TRY(cudaDeviceGetAttribute(&e->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
    {
        const char* k = getenv("LC_B200_REGEX_KERNEL");
        e->force_basic_regex = k && !strcmp(k, "basic");
        const char* mw = getenv("LC_B200_MAX_WARPS");
        if (mw && atoi(mw) >= 4 && atoi(mw) <= 32)
            e->max_warps = (uint32_t)atoi(mw);
        const char* lo = getenv("LC_B200_LENGTH_ORDER");
        e->length_order = !lo ? -1 : (!strcmp(lo, "1") ? 1 : 0);
        e->regex_variant = !k ? 0 : (!strcmp(k, "fast") ? 1 : (!strcmp(k, "fast2") ? 2 : (!strcmp(k, "generic") ? 3 : (!strcmp(k, "tdfa") ? 4 : (!strcmp(k, "tdfa_direct") ? 5 : 0)))));
    }
    CU_TRY(e->small.ensure(sizeof(Small)));
    CU_TRY(cudaMallocHost(&e->h_small, sizeof(Small)));
    *out = e;
    return LC_OK;
}

void lc_engine_destroy(lc_engine_t* e) {
    if (!e)
        return;
    cudaSetDevice(e->device);
    if (e->stream)
        cudaStreamSynchronize(e->stream);
    DevBuf* bufs[] = {&e->in, &e->ev_off, &e->ev_len, &e->out_a, &e->out_b, &e->out_c, &e->out_d, &e->out_e,
                      &e->lines_off, &e->lines_len, &e->flags, &e->state, &e->cnt, &e->pos, &e->lab_sizes,
                      &e->lab_off, &e->lab, &e->order, &e->desc, &e->small};
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
    if (e->s_h2d)
        cudaStreamDestroy(e->s_h2d);
    if (e->s_d2h)
        cudaStreamDestroy(e->s_d2h);
    if (e->stream)
        cudaStreamDestroy(e->stream);
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

void lc_regex_free(lc_regex_t* r) { delete r; }
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
    lck::launch_split(d_buf, (uint32_t)len, split_char, d_out_off, d_out_len, cap32, plan.r[0], &ds->tickets[0],
                      &ds->n_out, e->stream);
    e->launches++;
    CU_TRY(cudaGetLastError());
    Small* hs = (Small*)e->h_small;
    CU_TRY(cudaMemcpyAsync(&hs->n_out, &ds->n_out, 4, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
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
static int regex_parse_dev_impl(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                                const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint32_t nkeys,
                                uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len, bool bool_only);

int lc_regex_parse_dev(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                       const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint32_t nkeys,
                       uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len) {
    return regex_parse_dev_impl(e, re, d_base, base_len, d_ev_off, d_ev_len, n, nkeys, d_status, d_cap_off, d_cap_len,
                                false);
}

int lc_regex_match_dev(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                       const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint8_t* d_out_match) {
    return regex_parse_dev_impl(e, re, d_base, base_len, d_ev_off, d_ev_len, n, 0, d_out_match, nullptr, nullptr, true);
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

static int regex_parse_dev_impl(lc_engine_t* e, const lc_regex_t* re, const uint8_t* d_base, uint64_t base_len,
                                const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint32_t nkeys,
                                uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len, bool bool_only) {
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
    // ---- single-pass tagged DFA: the preferred kernel whenever the pattern's TDFA fits shared memory
    if (!force_basic && (e->regex_variant == 0 || e->regex_variant == 4 || e->regex_variant == 5) &&
        !re->res.tdfa_blob.empty()) {
        const LcTdfaHeader* th = reinterpret_cast<const LcTdfaHeader*>(re->res.tdfa_blob.data());
        const uint32_t tb = (uint32_t)re->res.tdfa_blob.size();
        const bool staged = e->regex_variant != 5;
        auto smem_need = [&](uint32_t warps) {
            return staged ? lck::tdfa_staged_smem_bytes(tb, th->nregs, warps * 32)
                          : lck::tdfa_smem_bytes(tb, th->nregs, warps * 32);
        };
        uint32_t warps = e->max_warps;
        while (warps > 4 && smem_need(warps) > smem_max)
            warps -= 2;
        bool usable = smem_need(warps) <= smem_max;
        if (usable && !staged && base_len >= 65535) { // capture registers are 16-bit: every event must be < 65535 bytes
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
            // bucket.  Only looked at when the mean length makes the pre-pass (one small kernel + a sync) negligible;
            // LC_B200_LENGTH_ORDER=1 forces it, =0 disables it.
            const uint32_t* d_order = nullptr;
            if (staged && n >= 4096 && e->length_order != 0 && (e->length_order == 1 || base_len / n >= 1024)) {
                CU_TRY(cudaMemsetAsync(ds->counters, 0, sizeof ds->counters, e->stream));
                lck::launch_len_stats(d_ev_len, n, ds->counters, e->stream);
                e->launches++;
                CU_TRY(cudaMemcpyAsync(hs->counters, ds->counters, sizeof ds->counters, cudaMemcpyDeviceToHost,
                                       e->stream));
                CU_TRY(cudaStreamSynchronize(e->stream));
                const uint64_t mx = hs->counters[0], avg = hs->counters[1] / n + 1;
                if (mx > avg + avg / 2 + 64) {
                    CU_TRY(e->order.ensure(n * 4 + 256));
                    uint32_t* hist = e->order.as<uint32_t>() + n;
                    lck::launch_length_order(d_ev_len, n, hist, e->order.as<uint32_t>(), e->stream);
                    e->launches += 3;
                    d_order = e->order.as<uint32_t>();
                }
            }
            CU_TRY(cudaMemsetAsync(&ds->overflow, 0, sizeof(Small) - offsetof(Small, overflow), e->stream));
            int er;
            if (staged)
                er = lck::launch_regex_tdfa_staged(d_tblob, tb, th->has_slow != 0, th->nregs, d_base, d_ev_off,
                                                   d_ev_len, n, nkeys, d_status, bool_only ? nullptr : d_cap_off,
                                                   bool_only ? nullptr : d_cap_len, threads, grid, &ds->next_batch,
                                                   &ds->overflow, d_order, e->stream);
            else
                er = lck::launch_regex_tdfa(d_tblob, tb, th->has_slow != 0, th->nregs, d_base, d_ev_off, d_ev_len, n,
                                            nkeys, d_status, bool_only ? nullptr : d_cap_off,
                                            bool_only ? nullptr : d_cap_len, threads, grid, &ds->next_batch, nullptr,
                                            e->stream);
            e->launches++;
            if (er)
                return fail(LC_ERR_CUDA, std::string("regex kernel launch: ") + cudaGetErrorString((cudaError_t)er));
            if (!staged || base_len < 65535)
                return LC_OK;
            // the kernel itself reports events too long for its 16-bit capture registers (no length pre-pass)
            CU_TRY(cudaMemcpyAsync(&hs->overflow, &ds->overflow, 4, cudaMemcpyDeviceToHost, e->stream));
            CU_TRY(cudaStreamSynchronize(e->stream));
            if (!hs->overflow)
                return LC_OK;
            // fall through: redo the whole call on a kernel with 32-bit slots
        }
    }
    if (!force_basic) {
        // ---- pick the kernel variant: stride-2 layout > stride-1 fast layout > generic shared-memory interpreter
        uint64_t mx = 0, avg = base_len / n + 1;
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
            uint64_t full = base_len / per + 2 * n + 1024;
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
                CU_TRY(e->order.ensure(n * 4 + 256));
                uint32_t* hist = e->order.as<uint32_t>() + n;
                lck::launch_length_order(d_ev_len, n, hist, e->order.as<uint32_t>(), e->stream);
                e->launches += 3;
                d_order = e->order.as<uint32_t>();
            }
            for (int attempt = 0; attempt < 8; ++attempt) {
                if (h->mode == LC_MODE_TWOPASS)
                    CU_TRY(e->lab.ensure(scratch_words * 4));
                CU_TRY(cudaMemsetAsync(&ds->overflow, 0, sizeof(Small) - offsetof(Small, overflow), e->stream));
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
    rc = bind(e);
    if (rc)
        return rc;
    const uint32_t G = re->res.ngroups;
    CU_TRY(e->in.ensure(base_len + 16));
    CU_TRY(e->ev_off.ensure(n * 4));
    CU_TRY(e->ev_len.ensure(n * 4));
    CU_TRY(e->out_a.ensure(n));
    CU_TRY(e->out_b.ensure(n * G * 4 + 4));
    CU_TRY(e->out_c.ensure(n * G * 4 + 4));
    // Large batches are pipelined: the arena is cut into chunks of whole events; all H2D copies are queued up front
    // on a copy stream, each chunk's kernels run on the engine stream as soon as its bytes have landed, and its
    // result tables travel back on a third stream while later chunks are still being uploaded (PCIe is full duplex).
    const uint64_t kChunkBytes = 48ull << 20;
    uint64_t nchunks = (base_len + kChunkBytes - 1) / kChunkBytes;
    if (nchunks > 64)
        nchunks = 64;
    if (nchunks < 2 || n < nchunks * 1024) {
        CU_TRY(cudaMemcpyAsync(e->in.p, base, base_len, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->ev_off.p, ev_off, n * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->ev_len.p, ev_len, n * 4, cudaMemcpyHostToDevice, e->stream));
        rc = lc_regex_parse_dev(e, re, e->in.as<uint8_t>(), base_len, e->ev_off.as<uint32_t>(),
                                e->ev_len.as<uint32_t>(), n, nkeys, e->out_a.as<uint8_t>(), e->out_b.as<uint32_t>(),
                                e->out_c.as<uint32_t>());
        if (rc)
            return rc;
        CU_TRY(cudaMemcpyAsync(status, e->out_a.p, n, cudaMemcpyDeviceToHost, e->stream));
        if (G) {
            CU_TRY(cudaMemcpyAsync(cap_off, e->out_b.p, n * G * 4, cudaMemcpyDeviceToHost, e->stream));
            CU_TRY(cudaMemcpyAsync(cap_len, e->out_c.p, n * G * 4, cudaMemcpyDeviceToHost, e->stream));
        }
        CU_TRY(cudaStreamSynchronize(e->stream));
        return LC_OK;
    }
    rc = ensure_copy_streams(e, (int)nchunks);
    if (rc)
        return rc;
    // event ranges per chunk (equal event counts) and the byte span each one touches
    std::vector<uint64_t> first(nchunks + 1), lo(nchunks), hi(nchunks);
    for (uint64_t c = 0; c <= nchunks; ++c)
        first[c] = n * c / nchunks;
    for (uint64_t c = 0; c < nchunks; ++c) {
        uint64_t l = ~0ull, h = 0;
        for (uint64_t i = first[c]; i < first[c + 1]; ++i) {
            uint64_t o = ev_off[i], en = o + ev_len[i];
            l = o < l ? o : l;
            h = en > h ? en : h;
        }
        if (h > base_len)
            return fail(LC_ERR_INVALID_ARG, "lc_regex_parse: event beyond base_len");
        lo[c] = l == ~0ull ? 0 : (l & ~15ull);
        hi[c] = h;
    }
    uint8_t* d_in = e->in.as<uint8_t>();
    uint32_t* d_off = e->ev_off.as<uint32_t>();
    uint32_t* d_len = e->ev_len.as<uint32_t>();
    for (uint64_t c = 0; c < nchunks; ++c) {
        uint64_t i0 = first[c], cnt = first[c + 1] - first[c];
        if (hi[c] > lo[c])
            CU_TRY(cudaMemcpyAsync(d_in + lo[c], base + lo[c], hi[c] - lo[c], cudaMemcpyHostToDevice, e->s_h2d));
        CU_TRY(cudaMemcpyAsync(d_off + i0, ev_off + i0, cnt * 4, cudaMemcpyHostToDevice, e->s_h2d));
        CU_TRY(cudaMemcpyAsync(d_len + i0, ev_len + i0, cnt * 4, cudaMemcpyHostToDevice, e->s_h2d));
        CU_TRY(cudaEventRecord(e->ev_h2d[c], e->s_h2d));
    }
    for (uint64_t c = 0; c < nchunks; ++c) {
        uint64_t i0 = first[c], cnt = first[c + 1] - first[c];
        CU_TRY(cudaStreamWaitEvent(e->stream, e->ev_h2d[c], 0));
        rc = lc_regex_parse_dev(e, re, d_in, base_len, d_off + i0, d_len + i0, cnt, nkeys, e->out_a.as<uint8_t>() + i0,
                                e->out_b.as<uint32_t>() + i0 * G, e->out_c.as<uint32_t>() + i0 * G);
        if (rc)
            return rc;
        CU_TRY(cudaEventRecord(e->ev_comp[c], e->stream));
        CU_TRY(cudaStreamWaitEvent(e->s_d2h, e->ev_comp[c], 0));
        CU_TRY(cudaMemcpyAsync(status + i0, e->out_a.as<uint8_t>() + i0, cnt, cudaMemcpyDeviceToHost, e->s_d2h));
        if (G) {
            CU_TRY(cudaMemcpyAsync(cap_off + i0 * G, e->out_b.as<uint32_t>() + i0 * G, cnt * G * 4,
                                   cudaMemcpyDeviceToHost, e->s_d2h));
            CU_TRY(cudaMemcpyAsync(cap_len + i0 * G, e->out_c.as<uint32_t>() + i0 * G, cnt * G * 4,
                                   cudaMemcpyDeviceToHost, e->s_d2h));
        }
    }
    CU_TRY(cudaStreamSynchronize(e->s_d2h));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return LC_OK;
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
    CU_TRY(engine_blob(e, start, &cfg.blob_start));
    CU_TRY(engine_blob(e, cont, &cfg.blob_cont));
    CU_TRY(engine_blob(e, end, &cfg.blob_end));
    cfg.discard = discard_unmatched;

    Small* ds = e->small.as<Small>();
    Small* hs = (Small*)e->h_small;
    // 1. line table (grow the workspace until it fits; typical logs fit the first estimate)
    uint64_t lcap = len / 24 + 4096;
    uint64_t n = 0;
    uint32_t shift = (uint32_t)((uintptr_t)d_buf & 15u);
    for (;;) {
        if (lcap > len)
            lcap = len;
        CU_TRY(e->lines_off.ensure((lcap + 1) * 4));
        CU_TRY(e->lines_len.ensure((lcap + 1) * 4));
        DescPlan plan;
        rc = prep_desc(e, lck::split_tiles(len, shift), 0, 0, plan);
        if (rc)
            return rc;
        lck::launch_split(d_buf, (uint32_t)len, '\n', e->lines_off.as<uint32_t>(), e->lines_len.as<uint32_t>(),
                          (uint32_t)(lcap > 0x3FFFFFFFull ? 0x3FFFFFFFull : lcap), plan.r[0], &ds->tickets[0],
                          &ds->n_out, e->stream);
        e->launches++;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(&hs->n_out, &ds->n_out, 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
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

// ------------------------------------------------------------------------------------------------ delimiter
int lc_delim_parse_dev(lc_engine_t* e, const uint8_t* d_base, uint64_t base_len, const uint32_t* d_ev_off,
                       const uint32_t* d_ev_len, uint64_t n, const uint8_t* sep, uint32_t sep_len, uint8_t quote,
                       uint32_t nkeys, int extend, int allow_short, uint32_t max_fields, uint8_t* d_status,
                       uint32_t* d_nfields, uint32_t* d_f_off, uint32_t* d_f_len, uint32_t* d_f_dq) {
    if (!e || !sep || sep_len < 1 || sep_len > 4 || max_fields == 0)
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
    lck::launch_delim(cfg, d_base, d_ev_off, d_ev_len, n, d_status, d_nfields, d_f_off, d_f_len, d_f_dq, e->stream);
    e->launches++;
    CU_TRY(cudaGetLastError());
    return LC_OK;
}

int lc_delim_parse(lc_engine_t* e, const uint8_t* base, uint64_t base_len, const uint32_t* ev_off,
                   const uint32_t* ev_len, uint64_t n, const uint8_t* sep, uint32_t sep_len, uint8_t quote,
                   uint32_t nkeys, int extend, int allow_short, uint32_t max_fields, uint8_t* status,
                   uint32_t* nfields, uint32_t* f_off, uint32_t* f_len, uint32_t* f_dq) {
    if (!e || (n && (!ev_off || !ev_len || !status || !nfields || !f_off || !f_len || !f_dq)) || (base_len && !base))
        return fail(LC_ERR_INVALID_ARG, "lc_delim_parse: bad arguments");
    if (n == 0)
        return LC_OK;
    int rc = bind(e);
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
    int rc = bind(e);
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

} // extern "C"
