// lc_kernels.cuh -- launch interface between the C-ABI layer (lc_capi.cu) and the sm_100a kernels
// (lc_kernels.cu).  Device pointers + stream in, nothing else.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lck {

// smallest split tile (16 KiB): look-back descriptors are sized for it; the kernel normally runs 128 KiB tiles
constexpr uint32_t kSplitTileBytes = 256 * 4 * 16;

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr uint32_t kScanTile = kScanThreads * kScanItems;

inline uint32_t split_tiles(uint64_t len, uint32_t shift) {
    return (uint32_t)((len + shift + kSplitTileBytes - 1) / kSplitTileBytes);
}
inline uint32_t scan_tiles(uint64_t n) { return (uint32_t)((n + kScanTile - 1) / kScanTile); }

// a1: newline split.  desc: >= split_tiles() u64 (zeroed), ticket: u32 (zeroed), n_out: u32 device counter.
// d_total: u64 device counter (zeroed) that receives the un-truncated number of split chars (the look-back payload
// keeps 30 bits of count: the caller reports LC_ERR_TOO_LARGE beyond that).
// d_scratch: split_scratch_bytes(len, probe) bytes (not initialised) for the three-pass formulation (masks, per-tile
// counts and prefixes); nullptr selects the single-pass look-back kernel.  Returns the number of kernels launched.
uint64_t split_scratch_bytes(uint64_t len, bool probe);
int launch_split(const uint8_t* d_buf, uint32_t len, uint8_t split_char, uint32_t* d_off, uint32_t* d_len,
                 uint32_t cap, uint64_t* d_desc, uint32_t* d_ticket, uint32_t* d_n_out, unsigned long long* d_total,
                 uint64_t* d_scratch, cudaStream_t st);

// exclusive sum of u32 -> u64 (out has n entries; *d_total receives the grand total)
void launch_exclusive_sum(const uint32_t* d_in, uint64_t n, uint64_t* d_out, uint64_t* d_total, uint64_t* d_desc,
                          uint32_t* d_ticket, cudaStream_t st);

// a3: regex_match + capture groups, one thread per event (baseline kernel, tables read from global memory)
// d_lab_off: per-event byte offset into d_lab (two-pass label scratch, u16 labels); unused for forward-only.
void launch_regex_parse_basic(const void* d_blob, uint32_t mode, uint32_t ngroups, const uint8_t* d_base,
                              const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint32_t nkeys,
                              uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len, const uint64_t* d_lab_off,
                              uint16_t* d_lab, cudaStream_t st);
// per-event scratch need of the two-pass matcher: len + 1 labels, rounded up to 8 labels (16 B)
void launch_label_sizes(const uint32_t* d_ev_len, uint64_t n, uint32_t* d_sizes, cudaStream_t st);

// d_out[0] = max(ev_len), d_out[1] = sum(ev_len); d_out must be zeroed by the caller
void launch_len_stats(const uint32_t* d_ev_len, uint64_t n, unsigned long long* d_out, cudaStream_t st);

// order[] = event indices sorted by descending length bucket (d_hist64: 64-word scratch + 1 flag word); 3 launches.
// d_hist64[64] = 1 when the lengths span more than two adjacent buckets (a ragged batch: use the order), else 0.
void launch_length_order(const uint32_t* d_ev_len, uint64_t n, uint32_t* d_hist64, uint32_t* d_order,
                         cudaStream_t st);

// a3 fast path: persistent kernel, automaton staged in shared memory.  Labels of events needing
// <= lab_words 32-bit words stay in shared memory, longer events bump-allocate from d_scratch.
// d_bump, d_overflow and d_next_batch must be zeroed by the caller.  Dynamic shared memory =
// blob_bytes + (threads / 32) * lab_words * 128.  Returns a cudaError_t value (0 on success).
int launch_regex_parse_fast(const void* d_blob, uint32_t blob_bytes, uint32_t rev_label_bytes, uint32_t ngroups,
                            const uint8_t* d_base, const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n,
                            uint32_t nkeys, uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len,
                            uint32_t lab_words, uint32_t threads, uint32_t grid, uint32_t* d_scratch,
                            uint64_t scratch_words, unsigned long long* d_bump, uint32_t* d_overflow,
                            unsigned long long* d_next_batch, const uint32_t* d_order, cudaStream_t st);

// a3 fastest path: two-pass automaton in the host-built fast layout (LcFastHeader).  Shared memory per block =
// blob + labels (threads/32 * lab_words * 128 B) + capture slots (threads * slot_pitch * 4 B).
inline uint32_t fast_slot_pitch(uint32_t ngroups) { return (2 * ngroups + 1) | 1u; } // odd word pitch
inline size_t fast_smem_bytes(uint32_t blob_bytes, uint32_t ngroups, uint32_t lab_words, uint32_t threads) {
    return (size_t)blob_bytes + (size_t)(threads / 32) * lab_words * 128 + (size_t)threads * fast_slot_pitch(ngroups) * 4;
}
int launch_regex_twopass_fast(const void* d_fast_blob, uint32_t blob_bytes, bool multi, uint32_t ngroups,
                              const uint8_t* d_base, const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n,
                              uint32_t nkeys, uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len,
                              uint32_t lab_words, uint32_t threads, uint32_t grid, uint32_t* d_scratch,
                              uint64_t scratch_words, unsigned long long* d_bump, uint32_t* d_overflow,
                              unsigned long long* d_next_batch, const uint32_t* d_order, cudaStream_t st);

// a3 stride-2 path (LcFast2Header): labels take (len + 15) / 8 + 1 words, capture slots are u16.
// Only valid when every event is shorter than 65535 bytes.
inline uint32_t fast2_slot_pitch(uint32_t ngroups) { return ((ngroups + 1) | 1u) * 2; } // halfwords; odd WORD pitch
inline size_t fast2_smem_bytes(uint32_t blob_bytes, uint32_t ngroups, uint32_t lab_words, uint32_t threads) {
    return (size_t)blob_bytes + (size_t)(threads / 32) * lab_words * 128 + (size_t)threads * fast2_slot_pitch(ngroups) * 2;
}
int launch_regex_fast2(const void* d_blob, uint32_t blob_bytes, bool multi, bool compact, uint32_t ngroups,
                       const uint8_t* d_base,
                       const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint32_t nkeys,
                       uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len, uint32_t lab_words,
                       uint32_t threads, uint32_t grid, uint32_t* d_scratch, uint64_t scratch_words,
                       unsigned long long* d_bump, uint32_t* d_overflow, unsigned long long* d_next_batch,
                       const uint32_t* d_order, cudaStream_t st);

// a3 single-pass tagged-DFA path (LcTdfaHeader): no labels; per thread only nregs u16 registers in shared memory.
// Only valid when every event is shorter than 65535 bytes (the staged kernel raises *d_overflow otherwise and the
// host repeats the call on a kernel with 32-bit slots; the direct kernel relies on a host-side length check).
inline uint32_t tdfa_reg_pitch(uint32_t nregs) { return (((nregs + 1) / 2) | 1u) * 2; } // halfwords; odd WORD pitch
inline size_t tdfa_smem_bytes(uint32_t blob_bytes, uint32_t nregs, uint32_t threads) {
    return (size_t)blob_bytes + (size_t)threads * tdfa_reg_pitch(nregs) * 2;
}
int launch_regex_tdfa(const void* d_blob, uint32_t blob_bytes, bool slow, uint32_t nregs, const uint8_t* d_base,
                      const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n, uint32_t nkeys,
                      uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len, uint32_t threads, uint32_t grid,
                      unsigned long long* d_next_batch, const uint32_t* d_order, cudaStream_t st);

// staged variant: lines are fetched cooperatively (cp.async, 4 full 128-byte lines per instruction) into a per-warp
// 4 KB tile; carve-out = 512 B (aligned class table) + blob + register files + 256 B line info and 4 KB tile per warp
inline size_t tdfa_staged_smem_bytes(uint32_t blob_bytes, uint32_t nregs, uint32_t threads) {
    return 512 + (size_t)blob_bytes + 16 + (size_t)threads * tdfa_reg_pitch(nregs) * 2 +
           (size_t)(threads / 32) * (256 + 4096);
}
int launch_regex_tdfa_staged(const void* d_blob, uint32_t blob_bytes, bool slow, uint32_t nregs, const uint8_t* d_base,
                             const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint32_t ev_stride /* elements */,
                             uint64_t n, uint32_t nkeys, uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len,
                             uint32_t threads, uint32_t grid, unsigned long long* d_next_batch, uint32_t* d_overflow,
                             const uint32_t* d_order /* or nullptr */, const uint32_t* d_order_flag /* or nullptr */,
                             cudaStream_t st);

// producer / consumer variant (A/B): the first 4 warps of the block only fill the other warps' tiles (LDG.128 + STS.128);
// threads counts ALL warps, consumers = threads / 32 - 4
inline size_t tdfa_pc_smem_bytes(uint32_t blob_bytes, uint32_t nregs, uint32_t threads) {
    const size_t nc = threads / 32 - 4;
    return 512 + (size_t)blob_bytes + 16 + nc * 32 * tdfa_reg_pitch(nregs) * 2 + nc * (256 + 4096 + 8 + 4) + 16 + 64;
}
int launch_regex_tdfa_pc(const void* d_blob, uint32_t blob_bytes, bool slow, uint32_t nregs, const uint8_t* d_base,
                         const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint32_t ev_stride, uint64_t n,
                         uint32_t nkeys, uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len, uint32_t threads,
                         uint32_t grid, unsigned long long* d_next_batch, uint32_t* d_overflow, cudaStream_t st);

// several patterns in one grid: all tagged-DFA blobs co-resident in shared memory, tried per line in array order
constexpr uint32_t LC_MULTI_MAX = 8;
struct TdfaMultiArgs {
    const void* blob[LC_MULTI_MAX]; // device tdfa blobs
    uint32_t blob_bytes[LC_MULTI_MAX];
    uint32_t nkeys[LC_MULTI_MAX];
    uint32_t npat;
};
inline size_t tdfa_multi_table_bytes(const TdfaMultiArgs& a) {
    size_t t = 256;
    for (uint32_t p = 0; p < a.npat; ++p)
        t += 256 + ((a.blob_bytes[p] + 255u) & ~255u);
    return t;
}
inline size_t tdfa_multi_smem_bytes(const TdfaMultiArgs& a, uint32_t max_nregs, uint32_t threads) {
    return tdfa_multi_table_bytes(a) + 16 + (size_t)threads * tdfa_reg_pitch(max_nregs) * 2 +
           (size_t)(threads / 32) * (256 + 4096);
}
int launch_regex_tdfa_multi(const TdfaMultiArgs& a, uint32_t p_base, bool resume, bool slow, uint32_t max_nregs,
                            const uint8_t* d_base, const uint32_t* d_ev_off, const uint32_t* d_ev_len, uint64_t n,
                            const uint8_t* d_sel, uint8_t* d_which, uint8_t* d_status, uint32_t* d_cap_off,
                            uint32_t* d_cap_len, uint32_t gpitch, uint32_t threads, uint32_t grid,
                            unsigned long long* d_next_batch, uint32_t* d_overflow, const uint32_t* d_order,
                            const uint32_t* d_order_flag, cudaStream_t st);
// events >= 65535 bytes (skipped by the staged kernels, which raise *d_overflow): 32-bit registers, no-op otherwise
void launch_regex_tdfa_long(const TdfaMultiArgs& a, const uint8_t* d_base, const uint32_t* d_ev_off,
                            const uint32_t* d_ev_len, uint32_t ev_stride, uint64_t n, const uint8_t* d_sel,
                            uint8_t* d_which, uint8_t* d_status, uint32_t* d_cap_off, uint32_t* d_cap_len,
                            uint32_t gpitch, const uint32_t* d_overflow, bool bool_only, cudaStream_t st);

// parse status -> boolean (1 = the whole value matched)
void launch_status_to_bool(uint8_t* d_status, uint64_t n, cudaStream_t st);

// anchored prefix probe, one bool per event
void launch_prefix_match(const void* d_blob, const uint8_t* d_base, const uint32_t* d_ev_off,
                         const uint32_t* d_ev_len, uint64_t n, uint8_t* d_out, cudaStream_t st);

// a2: multiline.  Lines (d_off,d_len,n) come from launch_split over the same buffer.
struct MlConfig {
    const void* blob_start; // device blobs or nullptr
    const void* blob_cont;
    const void* blob_end;
    int discard;
    // split + probe pass only (host-built from the prefix DFAs): first[p] bit b = a line starting with byte b can
    // match pattern p; empty_flags = flag bits of an empty line
    uint32_t first[3][8];
    uint32_t empty_flags;
};
// flags[i] bit0/1/2 = start/continue/end pattern matches a prefix of line i
void launch_ml_probe(const MlConfig& cfg, const uint8_t* d_buf, const uint32_t* d_off, const uint32_t* d_len,
                     uint64_t n, uint8_t* d_flags, cudaStream_t st);
// state scan over n lines + 1 virtual EOF element: d_state[i] = (s_in << 31) | lb_in, d_cnt[i] = output events
void launch_ml_state(const MlConfig& cfg, const uint8_t* d_flags, const uint32_t* d_len, uint64_t n, uint32_t* d_state,
                     uint32_t* d_cnt, uint64_t* d_desc, uint32_t* d_ticket, cudaStream_t st);
// emission: d_pos = exclusive sum of d_cnt; d_counters[0..1] += matched_events, unmatch_lines
void launch_ml_emit(const MlConfig& cfg, const uint8_t* d_flags, const uint32_t* d_off, const uint32_t* d_len,
                    uint64_t n, uint32_t total_len, const uint32_t* d_state, const uint64_t* d_pos, uint32_t* d_out_off,
                    uint32_t* d_out_len, uint8_t* d_out_flags, uint64_t cap, unsigned long long* d_counters,
                    cudaStream_t st);

// fused multiline path: split + per-line probes in one pass (flags[line]), then state scan + counts + slots + emission
// in one kernel that reads the line count from d_n_lines (no host round trip in between).  d_desc_state / d_desc_sum:
// ml_fused_tiles(line_cap) + 1 zeroed u64 each.
int launch_split_probe(const MlConfig& cfg, const uint8_t* d_buf, uint32_t len, uint32_t* d_off, uint32_t* d_len,
                       uint8_t* d_flags, uint32_t cap, uint64_t* d_desc, uint32_t* d_ticket, uint32_t* d_n_out,
                       unsigned long long* d_total, uint64_t* d_scratch, cudaStream_t st);
uint32_t ml_fused_tiles(uint64_t line_cap);
// the same result without look-backs (five launches; d_scratch: ml_pass_scratch_bytes(line_cap), not initialised)
uint32_t ml_pass_tiles(uint64_t line_cap);
uint64_t ml_pass_scratch_bytes(uint64_t line_cap);
int launch_ml_passes(const MlConfig& cfg, const uint8_t* d_flags, const uint32_t* d_off, const uint32_t* d_len,
                     const uint32_t* d_n_lines, uint32_t line_cap, uint32_t total_len, uint32_t* d_out_off,
                     uint32_t* d_out_len, uint8_t* d_out_flags, uint64_t cap, uint64_t* d_scratch,
                     unsigned long long* d_counters, uint64_t* d_total, cudaStream_t st);
void launch_ml_fused(const MlConfig& cfg, const uint8_t* d_flags, const uint32_t* d_off, const uint32_t* d_len,
                     const uint32_t* d_n_lines, uint32_t line_cap, uint32_t total_len, uint32_t* d_out_off,
                     uint32_t* d_out_len, uint8_t* d_out_flags, uint64_t cap, uint64_t* d_desc_state,
                     uint64_t* d_desc_sum, uint32_t* d_ticket, unsigned long long* d_counters, uint64_t* d_total,
                     cudaStream_t st);

// f3: last complete record of a chunk from the split pass's line table + flags; d_out[0] = keep bytes, [1] = rollback
void launch_last_record(const uint8_t* d_flags, const uint32_t* d_off, const uint32_t* d_len, const uint32_t* d_n_lines,
                        uint32_t line_cap, uint32_t size, bool has_start, bool has_end, unsigned long long* d_out,
                        cudaStream_t st);

// a4: delimiter
struct DelimConfig {
    uint8_t sep[4];
    uint32_t sep_len;
    uint8_t quote;
    uint32_t nkeys;
    int extend;
    int allow_short;
    uint32_t max_fields;
    // optional column tap: the (off, len) of column tap_col of every line also leave as two DENSE tables (the event
    // table of a processor chained on that column); tap_col >= max_fields or null pointers = no tap
    uint32_t tap_col;
    uint32_t* tap_off;
    uint32_t* tap_len;
};
// d_next_batch: zeroed u64 batch counter of the persistent tiled kernel (quote-FSM mode); nullptr = thread-per-line
void launch_delim(const DelimConfig& cfg, const uint8_t* d_base, const uint32_t* d_ev_off, const uint32_t* d_ev_len,
                  uint64_t n, uint8_t* d_status, uint32_t* d_nfields, uint32_t* d_f_off, uint32_t* d_f_len,
                  uint32_t* d_f_dq, unsigned long long* d_next_batch, cudaStream_t st);

// next row (rank 4): SLS wire format of LOG events.  ev_ns may be null; 0xFFFFFFFF = no nanosecond part.
// rec_size[i] = bytes of event i's Log record (0 = empty event, skipped); body_size[i] = bytes inside its Logs field.
void launch_sls_sizes(const uint64_t* d_ent_begin, const uint32_t* d_klen, const uint32_t* d_vlen,
                      const uint32_t* d_ev_ns, uint64_t n, uint32_t* d_rec_size, uint32_t* d_body_size,
                      cudaStream_t st);
void launch_sls_emit(const uint8_t* d_base, const uint32_t* d_ev_time, const uint32_t* d_ev_ns,
                     const uint64_t* d_ent_begin, const uint32_t* d_koff, const uint32_t* d_klen,
                     const uint32_t* d_voff, const uint32_t* d_vlen, uint64_t n, const uint64_t* d_rec_off,
                     const uint32_t* d_body_size, uint8_t* d_out, cudaStream_t st);

// f4, device-fed: Log records from the regex stage's result tables + constant key strings (all device pointers)
struct SlsParsedArgs {
    const uint8_t* base;
    const uint32_t* ev_off;
    const uint32_t* ev_len;
    const uint8_t* status;
    const uint32_t* cap_off;
    const uint32_t* cap_len;
    uint32_t pitch;
    const uint8_t* keys;
    const uint32_t* key_at; // [nkeys + 2]
    uint32_t nkeys;
    uint32_t has_fail_key;
};
void launch_sls_parsed_sizes(const SlsParsedArgs& a, const uint32_t* d_ev_ns, uint64_t n, uint32_t* d_rec_size,
                             uint32_t* d_body_size, cudaStream_t st);
void launch_sls_parsed_emit(const SlsParsedArgs& a, const uint32_t* d_ev_time, const uint32_t* d_ev_ns, uint64_t n,
                            const uint64_t* d_rec_off, const uint32_t* d_body_size, uint8_t* d_out, cudaStream_t st);

} // namespace lck
