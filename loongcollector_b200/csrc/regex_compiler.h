// regex_compiler.h -- host-side compiler: boost-Perl pattern -> prioritised NFA -> automaton blob.
//
// Replaces, for the GPU path, the `boost::regex(pattern)` construction done at
//   core/plugin/processor/ProcessorParseRegexNative.cpp:64-67  (full match with capture groups)
//   core/plugin/processor/inner/ProcessorSplitMultilineLogStringNative.cpp:70-80 (anchored prefix probes)
// Semantics reproduced: Boost.Regex 1.68 Perl syntax defaults on `char` / C locale -- '.' matches
// '\n', '^' '$' are line anchors, leftmost-first (backtracking priority) disambiguation, greedy and
// lazy quantifiers, captures numbered by '(' order.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "lc_tables.h"

namespace lcb200 {

struct CompileResult {
    bool valid = false;       // pattern parsed (boost would construct it, as far as we implement the syntax)
    bool supported = false;   // automaton tables were built (no backrefs / look-around / nullable loops ...)
    std::string error;        // why not valid / not supported
    uint32_t ngroups = 0;
    std::vector<uint8_t> blob; // LcRegexHeader + arrays (lc_tables.h); empty unless supported
    std::vector<uint8_t> fast_blob; // LcFastHeader + kernel-ready tables; empty when the fast layout does not apply
    std::vector<uint8_t> fast2_blob; // LcFast2Header + stride-2 tables; empty when that layout does not apply
    std::vector<uint8_t> tdfa_blob;  // LcTdfaHeader + single-pass tagged-DFA tables; empty when the TDFA is too large
    // diagnostics
    uint32_t n_insts = 0, n_walkers = 0, n_rev = 0, n_prefix = 0;
};

// max_table_bytes bounds the total blob size (tables beyond it => unsupported, never truncated).
CompileResult compile_regex(const char* pattern, size_t len, size_t max_table_bytes = (8u << 20));

} // namespace lcb200
