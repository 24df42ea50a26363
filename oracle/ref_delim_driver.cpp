// ref_delim_driver.cpp -- C entry point over the REFERENCE's own delimiter FSM translation unit
// (core/parser/DelimiterModeFsmParser.cpp, compiled in place from /root/reference by oracle/build_ref.sh).
// TEST INFRASTRUCTURE ONLY: used to validate oracle/lc_oracle.c's restatement of the FSM.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "parser/DelimiterModeFsmParser.h"

namespace logtail {
// The zero-copy overload in the same TU references this symbol; the string overload used here never calls it.
std::shared_ptr<SourceBuffer>& PipelineEvent::GetSourceBuffer() {
    abort();
}
} // namespace logtail

extern "C" {
// Runs DelimiterModeFsmParser::ParseDelimiterLine(const char*, int, int, std::vector<std::string>&)
// (DelimiterModeFsmParser.cpp:225-258).  Returns the column count (-1 on FSM error); the columns are written
// back-to-back into out (each preceded by a u32 length), at most out_cap bytes.
int64_t ref_delim_fsm(const char* buffer, int32_t begin, int32_t end, char sep, char quote, uint8_t* out,
                      uint64_t out_cap, uint64_t* out_used) {
    logtail::DelimiterModeFsmParser p(quote, sep);
    std::vector<std::string> cols;
    bool ok = p.ParseDelimiterLine(buffer, begin, end, cols);
    *out_used = 0;
    if (!ok)
        return -1;
    uint64_t at = 0;
    for (auto& c : cols) {
        uint32_t l = (uint32_t)c.size();
        if (at + 4 + l > out_cap)
            return -2;
        memcpy(out + at, &l, 4);
        memcpy(out + at + 4, c.data(), l);
        at += 4 + l;
    }
    *out_used = at;
    return (int64_t)cols.size();
}
}
