// ref_plugin.cpp -- CPU ORACLE at the plugin boundary.  TEST INFRASTRUCTURE ONLY (bench.py --impl reference and its
// cpu_baseline leg; tests/).  The product library never links, loads or calls it.
//
// A restatement of ProcessorParseRegexNative as the reference runs it on host cores:
//   Init ............ core/plugin/processor/ProcessorParseRegexNative.cpp:34-98 (SourceKey / Regex / Keys, one regex copy
//                     per processing thread :64-67) + CommonParserOptions.cpp:28-89
//   ProcessEvent .... :132-168;  RegexLogLineParser :186-253 (regex_match per event, AddLog per capture)
//   policy .......... CommonParserOptions.cpp:91-117
//   threading ....... process_thread_count ProcessorRunner threads, each popping whole groups and calling
//                     Process(group) on the shared instance (runner/ProcessorRunner.cpp:48-53,128-143)
// over the same event model (loongcollector_b200/host/Models.{h,cpp}, compiled INTO this library -- no product .so is
// loaded) and the same group builder (host/PluginBench.h) as the GPU arm, so that both arms pay the identical
// per-event object costs (SetContentNoCopy per capture, tombstones, erase) and differ only in where regex_match runs.
// Regex arithmetic = the flat oracle's PCRE2 wrapper (lc_oracle.c; boost.regex is not installable here).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../loongcollector_b200/host/Models.h"
#include "../loongcollector_b200/host/PluginBench.h"

extern "C" {
// lc_oracle.c (linked into this library)
typedef struct orc_regex orc_regex;
typedef struct orc_matcher orc_matcher;
orc_regex* orc_regex_compile(const char* pattern, uint64_t len, int jit);
void orc_regex_free(orc_regex* r);
uint32_t orc_regex_ngroups(const orc_regex* r);
orc_matcher* orc_matcher_create(const orc_regex* re);
void orc_matcher_free(orc_matcher* m);
int orc_regex_full_match(orc_matcher* m, const uint8_t* buf, uint64_t len, uint32_t* cap_off, uint32_t* cap_len);
}

using namespace logtail;

namespace {

struct CpuParseRegex {
    std::string sourceKey, pattern, renamedSourceKey;
    std::vector<std::string> keys;
    bool keepFail = false, keepOk = false, copyRaw = false, sourceKeyOverwritten = false, wholeLine = false;
    orc_regex* re = nullptr;
    uint32_t groups = 0;
    std::atomic<uint64_t> discarded{0}, failed{0}, keyNotFound{0}, successful{0};

    ~CpuParseRegex() {
        if (re)
            orc_regex_free(re);
    }

    void Init(const Json::Value& cfg) {
        sourceKey = cfg["SourceKey"].asString();
        pattern = cfg["Regex"].asString();
        for (size_t i = 0; i < cfg["Keys"].size(); ++i)
            keys.push_back(cfg["Keys"][i].asString());
        if (cfg.isMember("KeepingSourceWhenParseFail"))
            keepFail = cfg["KeepingSourceWhenParseFail"].asBool();
        if (cfg.isMember("KeepingSourceWhenParseSucceed"))
            keepOk = cfg["KeepingSourceWhenParseSucceed"].asBool();
        if (cfg.isMember("CopingRawLog"))
            copyRaw = cfg["CopingRawLog"].asBool();
        renamedSourceKey = cfg.isMember("RenamedSourceKey") ? cfg["RenamedSourceKey"].asString() : sourceKey;
        wholeLine = pattern == "(.*)";
        for (auto& k : keys)
            if (k == sourceKey)
                sourceKeyOverwritten = true;
        re = orc_regex_compile(pattern.data(), pattern.size(), 0);
        if (!re)
            throw std::runtime_error("regex does not compile");
        groups = orc_regex_ngroups(re);
    }

    static void AddLog(LogEvent& ev, StringView key, StringView value, bool overwritten = true) {
        if (!overwritten && ev.HasContent(key))
            return;
        ev.SetContentNoCopy(key, value);
    }

    // CommonParserOptions::ShouldEraseEvent (:99-117) for groups without file-offset / container metadata
    bool ShouldErase(bool ok, const LogEvent& ev) const {
        if (ok || keepFail)
            return false;
        return ev.Empty();
    }

    // Process(PipelineEventGroup&) with this thread's matcher (mReg[threadNo])
    void Process(PipelineEventGroup& group, orc_matcher* m, std::vector<uint32_t>& co, std::vector<uint32_t>& cl) {
        EventsContainer& events = group.MutableEvents();
        size_t wIdx = 0;
        uint64_t nDisc = 0, nFail = 0, nKnf = 0, nOk = 0;
        for (size_t rIdx = 0; rIdx < events.size(); ++rIdx) {
            bool keep = true;
            PipelineEventPtr& e = events[rIdx];
            if (!e.Is<LogEvent>()) {
                ++nFail;
            } else {
                LogEvent& ev = e.Cast<LogEvent>();
                if (!ev.HasContent(sourceKey)) {
                    ++nKnf;
                } else {
                    StringView raw = ev.GetContent(sourceKey);
                    bool ok = true;
                    if (wholeLine) {
                        AddLog(ev, keys.empty() ? StringView("content") : StringView(keys[0]), raw);
                    } else if (!orc_regex_full_match(m, reinterpret_cast<const uint8_t*>(raw.data()), raw.size(),
                                                     co.data(), cl.data())) {
                        ++nFail;
                        ok = false;
                    } else if (groups + 1 <= keys.size()) {
                        ok = false; // what.size() <= keys.size(): fail without out_failed++ (:227-244)
                    } else {
                        for (uint32_t k = 0; k < keys.size(); ++k)
                            AddLog(ev, keys[k], StringView(raw.data() + co[k], cl[k]));
                    }
                    if (!ok || !sourceKeyOverwritten)
                        ev.DelContent(sourceKey);
                    if ((ok && keepOk) || (!ok && keepFail))
                        AddLog(ev, renamedSourceKey, raw, false);
                    if (!ok && keepFail && copyRaw)
                        AddLog(ev, StringView("__raw_log__"), raw, false);
                    if (ShouldErase(ok, ev)) {
                        ++nDisc;
                        keep = false;
                    } else {
                        ++nOk;
                    }
                }
            }
            if (keep) {
                if (wIdx != rIdx)
                    events[wIdx] = std::move(events[rIdx]);
                ++wIdx;
            }
        }
        events.resize(wIdx);
        discarded += nDisc;
        failed += nFail;
        keyNotFound += nKnf;
        successful += nOk;
    }
};

char* dupstr(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

} // namespace

extern "C" {

// Same contract as lc_host_bench_plugin (include/lc_b200_host.h) for type "processor_parse_regex_native", run on
// `threads` host threads that take whole groups round-robin (the reference's ProcessorRunner threads).
int orc_bench_plugin(const char* config_json, const uint8_t* data, const uint32_t* line_off, const uint32_t* line_len,
                     uint64_t n_lines, uint32_t group_bytes, int threads, int reps, double* seconds_out,
                     uint64_t stats_out[12], char** err_out) {
    if (err_out)
        *err_out = nullptr;
    try {
        Json::Value cfg(Json::objectValue);
        std::string err;
        const char* cj = config_json ? config_json : "{}";
        if (!Json::Value::parse(cj, cj + strlen(cj), cfg, err))
            throw std::runtime_error("config is not valid JSON: " + err);
        CpuParseRegex proc;
        proc.Init(cfg);
        if (threads < 1)
            threads = 1;
        std::vector<orc_matcher*> matchers(threads);
        for (auto& m : matchers)
            m = orc_matcher_create(proc.re);
        PluginBench bench(data, line_off, line_len, n_lines, group_bytes, (unsigned)std::min(threads, 32));
        std::atomic<uint64_t> inEv{0}, outEv{0}, inBytes{0}, outBytes{0};
        PluginBenchResult r = bench.Run(reps, [&](std::vector<PipelineEventGroup>& groups) {
            std::atomic<size_t> next{0};
            auto work = [&](int t) {
                std::vector<uint32_t> co(proc.groups + 1), cl(proc.groups + 1);
                for (;;) {
                    const size_t g = next.fetch_add(1);
                    if (g >= groups.size())
                        break;
                    // ProcessorInstance::Process around the plugin call (ProcessorInstance.cpp:46-63)
                    inEv += groups[g].GetEvents().size();
                    inBytes += groups[g].DataSize();
                    proc.Process(groups[g], matchers[t], co, cl);
                    outEv += groups[g].GetEvents().size();
                    outBytes += groups[g].DataSize();
                }
            };
            std::vector<std::thread> th;
            for (int t = 1; t < threads; ++t)
                th.emplace_back(work, t);
            work(0);
            for (auto& x : th)
                x.join();
        });
        for (auto m : matchers)
            orc_matcher_free(m);
        for (int k = 0; k < reps && seconds_out; ++k)
            seconds_out[k] = r.seconds[k];
        if (stats_out) {
            stats_out[0] = r.groups;
            stats_out[1] = r.inEvents;
            stats_out[2] = r.outEvents;
            stats_out[3] = r.liveContents;
            stats_out[4] = r.checksum;
            stats_out[5] = r.arenaBytes;
            stats_out[6] = inEv;
            stats_out[7] = outEv;
            stats_out[8] = inBytes;
            stats_out[9] = outBytes;
            stats_out[10] = 0;
            stats_out[11] = 0;
        }
        return 0;
    } catch (const std::exception& e) {
        if (err_out)
            *err_out = dupstr(e.what());
        return 1;
    }
}

void orc_string_free(char* s) { free(s); }
}
