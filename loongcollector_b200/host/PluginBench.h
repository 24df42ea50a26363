// PluginBench.h -- end-to-end driver at the plugin boundary (bench.py's `e2e`): builds event groups the way the file
// reader + ProcessorSplitLogStringNative leave them (one arena chunk of <= group_bytes per group, LogFileReader.cpp:97;
// one LogEvent per line whose `content` value aliases that chunk) and times calls of the processor under test.
// Header-only and independent of the engine: the GPU arm (host_capi.cpp) and the CPU reference arm
// (oracle/ref_plugin.cpp) drive their processors through the very same code.
#pragma once
#include <chrono>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "Models.h"

namespace logtail {

struct PluginBenchResult {
    std::vector<double> seconds; // per repetition: wall time of the Process call(s) only
    uint64_t inEvents = 0, outEvents = 0, groups = 0;
    uint64_t liveContents = 0;   // after the last repetition
    uint64_t checksum = 0;       // sum over live contents of key.size * 131 + value.size * 31 + first value byte
    uint64_t arenaBytes = 0;     // bytes of all group arenas (what travels host -> device)
};

class PluginBench {
public:
    // lines: (off, len) into data, each followed by one separator byte that is copied along (the reader's buffer
    // keeps the '\n's); a group takes whole lines until group_bytes would be exceeded.
    PluginBench(const uint8_t* data, const uint32_t* lineOff, const uint32_t* lineLen, uint64_t nLines,
                uint32_t groupBytes, unsigned buildThreads)
        : mData(data), mOff(lineOff), mLen(lineLen), mN(nLines), mThreads(buildThreads ? buildThreads : 1) {
        uint64_t i = 0;
        while (i < nLines) {
            GroupPlan g;
            g.first = i;
            uint64_t bytes = 0;
            while (i < nLines && (bytes == 0 || bytes + lineLen[i] + 1 <= groupBytes)) {
                bytes += (uint64_t)lineLen[i] + 1;
                ++i;
            }
            g.count = i - g.first;
            g.bytes = bytes;
            mPlan.push_back(g);
        }
        // arenas are allocated once (pinned if the process installed the engine's chunk allocator) and keep the bytes
        mArenas.resize(mPlan.size());
        mChunk.resize(mPlan.size());
        Parallel(mPlan.size(), [&](size_t a, size_t b) {
            for (size_t g = a; g < b; ++g) {
                mArenas[g] = std::make_shared<SourceBuffer>();
                StringBuffer sb = mArenas[g]->AllocateStringBuffer(mPlan[g].bytes);
                char* p = sb.data;
                for (uint64_t k = mPlan[g].first; k < mPlan[g].first + mPlan[g].count; ++k) {
                    memcpy(p, mData + mOff[k], (size_t)mLen[k] + 1);
                    p += (size_t)mLen[k] + 1;
                }
                mChunk[g] = sb.data;
            }
        });
    }

    size_t Groups() const { return mPlan.size(); }

    // fresh groups over the persistent arenas (Process rewrites the events, never the arena bytes)
    void Build(std::vector<PipelineEventGroup>& out) {
        out.clear();
        out.reserve(mPlan.size());
        for (size_t g = 0; g < mPlan.size(); ++g)
            out.emplace_back(mArenas[g]);
        static const std::string kKey = "content";
        Parallel(mPlan.size(), [&](size_t a, size_t b) {
            for (size_t g = a; g < b; ++g) {
                EventsContainer& ev = out[g].MutableEvents();
                ev.reserve(mPlan[g].count);
                const char* p = mChunk[g];
                for (uint64_t k = mPlan[g].first; k < mPlan[g].first + mPlan[g].count; ++k) {
                    auto e = out[g].CreateLogEvent(true);
                    e->SetTimestamp(1700000000 + (time_t)(k & 0xFFFF));
                    e->SetContentNoCopy(StringView(kKey), StringView(p, mLen[k]));
                    ev.emplace_back(std::move(e), true, nullptr);
                    p += (size_t)mLen[k] + 1;
                }
            }
        });
    }

    // process(groups) is the timed call; it receives ALL groups of one repetition
    PluginBenchResult Run(int reps, const std::function<void(std::vector<PipelineEventGroup>&)>& process) {
        PluginBenchResult r;
        r.groups = mPlan.size();
        for (auto& g : mPlan)
            r.arenaBytes += g.bytes;
        std::vector<PipelineEventGroup> groups;
        for (int rep = 0; rep < reps; ++rep) {
            Build(groups);
            r.inEvents = mN;
            const auto t0 = std::chrono::steady_clock::now();
            process(groups);
            r.seconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        for (auto& g : groups) {
            r.outEvents += g.GetEvents().size();
            for (const auto& e : g.GetEvents()) {
                if (!e.Is<LogEvent>())
                    continue;
                for (const auto& c : e.Cast<LogEvent>().RawContents()) {
                    if (!c.second)
                        continue;
                    ++r.liveContents;
                    r.checksum += (uint64_t)c.first.first.size() * 131 + (uint64_t)c.first.second.size() * 31 +
                                  (c.first.second.empty() ? 0 : (uint8_t)c.first.second[0]);
                }
            }
        }
        return r;
    }

private:
    struct GroupPlan {
        uint64_t first = 0, count = 0, bytes = 0;
    };
    template <class Fn>
    void Parallel(size_t n, Fn fn) {
        unsigned t = mThreads;
        if (n < t)
            t = n ? (unsigned)n : 1;
        std::vector<std::thread> th;
        for (unsigned k = 1; k < t; ++k)
            th.emplace_back([&, k] { fn(n * k / t, n * (k + 1) / t); });
        fn(0, n / t);
        for (auto& x : th)
            x.join();
    }
    const uint8_t* mData;
    const uint32_t* mOff;
    const uint32_t* mLen;
    uint64_t mN;
    unsigned mThreads;
    std::vector<GroupPlan> mPlan;
    std::vector<std::shared_ptr<SourceBuffer>> mArenas;
    std::vector<const char*> mChunk;
};

} // namespace logtail
