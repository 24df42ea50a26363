#include "Processors.h"

#include <stdlib.h>
#include <string.h>

#include <sched.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace logtail {

// ------------------------------------------------------------------------------------------------ engine per thread
namespace {

// One engine per (GPU, host thread): the reference keeps one regex copy per ProcessorRunner thread
// (ProcessorParseRegexNative.cpp:64-67); here the per-thread state is the engine's stream and workspace.
struct ThreadEngine {
    lc_engine_t* e = nullptr;
    ~ThreadEngine() {
        if (e)
            lc_engine_destroy(e);
    }
};

lc_engine_t* Engine() {
    static thread_local ThreadEngine t;
    if (!t.e) {
        const char* d = getenv("LC_B200_DEVICE");
        int dev = d ? atoi(d) : 0;
        if (lc_engine_create(dev, &t.e) != LC_OK)
            throw std::runtime_error(std::string("loongcollector_b200: ") + lc_last_error());
    }
    return t.e;
}

void Check(int rc, const char* what) {
    if (rc != LC_OK)
        throw std::runtime_error(std::string(what) + ": " + lc_last_error());
}

// Grow-only pinned host table (lc_host_alloc): result tables land here by DMA without a staging copy.
template <class T>
struct PinnedVec {
    T* p = nullptr;
    size_t cap = 0;
    ~PinnedVec() { lc_host_free(p); }
    T* ensure(size_t n) {
        if (n > cap) {
            lc_host_free(p);
            cap = n + n / 4 + 64;
            p = static_cast<T*>(lc_host_alloc(cap * sizeof(T)));
            if (!p) {
                cap = 0;
                throw std::runtime_error(std::string("loongcollector_b200: ") + lc_last_error());
            }
        }
        return p;
    }
};

} // namespace

struct ThreadScratch {
    PinnedVec<uint32_t> off, len, capOff, capLen;
    PinnedVec<uint8_t> status, staging;
};

namespace {
ThreadScratch& Scratch() {
    static thread_local ThreadScratch s;
    return s;
}

// Host threads used for the gather / epilogue of a batched Process call (env LC_B200_HOST_THREADS; default: the CPUs of
// the process's affinity mask, at most 64).  The reference spends this work on its process_thread_count ProcessorRunner threads.
unsigned HostThreads() {
    static const unsigned n = [] {
        const char* e = getenv("LC_B200_HOST_THREADS");
        unsigned hw = std::thread::hardware_concurrency();
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0)
            hw = (unsigned)CPU_COUNT(&set); // the CPUs this process may use (e.g. the GPU's NUMA node)
        unsigned want = e ? (unsigned)atoi(e) : 64u;
        if (want < 1)
            want = 1;
        if (hw && want > hw)
            want = hw;
        return want;
    }();
    return n;
}

// Persistent worker pool: a batched Process call runs several short parallel regions (gather, one epilogue per engine
// chunk), so the threads are created once per process, not per region.  One region at a time; a caller that finds the
// pool busy (another ProcessorRunner thread inside its own batch) simply runs its region on its own thread.
class HostPool {
public:
    using Fn = std::function<void(size_t, size_t, unsigned)>;
    static HostPool& Get() {
        static HostPool p(HostThreads());
        return p;
    }
    unsigned Size() const { return mN; }
    void Run(size_t n, size_t minPerThread, const Fn& fn) {
        unsigned t = mN;
        if (minPerThread && n / minPerThread < t)
            t = (unsigned)std::max<size_t>(1, n / minPerThread);
        std::unique_lock<std::mutex> busy(mBusy, std::try_to_lock);
        if (t <= 1 || !busy.owns_lock()) {
            fn(0, n, 0u);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mMu);
            mFn = &fn;
            mTotal = n;
            mParts = t;
            mPending = t - 1;
            ++mGen;
        }
        mCv.notify_all();
        fn(0, n / t, 0u);
        std::unique_lock<std::mutex> lk(mMu);
        mDone.wait(lk, [&] { return mPending == 0; });
        mFn = nullptr;
    }

private:
    explicit HostPool(unsigned n) : mN(n ? n : 1) {
        for (unsigned k = 1; k < mN; ++k)
            mThreads.emplace_back([this, k] { Work(k); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mMu);
            mStop = true;
            ++mGen;
        }
        mCv.notify_all();
        for (auto& t : mThreads)
            t.join();
    }
    void Work(unsigned k) {
        uint64_t seen = 0;
        for (;;) {
            const Fn* fn;
            size_t total;
            unsigned parts;
            {
                std::unique_lock<std::mutex> lk(mMu);
                mCv.wait(lk, [&] { return mGen != seen; });
                seen = mGen;
                if (mStop)
                    return;
                fn = mFn;
                total = mTotal;
                parts = mParts;
            }
            if (k < parts && fn)
                (*fn)(total * k / parts, total * (k + 1) / parts, k);
            if (k < parts) {
                std::lock_guard<std::mutex> lk(mMu);
                if (--mPending == 0)
                    mDone.notify_one();
            }
        }
    }
    unsigned mN;
    std::vector<std::thread> mThreads;
    std::mutex mMu, mBusy;
    std::condition_variable mCv, mDone;
    const Fn* mFn = nullptr;
    size_t mTotal = 0;
    unsigned mParts = 0, mPending = 0;
    uint64_t mGen = 0;
    bool mStop = false;
};

// fn(begin, end, thread) over [0, n) in contiguous slices
template <class Fn>
void ParallelFor(size_t n, size_t minPerThread, Fn fn) {
    HostPool::Get().Run(n, minPerThread, fn);
}

// Flattens the source values of the events to be parsed into (base, off[], len[]).  If every value lies
// inside ONE arena chunk (the production shape: all lines alias the file read buffer) the span is handed to
// the engine as is; otherwise the values are packed into a staging buffer.
struct FlatBatch {
    std::vector<uint32_t> off, len;
    std::vector<size_t> eventIndex;
    std::vector<const char*> origin; // original data pointer of each value
    const uint8_t* base = nullptr;
    uint64_t baseLen = 0;
    std::vector<uint8_t> packed;

    void Add(size_t idx, StringView v) {
        eventIndex.push_back(idx);
        origin.push_back(v.data());
        len.push_back((uint32_t)v.size());
    }
    void Finish(SourceBuffer& sb) {
        size_t n = origin.size();
        off.resize(n);
        if (!n)
            return;
        const char* lo = origin[0];
        const char* hi = origin[0] + len[0];
        for (size_t i = 1; i < n; ++i) {
            lo = std::min(lo, origin[i]);
            hi = std::max(hi, origin[i] + len[i]);
        }
        size_t chunkSize = 0;
        if ((uint64_t)(hi - lo) < 0xFFFFFFF0ull && sb.ChunkContaining(lo, (size_t)(hi - lo), &chunkSize)) {
            base = reinterpret_cast<const uint8_t*>(lo);
            baseLen = (uint64_t)(hi - lo);
            for (size_t i = 0; i < n; ++i)
                off[i] = (uint32_t)(origin[i] - lo);
            return;
        }
        size_t total = 0;
        for (size_t i = 0; i < n; ++i)
            total += len[i];
        packed.resize(total + 1);
        size_t at = 0;
        for (size_t i = 0; i < n; ++i) {
            off[i] = (uint32_t)at;
            if (len[i])
                memcpy(packed.data() + at, origin[i], len[i]);
            at += len[i];
        }
        base = packed.data();
        baseLen = total;
    }
    // view of [o, o + l) (offsets relative to base) for event i, mapped back onto the original bytes
    StringView View(size_t i, uint32_t o, uint32_t l) const { return StringView(origin[i] + (o - off[i]), l); }
};

bool GetString(const Json::Value& cfg, const char* key, std::string& out) {
    if (cfg.isMember(key) && cfg[key].isString()) {
        out = cfg[key].asString();
        return true;
    }
    return false;
}
void GetBool(const Json::Value& cfg, const char* key, bool& out) {
    if (cfg.isMember(key) && cfg[key].isBool())
        out = cfg[key].asBool();
}

void AddLog(LogEvent& ev, StringView key, StringView value, bool overwritten = true) {
    if (!overwritten && ev.HasContent(key))
        return;
    ev.SetContentNoCopy(key, value);
}

std::string ToString(uint64_t v) {
    return std::to_string(v);
}

// CreateNewEvent of both splitters (ProcessorSplitLogStringNative.cpp:135-157,
// ProcessorSplitMultilineLogStringNative.cpp:311-340)
void EmitSplitEvent(PipelineEventGroup& group, const LogEvent& src, StringView sourceVal, const StringBuffer& sourceKey,
                    StringView content, bool isLast, bool raw, EventsContainer& out) {
    if (raw) {
        auto t = group.CreateRawEvent(true);
        t->SetContentNoCopy(content);
        t->SetTimestamp(src.GetTimestamp(), src.GetTimestampNanosecond());
        out.emplace_back(std::move(t), true, nullptr);
        return;
    }
    auto t = group.CreateLogEvent(true);
    t->SetContentNoCopy(StringView(sourceKey.data, sourceKey.size), content);
    t->SetTimestamp(src.GetTimestamp(), src.GetTimestampNanosecond());
    uint64_t rel = (uint64_t)(content.data() - sourceVal.data());
    uint64_t offset = src.GetPosition().first + rel;
    uint64_t length = isLast ? src.GetPosition().second - rel : content.size() + 1;
    t->SetPosition(offset, length);
    if (group.HasMetadata(EventGroupMetaKey::LOG_FILE_OFFSET_KEY)) {
        StringBuffer offStr = group.GetSourceBuffer()->CopyString(ToString(offset));
        t->SetContentNoCopy(group.GetMetadata(EventGroupMetaKey::LOG_FILE_OFFSET_KEY),
                            StringView(offStr.data, offStr.size));
    }
    out.emplace_back(std::move(t), true, nullptr);
}

} // namespace

// ------------------------------------------------------------------------------------------------ options
const std::string CommonParserOptions::legacyUnmatchedRawLogKey = "__raw_log__";

bool CommonParserOptions::Init(const Json::Value& config) {
    GetBool(config, "KeepingSourceWhenParseFail", mKeepingSourceWhenParseFail);
    GetBool(config, "KeepingSourceWhenParseSucceed", mKeepingSourceWhenParseSucceed);
    GetString(config, "RenamedSourceKey", mRenamedSourceKey);
    if (mRenamedSourceKey.empty())
        mRenamedSourceKey = config["SourceKey"].asString();
    GetBool(config, "CopingRawLog", mCopingRawLog);
    return true;
}
bool CommonParserOptions::ShouldAddSourceContent(bool ok) const {
    return (ok && mKeepingSourceWhenParseSucceed) || (!ok && mKeepingSourceWhenParseFail);
}
bool CommonParserOptions::ShouldAddLegacyUnmatchedRawLog(bool ok) const {
    return !ok && mKeepingSourceWhenParseFail && mCopingRawLog;
}
bool CommonParserOptions::ShouldEraseEvent(bool ok, const LogEvent& ev, const GroupMetadata& md) const {
    if (!ok && !mKeepingSourceWhenParseFail) {
        if (ev.Empty())
            return true;
        size_t size = ev.Size();
        auto offsetKey = md.find(EventGroupMetaKey::LOG_FILE_OFFSET_KEY);
        if (size == 1 && offsetKey != md.end() && ev.FirstLive()->first.first == offsetKey->second)
            return true;
        if (size == 2 && ev.HasContent("_time_") && ev.HasContent("_source_"))
            return true;
    }
    return false;
}

CompiledRegex::~CompiledRegex() {
    if (mRe)
        lc_regex_free(mRe);
}
bool CompiledRegex::Compile(const std::string& pattern, std::string& err) {
    if (mRe) {
        lc_regex_free(mRe);
        mRe = nullptr;
    }
    int rc = lc_regex_compile(pattern.data(), pattern.size(), &mRe);
    if (rc != LC_OK) {
        err = lc_last_error();
        if (mRe) {
            lc_regex_free(mRe);
            mRe = nullptr;
        }
        return false;
    }
    return true;
}

static bool EndsWith(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    return s.size() >= n && !s.compare(s.size() - n, n, suf);
}

bool MultilineOptions::Init(const Json::Value& config, std::string& err) {
    // dotted parameter names resolve to their last segment (ParamExtractor.cpp:23-29)
    struct {
        const char* key;
        std::string* dst;
    } pats[] = {{"StartPattern", &mStartPattern}, {"ContinuePattern", &mContinuePattern}, {"EndPattern", &mEndPattern}};
    bool compiled[3] = {false, false, false};
    int k = 0;
    for (auto& p : pats) {
        std::string pattern;
        if (GetString(config, p.key, pattern)) {
            // validation strips a trailing '$' and trailing ".*"s, yet the ORIGINAL pattern is stored (:205-222)
            std::string probe = pattern;
            if (!probe.empty() && EndsWith(probe, "$"))
                probe.pop_back();
            while (!probe.empty() && EndsWith(probe, ".*"))
                probe.resize(probe.size() - 2);
            bool valid = true;
            if (!probe.empty()) {
                CompiledRegex tmp;
                std::string e;
                lc_regex_t* raw = nullptr;
                int rc = lc_regex_compile(probe.data(), probe.size(), &raw);
                if (raw)
                    lc_regex_free(raw);
                valid = rc != LC_ERR_REGEX_INVALID;
                compiled[k] = valid;
            }
            if (valid)
                *p.dst = pattern;
        }
        ++k;
    }
    if (compiled[0] || compiled[2])
        mIsMultiline = true;
    std::string t;
    if (GetString(config, "UnmatchedContentTreatment", t) && t == "discard")
        mUnmatchedContentTreatment = UnmatchedContentTreatment::DISCARD;
    GetBool(config, "IgnoringUnmatchWarning", mIgnoringUnmatchWarning);
    (void)err;
    return true;
}

// ------------------------------------------------------------------------------------------------ split
const std::string ProcessorSplitLogStringNative::sName = "processor_split_string_native";

bool ProcessorSplitLogStringNative::Init(const Json::Value& config) {
    GetString(config, "SourceKey", mSourceKey);
    if (config.isMember("SplitChar") && config["SplitChar"].isInt())
        mSplitChar = (char)config["SplitChar"].asInt();
    GetBool(config, "EnableRawContent", mEnableRawContent);
    return true;
}

void ProcessorSplitLogStringNative::Process(PipelineEventGroup& group) {
    try {
        ProcessImpl(group);
    } catch (const std::exception& ex) {
        EngineFailed(ex.what()); // the reference's Process never throws: the group stays as it was
    }
}

void ProcessorSplitLogStringNative::ProcessImpl(PipelineEventGroup& group) {
    if (group.GetEvents().empty())
        return;
    EventsContainer newEvents;
    std::vector<uint32_t> off, len;
    for (PipelineEventPtr& e : group.MutableEvents()) {
        if (!IsSupportedEvent(e)) {
            newEvents.emplace_back(std::move(e));
            continue;
        }
        LogEvent& src = e.Cast<LogEvent>();
        if (src.Size() != 1 || !src.HasContent(mSourceKey)) {
            newEvents.emplace_back(std::move(e));
            continue;
        }
        StringView val = src.GetContent(mSourceKey);
        StringBuffer sourceKey = group.GetSourceBuffer()->CopyString(mSourceKey);
        if (val.empty())
            continue;
        // a1 on the GPU: one (offset, length) per piece
        uint64_t n = 0;
        uint64_t cap = std::max<uint64_t>(1024, val.size() / 16);
        for (;;) {
            off.resize(cap);
            len.resize(cap);
            int rc = lc_split_lines(Engine(), reinterpret_cast<const uint8_t*>(val.data()), val.size(),
                                    (uint8_t)mSplitChar, off.data(), len.data(), cap, &n);
            if (rc == LC_ERR_CAPACITY) {
                cap = n;
                continue;
            }
            Check(rc, "lc_split_lines");
            break;
        }
        for (uint64_t k = 0; k < n; ++k) {
            StringView content(val.data() + off[k], len[k]);
            bool isLast = (uint64_t)off[k] + len[k] == val.size();
            EmitSplitEvent(group, src, val, sourceKey, content, isLast, mEnableRawContent, newEvents);
        }
    }
    group.SwapEvents(newEvents);
}

// ------------------------------------------------------------------------------------------------ multiline
const std::string ProcessorSplitMultilineLogStringNative::sName = "processor_split_multiline_log_string_native";

bool ProcessorSplitMultilineLogStringNative::Init(const Json::Value& config) {
    GetString(config, "SourceKey", mSourceKey);
    std::string err;
    if (!mMultiline.Init(config, err))
        return Fail(err);
    GetBool(config, "EnableRawContent", mEnableRawContent);
    // the processor compiles the ORIGINAL pattern strings (:70-80); an unsupported pattern fails Init loudly
    if (!mMultiline.mStartPattern.empty() && !mStart.Compile(mMultiline.mStartPattern, err))
        return Fail("Multiline.StartPattern: " + err);
    if (!mMultiline.mContinuePattern.empty() && !mContinue.Compile(mMultiline.mContinuePattern, err))
        return Fail("Multiline.ContinuePattern: " + err);
    if (!mMultiline.mEndPattern.empty() && !mEnd.Compile(mMultiline.mEndPattern, err))
        return Fail("Multiline.EndPattern: " + err);
    return true;
}

std::vector<std::pair<std::string, uint64_t>> ProcessorSplitMultilineLogStringNative::Counters() const {
    return {{"matched_events", mMatchedEventsTotal.GetValue()},
            {"matched_lines", mMatchedLinesTotal.GetValue()},
            {"unmatched_lines", mUnmatchedLinesTotal.GetValue()}};
}

void ProcessorSplitMultilineLogStringNative::Process(PipelineEventGroup& group) {
    try {
        ProcessImpl(group);
    } catch (const std::exception& ex) {
        EngineFailed(ex.what()); // the reference's Process never throws: the group stays as it was
    }
}

void ProcessorSplitMultilineLogStringNative::ProcessImpl(PipelineEventGroup& group) {
    if (group.GetEvents().empty())
        return;
    EventsContainer newEvents;
    uint64_t inputLines = 0, unmatchLines = 0;
    std::vector<uint32_t> off, len;
    std::vector<uint8_t> flags;
    for (PipelineEventPtr& e : group.MutableEvents()) {
        if (!IsSupportedEvent(e)) {
            newEvents.emplace_back(std::move(e));
            continue;
        }
        LogEvent& src = e.Cast<LogEvent>();
        if (src.Size() != 1 || !src.HasContent(mSourceKey)) {
            newEvents.emplace_back(std::move(e));
            continue;
        }
        StringView val = src.GetContent(mSourceKey);
        StringBuffer sourceKey = group.GetSourceBuffer()->CopyString(mSourceKey);
        if (val.empty())
            continue;
        uint64_t n = 0, ctr[3] = {0, 0, 0};
        uint64_t cap = std::max<uint64_t>(1024, val.size() / 16);
        for (;;) {
            off.resize(cap);
            len.resize(cap);
            flags.resize(cap);
            uint64_t c[3] = {0, 0, 0};
            int rc = lc_multiline_split(
                Engine(), reinterpret_cast<const uint8_t*>(val.data()), val.size(), mStart.get(), mContinue.get(),
                mEnd.get(), mMultiline.mUnmatchedContentTreatment == MultilineOptions::UnmatchedContentTreatment::DISCARD,
                off.data(), len.data(), flags.data(), cap, &n, c);
            if (rc == LC_ERR_CAPACITY) {
                cap = n;
                continue;
            }
            Check(rc, "lc_multiline_split");
            memcpy(ctr, c, sizeof ctr);
            break;
        }
        mMatchedEventsTotal.Add(ctr[0]);
        inputLines += ctr[1];
        unmatchLines += ctr[2];
        for (uint64_t k = 0; k < n; ++k)
            EmitSplitEvent(group, src, val, sourceKey, StringView(val.data() + off[k], len[k]),
                           (flags[k] & LC_ML_IS_LAST) != 0, mEnableRawContent, newEvents);
    }
    mMatchedLinesTotal.Add(inputLines - unmatchLines);
    mUnmatchedLinesTotal.Add(unmatchLines);
    group.SwapEvents(newEvents);
}

// ------------------------------------------------------------------------------------------------ regex parse
const std::string ProcessorParseRegexNative::sName = "processor_parse_regex_native";

static bool GetKeys(const Json::Value& config, std::vector<std::string>& keys) {
    if (!config.isMember("Keys") || !config["Keys"].isArray())
        return false;
    keys.clear();
    for (const auto& k : config["Keys"]) {
        if (!k.isString())
            return false;
        keys.push_back(k.asString());
    }
    return true;
}

bool ProcessorParseRegexNative::Init(const Json::Value& config) {
    if (!GetString(config, "SourceKey", mSourceKey))
        return Fail("mandatory string param SourceKey is missing");
    if (!GetString(config, "Regex", mRegex))
        return Fail("mandatory string param Regex is missing");
    mIsWholeLineMode = mRegex == "(.*)";
    std::string err;
    if (!mIsWholeLineMode && !mReg.Compile(mRegex, err))
        return Fail("mandatory string param Regex is not usable: " + err);
    if (!GetKeys(config, mKeys))
        return Fail("mandatory list param Keys is missing");
    // legacy ["k1,k2"] form (:80-88)
    if (mKeys.size() == 1 && mKeys[0].find(',') != std::string::npos) {
        std::vector<std::string> parts;
        size_t start = 0;
        const std::string joined = mKeys[0];
        for (;;) {
            size_t pos = joined.find(',', start);
            parts.push_back(joined.substr(start, pos == std::string::npos ? std::string::npos : pos - start));
            if (pos == std::string::npos)
                break;
            start = pos + 1;
        }
        mKeys = parts;
    }
    for (const auto& k : mKeys)
        if (k == mSourceKey)
            mSourceKeyOverwritten = true;
    mKeysDistinct = true;
    for (size_t a = 0; a < mKeys.size(); ++a)
        for (size_t b = a + 1; b < mKeys.size(); ++b)
            if (mKeys[a] == mKeys[b])
                mKeysDistinct = false;
    return mCommonParserOptions.Init(config);
}

std::vector<std::pair<std::string, uint64_t>> ProcessorParseRegexNative::Counters() const {
    return {{"discarded", mDiscardedEventsTotal.GetValue()},
            {"out_failed", mOutFailedEventsTotal.GetValue()},
            {"out_key_not_found", mOutKeyNotFoundEventsTotal.GetValue()},
            {"out_successful", mOutSuccessfulEventsTotal.GetValue()},
            {"b200_gather_ns", mGatherNs.GetValue()},
            {"b200_engine_ns", mEngineNs.GetValue()},
            {"b200_epilogue_ns", mEpilogueNs.GetValue()}};
}

void ProcessorParseRegexNative::AddCounters(const LocalCounters& c) {
    if (c.discarded)
        mDiscardedEventsTotal.Add(c.discarded);
    if (c.failed)
        mOutFailedEventsTotal.Add(c.failed);
    if (c.keyNotFound)
        mOutKeyNotFoundEventsTotal.Add(c.keyNotFound);
    if (c.successful)
        mOutSuccessfulEventsTotal.Add(c.successful);
}

// ProcessEvent (:132-168) with the regex verdict already known.  r == nullptr: the event never reached the engine
// (unsupported type, source key absent, or whole-line mode).
bool ProcessorParseRegexNative::FinishEvent(PipelineEventGroup& group, PipelineEventPtr& e,
                                            const LogEvent::Content* src, const EventResult* r,
                                            LocalCounters& c) const {
    if (!IsSupportedEvent(e)) {
        ++c.failed;
        return true;
    }
    LogEvent& ev = e.Cast<LogEvent>();
    if (!src) {
        ++c.keyNotFound;
        return true;
    }
    StringView rawContent = src->first.second;
    bool ok = true;
    if (mIsWholeLineMode) {
        AddLog(ev, mKeys.empty() ? StringView("content") : StringView(mKeys[0]), rawContent);
    } else if (r->status == LC_REGEX_NOMATCH) {
        ++c.failed;
        ok = false;
    } else if (r->status == LC_REGEX_KEYS_MISMATCH) {
        ok = false;
    } else if (mKeysDistinct && !mSourceKeyOverwritten && ev.RawContents().size() == 1) {
        // the event holds nothing but the source key and no key can collide: SetContentNoCopy's look-up per key
        // (LogEvent.cpp:83-95) would find nothing, so the fields are appended directly (same resulting contents)
        for (uint32_t k = 0; k < mKeys.size(); ++k)
            ev.AppendContentNoCopy(mKeys[k], StringView(r->origin + (r->capOff[k] - r->originOff), r->capLen[k]));
    } else {
        for (uint32_t k = 0; k < mKeys.size(); ++k)
            AddLog(ev, mKeys[k], StringView(r->origin + (r->capOff[k] - r->originOff), r->capLen[k]));
    }
    if (!ok || !mSourceKeyOverwritten)
        ev.DelContent(mSourceKey);
    if (mCommonParserOptions.ShouldAddSourceContent(ok))
        AddLog(ev, mCommonParserOptions.mRenamedSourceKey, rawContent, false);
    if (mCommonParserOptions.ShouldAddLegacyUnmatchedRawLog(ok))
        AddLog(ev, CommonParserOptions::legacyUnmatchedRawLogKey, rawContent, false);
    if (mCommonParserOptions.ShouldEraseEvent(ok, ev, group.GetAllMetadata())) {
        ++c.discarded;
        return false;
    }
    ++c.successful;
    return true;
}

// the epilogue of every event of one group; the group's parsed events own rows firstEv, firstEv + 1, ... of the tables
void ProcessorParseRegexNative::EpilogueGroup(PipelineEventGroup& group, uint64_t firstEv, const ThreadScratch& sc,
                                              uint32_t G, LocalCounters& c) const {
    EventsContainer& events = group.MutableEvents();
    size_t wIdx = 0;
    const size_t nev = events.size();
    for (size_t rIdx = 0; rIdx < nev; ++rIdx) {
        // the walk is bound by cache misses on the event objects and their contents arrays (one heap block each, touched
        // once): pull the object 16 events ahead and its contents array 8 events ahead
        if (rIdx + 16 < nev && events[rIdx + 16])
            __builtin_prefetch(events[rIdx + 16].operator->(), 1, 1);
        if (rIdx + 8 < nev && IsSupportedEvent(events[rIdx + 8])) {
            const auto& rc = events[rIdx + 8].Cast<LogEvent>().RawContents();
            if (!rc.empty()) {
                __builtin_prefetch(rc.data(), 1, 1);
                __builtin_prefetch(reinterpret_cast<const char*>(rc.data()) + 256, 1, 1);
            }
        }
        EventResult r{};
        const EventResult* rp = nullptr;
        const LogEvent::Content* src = nullptr;
        const uint64_t i = firstEv + rIdx; // the event's row of the result tables
        if (IsSupportedEvent(events[rIdx])) {
            src = events[rIdx].Cast<LogEvent>().FindContent(mSourceKey);
            if (src && !mIsWholeLineMode) {
                r.status = sc.status.p[i];
                r.capOff = sc.capOff.p + i * G;
                r.capLen = sc.capLen.p + i * G;
                r.origin = src->first.second.data(); // captures map back onto the ORIGINAL bytes, staged or not
                r.originOff = sc.off.p[i];
                rp = &r;
            }
        }
        if (FinishEvent(group, events[rIdx], src, rp, c)) {
            if (wIdx != rIdx)
                events[wIdx] = std::move(events[rIdx]);
            ++wIdx;
        }
    }
    events.resize(wIdx);
}

void ProcessorParseRegexNative::Process(PipelineEventGroup& group) {
    if (group.GetEvents().empty())
        return;
    ProcessBatch(&group, 1);
}

void ProcessorParseRegexNative::Process(std::vector<PipelineEventGroup>& groups) {
    // sub-batches of at most 2048 groups (<= 512 KB each: <= 1 GiB of arena bytes, far below the 4 GiB / 2^30-event
    // limits of one engine call); groups that are larger than the reader's 512 KB make ProcessBatch split further
    const size_t kMaxGroups = 2048;
    for (size_t g0 = 0; g0 < groups.size(); g0 += kMaxGroups)
        ProcessBatch(groups.data() + g0, std::min(kMaxGroups, groups.size() - g0));
}

// One engine call for `ngroups` groups.  Per group the values to parse normally alias ONE arena chunk (all lines of a
// file read): that chunk range is the group's span and goes to the GPU in place; a group whose values are scattered
// over several chunks is packed into pinned staging first.
void ProcessorParseRegexNative::ProcessBatch(PipelineEventGroup* groups, size_t ngroups) {
    struct GroupPlan {
        uint64_t firstEv = 0, nEv = 0; // slice of the flat event table
        const char* lo = nullptr;      // span = [lo, lo + spanLen) in host memory
        uint32_t spanLen = 0, spanDst = 0;
        bool staged = false;
        uint64_t stagedAt = 0;
    };
    std::vector<GroupPlan> plan(ngroups);
    using Clock = std::chrono::steady_clock;
    auto since = [](Clock::time_point t0) {
        return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t0).count();
    };
    try {
        uint64_t nb = 0;
        const auto tGather = Clock::now();
        if (!mIsWholeLineMode) {
            // Rows of the flat event table map 1:1 onto the events of the groups (row = firstEv of the group + index of
            // the event): an event that does not reach RegexLogLineParser gets an empty row the engine parses for
            // nothing and the epilogue ignores.  That makes the gather ONE pass over the events.
            for (size_t g = 0; g < ngroups; ++g) {
                plan[g].firstEv = nb;
                plan[g].nEv = groups[g].GetEvents().size();
                nb += plan[g].nEv;
            }
            if (nb) {
                ThreadScratch& sc = Scratch();
                const uint32_t G = mReg.groups();
                uint32_t* off = sc.off.ensure(nb);
                uint32_t* len = sc.len.ensure(nb);
                uint8_t* status = sc.status.ensure(nb);
                uint32_t* capOff = sc.capOff.ensure(nb * G + 1);
                uint32_t* capLen = sc.capLen.ensure(nb * G + 1);
                // ---- pass 1 (parallel over groups): span = the arena chunk that holds the group's values (the reader's
                // <= 512 KB buffer); offsets relative to it for now
                ParallelFor(ngroups, 8, [&](size_t a, size_t b, unsigned) {
                    for (size_t g = a; g < b; ++g) {
                        GroupPlan& p = plan[g];
                        const char* base = nullptr;
                        size_t chunkSize = 0;
                        uint64_t total = 0;
                        uint64_t i = p.firstEv;
                        const EventsContainer& gev = groups[g].GetEvents();
                        for (size_t x = 0; x < gev.size(); ++x) {
                            const PipelineEventPtr& e = gev[x];
                            // (cache misses on the event objects and their contents arrays: pull them in ahead)
                            if (x + 16 < gev.size() && gev[x + 16])
                                __builtin_prefetch(gev[x + 16].operator->(), 0, 1);
                            if (x + 8 < gev.size() && IsSupportedEvent(gev[x + 8])) {
                                const auto& rc = gev[x + 8].Cast<LogEvent>().RawContents();
                                if (!rc.empty())
                                    __builtin_prefetch(rc.data(), 0, 1);
                            }
                            const LogEvent::Content* src =
                                IsSupportedEvent(e) ? e.Cast<LogEvent>().FindContent(mSourceKey) : nullptr;
                            if (!src) {
                                off[i] = 0;
                                len[i] = 0;
                                ++i;
                                continue;
                            }
                            StringView v = src->first.second;
                            if (!base && !p.staged) {
                                base = groups[g].GetSourceBuffer()->ChunkContaining(v.data(), v.size(), &chunkSize);
                                if (!base || chunkSize >= (1ull << 31))
                                    p.staged = true;
                            }
                            if (!p.staged && !(v.data() >= base && v.data() + v.size() <= base + chunkSize))
                                p.staged = true;
                            off[i] = p.staged ? 0u : (uint32_t)(v.data() - base);
                            len[i] = (uint32_t)v.size();
                            total += v.size();
                            ++i;
                        }
                        if (p.staged) {
                            p.spanLen = (uint32_t)total;
                        } else {
                            p.lo = base;
                            p.spanLen = base ? (uint32_t)chunkSize : 0u;
                        }
                    }
                });
                {
                    // (oversized batch: groups far larger than the reader's chunks) split and recurse
                    uint64_t bytes = 0;
                    for (auto& p : plan)
                        bytes += ((uint64_t)p.spanLen + 15) & ~15ull;
                    if (ngroups > 1 && (bytes >= (3ull << 30) || nb >= (1ull << 29))) {
                        ProcessBatch(groups, ngroups / 2);
                        ProcessBatch(groups + ngroups / 2, ngroups - ngroups / 2);
                        return;
                    }
                }
                uint64_t dst = 0, stagedBytes = 0;
                for (auto& p : plan) {
                    p.spanDst = (uint32_t)dst;
                    dst += ((uint64_t)p.spanLen + 15) & ~15ull;
                    if (p.staged) {
                        p.stagedAt = stagedBytes;
                        stagedBytes += ((uint64_t)p.spanLen + 15) & ~15ull;
                    }
                }
                uint8_t* staging = stagedBytes ? sc.staging.ensure(stagedBytes) : nullptr;
                // ---- pass 2 (parallel): packed-arena coordinates; groups whose values are scattered over several
                // chunks are packed into pinned staging here
                ParallelFor(ngroups, 8, [&](size_t a, size_t b, unsigned) {
                    for (size_t g = a; g < b; ++g) {
                        GroupPlan& p = plan[g];
                        uint64_t i = p.firstEv;
                        if (!p.staged) {
                            for (uint64_t k = 0; k < p.nEv; ++k)
                                off[i + k] += p.spanDst;
                            continue;
                        }
                        uint64_t at = 0;
                        p.lo = reinterpret_cast<const char*>(staging + p.stagedAt);
                        for (const auto& e : groups[g].GetEvents()) {
                            const LogEvent::Content* src =
                                IsSupportedEvent(e) ? e.Cast<LogEvent>().FindContent(mSourceKey) : nullptr;
                            if (src) {
                                StringView v = src->first.second;
                                if (v.size())
                                    memcpy(staging + p.stagedAt + at, v.data(), v.size());
                                off[i] = p.spanDst + (uint32_t)at;
                                at += v.size();
                            } else {
                                off[i] = p.spanDst;
                            }
                            ++i;
                        }
                    }
                });
                std::vector<const uint8_t*> spanPtr(ngroups);
                std::vector<uint32_t> spanLen(ngroups), spanDst(ngroups);
                std::vector<uint64_t> spanFirst(ngroups + 1);
                for (size_t g = 0; g < ngroups; ++g) {
                    spanPtr[g] = reinterpret_cast<const uint8_t*>(plan[g].lo);
                    spanLen[g] = plan[g].nEv ? plan[g].spanLen : 0;
                    spanDst[g] = plan[g].spanDst;
                    spanFirst[g] = plan[g].firstEv;
                }
                spanFirst[ngroups] = nb;
                mGatherNs.Add(since(tGather));
                // ---- pass 3: the per-event epilogue of a range of groups (parallel over the groups)
                const uint32_t Gc = mReg.groups();
                std::vector<LocalCounters> local(HostThreads());
                uint64_t epilogueNs = 0;
                auto epilogue = [&](size_t gBegin, size_t gCount) {
                    const auto t0 = Clock::now();
                    ParallelFor(gCount, 4, [&](size_t a, size_t b, unsigned tid) {
                        LocalCounters& c = local[tid];
                        for (size_t g = gBegin + a; g < gBegin + b; ++g)
                            EpilogueGroup(groups[g], plan[g].firstEv, sc, Gc, c);
                    });
                    epilogueNs += since(t0);
                };
                struct Ctx {
                    decltype(epilogue)* fn;
                } cbctx{&epilogue};
                const auto tEngine = Clock::now();
                Check(lc_regex_parse_packed_cb(
                          Engine(), mReg.get(), ngroups, spanPtr.data(), spanLen.data(), spanDst.data(),
                          spanFirst.data(), dst, off, len, nb, (uint32_t)mKeys.size(), status, capOff, capLen,
                          [](void* ctx, uint64_t first, uint64_t count) {
                              (*static_cast<Ctx*>(ctx)->fn)((size_t)first, (size_t)count);
                          },
                          &cbctx),
                      "lc_regex_parse_packed");
                mEngineNs.Add(since(tEngine) - epilogueNs); // the call's own share (the epilogue runs inside it)
                mEpilogueNs.Add(epilogueNs);
                for (const auto& c : local)
                    AddCounters(c);
                return;
            }
        }
        // nothing reached the engine (whole-line mode, or no event carries the source key): epilogue only
        const auto tEpilogue = Clock::now();
        ThreadScratch& sc = Scratch();
        std::vector<LocalCounters> local(HostThreads());
        ParallelFor(ngroups, 8, [&](size_t a, size_t b, unsigned tid) {
            for (size_t g = a; g < b; ++g)
                EpilogueGroup(groups[g], plan[g].firstEv, sc, mReg.groups(), local[tid]);
        });
        for (const auto& c : local)
            AddCounters(c);
        mEpilogueNs.Add(since(tEpilogue));
    } catch (const std::exception& ex) {
        EngineFailed(ex.what()); // groups not yet rewritten stay untouched (the reference never throws out of Process)
    }
}

// ------------------------------------------------------------------------------------------------ instance wrapper
void ProcessorInstance::Process(std::vector<PipelineEventGroup>& eventGroupList) {
    if (eventGroupList.empty())
        return;
    // the two DataSize() sweeps run on the host threads of the batch when the call carries many groups (the reference
    // spreads them over its ProcessorRunner threads, one group per call)
    auto sweep = [&](Counter& events, Counter& bytes) {
        ParallelFor(eventGroupList.size(), 64, [&](size_t a, size_t b, unsigned) {
            uint64_t ev = 0, by = 0;
            for (size_t g = a; g < b; ++g) {
                ev += eventGroupList[g].GetEvents().size();
                by += eventGroupList[g].DataSize();
            }
            events.Add(ev);
            bytes.Add(by);
        });
    };
    sweep(mInEventsTotal, mInSizeBytes);
    const auto before = std::chrono::steady_clock::now();
    mPlugin->Process(eventGroupList);
    const auto ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - before);
    mTotalProcessTimeNs.Add((uint64_t)ns.count());
    mTotalProcessTimeMs.Add((uint64_t)(ns.count() / 1000000));
    sweep(mOutEventsTotal, mOutSizeBytes);
}

// ------------------------------------------------------------------------------------------------ delimiter
const std::string ProcessorParseDelimiterNative::sName = "processor_parse_delimiter_native";

bool ProcessorParseDelimiterNative::Init(const Json::Value& config) {
    if (!GetString(config, "SourceKey", mSourceKey))
        return Fail("mandatory string param SourceKey is missing");
    if (!GetString(config, "Separator", mSeparator) || mSeparator.empty())
        return Fail("mandatory string param Separator is missing");
    if (mSeparator.size() > 4)
        return Fail("mandatory string param Separator has more than 4 chars");
    if (mSeparator == "\\t")
        mSeparator = "\t";
    std::string quote;
    bool hasQuote = GetString(config, "Quote", quote);
    if (mSeparator.size() == 1) {
        if (hasQuote && quote.size() > 1)
            return Fail("string param Quote is not a single char");
        if (hasQuote && !quote.empty())
            mQuote = quote[0];
    } // multi-char separator: a configured Quote is ignored (warning upstream)
    if (!GetKeys(config, mKeys))
        return Fail("mandatory list param Keys is missing");
    for (const auto& k : mKeys)
        if (k == mSourceKey)
            mSourceKeyOverwritten = true;
    GetBool(config, "AllowingShortenedFields", mAllowingShortenedFields);
    std::string t;
    if (GetString(config, "OverflowedFieldsTreatment", t)) {
        if (t == "keep")
            mOverflowedFieldsTreatment = OverflowedFieldsTreatment::KEEP;
        else if (t == "discard")
            mOverflowedFieldsTreatment = OverflowedFieldsTreatment::DISCARD;
    }
    mExtractingPartialFields = mOverflowedFieldsTreatment == OverflowedFieldsTreatment::DISCARD;
    return mCommonParserOptions.Init(config);
}

std::vector<std::pair<std::string, uint64_t>> ProcessorParseDelimiterNative::Counters() const {
    return {{"discarded", mDiscardedEventsTotal.GetValue()},
            {"out_failed", mOutFailedEventsTotal.GetValue()},
            {"out_key_not_found", mOutKeyNotFoundEventsTotal.GetValue()},
            {"out_successful", mOutSuccessfulEventsTotal.GetValue()}};
}

void ProcessorParseDelimiterNative::Process(PipelineEventGroup& group) {
    try {
        ProcessImpl(group);
    } catch (const std::exception& ex) {
        EngineFailed(ex.what()); // the reference's Process never throws: the group stays as it was
    }
}

void ProcessorParseDelimiterNative::ProcessImpl(PipelineEventGroup& group) {
    if (group.GetEvents().empty())
        return;
    EventsContainer& events = group.MutableEvents();
    FlatBatch batch;
    for (size_t i = 0; i < events.size(); ++i) {
        if (!IsSupportedEvent(events[i]))
            continue;
        const LogEvent& ev = events[i].Cast<LogEvent>();
        if (ev.HasContent(mSourceKey))
            batch.Add(i, ev.GetContent(mSourceKey));
    }
    batch.Finish(*group.GetSourceBuffer());
    const size_t nb = batch.eventIndex.size();
    const bool extend = mOverflowedFieldsTreatment == OverflowedFieldsTreatment::EXTEND;
    const bool useQuote = mSeparator.size() == 1 && mQuote != mSeparator[0];
    // Dense [n][MF] field tables with MF = keys + 16 serve every ordinary line.  A line with more columns than that
    // (log content is untrusted: one 256 KB line of separators must not size a table for the whole group) is parsed
    // again on its own, in small sub-batches whose tables hold exactly its columns -- memory stays O(bytes of those
    // lines), like the reference's per-line vectors (:246-248).
    const uint32_t MF = (uint32_t)mKeys.size() + 16;
    std::vector<uint8_t> status(nb);
    std::vector<uint32_t> nf(nb), fo((size_t)nb * MF), fl((size_t)nb * MF), fd((size_t)nb * MF);
    std::vector<uint32_t> wfo, wfl, wfd;
    std::vector<size_t> wideStart(nb, (size_t)-1);
    if (nb) {
        Check(lc_delim_parse(Engine(), batch.base, batch.baseLen, batch.off.data(), batch.len.data(), nb,
                             reinterpret_cast<const uint8_t*>(mSeparator.data()), (uint32_t)mSeparator.size(),
                             (uint8_t)mQuote, (uint32_t)mKeys.size(), extend, mAllowingShortenedFields, MF,
                             status.data(), nf.data(), fo.data(), fl.data(), fd.data()),
              "lc_delim_parse");
        std::vector<size_t> over;
        for (size_t i = 0; i < nb; ++i)
            if (nf[i] > MF && status[i] != LC_DELIM_PARSE_FAIL && status[i] != LC_DELIM_BLANK)
                over.push_back(i);
        const size_t kMaxEntries = 4u << 20; // 48 MB of tables per sub-batch at most (+ one oversize line alone)
        size_t at = 0;
        while (at < over.size()) {
            size_t cnt = 0;
            uint32_t mx = 0;
            while (at + cnt < over.size()) {
                uint32_t m2 = std::max(mx, nf[over[at + cnt]]);
                if (cnt && (cnt + 1) * (size_t)m2 > kMaxEntries)
                    break;
                mx = m2;
                ++cnt;
            }
            std::vector<uint32_t> so(cnt), sl(cnt), snf(cnt), sfo(cnt * (size_t)mx), sfl(cnt * (size_t)mx),
                sfd(cnt * (size_t)mx);
            std::vector<uint8_t> sst(cnt);
            for (size_t k = 0; k < cnt; ++k) {
                so[k] = batch.off[over[at + k]];
                sl[k] = batch.len[over[at + k]];
            }
            Check(lc_delim_parse(Engine(), batch.base, batch.baseLen, so.data(), sl.data(), cnt,
                                 reinterpret_cast<const uint8_t*>(mSeparator.data()), (uint32_t)mSeparator.size(),
                                 (uint8_t)mQuote, (uint32_t)mKeys.size(), extend, mAllowingShortenedFields, mx,
                                 sst.data(), snf.data(), sfo.data(), sfl.data(), sfd.data()),
                  "lc_delim_parse");
            for (size_t k = 0; k < cnt; ++k) {
                const size_t i = over[at + k];
                wideStart[i] = wfo.size();
                wfo.insert(wfo.end(), sfo.begin() + k * (size_t)mx, sfo.begin() + k * (size_t)mx + nf[i]);
                wfl.insert(wfl.end(), sfl.begin() + k * (size_t)mx, sfl.begin() + k * (size_t)mx + nf[i]);
                wfd.insert(wfd.end(), sfd.begin() + k * (size_t)mx, sfd.begin() + k * (size_t)mx + nf[i]);
            }
            at += cnt;
        }
    }

    size_t wIdx = 0, b = 0;
    std::vector<StringView> cols;
    for (size_t rIdx = 0; rIdx < events.size(); ++rIdx) {
        bool keep = true;
        PipelineEventPtr& e = events[rIdx];
        if (!IsSupportedEvent(e)) {
            mOutFailedEventsTotal.Add(1);
        } else {
            LogEvent& ev = e.Cast<LogEvent>();
            if (!ev.HasContent(mSourceKey)) {
                mOutKeyNotFoundEventsTotal.Add(1);
            } else {
                StringView buffer = ev.GetContent(mSourceKey);
                const uint8_t st = status[b];
                if (st == LC_DELIM_BLANK) {
                    // empty / blank value: out_failed++, event untouched (:220-242)
                    mOutFailedEventsTotal.Add(1);
                } else {
                    bool ok = st == LC_DELIM_OK;
                    if (ok) {
                        cols.clear();
                        SourceBuffer& sb = *group.GetSourceBuffer();
                        const bool wide = wideStart[b] != (size_t)-1;
                        const uint32_t* ro = wide ? wfo.data() + wideStart[b] : fo.data() + (size_t)b * MF;
                        const uint32_t* rl = wide ? wfl.data() + wideStart[b] : fl.data() + (size_t)b * MF;
                        const uint32_t* rd = wide ? wfd.data() + wideStart[b] : fd.data() + (size_t)b * MF;
                        for (uint32_t j = 0; j < nf[b]; ++j) {
                            uint32_t o = ro[j], l = rl[j], dq = rd[j];
                            StringView raw = batch.View(b, o, l);
                            if (useQuote && dq) {
                                // AddFieldWithUnQuote (:83-113): collapse doubled quotes into a fresh arena string
                                StringBuffer f = sb.AllocateStringBuffer(l - dq);
                                size_t w = 0;
                                for (size_t i = 0; i < raw.size(); ++i) {
                                    if (raw[i] == mQuote) {
                                        if (i + 1 < raw.size() && raw[i + 1] == mQuote) {
                                            f.data[w++] = mQuote;
                                            ++i;
                                        }
                                    } else {
                                        f.data[w++] = raw[i];
                                    }
                                }
                                cols.emplace_back(f.data, l - dq);
                            } else {
                                cols.push_back(raw);
                            }
                        }
                        if (useQuote && !extend && cols.size() > mKeys.size()) {
                            // overflow columns re-joined as sep + value each (:258-275)
                            size_t need = 0;
                            for (size_t j = mKeys.size(); j < cols.size(); ++j)
                                need += 1 + cols[j].size();
                            StringBuffer x = sb.AllocateStringBuffer(need);
                            char* p = x.data;
                            for (size_t j = mKeys.size(); j < cols.size(); ++j) {
                                *p++ = mSeparator[0];
                                memcpy(p, cols[j].data(), cols[j].size());
                                p += cols[j].size();
                            }
                            cols.resize(mKeys.size());
                            cols.emplace_back(x.data, need);
                        }
                        for (uint32_t idx = 0; idx < cols.size(); ++idx) {
                            if (idx < mKeys.size()) {
                                if (mExtractingPartialFields && mKeys[idx] == "_")
                                    continue;
                                AddLog(ev, mKeys[idx], cols[idx]);
                            } else {
                                if (mExtractingPartialFields)
                                    continue;
                                std::string key = "__column" + ToString(idx) + "__";
                                StringBuffer kb = sb.CopyString(key);
                                AddLog(ev, StringView(kb.data, kb.size), cols[idx]);
                            }
                        }
                        mOutSuccessfulEventsTotal.Add(1);
                    } else {
                        mOutFailedEventsTotal.Add(1);
                    }
                    if (!ok || !mSourceKeyOverwritten)
                        ev.DelContent(mSourceKey);
                    if (mCommonParserOptions.ShouldAddSourceContent(ok))
                        AddLog(ev, mCommonParserOptions.mRenamedSourceKey, buffer, false);
                    if (mCommonParserOptions.ShouldAddLegacyUnmatchedRawLog(ok))
                        AddLog(ev, CommonParserOptions::legacyUnmatchedRawLogKey, buffer, false);
                    if (mCommonParserOptions.ShouldEraseEvent(ok, ev, group.GetAllMetadata())) {
                        mDiscardedEventsTotal.Add(1);
                        keep = false;
                    }
                }
                ++b;
            }
        }
        if (keep) {
            if (wIdx != rIdx)
                events[wIdx] = std::move(events[rIdx]);
            ++wIdx;
        }
    }
    events.resize(wIdx);
}

// ------------------------------------------------------------------------------------------------ filter
const std::string ProcessorFilterNative::sName = "processor_filter_regex_native";

ProcessorFilterNative::~ProcessorFilterNative() {
    for (auto& l : mLeaves)
        delete l.reg;
}

static std::string Lower(std::string s) {
    for (auto& c : s)
        if (c >= 'A' && c <= 'Z')
            c = (char)(c + 32);
    return s;
}

// ParseExpressionFromJSON (:380-425): returns node index or -1
int ProcessorFilterNative::ParseExpression(const Json::Value& v, std::string& err) {
    if (!v.isObject())
        return -1;
    if (v["operator"].isString() && v["operands"].isArray()) {
        std::string op = Lower(v["operator"].asString());
        const Json::Value& ops = v["operands"];
        if (op == "not" && ops.size() == 1) {
            int c = ParseExpression(ops[(size_t)0], err);
            if (c < 0)
                return -1;
            Node n;
            n.op = 1;
            n.left = c;
            mNodes.push_back(n);
            return (int)mNodes.size() - 1;
        }
        if ((op == "and" || op == "or") && ops.size() == 2) {
            int l = ParseExpression(ops[(size_t)0], err);
            int r = ParseExpression(ops[(size_t)1], err);
            if (l < 0 || r < 0)
                return -1;
            Node n;
            n.op = op == "and" ? 2 : 3;
            n.left = l;
            n.right = r;
            mNodes.push_back(n);
            return (int)mNodes.size() - 1;
        }
        return -1;
    }
    if ((v["key"].isString() && v["exp"].isString()) || !v["type"].isString()) {
        if (Lower(v["type"].asString()) != "regex")
            return -1;
        Leaf leaf;
        leaf.key = v["key"].asString();
        leaf.reg = new CompiledRegex;
        if (!leaf.reg->Compile(v["exp"].asString(), err)) {
            delete leaf.reg;
            return -1;
        }
        mLeaves.push_back(leaf);
        Node n;
        n.op = 0;
        n.leaf = (int)mLeaves.size() - 1;
        mNodes.push_back(n);
        return (int)mNodes.size() - 1;
    }
    return -1;
}

bool ProcessorFilterNative::Init(const Json::Value& config) {
    std::string err;
    if (config.isMember("ConditionExp")) {
        if (!config["ConditionExp"].isObject())
            return Fail("object param ConditionExp is not of type object");
        mRoot = ParseExpression(config["ConditionExp"], err);
        if (mRoot < 0)
            return Fail("object param ConditionExp is not valid " + err);
        mFilterMode = Mode::EXPRESSION_MODE;
    }
    auto addRule = [&](const std::string& key, const std::string& pattern) -> bool {
        Leaf leaf;
        leaf.key = key;
        leaf.reg = new CompiledRegex;
        if (!leaf.reg->Compile(pattern, err)) {
            delete leaf.reg;
            return false;
        }
        mLeaves.push_back(leaf);
        return true;
    };
    if (mFilterMode == Mode::BYPASS_MODE && config.isMember("FilterKey") && config.isMember("FilterRegex")) {
        const Json::Value &ks = config["FilterKey"], &rs = config["FilterRegex"];
        if (!ks.isArray() || !rs.isArray() || ks.size() != rs.size())
            return Fail("param FilterKey and FilterRegex does not have the same size");
        for (size_t i = 0; i < ks.size(); ++i)
            if (!addRule(ks[i].asString(), rs[i].asString()))
                return Fail("value in list param FilterRegex is not a usable regex: " + err);
        if (ks.size())
            mFilterMode = Mode::RULE_MODE;
    }
    if (mFilterMode == Mode::BYPASS_MODE && config.isMember("Include") && config["Include"].isObject()) {
        for (const auto& k : config["Include"].getMemberNames())
            if (!addRule(k, config["Include"][k].asString()))
                return Fail("value in map param Include is not a usable regex: " + err);
        if (!mLeaves.empty())
            mFilterMode = Mode::RULE_MODE;
    }
    GetBool(config, "DiscardingNonUTF8", mDiscardingNonUTF8);
    return true;
}

bool ProcessorFilterNative::Eval(int node, const std::vector<std::vector<uint8_t>>& leafResult, size_t ev) const {
    const Node& n = mNodes[node];
    switch (n.op) {
        case 0:
            return leafResult[n.leaf][ev] != 0;
        case 1:
            return !Eval(n.left, leafResult, ev);
        case 2:
            return Eval(n.left, leafResult, ev) && Eval(n.right, leafResult, ev);
        default:
            return Eval(n.left, leafResult, ev) || Eval(n.right, leafResult, ev);
    }
}

// ProcessorFilterNative::noneUtf8 (:297-378): blank every byte that starts an invalid sequence; true if any
static bool BlankNoneUtf8(std::string& s, bool modify) {
    bool bad = false;
    size_t i = 0, n = s.size();
    auto cont = [&](size_t k) { return k < n && ((unsigned char)s[k] & 0xC0) == 0x80; };
    while (i < n) {
        unsigned char c = (unsigned char)s[i];
        size_t step = 1;
        bool inv = false;
        if ((c & 0x80) == 0) {
        } else if ((c & 0xE0) == 0xC0) {
            if (!cont(i + 1)) {
                inv = true;
            } else {
                uint32_t u = ((c & 0x1Fu) << 6) | ((unsigned char)s[i + 1] & 0x3Fu);
                inv = !(u >= 0x80 && u <= 0x7FF);
                step = 2;
            }
        } else if ((c & 0xF0) == 0xE0) {
            if (!cont(i + 1) || !cont(i + 2)) {
                inv = true;
            } else {
                uint32_t u = (((c & 0x0Fu) << 12) | (((unsigned char)s[i + 1] & 0x3Fu) << 6) |
                              ((unsigned char)s[i + 2] & 0x3Fu)) &
                             0xFFFFu;
                inv = !(u >= 0x800);
                step = 3;
            }
        } else if ((c & 0xF8) == 0xF0) {
            if (!cont(i + 1) || !cont(i + 2) || !cont(i + 3)) {
                inv = true;
            } else {
                uint32_t u = ((c & 0x07u) << 18) | (((unsigned char)s[i + 1] & 0x3Fu) << 12) |
                             (((unsigned char)s[i + 2] & 0x3Fu) << 6) | ((unsigned char)s[i + 3] & 0x3Fu);
                inv = !(u >= 0x10000 && u <= 0x10FFFF);
                step = 4;
            }
        } else {
            inv = true;
        }
        if (inv) {
            if (!modify)
                return true;
            s[i] = ' ';
            bad = true;
            i += 1;
        } else {
            i += step;
        }
    }
    return bad;
}

void ProcessorFilterNative::Process(PipelineEventGroup& group) {
    try {
        ProcessImpl(group);
    } catch (const std::exception& ex) {
        EngineFailed(ex.what()); // the reference's Process never throws: the group stays as it was
    }
}

void ProcessorFilterNative::ProcessImpl(PipelineEventGroup& group) {
    if (group.GetEvents().empty())
        return;
    EventsContainer& events = group.MutableEvents();
    const size_t ne = events.size();
    // one batched boolean regex_match per leaf over the events that carry its key
    std::vector<std::vector<uint8_t>> leafResult(mLeaves.size(), std::vector<uint8_t>(ne, 0));
    for (size_t li = 0; li < mLeaves.size(); ++li) {
        FlatBatch batch;
        for (size_t i = 0; i < ne; ++i) {
            if (!IsSupportedEvent(events[i]))
                continue;
            const LogEvent& ev = events[i].Cast<LogEvent>();
            if (ev.HasContent(mLeaves[li].key))
                batch.Add(i, ev.GetContent(mLeaves[li].key));
        }
        batch.Finish(*group.GetSourceBuffer());
        const size_t nb = batch.eventIndex.size();
        if (!nb)
            continue;
        std::vector<uint8_t> m(nb);
        Check(lc_regex_match(Engine(), mLeaves[li].reg->get(), batch.base, batch.baseLen, batch.off.data(),
                             batch.len.data(), nb, m.data()),
              "lc_regex_match");
        for (size_t b = 0; b < nb; ++b)
            leafResult[li][batch.eventIndex[b]] = m[b];
    }
    size_t wIdx = 0;
    for (size_t rIdx = 0; rIdx < ne; ++rIdx) {
        bool res = true;
        if (IsSupportedEvent(events[rIdx])) {
            LogEvent& ev = events[rIdx].Cast<LogEvent>();
            if (mFilterMode == Mode::EXPRESSION_MODE) {
                res = !ev.Empty() && Eval(mRoot, leafResult, rIdx);
            } else if (mFilterMode == Mode::RULE_MODE) {
                res = !ev.Empty();
                for (size_t li = 0; res && li < mLeaves.size(); ++li)
                    res = leafResult[li][rIdx] != 0; // a missing key left its slot at 0
            }
            if (res && mDiscardingNonUTF8) {
                std::vector<std::pair<StringView, StringView>> renamed;
                SourceBuffer& sb = *group.GetSourceBuffer();
                // contents are visited in place; keys needing repair are re-added after the walk (:190-211)
                std::vector<LogEvent::Content> snapshot = ev.RawContents();
                for (auto& c : snapshot) {
                    if (!c.second)
                        continue;
                    StringView key = c.first.first, val = c.first.second;
                    std::string v = val.to_string();
                    if (BlankNoneUtf8(v, true)) {
                        StringBuffer vb = sb.CopyString(v);
                        val = StringView(vb.data, vb.size);
                        ev.SetContentNoCopy(key, val);
                    }
                    std::string k = key.to_string();
                    if (BlankNoneUtf8(k, true)) {
                        StringBuffer kb = sb.CopyString(k);
                        renamed.emplace_back(StringView(kb.data, kb.size), val);
                        ev.DelContent(key);
                    }
                }
                for (auto& r : renamed)
                    ev.SetContentNoCopy(r.first, r.second);
            }
        }
        if (res) {
            if (wIdx != rIdx)
                events[wIdx] = std::move(events[rIdx]);
            ++wIdx;
        }
    }
    events.resize(wIdx);
}

// ------------------------------------------------------------------------------------------------ merge multiline
const std::string ProcessorMergeMultilineLogNative::sName = "processor_merge_multiline_log_native";
const std::string ProcessorMergeMultilineLogNative::PartLogFlag = "P";

bool ProcessorMergeMultilineLogNative::Init(const Json::Value& config) {
    GetString(config, "SourceKey", mSourceKey);
    std::string mergeType;
    if (!GetString(config, "MergeType", mergeType))
        return Fail("mandatory string param MergeType is missing");
    if (mergeType == "flag") {
        mMergeType = MergeType::BY_FLAG;
        return true;
    }
    if (mergeType != "regex")
        return Fail("string param MergeType is not valid");
    std::string err;
    if (!mMultiline.Init(config, err))
        return Fail(err);
    struct {
        const std::string* pattern;
        CompiledRegex* reg;
        bool* has;
    } regs[] = {{&mMultiline.mStartPattern, &mStartReg, &mHasStart},
                {&mMultiline.mContinuePattern, &mContinueReg, &mHasContinue},
                {&mMultiline.mEndPattern, &mEndReg, &mHasEnd}};
    for (auto& r : regs) {
        std::string p = *r.pattern;
        if (!p.empty() && EndsWith(p, "$"))
            p.pop_back();
        while (!p.empty() && EndsWith(p, ".*"))
            p.resize(p.size() - 2);
        if (p.empty())
            continue;
        if (!r.reg->Compile(p, err))
            return Fail("multiline pattern: " + err); // outside the automaton subset: never approximated
        *r.has = true;
    }
    if (!mHasStart && !mHasEnd && mHasContinue)
        mHasContinue = false;
    else if (mHasStart && mHasContinue && mHasEnd)
        mHasContinue = false;
    return true;
}

std::vector<std::pair<std::string, uint64_t>> ProcessorMergeMultilineLogNative::Counters() const {
    return {{"merged_events_total", mMergedEventsTotal.GetValue()},
            {"unmatched_events_total", mUnmatchedEventsTotal.GetValue()}};
}

void ProcessorMergeMultilineLogNative::Process(PipelineEventGroup& group) {
    try {
        ProcessImpl(group);
    } catch (const std::exception& ex) {
        EngineFailed(ex.what()); // the reference's Process never throws: the group stays as it was
    }
}

void ProcessorMergeMultilineLogNative::ProcessImpl(PipelineEventGroup& group) {
    if (group.GetEvents().empty())
        return;
    if (mMergeType == MergeType::BY_REGEX) {
        MergeLogsByRegex(group);
    } else if (group.HasMetadata(EventGroupMetaKey::HAS_PART_LOG)) {
        MergeLogsByFlag(group);
        group.DelMetadata(EventGroupMetaKey::HAS_PART_LOG);
    }
}

// :320-346.  The reference joins the values IN PLACE (it writes the line break over the byte that follows the target
// value and memmoves the next values down); that is only sound when the values lie in event order inside one
// arena chunk with nothing else in between, which is what the splitters produce.  Same here when that holds (values
// adjacent, at most one separator byte apart), else the joined value is built in a fresh arena allocation -- the
// resulting content is identical.
void ProcessorMergeMultilineLogNative::MergeEvents(PipelineEventGroup& group, std::vector<LogEvent*>& logEvents,
                                                   bool insertLineBreak) {
    if (logEvents.empty())
        return;
    mMergedEventsTotal.Add(logEvents.size());
    if (logEvents.size() == 1) {
        logEvents.clear();
        return;
    }
    LogEvent* target = logEvents[0];
    StringView targetValue = target->GetContent(mSourceKey);
    size_t total = targetValue.size();
    bool inPlace = true;
    const char* end = targetValue.data() + targetValue.size();
    for (size_t i = 1; i < logEvents.size(); ++i) {
        StringView cur = logEvents[i]->GetContent(mSourceKey);
        total += cur.size() + (insertLineBreak ? 1 : 0);
        const char* dst = end + (insertLineBreak ? 1 : 0);
        // adjacent values only (at most the one separator byte the splitter left between them): anything else
        // in the gap -- e.g. another content of the target event -- must not be overwritten
        if (cur.data() < dst || cur.data() > end + 1)
            inPlace = false;
        end = dst + cur.size();
    }
    size_t chunkSize = 0;
    SourceBuffer& sb = *group.GetSourceBuffer();
    if (inPlace && !sb.ChunkContaining(targetValue.data(), total, &chunkSize))
        inPlace = false;
    char* begin;
    if (inPlace) {
        begin = const_cast<char*>(targetValue.data());
    } else {
        StringBuffer b = sb.AllocateStringBuffer(total);
        begin = b.data;
        memcpy(begin, targetValue.data(), targetValue.size());
    }
    char* w = begin + targetValue.size();
    for (size_t i = 1; i < logEvents.size(); ++i) {
        if (insertLineBreak)
            *w++ = '\n';
        StringView cur = logEvents[i]->GetContent(mSourceKey);
        memmove(w, cur.data(), cur.size());
        w += cur.size();
    }
    // the key view must outlive the call: reuse the stored key of the target's content
    StringView key;
    for (auto& c : target->RawContents())
        if (c.second && c.first.first == StringView(mSourceKey))
            key = c.first.first;
    target->SetContentNoCopy(key, StringView(begin, (size_t)(w - begin)));
    logEvents.clear();
}

// :348-385 (alarms / log lines are not produced by this engine)
void ProcessorMergeMultilineLogNative::HandleUnmatchLogs(EventsContainer& logEvents, size_t& newSize, size_t begin,
                                                         size_t end) {
    mUnmatchedEventsTotal.Add(end - begin + 1);
    if (mMultiline.mUnmatchedContentTreatment == MultilineOptions::UnmatchedContentTreatment::DISCARD)
        return;
    for (size_t i = begin; i <= end; ++i)
        logEvents[newSize++] = std::move(logEvents[i]);
}

// :116-159.  A record is a run of parts: every part but the last carries the "P" content (the container runtime's
// partial-line marker); a part without it completes the run, and so does the end of the group.  The first part loses
// its marker, the joined value lands in the first part, and the output slot takes the event that FOLLOWS the previous
// record (an empty event there survives in place of the joined one -- the reference's behaviour, kept).
void ProcessorMergeMultilineLogNative::MergeLogsByFlag(PipelineEventGroup& group) {
    EventsContainer& all = group.MutableEvents();
    const size_t n = all.size();
    size_t kept = 0;  // events written back so far
    size_t head = 0;  // index right after the previous record
    bool open = false; // the previous part carried the marker
    std::vector<LogEvent*> parts;
    auto Complete = [&](size_t next) {
        MergeEvents(group, parts, false);
        all[kept++] = std::move(all[head]);
        head = next;
        open = false;
    };
    for (size_t cur = 0; cur < n; ++cur) {
        if (!IsSupportedEvent(all[cur])) {
            // ends the walk: the open run (or, without one, everything from here) is kept as it is
            if (parts.empty())
                head = cur;
            for (size_t i = head; i < n; ++i)
                all[kept++] = std::move(all[i]);
            all.resize(kept);
            return;
        }
        LogEvent* part = &all[cur].Cast<LogEvent>();
        if (part->Empty())
            continue;
        parts.push_back(part);
        const bool marked = part->HasContent(PartLogFlag);
        if (!marked) {
            Complete(cur + 1);
        } else if (!open) {
            part->DelContent(PartLogFlag);
            open = true;
        }
    }
    if (open)
        Complete(n);
    all.resize(kept);
}

// :161-318.  BoostRegexSearch(match_continuous) of every (event, pattern) pair is a pure function of the value, so all
// probes are evaluated up front on the GPU and the walk below only reads flags.
void ProcessorMergeMultilineLogNative::MergeLogsByRegex(PipelineEventGroup& group) {
    EventsContainer& sourceEvents = group.MutableEvents();
    const size_t ne = sourceEvents.size();
    std::vector<uint8_t> mS(ne, 0), mC(ne, 0), mE(ne, 0);
    {
        FlatBatch batch;
        for (size_t i = 0; i < ne; ++i) {
            if (!IsSupportedEvent(sourceEvents[i]))
                break; // the walk stops at the first unsupported event
            const LogEvent& ev = sourceEvents[i].Cast<LogEvent>();
            if (ev.Empty())
                continue;
            if (!ev.HasContent(mSourceKey))
                break;
            batch.Add(i, ev.GetContent(mSourceKey));
        }
        batch.Finish(*group.GetSourceBuffer());
        const size_t nb = batch.eventIndex.size();
        struct {
            bool has;
            CompiledRegex* reg;
            std::vector<uint8_t>* dst;
        } probes[] = {{mHasStart, &mStartReg, &mS}, {mHasContinue, &mContinueReg, &mC}, {mHasEnd, &mEndReg, &mE}};
        std::vector<uint8_t> m(nb);
        for (auto& p : probes) {
            if (!p.has || !nb)
                continue;
            Check(lc_regex_prefix_match(Engine(), p.reg->get(), batch.base, batch.baseLen, batch.off.data(),
                                        batch.len.data(), nb, m.data()),
                  "lc_regex_prefix_match");
            for (size_t b = 0; b < nb; ++b)
                (*p.dst)[batch.eventIndex[b]] = m[b];
        }
    }
    // The walk as a decision table: what a value does depends only on (which patterns exist, whether a record is
    // open, the three probe flags).  Decide() is that pure function; the loop below just executes its verdicts.
    enum class Act {
        OPEN,             // the value starts a record
        ALONE,            // continue + end mode, no record open, the value matches the end pattern: a record by itself
        UNMATCHED,        // no record open and nothing matches
        APPEND,           // joins the open record, which stays open
        APPEND_CLOSE,     // joins the open record and completes it
        APPEND_FAIL,      // joins the open record, which thereby fails as a whole (continue + end, end not matched)
        CLOSE_OPEN,       // the open record is complete WITHOUT this value, which starts the next one
        CLOSE_UNMATCHED,  // the open record is complete without this value, which matches nothing
    };
    const bool S = mHasStart, C = mHasContinue, E = mHasEnd;
    auto Decide = [S, C, E](bool open, bool mS, bool mC, bool mE) -> Act {
        if (!open) {
            if (S ? mS : mC)
                return Act::OPEN;
            return (!S && C && E && mE) ? Act::ALONE : Act::UNMATCHED;
        }
        if (C && mC)
            return Act::APPEND;
        if (E) {
            if (C)
                return mE ? Act::APPEND_CLOSE : Act::APPEND_FAIL;
            return mE ? Act::APPEND_CLOSE : Act::APPEND;
        }
        if (!C)
            return mS ? Act::CLOSE_OPEN : Act::APPEND;
        return mS ? Act::CLOSE_OPEN : Act::CLOSE_UNMATCHED;
    };
    size_t head = 0;  // index of the first event of the open record
    size_t kept = 0;  // events written back so far
    std::vector<LogEvent*> record;
    bool open = !S && !C && E; // with only an end pattern every value belongs to a record
    auto Complete = [&]() {   // the open record becomes one event (the head event carries the joined value)
        MergeEvents(group, record, true);
        sourceEvents[kept++] = std::move(sourceEvents[head]);
    };
    for (size_t cur = 0; cur < ne; ++cur) {
        LogEvent* ev = IsSupportedEvent(sourceEvents[cur]) ? &sourceEvents[cur].Cast<LogEvent>() : nullptr;
        if (ev && ev->Empty())
            continue;
        if (!ev || !ev->HasContent(mSourceKey)) {
            // an unsupported event or one without the source key ends the walk: everything from the head of the
            // open record (or from here) is kept as it is
            if (record.empty())
                head = cur;
            for (size_t i = head; i < ne; ++i)
                sourceEvents[kept++] = std::move(sourceEvents[i]);
            sourceEvents.resize(kept);
            return;
        }
        switch (Decide(open, mS[cur] != 0, mC[cur] != 0, mE[cur] != 0)) {
            case Act::OPEN:
                record.push_back(ev);
                head = cur;
                open = true;
                break;
            case Act::ALONE:
                mMergedEventsTotal.Add(1);
                sourceEvents[kept++] = std::move(sourceEvents[cur]);
                break;
            case Act::UNMATCHED:
                HandleUnmatchLogs(sourceEvents, kept, cur, cur);
                break;
            case Act::APPEND:
                record.push_back(ev);
                break;
            case Act::APPEND_CLOSE:
                record.push_back(ev);
                Complete();
                if (S || C)
                    open = false;
                else
                    head = cur + 1; // only an end pattern: the next record starts right away
                break;
            case Act::APPEND_FAIL:
                record.clear();
                HandleUnmatchLogs(sourceEvents, kept, head, cur);
                open = false;
                break;
            case Act::CLOSE_OPEN:
                Complete();
                head = cur;
                record.push_back(ev);
                break;
            case Act::CLOSE_UNMATCHED:
                Complete();
                HandleUnmatchLogs(sourceEvents, kept, cur, cur);
                open = false;
                break;
        }
    }
    // a record still open at the end of the group: complete when there is no end pattern to wait for, else unmatched
    if (open && head < ne) {
        if (!E)
            Complete();
        else
            HandleUnmatchLogs(sourceEvents, kept, head, ne - 1);
    }
    sourceEvents.resize(kept);
}

// ------------------------------------------------------------------------------------------------ SLS serialise
namespace {
void PutVarint(std::string& out, uint32_t v) {
    while (v >= 0x80u) {
        out.push_back((char)(v | 0x80u));
        v >>= 7;
    }
    out.push_back((char)v);
}
void PutString(std::string& out, StringView s) {
    PutVarint(out, (uint32_t)s.size());
    out.append(s.data(), s.size());
}
} // namespace

bool SLSEventGroupSerializer::Serialize(PipelineEventGroup& group, std::string& res, std::string& errorMsg) const {
    const EventsContainer& events = group.GetEvents();
    if (events.empty()) {
        errorMsg = "empty event group";
        return false;
    }
    if (!events[0].Is<LogEvent>()) {
        errorMsg = "unsupported event type in event group"; // metric / span / raw groups: not built yet
        return false;
    }
    // flatten the contents of every event into entry tables over the arena
    const size_t n = events.size();
    std::vector<uint32_t> evTime(n), evNs(n, LC_SLS_NO_NS);
    std::vector<uint64_t> entBegin(n + 1, 0);
    std::vector<const char*> kPtr, vPtr;
    std::vector<uint32_t> kLen, vLen;
    for (size_t i = 0; i < n; ++i) {
        const LogEvent& e = events[i].Cast<LogEvent>();
        evTime[i] = (uint32_t)e.GetTimestamp();
        if (mEnableTimestampNanosecond && e.GetTimestampNanosecond())
            evNs[i] = e.GetTimestampNanosecond().value();
        for (auto& c : e.RawContents()) {
            if (!c.second)
                continue;
            kPtr.push_back(c.first.first.data());
            kLen.push_back((uint32_t)c.first.first.size());
            vPtr.push_back(c.first.second.data());
            vLen.push_back((uint32_t)c.first.second.size());
        }
        entBegin[i + 1] = kPtr.size();
    }
    const size_t m = kPtr.size();
    if (m == 0) {
        errorMsg = "all empty logs";
        return false;
    }
    // one span of the arena if all keys and values live in the same chunk, else a packed copy
    const char* lo = kPtr[0];
    const char* hi = kPtr[0] + kLen[0];
    for (size_t k = 0; k < m; ++k) {
        lo = std::min(lo, std::min(kPtr[k], vPtr[k]));
        hi = std::max(hi, std::max(kPtr[k] + kLen[k], vPtr[k] + vLen[k]));
    }
    std::vector<uint32_t> kOff(m), vOff(m);
    std::vector<uint8_t> packed;
    const uint8_t* base;
    uint64_t baseLen;
    size_t chunkSize = 0;
    if ((uint64_t)(hi - lo) < 0xFFFFFFF0ull && group.GetSourceBuffer()->ChunkContaining(lo, (size_t)(hi - lo), &chunkSize)) {
        base = reinterpret_cast<const uint8_t*>(lo);
        baseLen = (uint64_t)(hi - lo);
        for (size_t k = 0; k < m; ++k) {
            kOff[k] = (uint32_t)(kPtr[k] - lo);
            vOff[k] = (uint32_t)(vPtr[k] - lo);
        }
    } else {
        size_t total = 0;
        for (size_t k = 0; k < m; ++k)
            total += (size_t)kLen[k] + vLen[k];
        packed.resize(total + 1);
        size_t at = 0;
        for (size_t k = 0; k < m; ++k) {
            kOff[k] = (uint32_t)at;
            memcpy(packed.data() + at, kPtr[k], kLen[k]);
            at += kLen[k];
            vOff[k] = (uint32_t)at;
            memcpy(packed.data() + at, vPtr[k], vLen[k]);
            at += vLen[k];
        }
        base = packed.data();
        baseLen = total;
    }
    // group-level fields in tag (map) order, :203-213,239-249
    std::string tail;
    for (auto& tag : group.GetTags()) {
        if (tag.first == StringView("__topic__")) {
            tail.push_back(0x1A);
            PutString(tail, tag.second);
        } else if (tag.first == StringView("__source__")) {
            tail.push_back(0x22);
            PutString(tail, tag.second);
        } else if (tag.first == StringView("__machine_uuid__")) {
            tail.push_back(0x2A);
            PutString(tail, tag.second);
        } else {
            std::string inner;
            inner.push_back(0x0A);
            PutString(inner, tag.first);
            inner.push_back(0x12);
            PutString(inner, tag.second);
            tail.push_back(0x32);
            PutVarint(tail, (uint32_t)inner.size());
            tail += inner;
        }
    }
    uint64_t need = 0;
    int rc = lc_sls_serialize_logs(Engine(), base, baseLen, n, evTime.data(), evNs.data(), entBegin.data(), kOff.data(),
                                   kLen.data(), vOff.data(), vLen.data(), nullptr, 0, &need);
    if (rc != LC_OK && rc != LC_ERR_CAPACITY)
        Check(rc, "lc_sls_serialize_logs");
    if ((int64_t)(need + tail.size()) > (int64_t)mMaxSendLogGroupSize) {
        errorMsg = "log group exceeds size limit\tgroup size: " + ToString(need + tail.size())
                   + "\tsize limit: " + ToString((uint64_t)mMaxSendLogGroupSize);
        return false;
    }
    res.resize(need);
    Check(lc_sls_serialize_logs(Engine(), base, baseLen, n, evTime.data(), evNs.data(), entBegin.data(), kOff.data(),
                                kLen.data(), vOff.data(), vLen.data(), reinterpret_cast<uint8_t*>(&res[0]), need, &need),
          "lc_sls_serialize_logs");
    res += tail;
    return true;
}

Processor* CreateProcessor(const std::string& type) {
    if (type == ProcessorMergeMultilineLogNative::sName)
        return new ProcessorMergeMultilineLogNative;
    if (type == ProcessorFilterNative::sName)
        return new ProcessorFilterNative;
    if (type == ProcessorSplitLogStringNative::sName)
        return new ProcessorSplitLogStringNative;
    if (type == ProcessorSplitMultilineLogStringNative::sName)
        return new ProcessorSplitMultilineLogStringNative;
    if (type == ProcessorParseRegexNative::sName)
        return new ProcessorParseRegexNative;
    if (type == ProcessorParseDelimiterNative::sName)
        return new ProcessorParseDelimiterNative;
    return nullptr;
}

} // namespace logtail
