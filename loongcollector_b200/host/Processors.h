// Processors.h -- B200-backed replacements of LoongCollector's four native log-parsing processors behind the
// reference's own plugin API: same class names, Init(const Json::Value&) parameters, Process(PipelineEventGroup&)
// side effects and counters.  The per-byte work (newline scan, regex automata, delimiter FSM) happens on the
// GPU through include/lc_b200.h; this file keeps only the host-side policy that needs the event object graph.
//   Processor interface ............ core/collection_pipeline/plugin/interface/Processor.h:27-37
//   CommonParserOptions ............ core/plugin/processor/CommonParserOptions.cpp:28-117
//   MultilineOptions ............... core/file_server/MultilineOptions.cpp:22-222
//   processors ..................... core/plugin/processor/{ProcessorParseRegexNative,ProcessorParseDelimiterNative}.cpp,
//                                    core/plugin/processor/inner/{ProcessorSplitLogStringNative,ProcessorSplitMultilineLogStringNative}.cpp
#pragma once
#include <atomic>
#include <chrono>
#include <memory>
#include <string>
#include <vector>

#include "../../include/lc_b200.h"
#include "Models.h"

namespace logtail {

struct ThreadScratch; // pinned per-thread tables of the batched paths (Processors.cpp)

// Process() runs concurrently on the same instance from every ProcessorRunner thread (ProcessorRunner.cpp:48-53):
// counters are atomic like the reference's (monitor/metric_models/MetricTypes.h, relaxed adds).
struct Counter {
    std::atomic<uint64_t> v{0};
    uint64_t GetValue() const { return v.load(std::memory_order_relaxed); }
    void Add(uint64_t d) { v.fetch_add(d, std::memory_order_relaxed); }
};

class Processor {
public:
    virtual ~Processor() = default;
    virtual const std::string& Name() const = 0;
    virtual bool Init(const Json::Value& config) = 0;
    virtual void Process(std::vector<PipelineEventGroup>& groups) {
        for (auto& g : groups)
            Process(g);
    }
    virtual void Process(PipelineEventGroup& group) = 0;
    const std::string& LastError() const { return mError; }
    // The reference's Process never throws (per-event failure = counters + policy): an engine failure (CUDA error,
    // out of memory) leaves the affected groups untouched, is counted here and its text kept in LastError().
    uint64_t EngineErrors() const { return mEngineErrors.GetValue(); }
    // name -> value of every counter the reference registers for this plugin
    virtual std::vector<std::pair<std::string, uint64_t>> Counters() const { return {}; }

protected:
    virtual bool IsSupportedEvent(const PipelineEventPtr& e) const = 0;
    bool Fail(const std::string& msg) {
        mError = msg;
        return false;
    }
    void EngineFailed(const char* what) {
        mEngineErrors.Add(1);
        mError = what;
    }
    std::string mError;
    Counter mEngineErrors;
};

// ProcessorInstance (core/collection_pipeline/plugin/instance/ProcessorInstance.cpp:29-63): the wrapper the pipeline
// calls -- in / out event and byte counters (two DataSize() sweeps per call) and the wall time spent in the plugin.
class ProcessorInstance {
public:
    explicit ProcessorInstance(Processor* plugin) : mPlugin(plugin) {}
    const std::string& Name() const { return mPlugin->Name(); }
    Processor* GetPlugin() { return mPlugin.get(); }
    bool Init(const Json::Value& config) { return mPlugin->Init(config); }
    void Process(std::vector<PipelineEventGroup>& eventGroupList);
    Counter mInEventsTotal, mOutEventsTotal, mInSizeBytes, mOutSizeBytes, mTotalProcessTimeMs;
    // finer-grained than the reference's millisecond counter (kept in addition to it)
    Counter mTotalProcessTimeNs;

private:
    std::unique_ptr<Processor> mPlugin;
};

struct CommonParserOptions {
    bool mKeepingSourceWhenParseFail = false;
    bool mKeepingSourceWhenParseSucceed = false;
    std::string mRenamedSourceKey;
    bool mCopingRawLog = false;
    static const std::string legacyUnmatchedRawLogKey;
    bool Init(const Json::Value& config);
    bool ShouldAddSourceContent(bool parseSuccess) const;
    bool ShouldAddLegacyUnmatchedRawLog(bool parseSuccess) const;
    bool ShouldEraseEvent(bool parseSuccess, const LogEvent& sourceEvent, const GroupMetadata& metadata) const;
};

struct MultilineOptions {
    enum class UnmatchedContentTreatment { DISCARD, SINGLE_LINE };
    std::string mStartPattern, mContinuePattern, mEndPattern;
    UnmatchedContentTreatment mUnmatchedContentTreatment = UnmatchedContentTreatment::SINGLE_LINE;
    bool mIgnoringUnmatchWarning = false;
    bool mIsMultiline = false;
    bool Init(const Json::Value& config, std::string& err);
};

class CompiledRegex {
public:
    CompiledRegex() = default;
    ~CompiledRegex();
    CompiledRegex(const CompiledRegex&) = delete;
    CompiledRegex& operator=(const CompiledRegex&) = delete;
    bool Compile(const std::string& pattern, std::string& err);
    lc_regex_t* get() const { return mRe; }
    uint32_t groups() const { return mRe ? lc_regex_ngroups(mRe) : 0; }

private:
    lc_regex_t* mRe = nullptr;
};

class ProcessorSplitLogStringNative : public Processor {
public:
    static const std::string sName;
    const std::string& Name() const override { return sName; }
    bool Init(const Json::Value& config) override;
    void Process(PipelineEventGroup& group) override;
    void ProcessImpl(PipelineEventGroup& group);
    using Processor::Process;
    std::string mSourceKey = "content";
    char mSplitChar = '\n';
    bool mEnableRawContent = false;

protected:
    bool IsSupportedEvent(const PipelineEventPtr& e) const override { return e.Is<LogEvent>(); }
};

class ProcessorSplitMultilineLogStringNative : public Processor {
public:
    static const std::string sName;
    const std::string& Name() const override { return sName; }
    bool Init(const Json::Value& config) override;
    void Process(PipelineEventGroup& group) override;
    void ProcessImpl(PipelineEventGroup& group);
    using Processor::Process;
    std::vector<std::pair<std::string, uint64_t>> Counters() const override;
    std::string mSourceKey = "content";
    MultilineOptions mMultiline;
    bool mEnableRawContent = false;
    Counter mMatchedEventsTotal, mMatchedLinesTotal, mUnmatchedLinesTotal;

protected:
    bool IsSupportedEvent(const PipelineEventPtr& e) const override { return e.Is<LogEvent>(); }

private:
    CompiledRegex mStart, mContinue, mEnd;
};

class ProcessorParseRegexNative : public Processor {
public:
    static const std::string sName;
    const std::string& Name() const override { return sName; }
    bool Init(const Json::Value& config) override;
    void Process(PipelineEventGroup& group) override;
    // Batched override of Processor.h:31: the arenas of all groups are packed into one device arena and parsed by
    // one launch sequence (lc_regex_parse_packed); gather and per-event epilogue run on a few host threads.
    void Process(std::vector<PipelineEventGroup>& groups) override;
    std::vector<std::pair<std::string, uint64_t>> Counters() const override;
    std::string mSourceKey, mRegex;
    std::vector<std::string> mKeys;
    CommonParserOptions mCommonParserOptions;
    Counter mDiscardedEventsTotal, mOutFailedEventsTotal, mOutKeyNotFoundEventsTotal, mOutSuccessfulEventsTotal;

protected:
    bool IsSupportedEvent(const PipelineEventPtr& e) const override { return e.Is<LogEvent>(); }

private:
    struct EventResult {
        uint8_t status;
        const uint32_t* capOff; // row of the capture tables (offsets relative to `origin - originOff`)
        const uint32_t* capLen;
        const char* origin;     // first byte of the source value
        uint32_t originOff;     // its offset in the table's coordinate system
    };
    struct LocalCounters {
        uint64_t discarded = 0, failed = 0, keyNotFound = 0, successful = 0;
    };
    // ProcessEvent epilogue (:135-167) of one event; returns false when the event is to be erased
    bool FinishEvent(PipelineEventGroup& group, PipelineEventPtr& e, const LogEvent::Content* src, const EventResult* r,
                     LocalCounters& c) const;
    void ProcessBatch(PipelineEventGroup* groups, size_t ngroups);
    void EpilogueGroup(PipelineEventGroup& group, uint64_t firstEv, const struct ThreadScratch& sc, uint32_t G,
                       LocalCounters& c) const;
    void AddCounters(const LocalCounters& c);
    bool mSourceKeyOverwritten = false;
    bool mIsWholeLineMode = false;
    bool mKeysDistinct = false; // no key repeats: an event that holds only the source key takes the parsed fields as
                                // plain appends (AppendContentNoCopy) instead of one look-up per key
    CompiledRegex mReg;

public:
    // wall time of the three phases of the batched path, summed over calls (ns): gather, engine call, epilogue
    Counter mGatherNs, mEngineNs, mEpilogueNs;
};

class ProcessorParseDelimiterNative : public Processor {
public:
    enum class OverflowedFieldsTreatment { EXTEND, KEEP, DISCARD };
    static const std::string sName;
    const std::string& Name() const override { return sName; }
    bool Init(const Json::Value& config) override;
    void Process(PipelineEventGroup& group) override;
    void ProcessImpl(PipelineEventGroup& group);
    using Processor::Process;
    std::vector<std::pair<std::string, uint64_t>> Counters() const override;
    std::string mSourceKey, mSeparator;
    char mQuote = '"';
    std::vector<std::string> mKeys;
    bool mAllowingShortenedFields = true;
    OverflowedFieldsTreatment mOverflowedFieldsTreatment = OverflowedFieldsTreatment::EXTEND;
    bool mExtractingPartialFields = false;
    CommonParserOptions mCommonParserOptions;
    Counter mDiscardedEventsTotal, mOutFailedEventsTotal, mOutKeyNotFoundEventsTotal, mOutSuccessfulEventsTotal;

protected:
    bool IsSupportedEvent(const PipelineEventPtr& e) const override { return e.Is<LogEvent>(); }

private:
    bool mSourceKeyOverwritten = false;
};

// First "next" row: regex include / exclude filter.  Every regex leaf of the rule / expression is
// evaluated for the whole group with one batched boolean regex_match on the GPU; the and/or/not tree and the
// non-UTF8 blanking stay on the host.  core/plugin/processor/ProcessorFilterNative.cpp:30-275,380-488
class ProcessorFilterNative : public Processor {
public:
    static const std::string sName;
    const std::string& Name() const override { return sName; }
    bool Init(const Json::Value& config) override;
    void Process(PipelineEventGroup& group) override;
    void ProcessImpl(PipelineEventGroup& group);
    using Processor::Process;
    bool mDiscardingNonUTF8 = false;
    ~ProcessorFilterNative() override;

protected:
    bool IsSupportedEvent(const PipelineEventPtr& e) const override { return e.Is<LogEvent>(); }

private:
    enum class Mode { BYPASS_MODE, EXPRESSION_MODE, RULE_MODE };
    struct Leaf {
        std::string key;
        CompiledRegex* reg;
    };
    struct Node { // expression tree; leaf >= 0 indexes mLeaves
        int op = 0; // 0 leaf, 1 not, 2 and, 3 or
        int leaf = -1;
        int left = -1, right = -1;
    };
    int ParseExpression(const Json::Value& v, std::string& err);
    bool Eval(int node, const std::vector<std::vector<uint8_t>>& leafResult, size_t ev) const;
    Mode mFilterMode = Mode::BYPASS_MODE;
    std::vector<Leaf> mLeaves;
    std::vector<Node> mNodes;
    int mRoot = -1;
};

// Second "next" row: merges already-split LogEvents of a group back into records -- by the docker
// partial-log flag or by start / continue / end patterns.  The anchored prefix probes of every event are evaluated for
// the whole group with one batched lc_regex_prefix_match per pattern; the sequential merge walk (which needs the
// event objects and joins the values in place in the arena) stays on the host.
// core/plugin/processor/inner/ProcessorMergeMultilineLogNative.cpp:33-420
class ProcessorMergeMultilineLogNative : public Processor {
public:
    enum class MergeType { BY_REGEX, BY_FLAG };
    static const std::string sName;
    static const std::string PartLogFlag;
    const std::string& Name() const override { return sName; }
    bool Init(const Json::Value& config) override;
    void Process(PipelineEventGroup& group) override;
    void ProcessImpl(PipelineEventGroup& group);
    using Processor::Process;
    std::vector<std::pair<std::string, uint64_t>> Counters() const override;
    std::string mSourceKey = "content";
    MergeType mMergeType = MergeType::BY_REGEX;
    MultilineOptions mMultiline;
    Counter mMergedEventsTotal, mUnmatchedEventsTotal;

protected:
    bool IsSupportedEvent(const PipelineEventPtr& e) const override { return e.Is<LogEvent>(); }

private:
    void MergeLogsByFlag(PipelineEventGroup& group);
    void MergeLogsByRegex(PipelineEventGroup& group);
    void MergeEvents(PipelineEventGroup& group, std::vector<LogEvent*>& logEvents, bool insertLineBreak);
    void HandleUnmatchLogs(EventsContainer& logEvents, size_t& newSize, size_t begin, size_t end);
    // the *RegPtr members of MultilineOptions (MultilineOptions.cpp:125-160,196-215): compiled from the pattern
    // with its trailing '$' / ".*" removed; unset when that is empty or dropped by the combination rules
    CompiledRegex mStartReg, mContinueReg, mEndReg;
    bool mHasStart = false, mHasContinue = false, mHasEnd = false;
};

// Fourth "next" row: the SLS wire format of a group of LOG events.  Mirrors
// SLSEventGroupSerializer::Serialize (core/collection_pipeline/serializer/SLSSerializer.cpp:162-252): the `Logs`
// fields -- all the bytes that scale with the data -- are written on the GPU straight from the arena spans of the
// events' contents (lc_sls_serialize_logs); Topic / Source / MachineUUID / LogTags are appended here.
class SLSEventGroupSerializer {
public:
    bool mEnableTimestampNanosecond = false;      // GlobalConfig::mEnableTimestampNanosecond
    int32_t mMaxSendLogGroupSize = 10 * 1024 * 1024; // flag max_send_log_group_size (FlusherSLS.cpp:61)
    bool Serialize(PipelineEventGroup& group, std::string& res, std::string& errorMsg) const;
};

// Factory by plugin type name (the names the reference registers, PluginRegistry.cpp:183-200).
Processor* CreateProcessor(const std::string& type);

} // namespace logtail
