// Models.h -- host-side event model kept by the engine's C++ host, mirroring the names and observable
// behaviour of the reference's L5 data model so that a B200-backed processor is a drop-in:
//   StringView / SourceBuffer     core/common/StringView.h, core/common/memory/SourceBuffer.h:98-181
//   LogEvent / RawEvent           core/models/LogEvent.{h,cpp} (contents with tombstones, :50-106), RawEvent.cpp
//   PipelineEventPtr / Group      core/models/PipelineEventPtr.h:32-96, PipelineEventGroup.{h,cpp}
// Independent implementation (std::string_view instead of boost::string_view, no event pool): only the
// behaviour the four processors and their unit-test fixtures observe is reproduced.
#pragma once
#include <stdint.h>
#include <string.h>

#include <map>
#include <memory>
#include <optional>
#include <string>
#include <string_view>
#include <vector>

#include "Json.h"

namespace logtail {

class StringView : public std::string_view {
public:
    using std::string_view::string_view;
    StringView() = default;
    StringView(const std::string_view& v) : std::string_view(v) {}
    StringView(const std::string& s) : std::string_view(s) {}
    std::string to_string() const { return std::string(data(), size()); }
};

struct StringBuffer {
    char* data = nullptr;
    size_t size = 0;
    size_t capacity = 0;
};

// Chunked bump arena.  Chunks double from 4 KiB up to 128 KiB; a request of at least half the current chunk
// size gets its own allocation (same growth rule as the reference so that a file read lands in ONE chunk).
// Every string buffer is NUL-terminated one past its end.
// Chunks come from operator new[] by default; a process that hands its arenas to the GPU engine installs the
// engine's pinned allocator (lc_host_alloc / lc_host_free) with SetChunkAllocator so that a group's bytes are
// DMA-able in place -- the "SourceBuffer zero-copy arena" stays the only copy of the log bytes on the host.
class SourceBuffer {
public:
    using AllocFn = void* (*)(size_t);
    using FreeFn = void (*)(void*);
    static void SetChunkAllocator(AllocFn a, FreeFn f) {
        sAlloc = a;
        sFree = f;
    }
    SourceBuffer() = default;
    SourceBuffer(const SourceBuffer&) = delete;
    SourceBuffer& operator=(const SourceBuffer&) = delete;

    StringBuffer AllocateStringBuffer(size_t size) {
        char* p = Allocate(size + 1);
        p[size] = '\0';
        StringBuffer b;
        b.data = p;
        b.size = size;
        b.capacity = size + 1;
        return b;
    }
    StringBuffer CopyString(const char* data, size_t len) {
        StringBuffer b = AllocateStringBuffer(len);
        if (len)
            memcpy(b.data, data, len);
        return b;
    }
    StringBuffer CopyString(StringView s) { return CopyString(s.data(), s.size()); }
    StringBuffer CopyString(const std::string& s) { return CopyString(s.data(), s.size()); }

    // chunk that contains [p, p + n) entirely, or nullptr (used to hand a zero-copy span to the engine)
    const char* ChunkContaining(const char* p, size_t n, size_t* chunk_size) const {
        for (auto& c : mChunks)
            if (p >= c.mem.get() && p + n <= c.mem.get() + c.used) {
                *chunk_size = c.used;
                return c.mem.get();
            }
        return nullptr;
    }

private:
    struct ChunkFree {
        FreeFn fn; // allocator that was active when the chunk was made (nullptr: operator new[])
        ChunkFree() : fn(nullptr) {}
        explicit ChunkFree(FreeFn f) : fn(f) {}
        void operator()(char* p) const {
            if (fn)
                fn(p);
            else
                delete[] p;
        }
    };
    struct Chunk {
        std::unique_ptr<char[], ChunkFree> mem;
        size_t cap = 0, used = 0;
    };
    static std::unique_ptr<char[], ChunkFree> NewChunk(size_t bytes) {
        if (sAlloc && sFree) {
            if (void* p = sAlloc(bytes))
                return std::unique_ptr<char[], ChunkFree>(static_cast<char*>(p), ChunkFree(sFree));
        }
        return std::unique_ptr<char[], ChunkFree>(new char[bytes], ChunkFree());
    }
    inline static AllocFn sAlloc = nullptr;
    inline static FreeFn sFree = nullptr;
    char* Allocate(size_t bytes) {
        if (bytes * 2 >= mNextChunk) { // oversize: own allocation
            Chunk c;
            c.mem = NewChunk(bytes);
            c.cap = c.used = bytes;
            mChunks.insert(mChunks.begin(), std::move(c)); // keep the current bump chunk last
            return mChunks.front().mem.get();
        }
        if (mChunks.empty() || mChunks.back().used + bytes > mChunks.back().cap || mChunks.back().cap != mCurCap) {
            Chunk c;
            c.mem = NewChunk(mNextChunk);
            c.cap = mNextChunk;
            mCurCap = mNextChunk;
            mChunks.push_back(std::move(c));
            if (mNextChunk < 128 * 1024)
                mNextChunk *= 2;
        }
        Chunk& c = mChunks.back();
        char* p = c.mem.get() + c.used;
        c.used += bytes;
        return p;
    }
    std::vector<Chunk> mChunks;
    size_t mNextChunk = 4096, mCurCap = 0;
};

class PipelineEventGroup;

class PipelineEvent {
public:
    enum class Type { NONE = 0, LOG = 1, METRIC = 2, SPAN = 3, RAW = 4 };
    virtual ~PipelineEvent() = default;
    Type GetType() const { return mType; }
    time_t GetTimestamp() const { return mTimestamp; }
    std::optional<uint32_t> GetTimestampNanosecond() const { return mTimestampNanosecond; }
    void SetTimestamp(time_t t) { mTimestamp = t; }
    void SetTimestamp(time_t t, std::optional<uint32_t> ns) {
        mTimestamp = t;
        mTimestampNanosecond = ns;
    }
    std::shared_ptr<SourceBuffer>& GetSourceBuffer();
    virtual Json::Value ToJson(bool enableEventMeta) const = 0;
    // PipelineEvent.h:62 -- what ProcessorInstance's in / out byte counters add up
    virtual size_t DataSize() const { return sizeof(mTimestamp) + sizeof(mTimestampNanosecond); }

protected:
    PipelineEvent(Type t, PipelineEventGroup* g) : mType(t), mGroup(g) {}
    Type mType;
    time_t mTimestamp = 0;
    std::optional<uint32_t> mTimestampNanosecond;
    PipelineEventGroup* mGroup;
};

class LogEvent : public PipelineEvent {
public:
    explicit LogEvent(PipelineEventGroup* g) : PipelineEvent(Type::LOG, g) { mContents.reserve(16); }
    using Content = std::pair<std::pair<StringView, StringView>, bool>; // ((key, value), live)

    StringView GetContent(StringView key) const;
    bool HasContent(StringView key) const;
    void SetContent(StringView key, StringView val); // copies both into the arena
    void SetContentNoCopy(StringView key, StringView val);
    void SetContentNoCopy(const StringBuffer& key, const StringBuffer& val) {
        SetContentNoCopy(StringView(key.data, key.size), StringView(val.data, val.size));
    }
    // LogEvent.cpp:159-163: no look-up of an existing key (the caller knows the key is not present)
    void AppendContentNoCopy(StringView key, StringView val) {
        ++mContentCnt;
        mAllocatedContentSize += key.size() + val.size();
        mContents.emplace_back(std::make_pair(key, val), true);
    }
    // one scan instead of HasContent + GetContent: the newest live entry with this key, or nullptr
    const Content* FindContent(StringView key) const {
        for (auto it = mContents.rbegin(); it != mContents.rend(); ++it)
            if (it->second && it->first.first == key)
                return &*it;
        return nullptr;
    }
    void DelContent(StringView key);
    bool Empty() const { return mContentCnt == 0; }
    size_t Size() const { return mContentCnt; }
    const std::vector<Content>& RawContents() const { return mContents; }
    // first live content (LogEvent::cbegin() skips tombstones)
    const Content* FirstLive() const {
        for (auto& c : mContents)
            if (c.second)
                return &c;
        return nullptr;
    }
    void SetPosition(uint64_t offset, uint64_t size) {
        mFileOffset = offset;
        mRawSize = size;
    }
    std::pair<uint64_t, uint64_t> GetPosition() const { return {mFileOffset, mRawSize}; }
    Json::Value ToJson(bool enableEventMeta) const override;
    bool FromJson(const Json::Value& root);
    size_t DataSize() const override;

private:
    std::vector<Content> mContents;
    size_t mContentCnt = 0;
    size_t mAllocatedContentSize = 0;
    uint64_t mFileOffset = 0, mRawSize = 0;
};

class RawEvent : public PipelineEvent {
public:
    explicit RawEvent(PipelineEventGroup* g) : PipelineEvent(Type::RAW, g) {}
    StringView GetContent() const { return mContent; }
    void SetContentNoCopy(StringView c) { mContent = c; }
    void SetContent(const std::string& c);
    Json::Value ToJson(bool enableEventMeta) const override;
    bool FromJson(const Json::Value& root);
    size_t DataSize() const override { return PipelineEvent::DataSize() + mContent.size(); } // RawEvent.cpp:47-49

private:
    StringView mContent;
};

// Metric / span events are outside the hot path: they are carried through untouched.
class OpaqueEvent : public PipelineEvent {
public:
    OpaqueEvent(Type t, PipelineEventGroup* g, const Json::Value& v) : PipelineEvent(t, g), mJson(v) {}
    Json::Value ToJson(bool) const override { return mJson; }

private:
    Json::Value mJson;
};

class PipelineEventPtr {
public:
    PipelineEventPtr() = default;
    explicit PipelineEventPtr(std::unique_ptr<PipelineEvent>&& p) : mData(std::move(p)) {}
    PipelineEventPtr(std::unique_ptr<PipelineEvent>&& p, bool /*fromPool*/, void* /*pool*/) : mData(std::move(p)) {}
    template <class T>
    bool Is() const;
    template <class T>
    T& Cast() {
        return static_cast<T&>(*mData);
    }
    template <class T>
    const T& Cast() const {
        return static_cast<const T&>(*mData);
    }
    PipelineEvent* operator->() { return mData.get(); }
    const PipelineEvent* operator->() const { return mData.get(); }
    explicit operator bool() const { return (bool)mData; }

private:
    std::unique_ptr<PipelineEvent> mData;
};
template <>
inline bool PipelineEventPtr::Is<LogEvent>() const {
    return mData && mData->GetType() == PipelineEvent::Type::LOG;
}
template <>
inline bool PipelineEventPtr::Is<RawEvent>() const {
    return mData && mData->GetType() == PipelineEvent::Type::RAW;
}

using EventsContainer = std::vector<PipelineEventPtr>;

enum class EventGroupMetaKey { UNKNOWN, LOG_FILE_PATH_RESOLVED, LOG_FILE_OFFSET_KEY, SOURCE_ID, HAS_PART_LOG };
using GroupMetadata = std::map<EventGroupMetaKey, StringView>;

class PipelineEventGroup {
public:
    explicit PipelineEventGroup(const std::shared_ptr<SourceBuffer>& sb) : mSourceBuffer(sb) {}
    PipelineEventGroup(PipelineEventGroup&&) = default;
    PipelineEventGroup& operator=(PipelineEventGroup&&) = default;

    const EventsContainer& GetEvents() const { return mEvents; }
    EventsContainer& MutableEvents() { return mEvents; }
    void SwapEvents(EventsContainer& other) { mEvents.swap(other); }
    std::unique_ptr<LogEvent> CreateLogEvent(bool /*fromPool*/ = false) { return std::make_unique<LogEvent>(this); }
    std::unique_ptr<RawEvent> CreateRawEvent(bool /*fromPool*/ = false) { return std::make_unique<RawEvent>(this); }
    LogEvent* AddLogEvent();
    RawEvent* AddRawEvent();
    std::shared_ptr<SourceBuffer>& GetSourceBuffer() { return mSourceBuffer; }

    void SetMetadata(EventGroupMetaKey key, const std::string& val);
    StringView GetMetadata(EventGroupMetaKey key) const;
    bool HasMetadata(EventGroupMetaKey key) const { return mMetadata.count(key) != 0; }
    void DelMetadata(EventGroupMetaKey key) { mMetadata.erase(key); }
    const GroupMetadata& GetAllMetadata() const { return mMetadata; }
    void SetTag(const std::string& key, const std::string& val);
    const std::map<StringView, StringView>& GetTags() const { return mTags; }

    // PipelineEventGroup.cpp:333-339: the events vector header + every event + the tags (SizedMap: header + bytes)
    size_t DataSize() const {
        size_t s = sizeof(mEvents);
        for (const auto& e : mEvents)
            s += e->DataSize();
        s += sizeof(mTags);
        for (const auto& kv : mTags)
            s += kv.first.size() + kv.second.size();
        return s;
    }

    Json::Value ToJson(bool enableEventMeta = false) const;
    bool FromJson(const Json::Value& root);
    std::string ToJsonString(bool enableEventMeta = false) const { return ToJson(enableEventMeta).toString(); }
    bool FromJsonString(const std::string& inJson);

private:
    EventsContainer mEvents;
    GroupMetadata mMetadata;
    std::map<StringView, StringView> mTags;
    std::shared_ptr<SourceBuffer> mSourceBuffer;
};

} // namespace logtail
