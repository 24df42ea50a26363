// host_capi.cpp -- JSON-driven C entry points over the host Processor classes (include/lc_b200_host.h).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <stdexcept>
#include <string>

#include "../../include/lc_b200_host.h"
#include "PluginBench.h"
#include "Processors.h"

using namespace logtail;

struct lc_host_processor {
    std::unique_ptr<Processor> proc;
};

static char* dup(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

extern "C" {

lc_host_processor_t* lc_host_processor_create(const char* type, const char* config_json, char** err_out) {
    if (err_out)
        *err_out = nullptr;
    std::unique_ptr<Processor> p(CreateProcessor(type ? type : ""));
    if (!p) {
        if (err_out)
            *err_out = dup(std::string("unknown processor type: ") + (type ? type : "(null)"));
        return nullptr;
    }
    Json::Value cfg(Json::objectValue);
    std::string err;
    const char* cj = config_json ? config_json : "{}";
    if (!Json::Value::parse(cj, cj + strlen(cj), cfg, err)) {
        if (err_out)
            *err_out = dup("config is not valid JSON: " + err);
        return nullptr;
    }
    try {
        if (!p->Init(cfg)) {
            if (err_out)
                *err_out = dup("Init failed: " + p->LastError());
            return nullptr;
        }
    } catch (const std::exception& e) {
        if (err_out)
            *err_out = dup(std::string("Init threw: ") + e.what());
        return nullptr;
    }
    auto* h = new lc_host_processor;
    h->proc = std::move(p);
    return h;
}

void lc_host_processor_destroy(lc_host_processor_t* p) { delete p; }

char* lc_host_processor_process(lc_host_processor_t* p, const char* group_json, int enable_event_meta, char** err_out) {
    if (err_out)
        *err_out = nullptr;
    try {
        auto sb = std::make_shared<SourceBuffer>();
        PipelineEventGroup group(sb);
        if (!group.FromJsonString(group_json ? group_json : "null"))
            throw std::runtime_error("group JSON does not parse");
        const uint64_t errs = p->proc->EngineErrors();
        p->proc->Process(group);
        if (p->proc->EngineErrors() != errs) // Process itself never throws (reference contract); the harness reports it
            throw std::runtime_error("engine error inside Process: " + p->proc->LastError());
        return dup(group.ToJsonString(enable_event_meta != 0));
    } catch (const std::exception& e) {
        if (err_out)
            *err_out = dup(e.what());
        return nullptr;
    }
}

char* lc_host_processor_counters(const lc_host_processor_t* p) {
    Json::Value v(Json::objectValue);
    for (auto& kv : p->proc->Counters())
        v[kv.first] = Json::Value((uint64_t)kv.second);
    return dup(v.toString());
}

void lc_host_string_free(char* s) { free(s); }

char* lc_host_sls_serialize(const char* group_json, int enable_ns, unsigned long long* len_out, char** err_out) {
    if (err_out)
        *err_out = nullptr;
    if (len_out)
        *len_out = 0;
    try {
        PipelineEventGroup group(std::make_shared<SourceBuffer>());
        if (group_json && strcmp(group_json, "null") != 0 && !group.FromJsonString(group_json)) {
            if (err_out)
                *err_out = dup("group is not valid JSON");
            return nullptr;
        }
        SLSEventGroupSerializer ser;
        ser.mEnableTimestampNanosecond = enable_ns != 0;
        std::string res, err;
        if (!ser.Serialize(group, res, err)) {
            if (err_out)
                *err_out = dup(err);
            return nullptr;
        }
        char* p = (char*)malloc(res.size() + 1);
        memcpy(p, res.data(), res.size());
        p[res.size()] = 0;
        if (len_out)
            *len_out = res.size();
        return p;
    } catch (const std::exception& e) {
        if (err_out)
            *err_out = dup(std::string("Serialize threw: ") + e.what());
        return nullptr;
    }
}


// What PluginRegistry::LoadProcessorPlugin + DynamicCProcessorProxy do with a dynamic plugin
// (PluginRegistry.cpp:218-275, DynamicCProcessorProxy.cpp:21-36), step by step, on the plugin at so_path.
char* lc_host_dynamic_plugin_roundtrip(const char* so_path, const char* config_json, const char* group_json,
                                       int enable_event_meta, int* version_out, char** name_out, char** err_out) {
    struct Iface { // CProcessor.h:23-45
        int version;
        const char* name;
        const char* language;
        int (*init)(void* ins, void* config, void* context);
        void (*finalize)(void* state);
        void (*process)(void* state, void* group);
    };
    struct Instance {
        const Iface* plugin;
        void* plugin_state;
    };
    if (err_out)
        *err_out = nullptr;
    void* h = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        if (err_out)
            *err_out = dup(std::string("dlopen: ") + dlerror());
        return nullptr;
    }
    char* result = nullptr;
    try {
        const Iface* plugin = static_cast<const Iface*>(dlsym(h, "processor_interface"));
        if (!plugin)
            throw std::runtime_error("symbol processor_interface not found");
        if (version_out)
            *version_out = plugin->version;
        if (name_out)
            *name_out = dup(plugin->name ? plugin->name : "");
        if (plugin->version != 100)
            throw std::runtime_error("plugin interface version mismatch");
        Json::Value cfg(Json::objectValue);
        std::string err;
        const char* cj = config_json ? config_json : "{}";
        if (!Json::Value::parse(cj, cj + strlen(cj), cfg, err))
            throw std::runtime_error("config is not valid JSON: " + err);
        Instance ins{plugin, reinterpret_cast<void*>(0xdeadbeef)}; // the proxy leaves plugin_state uninitialised
        int ctx = 0;
        if (plugin->init(&ins, &cfg, &ctx) != 0) {
            if (ins.plugin_state != nullptr)
                throw std::runtime_error("init failed and left plugin_state dangling");
            throw std::runtime_error("init returned non-zero");
        }
        if (group_json) {
            PipelineEventGroup group(std::make_shared<SourceBuffer>());
            if (!group.FromJsonString(group_json))
                throw std::runtime_error("group JSON does not parse");
            plugin->process(ins.plugin_state, &group);
            result = dup(group.ToJsonString(enable_event_meta != 0));
        } else {
            result = dup("null");
        }
        plugin->finalize(ins.plugin_state);
    } catch (const std::exception& e) {
        if (err_out)
            *err_out = dup(e.what());
        result = nullptr;
    }
    dlclose(h);
    return result;
}

void lc_host_use_pinned_arenas(int on) {
    if (on)
        SourceBuffer::SetChunkAllocator(&lc_host_alloc, &lc_host_free);
    else
        SourceBuffer::SetChunkAllocator(nullptr, nullptr);
}

int lc_host_bench_plugin(const char* type, const char* config_json, const uint8_t* data, const uint32_t* line_off,
                         const uint32_t* line_len, uint64_t n_lines, uint32_t group_bytes, int mode, int reps,
                         double* seconds_out, uint64_t stats_out[16], char** err_out) {
    if (err_out)
        *err_out = nullptr;
    try {
        std::unique_ptr<Processor> p(CreateProcessor(type ? type : ""));
        if (!p)
            throw std::runtime_error(std::string("unknown processor type: ") + (type ? type : "(null)"));
        Json::Value cfg(Json::objectValue);
        std::string err;
        const char* cj = config_json ? config_json : "{}";
        if (!Json::Value::parse(cj, cj + strlen(cj), cfg, err))
            throw std::runtime_error("config is not valid JSON: " + err);
        ProcessorInstance inst(p.release());
        if (!inst.Init(cfg))
            throw std::runtime_error("Init failed: " + inst.GetPlugin()->LastError());
        PluginBench bench(data, line_off, line_len, n_lines, group_bytes, 16);
        PluginBenchResult r = bench.Run(reps, [&](std::vector<PipelineEventGroup>& groups) {
            if (mode == 1) {
                inst.Process(groups); // ProcessorInstance::Process(vector<PipelineEventGroup>&), the pipeline's call
            } else {
                // one call per group, as a ProcessorRunner thread does with the groups it pops
                // (ProcessorRunner.cpp:128-143): the vector holds ONE group each time
                std::vector<PipelineEventGroup> one;
                for (auto& g : groups) {
                    one.clear();
                    one.emplace_back(std::move(g));
                    inst.Process(one);
                    g = std::move(one[0]);
                }
            }
        });
        if (inst.GetPlugin()->EngineErrors())
            throw std::runtime_error("engine error inside Process: " + inst.GetPlugin()->LastError());
        for (int k = 0; k < reps && seconds_out; ++k)
            seconds_out[k] = r.seconds[k];
        if (stats_out) {
            stats_out[0] = r.groups;
            stats_out[1] = r.inEvents;
            stats_out[2] = r.outEvents;
            stats_out[3] = r.liveContents;
            stats_out[4] = r.checksum;
            stats_out[5] = r.arenaBytes;
            stats_out[6] = inst.mInEventsTotal.GetValue();
            stats_out[7] = inst.mOutEventsTotal.GetValue();
            stats_out[8] = inst.mInSizeBytes.GetValue();
            stats_out[9] = inst.mOutSizeBytes.GetValue();
            stats_out[10] = inst.mTotalProcessTimeNs.GetValue();
            stats_out[11] = inst.mTotalProcessTimeMs.GetValue();
            stats_out[12] = stats_out[13] = stats_out[14] = stats_out[15] = 0;
            for (auto& kv : inst.GetPlugin()->Counters()) { // phase breakdown of the batched path, when the plugin has one
                if (kv.first == "b200_gather_ns")
                    stats_out[12] = kv.second;
                else if (kv.first == "b200_engine_ns")
                    stats_out[13] = kv.second;
                else if (kv.first == "b200_epilogue_ns")
                    stats_out[14] = kv.second;
            }
        }
        return 0;
    } catch (const std::exception& e) {
        if (err_out)
            *err_out = dup(e.what());
        return 1;
    }
}

}
