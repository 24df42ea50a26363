// dynamic_processor.cpp -- the B200-backed processors packaged as LoongCollector DYNAMIC plugins: one shared object
// per processor, each exporting the data symbol `processor_interface` that the agent's plugin loader resolves.
//
//   contract ...... core/collection_pipeline/plugin/creator/CProcessor.h:23-45 (processor_interface_t /
//                   processor_instance_t, PROCESSOR_INTERFACE_VERSION == 100)
//   loader ........ core/collection_pipeline/plugin/PluginRegistry.cpp:218-238 (dlopen of
//                   <execdir>/plugins + "lib" + name + ".so"), :255-275 (dlsym("processor_interface") + version check)
//   proxy ......... core/plugin/processor/DynamicCProcessorProxy.cpp:21-36: Init -> init(ins, &config, &context) == 0,
//                   Process -> process(plugin_state, &group), destructor -> finalize(plugin_state); the proxy leaves
//                   plugin_state uninitialised, so init must fill it on success AND on failure.
// The void*s carry live C++ objects (const Json::Value*, CollectionPipelineContext*, PipelineEventGroup*): this file is
// compiled against the host layer's models here and against the reference's own headers inside a LoongCollector build
// (INTEGRATION.md).  Static plugins register first and are never overwritten (PluginRegistry.cpp:277-283), hence the
// distinct `_b200` names.
//
// Built once per processor with -DLC_PLUGIN_NAME="processor_parse_regex_b200" -DLC_PLUGIN_TYPE="processor_parse_regex_native".
#include <memory>
#include <string>

#include "../host/Processors.h"

#ifndef LC_PLUGIN_NAME
#error "LC_PLUGIN_NAME / LC_PLUGIN_TYPE must be defined"
#endif

extern "C" {

// restated from CProcessor.h:23-45 (same member order and types: this IS the binary contract)
struct processor_instance_t;
typedef int (*processor_init_func_t)(struct processor_instance_t* ins, void* config, void* context);
typedef void (*processor_finialize_func_t)(void* plugin_state);
typedef void (*processor_process_func_t)(void* plugin_state, void* logGroup);
typedef struct processor_interface_t {
    int version;
    const char* name;
    const char* language;
    processor_init_func_t init;
    processor_finialize_func_t finalize;
    processor_process_func_t process;
} processor_interface_t;
typedef struct processor_instance_t {
    const processor_interface_t* plugin;
    void* plugin_state;
} processor_instance_t;

static int b200_init(processor_instance_t* ins, void* config, void* /*context*/) {
    if (!ins)
        return 1;
    ins->plugin_state = nullptr;
    if (!config)
        return 1;
    std::unique_ptr<logtail::Processor> p(logtail::CreateProcessor(LC_PLUGIN_TYPE));
    if (!p)
        return 1;
    try {
        if (!p->Init(*static_cast<const Json::Value*>(config)))
            return 1;
    } catch (...) {
        return 1;
    }
    ins->plugin_state = p.release();
    return 0;
}

static void b200_finalize(void* plugin_state) {
    delete static_cast<logtail::Processor*>(plugin_state);
}

static void b200_process(void* plugin_state, void* logGroup) {
    if (plugin_state && logGroup)
        static_cast<logtail::Processor*>(plugin_state)->Process(*static_cast<logtail::PipelineEventGroup*>(logGroup));
}

__attribute__((visibility("default"))) processor_interface_t processor_interface = {
    100, LC_PLUGIN_NAME, "C++ / CUDA sm_100a", b200_init, b200_finalize, b200_process,
};
}
