/*
 * lc_b200_host.h -- C entry points of the C++ host layer (loongcollector_b200/host): the B200-backed
 * replacements of the reference's Processor subclasses, driven through JSON event groups exactly like the
 * reference's unit tests drive them (PipelineEventGroup::FromJsonString / ToJsonString,
 * core/models/PipelineEventGroup.cpp:432-483).  A LoongCollector build links the C++ classes directly
 * (INTEGRATION.md); these functions exist so that tests in any language can replay the reference fixtures.
 */
#ifndef LC_B200_HOST_H
#define LC_B200_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lc_host_processor lc_host_processor_t;

/* type: "processor_split_string_native" | "processor_split_multiline_log_string_native" |
 *       "processor_parse_regex_native" | "processor_parse_delimiter_native" (the reference's plugin names).
 * Returns NULL when Init(config) fails; *err_out (if not NULL) then holds a malloc'd message. */
lc_host_processor_t* lc_host_processor_create(const char* type, const char* config_json, char** err_out);
void lc_host_processor_destroy(lc_host_processor_t* p);
/* Runs Processor::Process(PipelineEventGroup&) on the group described by group_json and returns the group's
 * ToJsonString(enable_event_meta) ("null" for an empty group) as a malloc'd string; NULL + *err_out on error. */
char* lc_host_processor_process(lc_host_processor_t* p, const char* group_json, int enable_event_meta, char** err_out);
/* {"counter": value, ...} with the reference's counter meanings. */
char* lc_host_processor_counters(const lc_host_processor_t* p);
void lc_host_string_free(char* s);

/* SLSEventGroupSerializer::Serialize (core/collection_pipeline/serializer/SLSSerializer.cpp:162-252) of the LOG
 * group described by group_json; enable_ns = GlobalConfig::mEnableTimestampNanosecond.  Returns the malloc'd wire
 * bytes (free with lc_host_string_free) and their length, or NULL + *err_out = the reference's error message. */
char* lc_host_sls_serialize(const char* group_json, int enable_ns, unsigned long long* len_out, char** err_out);

/* Loads a dynamic plugin the way the agent does (dlopen, dlsym("processor_interface"), version == 100 --
 * PluginRegistry.cpp:255-275) and drives it like DynamicCProcessorProxy (.cpp:21-36): init(ins, &config, &context),
 * process(plugin_state, &group), finalize(plugin_state).  Returns the processed group's JSON ("null" when group_json
 * is NULL: init / finalize only); NULL + *err_out on any failure.  *version_out / *name_out (malloc'd) report the
 * interface fields as soon as the symbol resolves. */
char* lc_host_dynamic_plugin_roundtrip(const char* so_path, const char* config_json, const char* group_json,
                                       int enable_event_meta, int* version_out, char** name_out, char** err_out);

/* Makes every SourceBuffer chunk created from now on pinned (lc_host_alloc): a group's arena is then DMA-able in
 * place -- the integration's replacement of SourceBuffer's `new char[]` (core/common/memory/SourceBuffer.h:98-131). */
void lc_host_use_pinned_arenas(int on);

/* End-to-end run at the plugin boundary (bench.py's `e2e`).  Builds event groups of <= group_bytes from the line
 * table (one arena chunk per group, one LogEvent {"content": line} per line: the state the reader + split processor
 * leave, LogFileReader.cpp:97) and times `reps` repetitions of
 *   mode 0: ProcessorInstance::Process with ONE group per call (a ProcessorRunner thread popping groups),
 *   mode 1: ProcessorInstance::Process(std::vector<PipelineEventGroup>&) with ALL groups (batched override).
 * data[line_off[i] + line_len[i]] must be readable (the separator byte travels with the line).  Groups are rebuilt,
 * untimed, before every repetition.  seconds_out[reps] = wall time of the Process calls of each repetition.
 * stats_out[12] = groups, in events, out events, live contents, content checksum (sum of key.size*131 +
 * value.size*31 + first value byte), arena bytes, then ProcessorInstance's counters: in events, out events, in
 * bytes, out bytes, process ns, process ms, then the plugin's own phase timers (gather ns, engine-call ns, epilogue ns;
 * 0 when it has none), one spare -- all summed over the repetitions.  Returns 0, or 1 + *err_out. */
int lc_host_bench_plugin(const char* type, const char* config_json, const uint8_t* data, const uint32_t* line_off,
                         const uint32_t* line_len, uint64_t n_lines, uint32_t group_bytes, int mode, int reps,
                         double* seconds_out, uint64_t stats_out[16], char** err_out);

#ifdef __cplusplus
}
#endif
#endif
