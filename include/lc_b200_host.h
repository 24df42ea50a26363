/*
 * lc_b200_host.h -- C entry points of the C++ host layer (loongcollector_b200/host): the B200-backed
 * replacements of the reference's Processor subclasses, driven through JSON event groups exactly like the
 * reference's unit tests drive them (PipelineEventGroup::FromJsonString / ToJsonString,
 * core/models/PipelineEventGroup.cpp:432-483).  A LoongCollector build links the C++ classes directly
 * (INTEGRATION.md); these functions exist so that tests in any language can replay the reference fixtures.
 */
#ifndef LC_B200_HOST_H
#define LC_B200_HOST_H
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

#ifdef __cplusplus
}
#endif
#endif
