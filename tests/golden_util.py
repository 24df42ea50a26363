"""Shared helpers to replay the reference's unit-test fixtures (tests/golden/ref_*.json).

The fixtures were extracted from core/unittest/processor/*Unittest.cpp by
tests/golden/extract_reference_vectors.py; each case = processors (type+config), input group JSON,
programmatic metadata, expected ``ToJsonString()`` and the counter assertions that followed."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

META_ENUM_TO_JSON = {
    "LOG_FILE_OFFSET_KEY": "log.file.offset",
    "LOG_FILE_PATH_RESOLVED": "log.file.path_resolved",
    "SOURCE_ID": "source.id",
    "HAS_PART_LOG": "has.part.log",
}

COUNTER_NAMES = {
    "mDiscardedEventsTotal": "discarded",
    "mOutFailedEventsTotal": "out_failed",
    "mOutKeyNotFoundEventsTotal": "out_key_not_found",
    "mOutSuccessfulEventsTotal": "out_successful",
    "mMatchedEventsTotal": "matched_events",
    "mMatchedLinesTotal": "matched_lines",
    "mUnmatchedLinesTotal": "unmatched_lines",
}


def load_cases(kind):
    with open(os.path.join(GOLDEN, "ref_%s.json" % kind), encoding="utf-8") as f:
        return json.load(f)["cases"]


def all_cases():
    out = []
    for k in ("split", "multiline", "regex", "delimiter", "filter", "merge"):
        out.extend(load_cases(k))
    return out


def input_with_metadata(case):
    root = json.loads(json.dumps(case["input"])) or {}
    for k, v in case.get("metadata", {}).items():
        root.setdefault("metadata", {})[META_ENUM_TO_JSON[k]] = v
    return root


def strip_event_meta(root):
    if not root:
        return root
    root = json.loads(json.dumps(root))
    for ev in root.get("events", []):
        ev.pop("fileOffset", None)
        ev.pop("rawSize", None)
    return root


def run_cases_of_function(cases, make_processor, run_processor, get_counters):
    """Replay the cases of ONE reference test function in order.

    make_processor(type, config) -> processor object
    run_processor(proc, group_json, enable_meta) -> (out_group_json_with_event_meta, n_events_out, n_events_in)
    get_counters(proc) -> dict of counter name -> value (names as in COUNTER_NAMES values)
    Processor instances persist across cases exactly when the reference test re-used the same object
    (``instance`` ids from the extractor), so cumulative counter assertions line up."""
    live = {}
    checked = 0
    for case in cases:
        root = input_with_metadata(case)
        n_in = n_out = None
        for step in case["pipeline"]:
            key = step["instance"]
            if key not in live:
                live[key] = make_processor(step["type"], step["config"])
            root, n_out, n_in = run_processor(live[key], root, True)
        got = root if case["enable_event_meta"] else strip_event_meta(root)
        if case["expected"] != "__unchecked__":
            assert json.dumps(got, sort_keys=True) == json.dumps(case["expected"], sort_keys=True), case["name"]
            checked += 1
        for c in case["counters"]:
            if c["on_instance"]:
                val = {"mInEventsTotal": n_in, "mOutEventsTotal": n_out}.get(c["counter"])
                if val is None:
                    continue
            else:
                p = live.get(c["instance"])
                name = COUNTER_NAMES.get(c["counter"])
                if p is None or name is None or name not in get_counters(p):
                    continue
                val = get_counters(p)[name]
            assert val == c["value"], (case["name"], c, val)
            checked += 1
    return checked
