#!/usr/bin/env python3
"""Extract golden vectors from the reference's own processor unit tests.

Run in the BUILD container (where /root/reference exists); the output JSON files are committed
under tests/golden/ and are what the test-suite reads (the GPU box has no /root/reference).

The reference tests build a PipelineEventGroup from a JSON string, run one or more processors and
compare ``ToJsonString()`` with an expected JSON (core/unittest/processor/*Unittest.cpp, see
SURVEY.md section 4).  This script walks those C++ files with a small tokenizer, evaluates the string
expressions (raw literals, ``<<`` / ``+`` chains, file-level constants, ``strlen``), and records for
every assertion: the processors that ran (type + config snapshot), the input group JSON, the
metadata set programmatically, the expected output JSON (or null), whether event meta
(fileOffset/rawSize) is part of the comparison, and the counter assertions that follow.

Only DATA (configs, inputs, expected outputs) is extracted -- no reference code is copied.
"""
import json
import os
import re
import sys

REF = os.environ.get("LC_REFERENCE", "/root/reference")
UT = os.path.join(REF, "core/unittest/processor")
OUT = os.path.dirname(os.path.abspath(__file__))

FILES = {
    "split": "ProcessorSplitLogStringNativeUnittest.cpp",
    "multiline": "ProcessorSplitMultilineLogStringNativeUnittest.cpp",
    "regex": "ProcessorParseRegexNativeUnittest.cpp",
    "delimiter": "ProcessorParseDelimiterNativeUnittest.cpp",
    "merge": "ProcessorMergeMultilineLogNativeUnittest.cpp",
}

PROC_TYPES = {
    "ProcessorSplitLogStringNative": "processor_split_string_native",
    "ProcessorSplitMultilineLogStringNative": "processor_split_multiline_log_string_native",
    "ProcessorParseRegexNative": "processor_parse_regex_native",
    "ProcessorParseDelimiterNative": "processor_parse_delimiter_native",
    "ProcessorMergeMultilineLogNative": "processor_merge_multiline_log_native",
}

BUILTIN_CONSTS = {
    "DEFAULT_LOG_TAG_FILE_OFFSET": "__file_offset__",
    "DEFAULT_CONTENT_KEY": "content",
}

TOK_RE = re.compile(
    r'''(?P<raw>R"(?P<delim>[^()\\ ]{0,16})\((?P<rawbody>.*?)\)(?P=delim)")'''
    r'''|(?P<str>"(?:\\.|[^"\\])*")'''
    r"""|(?P<chr>'(?:\\.|[^'\\])')"""
    r"""|(?P<num>\d+)"""
    r"""|(?P<id>[A-Za-z_][A-Za-z0-9_]*(?:::[A-Za-z_][A-Za-z0-9_]*)*)"""
    r"""|(?P<lcomment>//[^\n]*)"""
    r"""|(?P<bcomment>/\*.*?\*/)"""
    r"""|(?P<op><<|->|[{}()\[\];,+\-*&.=<>!?:~|^%/])"""
    r"""|(?P<ws>\s+)""", re.S)

_ESC = {"n": "\n", "t": "\t", "r": "\r", "0": "\0", "\\": "\\", '"': '"', "'": "'", "a": "\a", "b": "\b",
        "f": "\f", "v": "\v"}


def unescape_c(s):
    out = []
    i = 0
    while i < len(s):
        c = s[i]
        if c == "\\" and i + 1 < len(s):
            n = s[i + 1]
            if n == "x":
                j = i + 2
                while j < len(s) and s[j] in "0123456789abcdefABCDEF":
                    j += 1
                out.append(chr(int(s[i + 2:j], 16)))
                i = j
                continue
            out.append(_ESC.get(n, n))
            i += 2
        else:
            out.append(c)
            i += 1
    return "".join(out)


def tokenize(src):
    toks = []
    pos = 0
    while pos < len(src):
        m = TOK_RE.match(src, pos)
        if not m:
            pos += 1
            continue
        pos = m.end()
        k = m.lastgroup
        if k in ("ws", "lcomment", "bcomment", "delim", "rawbody"):
            if m.group("raw") is None:
                continue
        if m.group("raw") is not None:
            toks.append(("str", m.group("rawbody")))
        elif m.group("str") is not None:
            toks.append(("str", unescape_c(m.group("str")[1:-1])))
        elif m.group("chr") is not None:
            toks.append(("num", ord(unescape_c(m.group("chr")[1:-1]))))
        elif m.group("num") is not None:
            toks.append(("num", int(m.group("num"))))
        elif m.group("id") is not None:
            toks.append(("id", m.group("id")))
        elif m.group("op") is not None:
            toks.append(("op", m.group("op")))
    return toks


class Eval:
    """Evaluates the small string/number expression language the tests use."""

    def __init__(self, consts):
        self.vars = dict(BUILTIN_CONSTS)
        self.vars.update(consts)

    def term(self, toks, i):
        k, v = toks[i]
        if k == "str":
            # adjacent literals concatenate
            s = v
            i += 1
            while i < len(toks) and toks[i][0] == "str":
                s += toks[i][1]
                i += 1
            return s, i
        if k == "num":
            return v, i + 1
        if k == "id":
            if v in ("strlen",) and toks[i + 1] == ("op", "("):
                val, j = self.expr(toks, i + 2, stop=(")",))
                return len(val.encode("utf-8")), j + 1
            if v in ("std::string", "string", "std::to_string", "ToString") and i + 1 < len(toks) and toks[i + 1] == (
                    "op", "("):
                val, j = self.expr(toks, i + 2, stop=(")",))
                return (str(val) if v.endswith("to_string") or v == "ToString" else val), j + 1
            if v == "GetDefaultTagKeyString":
                j = i + 1
                depth = 0
                while True:
                    if toks[j] == ("op", "("):
                        depth += 1
                    if toks[j] == ("op", ")"):
                        depth -= 1
                        if depth == 0:
                            break
                    j += 1
                return "__file_offset__", j + 1
            if v in ("true", "false"):
                return v == "true", i + 1
            if v in self.vars:
                val = self.vars[v]
                j = i + 1
                # NAME.str() / NAME.c_str()
                while j + 3 < len(toks) + 1 and j < len(toks) and toks[j] == ("op", ".") and toks[j + 1][0] == "id" \
                        and toks[j + 1][1] in ("str", "c_str") and toks[j + 2] == ("op", "(") and toks[j + 3] == (
                        "op", ")"):
                    j += 4
                return val, j
            raise KeyError(v)
        if (k, v) == ("op", "("):
            val, j = self.expr(toks, i + 1, stop=(")",))
            return val, j + 1
        raise ValueError("unexpected token %r" % (toks[i],))

    def chain(self, toks, i, stop):
        """a + b + c (numeric sum if all numeric else string concat)"""
        vals = []
        val, i = self.term(toks, i)
        vals.append(val)
        while i < len(toks) and toks[i] == ("op", "+"):
            val, i = self.term(toks, i + 1)
            vals.append(val)
        if all(isinstance(v, int) and not isinstance(v, bool) for v in vals):
            return sum(vals), i
        if len(vals) == 1:
            return vals[0], i
        return "".join(str(v) for v in vals), i

    def expr(self, toks, i, stop=(";",)):
        parts = []
        val, i = self.chain(toks, i, stop)
        parts.append(val)
        while i < len(toks) and toks[i] == ("op", "<<"):
            val, i = self.chain(toks, i + 1, stop)
            parts.append(val)
        if i < len(toks) and not (toks[i][0] == "op" and toks[i][1] in stop):
            raise ValueError("trailing tokens at %r" % (toks[i:i + 4],))
        if len(parts) == 1:
            return parts[0], i
        return "".join(str(p) for p in parts), i


def split_statements(toks):
    """Split a function body into statements at ';' '{' '}' (top level of parentheses)."""
    stmts, cur, depth = [], [], 0
    for t in toks:
        if t == ("op", "("):
            depth += 1
        elif t == ("op", ")"):
            depth -= 1
        if depth == 0 and t[0] == "op" and t[1] in (";", "{", "}"):
            if cur:
                stmts.append(cur)
            if t[1] != ";":
                stmts.append([t])
            cur = []
        else:
            cur.append(t)
    if cur:
        stmts.append(cur)
    return stmts


def find_functions(toks):
    """yield (class, name, body_tokens) for 'void Class::Name() {' definitions"""
    i = 0
    while i < len(toks) - 4:
        if toks[i] == ("id", "void") and toks[i + 1][0] == "id" and "::" in toks[i + 1][1] and toks[i + 2] == (
                "op", "(") and toks[i + 3] == ("op", ")") and toks[i + 4] == ("op", "{"):
            name = toks[i + 1][1]
            j = i + 5
            depth = 1
            while depth:
                if toks[j] == ("op", "{"):
                    depth += 1
                elif toks[j] == ("op", "}"):
                    depth -= 1
                j += 1
            yield name.split("::")[0], name.split("::")[-1], toks[i + 5:j - 1]
            i = j
        else:
            i += 1


def file_consts(toks, ev):
    i = 0
    while i < len(toks) - 4:
        if toks[i] == ("id", "const") and toks[i + 1][1] in ("std::string", "string") and toks[i + 2][0] == "id" \
                and toks[i + 3] == ("op", "="):
            j = i + 4
            k = j
            while toks[k] != ("op", ";"):
                k += 1
            try:
                val, _ = ev.expr(toks[j:k + 1], 0)
                ev.vars[toks[i + 2][1]] = val
            except Exception:
                pass
            i = k
        i += 1


def is_proc(name):
    return name.startswith("Processor") and name != "ProcessorInstance" and "::" not in name and \
        not name.endswith("Unittest")


def extract_file(kind, path):
    src = open(path, encoding="utf-8").read()
    toks = tokenize(src)
    ev0 = Eval({})
    file_consts(toks, ev0)
    cases = []
    skipped = []
    for cls, fname, body in find_functions(toks):
        ev = Eval(ev0.vars)
        config = {}
        var_type = {}  # var -> processor type name
        decl_no = {}  # var -> how many times it has been (re)declared: identifies the live instance
        inst_of = {}  # ProcessorInstance var -> processor var
        init_cfg = {}  # processor var -> config snapshot
        cur = None  # current case being built

        def new_case(in_json):
            return {"input": in_json, "pipeline": [], "metadata": dict(pending_meta), "counters": []}

        pending_meta = {}
        group_no = 0
        pre_events = []
        last_case = None
        cur_emitted = False
        enable_meta = False
        for st in split_statements(body):
            try:
                ids = [t[1] for t in st if t[0] == "id"]
                # declarations of processors
                if len(st) >= 2 and st[0][0] == "id" and is_proc(st[0][1]) and st[1][0] == "id" and len(st) == 2:
                    var_type[st[1][1]] = st[0][1]
                    decl_no[st[1][1]] = decl_no.get(st[1][1], 0) + 1
                    continue
                if len(st) >= 4 and st[0][0] == "id" and is_proc(st[0][1]) and st[1] == ("op", "&") and st[2][
                        0] == "id":
                    var_type[st[2][1]] = st[0][1]
                    decl_no[st[2][1]] = decl_no.get(st[2][1], 0) + 1
                    continue
                if ids[:1] == ["ProcessorInstance"] and len(st) > 4 and st[1][0] == "id" and st[2] == ("op", "("):
                    # ProcessorInstance inst(&proc, ...)
                    k = 3
                    if st[k] == ("op", "&"):
                        k += 1
                    inst_of[st[1][1]] = st[k][1]
                    continue
                if ids[:2] == ["Json::Value", "config"] and len(st) == 2:
                    config = {}
                    continue
                # config["K"] = expr   /  config["K"].append(expr)
                if st[0] == ("id", "config") and st[1] == ("op", "[") and st[2][0] == "str" and st[3] == ("op", "]"):
                    key = st[2][1]
                    if st[4] == ("op", "="):
                        if st[5] == ("id", "Json::arrayValue"):
                            config[key] = []
                        else:
                            val, _ = ev.expr(st[5:] + [("op", ";")], 0)
                            config[key] = val
                    elif st[4] == ("op", ".") and st[5] == ("id", "append"):
                        val, _ = ev.expr(st[7:-1] + [("op", ";")], 0)
                        config.setdefault(key, []).append(val)
                    continue
                if "ToJsonString" in ids:
                    enable_meta = ("id", "true") in st
                    continue
                # string variables
                if ids[:1] in (["std::string"], ["string"]) and len(st) >= 4 and st[1][0] == "id" and st[2] == (
                        "op", "="):
                    val, _ = ev.expr(st[3:] + [("op", ";")], 0)
                    ev.vars[st[1][1]] = val
                    continue
                if ids[:1] == ["std::stringstream"] and len(st) == 2:
                    ev.vars[st[1][1]] = ""
                    continue
                if st[0][0] == "id" and st[0][1] in ev.vars and len(st) > 2 and st[1] == ("op", "<<"):
                    val, _ = ev.expr(st[2:] + [("op", ";")], 0)
                    ev.vars[st[0][1]] = str(ev.vars[st[0][1]]) + str(val)
                    continue
                if st[0][0] == "id" and st[0][1] in ev.vars and len(st) > 2 and st[1] == ("op", "=") and \
                        "ToJsonString" not in ids:
                    val, _ = ev.expr(st[2:] + [("op", ";")], 0)
                    ev.vars[st[0][1]] = val
                    continue
                # metadata
                if "SetMetadata" in ids:
                    k = [t for t in st if t[0] == "id" and t[1].startswith("EventGroupMetaKey::")]
                    idx = st.index(("op", ","))
                    val, _ = ev.expr(st[idx + 1:-1] + [("op", ";")], 0)
                    pending_meta[k[0][1].split("::")[1]] = val
                    if cur is not None:
                        cur["metadata"][k[0][1].split("::")[1]] = val
                    continue
                if "make_shared" in " ".join(ids) or ids[:1] == ["PipelineEventGroup"]:
                    if ids[:1] == ["PipelineEventGroup"]:
                        pending_meta = {}
                        group_no += 1
                        pre_events = []
                    continue
                if "FromJsonString" in ids:
                    idx = ids.index("FromJsonString")
                    k = [n for n, t in enumerate(st) if t == ("id", "FromJsonString")][0]
                    depth, e = 0, k + 1
                    while True:
                        if st[e] == ("op", "("):
                            depth += 1
                        elif st[e] == ("op", ")"):
                            depth -= 1
                            if depth == 0:
                                break
                        e += 1
                    val, _ = ev.expr(st[k + 2:e] + [("op", ";")], 0)
                    if cur is not None and not cur["pipeline"] and cur.get("_group") == group_no:
                        cur.setdefault("more_inputs", []).append(val)  # a second FromJsonString on the same group
                    else:
                        cur = new_case(val)
                        cur["_group"] = group_no
                        if pre_events:
                            cur["pre_inputs"] = list(pre_events)
                            pre_events = []
                    cur_emitted = False
                    continue
                if "AddMetricEvent" in ids:
                    if cur is not None and not cur["pipeline"] and cur.get("_group") == group_no:
                        cur.setdefault("more_inputs", []).append("__metric__")
                    else:
                        pre_events.append("__metric__")  # added before the group's first FromJsonString
                    continue
                # Init
                if "Init" in ids and "config" in ids:
                    k = [n for n, t in enumerate(st) if t == ("id", "Init")][0]
                    v = st[k - 2][1]
                    v = inst_of.get(v, v)
                    init_cfg[v] = json.loads(json.dumps(config))
                    continue
                # Process
                if "Process" in ids and st[1] == ("op", ".") and st[2] == ("id", "Process"):
                    v = inst_of.get(st[0][1], st[0][1])
                    if cur is not None and v in var_type:
                        if var_type[v] not in PROC_TYPES:
                            cur["pipeline"].append({"type": "UNSUPPORTED:" + var_type[v], "config": {}, "var": v})
                        else:
                            cur["pipeline"].append({"type": PROC_TYPES[var_type[v]], "config": init_cfg.get(v, {}),
                                                    "var": v, "instance": "%s#%d" % (v, decl_no.get(v, 0))})
                    continue
                if "ToJsonString" in ids:
                    enable_meta = ("id", "true") in st
                    continue
                # assertions
                if ids and ids[0].startswith("APSARA_TEST_STREQ"):
                    if cur is None:
                        continue
                    if st[2][0] == "str" and st[2][1] == "null":
                        exp = None
                    else:
                        # CompactJson(NAME[.str()]).c_str()
                        k = [n for n, t in enumerate(st) if t == ("id", "CompactJson")]
                        if not k:
                            continue
                        name = st[k[0] + 2][1]
                        exp = ev.vars[name]
                    case = dict(cur)
                    case["expected"] = exp
                    case["enable_event_meta"] = enable_meta
                    case["name"] = "%s.%s#%d" % (cls, fname, len([c for c in cases if c["fn"] == fname]))
                    case["fn"] = fname
                    case["kind"] = kind
                    case["counters"] = []
                    cases.append(case)
                    last_case = case
                    cur_emitted = True
                    enable_meta = False
                    continue
                if ids and ids[0].startswith("APSARA_TEST_EQUAL") and "GetValue" in ids and cur is not None \
                        and not cur_emitted and cur["pipeline"]:
                    case = dict(cur)
                    case["expected"] = "__unchecked__"
                    case["enable_event_meta"] = False
                    case["name"] = "%s.%s#%d" % (cls, fname, len([c for c in cases if c["fn"] == fname]))
                    case["fn"] = fname
                    case["kind"] = kind
                    case["counters"] = []
                    cases.append(case)
                    last_case = case
                    cur_emitted = True
                if ids and ids[0].startswith("APSARA_TEST_EQUAL") and last_case is not None and "GetValue" in ids:
                    # APSARA_TEST_EQUAL_FATAL(N, var.mCounter->GetValue())
                    try:
                        comma = st.index(("op", ","))
                        val, _ = ev.expr(st[2:comma] + [("op", ";")], 0)
                        rest = [t[1] for t in st[comma + 1:] if t[0] == "id"]
                        holder = inst_of.get(rest[0], rest[0])
                        last_case["counters"].append({"var": holder, "counter": rest[1], "value": val,
                                                      "instance": "%s#%d" % (holder, decl_no.get(holder, 0)),
                                                      "on_instance": rest[0] in inst_of})
                    except Exception:
                        pass
                    continue
                if ids[:1] == ["int"] and len(st) == 4 and st[2] == ("op", "="):
                    ev.vars[st[1][1]] = st[3][1]
                    continue
            except Exception as e:  # noqa
                skipped.append((fname, " ".join(str(t[1]) for t in st[:8]), repr(e)))
    # validate JSON
    good = []
    for c in cases:
        try:
            c["input"] = json.loads(c["input"], strict=False)
            c.pop("_group", None)
            metric = {"name": "", "timestamp": 0, "type": 2, "value": {"type": "unknown"}}
            if c.get("pre_inputs"):
                c["input"]["events"] = [metric] * len(c.pop("pre_inputs")) + c["input"]["events"]
            for more in c.pop("more_inputs", []):
                # what eventGroup.AddMetricEvent() serialises to (core/models/MetricEvent.cpp ToJson, empty event)
                extra = [{"name": "", "timestamp": 0, "type": 2, "value": {"type": "unknown"}}] if more == "__metric__" \
                    else json.loads(more, strict=False)["events"]
                c["input"]["events"] = c["input"]["events"] + extra
            if c["expected"] is not None and c["expected"] != "__unchecked__":
                c["expected"] = json.loads(c["expected"], strict=False)
            if not c["pipeline"]:
                raise ValueError("no pipeline")
            if any(p["type"].startswith("UNSUPPORTED") for p in c["pipeline"]):
                raise ValueError("out-of-scope processor in pipeline: %s" % [p["type"] for p in c["pipeline"]])
            good.append(c)
        except Exception as e:
            skipped.append((c["name"], "json", repr(e)))
    return good, skipped


def main():
    total = 0
    for kind, fn in FILES.items():
        cases, skipped = extract_file(kind, os.path.join(UT, fn))
        with open(os.path.join(OUT, "ref_%s.json" % kind), "w", encoding="utf-8") as f:
            json.dump({"source": "core/unittest/processor/" + fn, "cases": cases}, f, indent=1, ensure_ascii=True,
                      sort_keys=True)
        total += len(cases)
        print("%-10s %3d cases, %d skipped statements" % (kind, len(cases), len(skipped)))
        if "-v" in sys.argv:
            for s in skipped:
                print("   skipped:", s)
    print("total", total)


if __name__ == "__main__":
    main()
