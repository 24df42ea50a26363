#!/usr/bin/env python3
"""Extracts the reference's RemoveLastIncompleteLog unit-test cases into tests/golden/ref_rollback.json.

Source: core/unittest/reader/RemoveLastIncompleteLogUnittest.cpp (TestSingleline, TestMultiline and the five
RemoveLastIncompleteLogMultilineUnittest functions).  Every `{ // case ... }` block builds its input with std::string
concatenations of the file's constants; this script evaluates those expressions and records
(config, input, expected return value, expected rollbackLineFeedCount).  Run in the build container only
(/root/reference does not exist on the GPU box); the JSON it writes is what the tests read.
"""
import json
import os
import re
import sys

SRC = "/root/reference/core/unittest/reader/RemoveLastIncompleteLogUnittest.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_rollback.json")


def cxx_expr(e, env):
    e = " ".join(e.strip().rstrip(";").split())
    e = re.sub(r"'\\n'", '"\\\\n"', e)
    e = re.sub(r'R"\((.*?)\)"', lambda m: repr(m.group(1)), e)
    e = re.sub(r"std::string\((\w+)\.data\(\)\)", r"\1", e)
    return eval(e, {}, env)


def main():
    src = open(SRC, encoding="utf-8").read()
    env = {}
    for m in re.finditer(r"const std::string (\w+) = (.*?);\n", src):
        env[m.group(1)] = cxx_expr(m.group(2), env)
    cases = []
    for fm in re.finditer(r"void (\w+)::(\w+)\(\) \{\n(.*?)\n\}\n", src, re.S):
        cls, fn, body = fm.groups()
        if not fn.startswith("Test") or "RemoveLastIncompleteLog(" not in body:
            continue
        if re.search(r"ContainerdTextParser|DockerJsonFileParser|mFileLogFormat|GetParser<", body):
            continue  # container stdout parsers: outside the raw-text path restated here
        cfg = {}
        for m in re.finditer(r'config\["(\w+)"\] = (\w+);', body):
            cfg[m.group(1)] = env[m.group(2)]
        loc = dict(env)
        title = ""
        k = 0
        tok = re.compile(r"// case([^\n]*)\n|std::string (\w+)\s*=\s*([^;]*?);\n|"
                         r"RemoveLastIncompleteLog\(\s*const_cast<char\*>\((\w+)\.data\(\)\)|"
                         r"APSARA_TEST_EQUAL(?:_FATAL)?\(([^\n;]*?), matchSize\);|"
                         r"APSARA_TEST_EQUAL(?:_FATAL)?\((\d+), rollbackLineFeedCount\);", re.S)
        cur = None
        for m in tok.finditer(body):
            if m.group(1) is not None:
                title = m.group(1).strip(" :")
            elif m.group(2):
                loc[m.group(2)] = cxx_expr(m.group(3), loc)
            elif m.group(4):
                cur = {"name": "%s#%d" % (fn, k), "title": title, "config": cfg, "input": loc[m.group(4)]}
                k += 1
            elif m.group(5) is not None and cur is not None:
                e = m.group(5)
                sz = re.search(r"(\w+)\.size\(\)", e)
                cur["expect_size"] = len(loc[sz.group(1)]) if sz else int(e)
            elif m.group(6) is not None and cur is not None:
                cur["expect_rollback"] = int(m.group(6))
                cases.append(cur)
                cur = None
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(cases, f, indent=1, ensure_ascii=False)
    print("wrote %d cases to %s" % (len(cases), OUT))


if __name__ == "__main__":
    sys.exit(main())
