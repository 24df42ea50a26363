"""Pins the CPU oracle against every golden vector the reference's own unit tests hold for the path."""
import json
from collections import OrderedDict

import pytest

from oracle import oracle as orc
from tests.golden_util import all_cases, run_cases_of_function

BY_FN = OrderedDict()
for c in all_cases():
    BY_FN.setdefault(c["name"].split("#")[0], []).append(c)


@pytest.mark.parametrize("fn", list(BY_FN), ids=list(BY_FN))
def test_oracle_matches_reference_fixture(fn):
    def make(ptype, cfg):
        return orc.PROCESSORS[ptype](cfg)

    def run(proc, group_json, enable_meta):
        g = orc.Group.from_json(group_json)
        n_in = len(g.events)
        proc.process(g)
        return g.to_json(True), len(g.events), n_in

    n = run_cases_of_function(BY_FN[fn], make, run, lambda p: p.counters)
    assert n > 0


def test_oracle_misc_vectors():
    """Doc / StringTools golden vectors (tests/golden/ref_misc.json)."""
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "ref_misc.json"), encoding="utf-8") as f:
        d = json.load(f)
    for c in d["prefix_search"] + d["multiline_start"]:
        r = orc.Regex(c["pattern"])
        assert [r.prefix_match(i.encode()) for i in c["inputs"]] == c["expected"], c["source"]
    for c in d["full_match_fields"]:
        r = orc.Regex(c["pattern"])
        b = c["input"].encode()
        caps = r.full_match(b)
        assert caps is not None and [b[o:o + l].decode() for o, l in caps] == c["fields"], c["source"]
    # python `re` (independent Perl-semantics engine) agrees on the same vectors
    import re
    for c in d["full_match_fields"]:
        m = re.compile(c["pattern"].encode(), re.S | re.M).fullmatch(c["input"].encode())
        assert [g.decode() for g in m.groups()] == c["fields"]
