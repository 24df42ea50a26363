"""GPU tier: the reference's own unit-test fixtures (tests/golden/ref_*.json) replayed through the C++ host
layer (B200-backed Processor classes behind the reference's plugin API) -- the drop-in claim."""
from collections import OrderedDict

import pytest

pytestmark = pytest.mark.gpu

from tests.golden_util import all_cases, run_cases_of_function  # noqa: E402

BY_FN = OrderedDict()
for c in all_cases():
    BY_FN.setdefault(c["name"].split("#")[0], []).append(c)


@pytest.mark.parametrize("fn", list(BY_FN), ids=list(BY_FN))
def test_host_processors_match_reference_fixture(fn):
    import loongcollector_b200 as lc

    def make(ptype, cfg):
        return lc.HostProcessor(ptype, cfg)

    def run(proc, group_json, enable_meta):
        n_in = len((group_json or {}).get("events", []))
        out = proc.process(group_json, True)
        return out, len((out or {}).get("events", [])), n_in

    n = run_cases_of_function(BY_FN[fn], make, run, lambda p: p.counters())
    assert n > 0
