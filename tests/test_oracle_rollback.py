"""CPU tier: the oracle's restatement of LogFileReader::RemoveLastIncompleteLog against every case of the reference's
own unit test (tests/golden/ref_rollback.json, extracted by tests/golden/extract_rollback_vectors.py)."""
import json
import os

import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "ref_rollback.json"), encoding="utf-8"))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_fixture(case):
    start, end = orc.multiline_regs(case["config"])
    keep, rb = orc.remove_last_incomplete_log(case["input"].encode(), start, end, True)
    assert (keep, rb) == (case["expect_size"], case["expect_rollback"]), case["title"]


def test_rollback_not_allowed_keeps_everything():
    keep, rb = orc.remove_last_incomplete_log(b"a\nb", None, None, False)
    assert (keep, rb) == (3, 0)
