"""SELF-regression vectors, not reference pins (those are tests/golden/*_known_answer.py, derived by hand from the
reference).  The oracle (and the numpy generators) still give the digests frozen in tests/golden/self_regression_digests.json
(written by tests/golden/make_golden.py; regression vectors of the pinned oracle, see that script's header)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_the_frozen_digests(oracle_api):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    want = json.load(open(os.path.join(HERE, "golden", "self_regression_digests.json")))
    assert make_golden.compute() == want
