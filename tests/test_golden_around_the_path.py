"""The oracle (and the numpy generators) still give the digests frozen in tests/golden/around_the_path.json
(written by tests/golden/make_golden.py; regression vectors of the pinned oracle, see that script's header)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_the_frozen_digests(oracle_api):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    want = json.load(open(os.path.join(HERE, "golden", "around_the_path.json")))
    assert make_golden.compute() == want
