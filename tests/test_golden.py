"""The oracle against the committed golden vectors (tests/golden/golden_v1.json,
generated from the unmodified reference by tests/golden/make_golden.py).  This
is what pins the oracle on boxes where /root/reference does not exist."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G   # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden_v1.json")))


@pytest.mark.parametrize("name", sorted(G.cases()))
def test_oracle_matches_golden(orc, name):
    got = G.evaluate(orc, name, G.cases()[name])
    want = GOLDEN[name]
    got = json.loads(json.dumps(got))          # same int/str key normalisation as the file
    assert got == want


def test_golden_known_constants():
    assert GOLDEN["kat131"]["pos"] == [3, 41, 49, 84]
    assert GOLDEN["kat131"]["hasher"]["factor1"] == 0x49308bb9003cb3ad
    assert GOLDEN["kat_short"]["n"] == 0
