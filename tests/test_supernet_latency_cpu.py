"""Expected-latency model of the supernet (`Network_Multi_Path.forward_latency`, search/model_search.py:361-475) against
goldens from the UNMODIFIED reference over a synthetic lookup table (oracle/make_golden_latency.py): 6 supernets x 2
architectures x 8 switch combinations.  Pure host arithmetic -- no kernels."""
import pytest
import torch

from oracle import make_golden_latency as mk
from oracle.make_golden_decode import SyntheticLatencyTable
from tests import helpers as H

GOLD = H.load_json("supernet_latency.json")


@pytest.mark.parametrize("case", mk.CASES, ids=lambda c: "L%d-%dx%d" % (c["layers"], c["hw"][0], c["hw"][1]))
def test_supernet_forward_latency_matches_reference(case, monkeypatch):
    from fasterseg_b200 import operations
    from fasterseg_b200.model_search import Network_Multi_Path
    monkeypatch.setattr(operations, "latency_lookup_table", SyntheticLatencyTable())
    model = mk.build(Network_Multi_Path, case["layers"])
    mk.randomise_arch(model, case["seed"])
    got = mk.evaluate(model, case)
    want = GOLD[str(case["seed"])]
    assert got.keys() == want.keys()
    for k in want:
        assert got[k] == pytest.approx(want[k], rel=2e-6), k   # float32 scalar arithmetic on the arch parameters
