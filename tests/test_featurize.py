"""N4 — trace featurizer against vectors minted by running the reference's featurize.py functions
(oracle/make_golden.py::featurize_golden): synthetic call trees and the reference's own shipped sample."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from deeprest_b200 import featurize as fz

CASES = json.load(open(os.path.join(GOLDEN_DIR, "g9_featurize.json")))


@pytest.mark.parametrize("name", sorted(CASES))
def test_featurizer_matches_reference(name):
    g = CASES[name]
    traffic, resources, invocations, space = fz.featurize(g["raw"])
    assert space.keys() == g["keys"]                              # same features, same first-seen DFS numbering
    assert np.array_equal(traffic, np.asarray(g["traffic"]))
    for comp, series in invocations.items():                      # reference lists only components it saw; ours adds zeros
        assert list(series) == g["invocations"].get(comp, [0] * len(g["raw"])), comp
    assert set(g["invocations"]) <= set(invocations)
    first = g["raw"][0]["metrics"][0]
    key = "%s_%s" % (first["component"], first["resource"])
    assert resources[key][0] == first["value"] and len(resources[key]) == len(g["raw"])


def test_unseen_call_path_is_an_error_like_the_reference():
    g = CASES["synthetic"]
    space = fz.FeatureSpace().fit(g["raw"][:1])
    novel = {"traces": [{"component": "brand", "operation": "New", "children": []}], "metrics": []}
    with pytest.raises(KeyError):
        space.transform([novel])


def test_featurized_traffic_feeds_the_windowing_helpers():
    from deeprest_b200 import sliding_window
    g = CASES["synthetic"]
    traffic, _, _, _ = fz.featurize(g["raw"])
    win = sliding_window(traffic, 2)
    assert win.shape == (len(g["raw"]) - 2, 2, traffic.shape[1])   # utils.py:4-5 drops the last window
