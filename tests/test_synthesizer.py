"""N4 — trace synthesizer against vectors minted by running the reference's TraceSynthesizer
(oracle/make_golden.py::synthesizer_golden): learned endpoint distributions and seeded synthesized vectors."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from deeprest_b200.synthesizer import TraceSynthesizer

CASES = json.load(open(os.path.join(GOLDEN_DIR, "g11_synthesizer.json")))


@pytest.mark.parametrize("name", sorted(CASES))
def test_learned_distributions_match_reference(name):
    g = CASES[name]
    syn = TraceSynthesizer().fit(g["raw"])
    assert list(syn.M.keys()) == g["keys"]
    assert syn.endpoints() == list(g["api2dist"])                 # same endpoints, same first-seen order
    for api, ref in g["api2dist"].items():
        cand, weights = syn.api2dist[api]
        assert cand.tolist() == ref["candidates"]                 # same candidate vectors in first-seen order
        assert weights.tolist() == ref["weights"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_seeded_synthesis_is_bit_identical_to_reference(name):
    g = CASES[name]
    syn = TraceSynthesizer().fit(g["raw"])
    for k, (req, want) in enumerate(zip(g["requests"], g["vectors"])):
        np.random.seed(g["seed0"] + k)                            # the reference draws from the global stream
        assert syn.synthesize(req).tolist() == want
        got = syn.synthesize(req, rng=np.random.RandomState(g["seed0"] + k))
        assert got.tolist() == want


def test_unknown_endpoint_fails_like_the_reference():
    syn = TraceSynthesizer().fit(CASES["synthetic"]["raw"])
    with pytest.raises(AssertionError, match="does not exist"):
        syn.synthesize({"no-such_Endpoint": 1})


def test_series_and_weighted_draws():
    g = CASES["synthetic"]
    syn = TraceSynthesizer().fit(g["raw"])
    api = syn.endpoints()[0]
    series = syn.synthesize_series([{api: 3}, {api: 0}, {api: 5}], rng=np.random.RandomState(1))
    assert series.shape == (3, len(syn.M)) and series.dtype == np.int64
    assert series[1].sum() == 0
    root = syn.M["['%s']" % api]
    assert series[0, root] == 3 and series[2, root] == 5          # every invocation counts its root path once
    w = syn.synthesize({api: 200}, rng=np.random.RandomState(2), weighted=True)
    assert w[root] == 200
