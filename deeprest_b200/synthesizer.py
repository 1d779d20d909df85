"""Trace synthesizer — turns *expected API traffic* into estimator input vectors, host side.

Restates ``resource-estimation/synthesizer.py``: the tool a DeepRest user runs to ask "what would the
resources look like under THIS traffic": for every API endpoint (a root ``component_operation``) it
learns, from observed traces, the set of call-path count vectors one invocation of that endpoint
produces; a hypothetical bucket ``{'endpoint': n_calls, ...}`` is then synthesized by drawing
``n_calls`` of those vectors per endpoint and summing them (synthesizer.py:15-52).  The result has the
featurizer's layout, so rows of it (windowed and normalised) go straight into ``QuantileRNN``.

Behaviour kept from the reference, including its quirks:

* endpoints are the length-1 call paths of the feature space, in first-seen order (synthesizer.py:21-25);
* the candidate vectors of an endpoint are numbered in first-seen order and their observed
  frequencies are recorded (synthesizer.py:28-37) — but **the draw is uniform over the distinct
  candidates**: the reference passes no ``p=`` to ``np.random.choice`` (synthesizer.py:48), so the
  learned frequencies are not used.  ``weighted=True`` is this package's opt-in to use them;
* an unknown endpoint is an ``AssertionError`` with the reference's message (synthesizer.py:43-44);
* draws come from numpy's global ``np.random`` stream unless ``rng`` is given, endpoint by endpoint in
  the order of the request dict, one ``choice(size=count)`` call each — so with the same seed the
  result is bit-identical to the reference's.

(The reference's ``synthesize`` itself no longer runs on numpy ≥ 1.24 — it uses ``np.int`` — the golden
vectors in tests/golden/g11_synthesizer.json were minted with that alias restored.)
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from .featurize import FeatureSpace, _walk, feature_key


class TraceSynthesizer:
    def __init__(self):
        self.space = None                   # FeatureSpace (call-path prefix -> feature id)
        self.api2dist = None                # endpoint -> (candidates int64 [C, F], weights int64 [C])

    # -- learning (synthesizer.py:15-40) -----------------------------------------------------------------------------
    def fit(self, data):
        space = FeatureSpace().fit(data)
        seen = OrderedDict((path[0], OrderedDict()) for path in space.index if len(path) == 1)
        for bucket in data:
            for trace in bucket["traces"]:
                vec = np.zeros(len(space), np.int64)
                for path, _ in _walk(trace):
                    vec[space.index[path]] += 1
                dist = seen[trace["component"] + "_" + trace["operation"]]
                key = vec.tobytes()
                if key in dist:
                    dist[key][1] += 1
                else:
                    dist[key] = [vec, 1]
        self.space = space
        self.api2dist = OrderedDict(
            (api, (np.stack([v for v, _ in dist.values()]) if dist else np.zeros((0, len(space)), np.int64),
                   np.asarray([c for _, c in dist.values()], np.int64)))
            for api, dist in seen.items())
        return self

    @property
    def M(self):
        """The reference's feature-space dict: str(call path) -> feature id."""
        return OrderedDict((feature_key(p), i) for p, i in self.space.index.items())

    def endpoints(self):
        return list(self.api2dist)

    # -- synthesis (synthesizer.py:42-52) ----------------------------------------------------------------------------
    def synthesize(self, expected_api_calls, rng=None, weighted=False):
        """{'endpoint': n_calls} -> feature vector int64 [F] for one time bucket."""
        for api in expected_api_calls:
            assert api in self.api2dist, "API endpoint `%s` does not exist." % api
        rng = np.random if rng is None else rng
        x = np.zeros(len(self.space), np.int64)
        for api, count in expected_api_calls.items():
            candidates, weights = self.api2dist[api]
            p = weights / weights.sum() if weighted else None
            picks = rng.choice(len(candidates), size=count, replace=True, p=p)
            if count:
                x += candidates[picks].sum(axis=0)
        return x

    def synthesize_series(self, expected_api_traffic, rng=None, weighted=False):
        """[{'endpoint': n_calls}, ...] (one dict per time step) -> traffic int64 [N, F], the featurizer's layout."""
        return np.stack([self.synthesize(calls, rng=rng, weighted=weighted) for calls in expected_api_traffic]) \
            if len(expected_api_traffic) else np.zeros((0, len(self.space)), np.int64)
