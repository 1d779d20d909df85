"""Trace featurizer — the producer of the estimator's input, host side.

Restates what ``resource-estimation/featurize.py`` computes from the bucketed raw data
(`raw_data`: one dict per time bucket with ``traces`` — call trees of
``{'component', 'operation', 'children'}`` — and ``metrics``):

* the feature space: one feature per distinct call-path PREFIX (root -> ... -> node), numbered in
  first-seen depth-first order over all buckets (featurize.py:11-24, :80-82);
* ``traffic[N, F]``: how many times each prefix occurs in each bucket (featurize.py:27-40, :84);
* ``resources``: ``'<component>_<resource>' -> series[N]`` (featurize.py:68-75);
* ``invocations``: per-component call counts per bucket plus ``'general'`` = number of traces
  (featurize.py:43-57, :89-101) — used by the reference's ComponentAware baseline.

Own implementation: iterative walks over tuple paths (no recursion, no ``copy.deepcopy`` per node); feature
keys are exposed in the reference's format (``str(list_of_'component_operation')``) so a feature space built by
either side is interchangeable.  Parity: tests/test_featurize.py against vectors minted by running the
reference functions (oracle/make_golden.py -> tests/golden/g9_featurize.npz).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def _walk(trace):
    """Yield (path_tuple, node) for every node of a call tree, parent before children, children in order."""
    stack = [(trace, ())]
    while stack:
        node, prefix = stack.pop()
        path = prefix + (node["component"] + "_" + node["operation"],)
        yield path, node
        for child in reversed(node["children"]):
            stack.append((child, path))


def feature_key(path):
    """The reference's dictionary key for a call path: str() of the list of 'component_operation' names."""
    return str(list(path))


class FeatureSpace:
    """Ordered call-path-prefix -> feature index map (featurize.py:11-24)."""

    def __init__(self):
        self.index = OrderedDict()          # path tuple -> feature id

    def fit(self, buckets):
        for bucket in buckets:
            for trace in bucket["traces"]:
                for path, _ in _walk(trace):
                    if path not in self.index:
                        self.index[path] = len(self.index)
        return self

    def __len__(self):
        return len(self.index)

    def keys(self):
        return [feature_key(p) for p in self.index]

    def transform(self, buckets):
        """traffic[N, F] int64 — occurrences of every known prefix per bucket (featurize.py:27-40)."""
        out = np.zeros((len(buckets), len(self.index)), np.int64)
        for i, bucket in enumerate(buckets):
            row = out[i]
            for trace in bucket["traces"]:
                for path, _ in _walk(trace):
                    row[self.index[path]] += 1          # KeyError for an unseen path, like the reference
        return out

    def components(self):
        comps = set()
        for path in self.index:
            for name in path:
                comps.add(name.split("_")[0])            # featurize.py:92-93 (component names carry no '_')
        return comps


def resources_of(buckets):
    """'<component>_<resource>' -> np.ndarray over buckets, in first-seen order (featurize.py:68-75)."""
    res = OrderedDict()
    for bucket in buckets:
        for metric in bucket["metrics"]:
            res.setdefault("%s_%s" % (metric["component"], metric["resource"]), []).append(metric["value"])
    return OrderedDict((k, np.asarray(v)) for k, v in res.items())


def invocations_of(buckets, components):
    """component -> calls per bucket, plus 'general' = traces per bucket (featurize.py:43-57, :95-101)."""
    names = set(components) | {"general"}
    counts = {c: np.zeros(len(buckets), np.int64) for c in names}
    for i, bucket in enumerate(buckets):
        counts["general"][i] = len(bucket["traces"])
        for trace in bucket["traces"]:
            for _, node in _walk(trace):
                c = node["component"]
                if c in counts:
                    counts[c][i] += 1
    return counts


def featurize(raw_data):
    """raw buckets -> [traffic, resources, invocations], the content of the reference's input.pkl (featurize.py:105-106)."""
    space = FeatureSpace().fit(raw_data)
    traffic = space.transform(raw_data)
    return traffic, resources_of(raw_data), invocations_of(raw_data, space.components()), space
