"""Evaluation stage of the reference's driver script, host side.

``resource-estimation/estimate.py:79-122`` — after every epoch the reference forecasts up to nine
non-overlapping test windows (every ``step_size``-th stride-1 window of the test split, batch size 1),
averages their quantile losses, clamps the forecasts at 1e-6, de-normalises the median quantile and the
labels with the per-metric ``(max-min, min)`` pairs, and prints median / 95-th / 99-th / max absolute
error per metric.  This module restates that stage on top of ``QuantileRNN.forward_series`` (windows
formed on the device from the raw series — no W-fold copy), for any object exposing
``forward_series(series, window, stride)`` and ``quantile_loss(outputs, labels)``.

The baselines the reference prints beside it (RESRC / COMP, ``baselines.py``) are CPU regressors outside
the hot path; ``summaries`` accepts their error lists so the same report can be printed.
"""
from __future__ import annotations

import numpy as np

MAX_CYCLES = 9          # estimate.py:86 `num_cycles >= 9`
CLAMP = 1e-6            # estimate.py:96


def eval_window_offsets(n_series, split, step_size, max_cycles=MAX_CYCLES):
    """Series offsets of the windows the reference evaluates: stride-1 windows X[i] = series[i:i+W] exist for
    i < n_series - W (utils.py:4-5 drops the last one); the test loader walks i = split + iv and keeps
    iv % step_size == 0 until nine were taken (estimate.py:85-88)."""
    n_test = n_series - step_size - split
    return [split + iv for iv in range(0, max(n_test, 0), step_size)][:max_cycles]


def error_summary(errs):
    """estimate.py:111-121 — (median, 95-th, 99-th, max) of absolute errors."""
    e = np.asarray(errs, np.float64)
    return {"median": float(np.median(e)), "p95": float(np.percentile(e, q=95)),
            "p99": float(np.percentile(e, q=99)), "max": float(np.max(e))}


def format_summary(tag, s):
    """One console line in the reference's format (estimate.py:112-120): tag is 'RESRC', 'COMP ' or 'DEEPR'."""
    return "   %s => Median: %.4f | 95-th: %.4f | 99-th: %.4f | Max: %.4f" % (tag, s["median"], s["p95"], s["p99"], s["max"])


def evaluate(model, traffic, resources, names, scales, split, step_size, max_cycles=MAX_CYCLES):
    """Test stage of one epoch.

    traffic   [N,F]  normalised input series (estimate.py:42)
    resources [N,M]  normalised label series, one column per entry of ``names`` (estimate.py:44-47)
    scales    [(max-min, min), ...] per metric (estimate.py:47)
    split     index of the first test window (estimate.py:30)

    Returns {'loss': mean test loss, 'errors': {name: abs errors}, 'summary': {name: {...}}, 'offsets': [...]}.
    """
    traffic = np.asarray(traffic, np.float32)
    resources = np.asarray(resources, np.float32)
    offsets = eval_window_offsets(len(traffic), split, step_size, max_cycles)
    errors = {name: [] for name in names}
    losses = []
    if offsets:
        # one call for all evaluated windows: they are step_size apart, i.e. a strided walk over the test series
        last = offsets[-1] + step_size
        out = model.forward_series(traffic[split:last + 1], step_size, stride=step_size)
        assert len(out) == len(offsets), (len(out), len(offsets))
        for k, off in enumerate(offsets):
            labels = resources[off:off + step_size]                              # [W,M]
            losses.append(float(model.quantile_loss(out[k:k + 1], labels[None])))  # batch of 1, as the reference
            clamped = np.maximum(out[k], CLAMP)
            for idx, name in enumerate(names):
                rng, lo = scales[idx]
                lab = labels[:, idx] * rng + lo
                med = clamped[:, idx, 1] * rng + lo
                errors[name] += list(np.abs(med - lab))
    summary = {name: error_summary(e) for name, e in errors.items() if len(e)}
    return {"loss": float(np.mean(losses)) if losses else float("nan"), "errors": errors, "summary": summary,
            "offsets": offsets}


def report(names, deeprest, resrc=None, comp=None, header=None):
    """The per-epoch console block of estimate.py:109-121.  Each argument maps name -> summary dict."""
    lines = [header] if header else []
    for name in names:
        lines.append("===== %s =====" % name)
        if resrc is not None:
            lines.append(format_summary("RESRC", resrc[name]))
        if comp is not None:
            lines.append(format_summary("COMP ", comp[name]))
        lines.append(format_summary("DEEPR", deeprest[name]))
    return "\n".join(lines)
