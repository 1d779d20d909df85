"""deeprest_b200 — B200-native (sm_100a) implementation of the DeepRest resource-estimator
hot path: the batched ``QuantileRNN`` forward (reference resource-estimation/qrnn.py).

Public surface:
  * ``QuantileRNN``       host mirror of the reference module, backed by libdeeprest_b200.so
  * ``layout``            weight-blob layout (reference state_dict order)
  * ``synth``             counter-based synthetic trace/weight generator
  * ``featurize`` / ``synthesizer`` / ``evaluation``   host code either side of the path:
                          trace featurizer, what-if trace synthesizer, the test stage with its error report
The CUDA library is loaded lazily on first use and there is no CPU fallback.
"""
from . import evaluation, featurize, layout, synth, synthesizer  # noqa: F401
from .estimator import QuantileRNN, sliding_window  # noqa: F401

__all__ = ["QuantileRNN", "sliding_window", "layout", "synth", "featurize", "synthesizer", "evaluation"]
