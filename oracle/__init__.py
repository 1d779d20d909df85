"""CPU oracle for the DeepRest hot path — test infrastructure only (see qrnn_numpy.py)."""
