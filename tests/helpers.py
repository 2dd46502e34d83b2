"""Shared test helpers: golden fixtures, synthetic weights, the oracle (test infrastructure)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import synth  # noqa: E402  (oracle/synth.py)
import gwdepth_oracle as oracle  # noqa: E402  (oracle/gwdepth_oracle.py)


def load_spec():
    with open(os.path.join(GOLDEN, "state_dict_spec.json")) as f:
        return json.load(f)


_weights = {}


def synth_weights(seed=0):
    """flat fp32 state dict with the reference's key names (structural int buffers included)"""
    if seed not in _weights:
        spec = [tuple(s) for s in load_spec()["keys"]]
        sd = synth.synth_state_dict(spec, seed=seed)
        _weights[seed] = synth.add_structural_buffers(sd, spec)
    return _weights[seed]


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
