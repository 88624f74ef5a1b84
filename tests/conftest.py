import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def case_inputs(manifest, name):
    """Regenerate (cfg, state_dict, mel) of a golden case from its seeds."""
    from tts_sambert_hifigan_b200 import synth
    c = manifest["cases"][name]
    cfg = c["cfg"]
    if c["extra"].get("weight_norm"):
        sd = synth.make_weightnorm_weights(cfg, c["weight_seed"])
    else:
        sd = synth.make_weights(cfg, c["weight_seed"], gain=c["extra"].get("gain", 1.0))
    mel = synth.make_mel(c["mel_seed"], c["B"], cfg["n_mels"], c["T"])
    return cfg, sd, mel
