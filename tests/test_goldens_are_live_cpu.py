"""The committed golden vectors ARE the live reference's output: tests/golden/make_golden.py is imported and its
cases re-run on /root/reference here; every array of the committed .npz (waveform, stage-boundary activations,
key count) and the manifest entry must come out again.  Same container, same torch: the waveforms are expected
bit for bit; the bound allows for ATen choosing another summation order on a box with a different core count.
CPU box only (the reference does not travel)."""
import importlib.util
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")),
                                reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def maker():
    import sys
    before = list(sys.path)
    spec = importlib.util.spec_from_file_location("_make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.path[:] = before                       # the script puts /root/reference on sys.path; the suite must not keep it
    return mod


@pytest.mark.parametrize("name", ["default_b2_t24", "default_stages_b1_t9", "default_weightnorm_b1_t16",
                                  "default_saturated_b1_t16", "default_ragged_b3_t7", "odd_upsample_b1_t20",
                                  "small_custom_b3_t33"])
def test_committed_golden_is_what_the_live_reference_computes(maker, manifest, name):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        out, meta = maker.compute_case(name, maker.CASES[name])
    g = load_golden(name)
    assert set(out) == set(g)
    worst = 0.0
    for key, want in g.items():
        got = np.asarray(out[key])
        assert got.shape == want.shape, key
        if key == "wav_fp64_maxdiff":
            continue
        d = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) if got.size else 0.0
        worst = max(worst, d)
        assert d <= 1e-6 * max(1.0, float(np.abs(want).max())), (key, d)
    committed = manifest["cases"][name]
    assert json.loads(json.dumps(meta["cfg"])) == committed["cfg"]
    assert (meta["weight_seed"], meta["mel_seed"], meta["B"], meta["T"]) == \
        (committed["weight_seed"], committed["mel_seed"], committed["B"], committed["T"])
    assert meta["wav_shape"] == committed["wav_shape"]
    print(f"{name}: max |live - committed| over all arrays = {worst:.3e}")
