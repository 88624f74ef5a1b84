"""The oracles against the LIVE reference on random geometries -- beyond the eight committed goldens.

For each drawn constructor configuration (incl. odd `k - u`, where T_out != T*hop) the unmodified
`models/hifigan.py::HiFiGANGenerator` is built from /root/reference, loaded with seeded weights in either
schema (plain / weight-normed, reference :263-283), run on a seeded mel, and both restatements --
`oracle/torch_port.py` and the plain-C `oracle/hifigan_oracle.c` -- must reproduce its waveform to fp32
round-off: 5e-7 (measured over a 150-example sweep: torch port 1.0e-7, C oracle 1.2e-7, signal peaks up to 0.37).  Runs on the CPU box only: the reference tree does not
travel to the GPU box.  Examples are derandomised (same draws on every run)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle
from tts_sambert_hifigan_b200 import synth

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")),
                                reason="reference tree not present (GPU box)")
TOL = 5e-7


@pytest.fixture(scope="module")
def ref_generator_cls():
    sys.path.insert(0, REF)
    try:
        from models.hifigan import HiFiGANGenerator
    finally:
        sys.path.remove(REF)
    return HiFiGANGenerator


@st.composite
def geometries(draw):
    n_up = draw(st.integers(1, 3))
    rates = [draw(st.sampled_from([2, 3, 4, 5, 8])) for _ in range(n_up)]
    kernels = [u + draw(st.integers(0, u + 2)) for u in rates]              # any k >= u: odd k - u included
    n_rb = draw(st.integers(1, 3))
    rks = [draw(st.sampled_from([3, 5, 7, 11])) for _ in range(n_rb)]
    dils = [[draw(st.integers(1, 5)) for _ in range(draw(st.integers(1, 3)))] for _ in range(n_rb)]
    return dict(n_mels=draw(st.sampled_from([5, 8, 16])), upsample_rates=rates, upsample_kernel_sizes=kernels,
                upsample_initial_channel=draw(st.sampled_from([8, 16, 24])) * 2 ** (n_up - 1),
                resblock_kernel_sizes=rks, resblock_dilation_sizes=dils)


@settings(max_examples=25, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
@given(geometries(), st.integers(1, 3), st.integers(1, 24), st.integers(0, 999), st.booleans())
def test_oracles_reproduce_the_live_reference(ref_generator_cls, cfg, batch, frames, seed, weight_norm):
    sd = synth.make_weightnorm_weights(cfg, seed) if weight_norm else synth.make_weights(cfg, seed)
    sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
    mel = synth.make_mel(seed + 1, batch, cfg["n_mels"], frames)

    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_generator_cls(**cfg).eval()
        if weight_norm:
            ref.apply_weight_norm()                                          # 232-key schema (the constructor builds the plain one)
        ref.load_state_dict(sd_t, strict=True)
        with torch.no_grad():
            want = ref(torch.from_numpy(mel)).numpy()

    got_t = oracle.forward_torch(cfg, sd_t, torch.from_numpy(mel)).numpy()
    assert got_t.shape == want.shape and np.abs(got_t - want).max() <= TOL, cfg

    plain = {k: v.numpy() for k, v in oracle.fold_weight_norm(sd_t).items()}
    names = [n for n, _ in synth.weight_shapes(cfg)]
    got_c = oracle.forward_c(cfg, plain, names, mel)
    assert got_c.shape == want.shape and np.abs(got_c - want).max() <= TOL, cfg


# ---- the KV-cached decoder restatement against the reference's O(T^2) loop ----

@pytest.fixture(scope="module")
def ref_decoder_cls():
    sys.path.insert(0, REF)
    try:
        from models.ar_decoder import PNCAARDecoder
    finally:
        sys.path.remove(REF)
    return PNCAARDecoder


@settings(max_examples=12, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
@given(st.sampled_from([(32, 2), (32, 4), (64, 8), (48, 3)]), st.integers(1, 3), st.sampled_from([32, 96]),
       st.sampled_from([4, 20]), st.integers(1, 3), st.integers(1, 10), st.integers(0, 999))
def test_kv_cached_decoder_oracle_reproduces_the_live_reference(ref_decoder_cls, dh, n_layers, d_ff, n_mels, batch, frames, seed):
    """Random small decoder geometries with the reference's own random initialisation: caching keys / values
    must not change a single frame (reference models/ar_decoder.py:167-238 recomputes the whole prefix per frame)."""
    from oracle import ar_decoder as ard_oracle
    d_model, n_heads = dh
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(seed)
        ref = ref_decoder_cls(d_model=d_model, n_mels=n_mels, n_layers=n_layers, n_heads=n_heads, d_ff=d_ff).eval()
        hvar = torch.randn(batch, frames, d_model)
        with torch.no_grad():
            want = ref(hvar)
    with torch.no_grad():
        got = ard_oracle.decode(ref.state_dict(), hvar, n_layers, n_heads)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max())), (dh, n_layers, d_ff)


# ---- the log-mel restatement against the reference's extract_mel ----

def test_log_mel_oracle_reproduces_the_live_extract_mel():
    """reference data/audio_processing.py:31-139 with the reference's own configs/config.yaml, on seeded waveforms
    of several lengths (incl. one shorter than a hop and one that is not a multiple of the hop)."""
    pytest.importorskip("torchaudio")
    import yaml
    from oracle import log_mel as olm
    sys.path.insert(0, REF)
    try:
        from data.audio_processing import extract_mel
    finally:
        sys.path.remove(REF)
    with open(os.path.join(REF, "configs", "config.yaml")) as f:
        config = yaml.safe_load(f)
    config.setdefault("debug", {})["print_shapes"] = False
    a = config["audio"]
    assert {k: a[k] for k in ("sample_rate", "n_fft", "hop_length", "win_length", "n_mels")} == \
        {k: olm.AUDIO[k] for k in ("sample_rate", "n_fft", "hop_length", "win_length", "n_mels")}
    for seed, samples in ((1, 22050), (2, 5000), (3, 700), (4, 256 * 40)):
        wav = (0.1 * synth.normal(seed, (samples,))).astype(np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            want = extract_mel(torch.from_numpy(wav), sample_rate=a["sample_rate"], config=config).numpy()
        got = olm.log_mel(wav[None])[0]
        assert got.shape == want.shape == (80, samples // 256 + 1)
        assert np.abs(got - want).max() <= 5e-5, (samples, float(np.abs(got - want).max()))     # fp32 FFT vs float64


# ---- the length-regulator restatement against the reference module ----

@settings(max_examples=40, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
@given(st.integers(1, 6), st.integers(1, 14), st.integers(1, 9), st.integers(0, 2 ** 31 - 1))
def test_length_regulator_oracle_reproduces_the_live_module(batch, n_ph, d_model, seed):
    """reference models/variance_adaptor.py:171-269 on random durations incl. zeros and negatives (clamped to 0 there)."""
    from oracle import length_regulator as lr_oracle
    sys.path.insert(0, REF)
    try:
        from models.variance_adaptor import LengthRegulator
    finally:
        sys.path.remove(REF)
    rng = np.random.default_rng(seed)
    henc = rng.standard_normal((batch, n_ph, d_model)).astype(np.float32)
    dur = rng.integers(-2, 7, size=(batch, n_ph)).astype(np.int64)
    if dur.clip(0).sum(axis=1).max() == 0:
        dur[0, 0] = 1                                 # an all-empty batch has no frame axis to compare
    with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
        want = LengthRegulator()(torch.from_numpy(henc), torch.from_numpy(dur)).numpy()
    got = lr_oracle.length_regulate(henc, dur)
    assert got.shape == want.shape and np.array_equal(got, want)
