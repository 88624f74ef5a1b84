"""on-device log-mel and log-mel L1 (include/hfg_mel.h) against the reference's
definition (torchaudio MelSpectrogram + log10, reference data/audio_processing.py:99-127 and
models/losses.py:708-797) -- through the numpy oracle, the torchaudio transform and the value the live reference's
VocoderLoss.mel_reconstruction_loss returned (tests/golden/manifest.json "logmel_l1_pin")."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from tts_sambert_hifigan_b200 import _capi, metrics, synth
import oracle.log_mel as olm


def _pin_pair(manifest):
    p = manifest["logmel_l1_pin"]
    a = synth.normal(p["seed_a"], p["shape"]) * p["scale_a"]
    b = a + synth.normal(p["seed_b"], p["shape"]) * p["scale_b"]
    return a.astype(np.float32), b.astype(np.float32), p["value"]


def test_header_symbols_exported():
    header = open(os.path.join(ROOT, "include", "hfg_mel.h")).read()
    declared = set(re.findall(r"\b(hfg_(?:mel_[a-z_]+|log_mel[a-z0-9_]*))\s*\(", header))
    assert declared == set(_capi.MEL_SYMBOLS)
    lib = ctypes.CDLL(_capi.lib_path())
    for name in declared:
        assert hasattr(lib, name), name


def test_oracle_matches_reference_value_and_torchaudio(manifest):
    a, b, want = _pin_pair(manifest)
    got = olm.log_mel_l1(a[:, 0], b[:, 0])
    assert abs(got - want) <= 1e-5 * want                   # measured 2e-7 relative
    pytest.importorskip("torchaudio")
    lt = metrics.log_mel(torch.from_numpy(a)).numpy()       # CPU tensor: torchaudio, as the reference builds it
    assert np.abs(olm.log_mel(a[:, 0]) - lt).max() <= 5e-5  # measured 6.5e-6 (float32 vs float64 FFT)


@pytest.mark.gpu
def test_cuda_log_mel_matches_oracle_and_reference_value(manifest):
    a, b, want = _pin_pair(manifest)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    lm = metrics.log_mel(ta)
    ref = olm.log_mel(a[:, 0])
    assert tuple(lm.shape) == ref.shape == (2, 80, 8192 // 256 + 1)
    err = float(np.abs(lm.cpu().numpy() - ref).max())
    loss = metrics.log_mel_l1(ta, tb)
    print(f"on-device log-mel vs float64 oracle: max-abs {err:.3e}; L1 {loss:.9f} vs reference's {want:.9f} "
          f"(rel {abs(loss - want) / want:.2e})")
    assert err <= 5e-5
    assert abs(loss - want) <= 2e-5 * want
    assert loss == metrics.log_mel_l1(ta, tb)               # deterministic reduction
    # a 2 s utterance batch with odd length and a generator-like amplitude: edges (reflect padding) included
    w = (synth.normal(77, (3, 1, 44100 + 37)) * 0.03).astype(np.float32)
    lm = metrics.log_mel(torch.from_numpy(w).cuda()).cpu().numpy()
    ref = olm.log_mel(w[:, 0])
    assert lm.shape == ref.shape and np.abs(lm - ref).max() <= 5e-5
    with pytest.raises(_capi.HfgError):
        metrics.log_mel(torch.zeros(1, 1, 100).cuda())      # shorter than the reflect padding


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [
    dict(sample_rate=16000, n_fft=512, hop_length=128, win_length=512, n_mels=40, fmin=50.0, fmax=7600.0),
    dict(sample_rate=22050, n_fft=2048, hop_length=512, win_length=2048, n_mels=128, fmin=0.0, fmax=11025.0),
])
def test_cuda_log_mel_other_configurations(cfg):
    """The kernel is parametrised like configs/config.yaml's `audio:` section; other FFT sizes / filterbanks against
    the float64 oracle (hfg_mel_create through the raw C ABI)."""
    lib = _capi.load()
    c = metrics._MelConfig(cfg["sample_rate"], cfg["n_fft"], cfg["hop_length"], cfg["win_length"], cfg["n_mels"],
                           cfg["fmin"], cfg["fmax"])
    h = ctypes.c_void_p()
    assert lib.hfg_mel_create(ctypes.byref(c), ctypes.byref(h)) == 0
    w = (synth.normal(91, (2, 9000)) * 0.05).astype(np.float32)
    x = torch.from_numpy(w).cuda()
    frames = 9000 // cfg["hop_length"] + 1
    out = torch.empty((2, cfg["n_mels"], frames), dtype=torch.float32, device="cuda")
    assert lib.hfg_log_mel(h, x.data_ptr(), 2, 9000, out.data_ptr(), None) == 0
    torch.cuda.synchronize()
    ref = olm.log_mel(w, cfg)
    err = float(np.abs(out.cpu().numpy() - ref).max())
    print(f"log-mel n_fft {cfg['n_fft']} n_mels {cfg['n_mels']}: max-abs vs float64 oracle {err:.3e}")
    assert ref.shape == tuple(out.shape) and err <= 1e-4
    bad = metrics._MelConfig(22050, 1000, 256, 1000, 80, 0.0, 8000.0)          # n_fft not a power of two
    assert lib.hfg_mel_create(ctypes.byref(bad), ctypes.byref(ctypes.c_void_p())) == _capi.ERR_INVALID
    lib.hfg_mel_destroy(h)
