"""Output-quality metrics for the reduced-precision modes (not on the product path).

log_mel_l1 restates the reference's `VocoderLoss.mel_reconstruction_loss`
(reference models/losses.py:708-797; mel definition = data/audio_processing.py:99-127,
parameters configs/config.yaml:4-14): torchaudio MelSpectrogram(22050 Hz, n_fft 1024,
hop 256, win 1024, 80 slaney mels, 0-8000 Hz, power 2) -> log10(. + 1e-10) -> L1.
The reference class cannot travel to the GPU box (and prints on construction), so the
metric is restated here; tests/golden pins it against the reference's own value."""
from __future__ import annotations

import torch

AUDIO = dict(sample_rate=22050, n_fft=1024, hop_length=256, win_length=1024, n_mels=80,
             fmin=0, fmax=8000, mel_scale="slaney", norm="slaney", log_base=10.0)

_cache = {}


def _mel_transform(device):
    import torchaudio
    key = str(device)
    if key not in _cache:
        _cache[key] = torchaudio.transforms.MelSpectrogram(
            sample_rate=AUDIO["sample_rate"], n_fft=AUDIO["n_fft"], hop_length=AUDIO["hop_length"],
            win_length=AUDIO["win_length"], n_mels=AUDIO["n_mels"], f_min=AUDIO["fmin"], f_max=AUDIO["fmax"],
            mel_scale=AUDIO["mel_scale"], norm=AUDIO["norm"], power=2.0).to(device)
    return _cache[key]


def log_mel(wav: torch.Tensor) -> torch.Tensor:
    """wav [B, 1, T] -> log10 mel [B, 80, T//256 + 1]."""
    assert wav.dim() == 3 and wav.size(1) == 1
    return torch.log10(_mel_transform(wav.device)(wav.squeeze(1)) + 1e-10)


def log_mel_l1(wav_ref: torch.Tensor, wav_new: torch.Tensor) -> float:
    return float(torch.nn.functional.l1_loss(log_mel(wav_new), log_mel(wav_ref)))
