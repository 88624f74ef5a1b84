"""Output-quality metric of the reduced-precision modes: the reference's log-mel and log-mel L1.

Reference definition: `VocoderLoss.mel_reconstruction_loss` (reference models/losses.py:708-797; mel =
data/audio_processing.py:99-127, parameters configs/config.yaml:4-14): MelSpectrogram(22050 Hz, n_fft 1024,
hop 256, win 1024, 80 slaney mels, 0-8000 Hz, power 2) -> log10(. + 1e-10) -> L1.

CUDA tensors go through the library's own kernel (include/hfg_mel.h: framing, FFT, filterbank, log and the L1
reduction on the device -- SURVEY.md section 8f row 4).  CPU tensors use torchaudio, the same transform the
reference builds (CPU-box tests pin both to the reference's value)."""
from __future__ import annotations

import ctypes

import torch

from . import _capi

AUDIO = dict(sample_rate=22050, n_fft=1024, hop_length=256, win_length=1024, n_mels=80,
             fmin=0, fmax=8000, mel_scale="slaney", norm="slaney", log_base=10.0)

_cache = {}


class _MelConfig(ctypes.Structure):
    _fields_ = [("sample_rate", ctypes.c_int32), ("n_fft", ctypes.c_int32), ("hop_length", ctypes.c_int32),
                ("win_length", ctypes.c_int32), ("n_mels", ctypes.c_int32), ("fmin", ctypes.c_float), ("fmax", ctypes.c_float)]


class _MelHandle:
    def __init__(self):
        self.lib = _capi.load()
        self.h = ctypes.c_void_p()
        cfg = _MelConfig(AUDIO["sample_rate"], AUDIO["n_fft"], AUDIO["hop_length"], AUDIO["win_length"], AUDIO["n_mels"],
                         float(AUDIO["fmin"]), float(AUDIO["fmax"]))
        rc = self.lib.hfg_mel_create(ctypes.byref(cfg), ctypes.byref(self.h))
        if rc != _capi.OK:
            raise _capi.HfgError(rc, "hfg_mel_create failed (no CUDA device?)")

    def check(self, rc):
        if rc != _capi.OK:
            raise _capi.HfgError(rc, self.lib.hfg_mel_last_error(self.h).decode())

    def __del__(self):
        try:
            if self.h.value:
                self.lib.hfg_mel_destroy(self.h)
                self.h = ctypes.c_void_p()
        except Exception:
            pass


def _device_handle(dev) -> _MelHandle:
    key = ("hfg", dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _cache:
        with torch.cuda.device(dev):
            _cache[key] = _MelHandle()
    return _cache[key]


def _mel_transform(device):
    import torchaudio
    key = str(device)
    if key not in _cache:
        _cache[key] = torchaudio.transforms.MelSpectrogram(
            sample_rate=AUDIO["sample_rate"], n_fft=AUDIO["n_fft"], hop_length=AUDIO["hop_length"],
            win_length=AUDIO["win_length"], n_mels=AUDIO["n_mels"], f_min=AUDIO["fmin"], f_max=AUDIO["fmax"],
            mel_scale=AUDIO["mel_scale"], norm=AUDIO["norm"], power=2.0).to(device)
    return _cache[key]


def _as_2d(wav: torch.Tensor) -> torch.Tensor:
    assert wav.dim() == 3 and wav.size(1) == 1, f"expected [B, 1, T], got {list(wav.shape)}"
    return wav.squeeze(1).contiguous().float()


def log_mel(wav: torch.Tensor) -> torch.Tensor:
    """wav [B, 1, T] -> log10 mel [B, 80, T//256 + 1]."""
    x = _as_2d(wav)
    if not x.is_cuda:
        return torch.log10(_mel_transform(x.device)(x) + 1e-10)
    B, T = x.shape
    with torch.cuda.device(x.device):
        h = _device_handle(x.device)
        out = torch.empty((B, AUDIO["n_mels"], T // AUDIO["hop_length"] + 1), dtype=torch.float32, device=x.device)
        h.check(h.lib.hfg_log_mel(h.h, x.data_ptr(), B, T, out.data_ptr(),
                                  ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return out


def log_mel_l1(wav_ref: torch.Tensor, wav_new: torch.Tensor) -> float:
    a, b = _as_2d(wav_ref), _as_2d(wav_new)
    assert a.shape == b.shape
    if not a.is_cuda:
        return float(torch.nn.functional.l1_loss(log_mel(wav_new), log_mel(wav_ref)))
    B, T = a.shape
    with torch.cuda.device(a.device):
        h = _device_handle(a.device)
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        scratch = torch.empty(B * (T // AUDIO["hop_length"] + 1), dtype=torch.float32, device=a.device)
        h.check(h.lib.hfg_log_mel_l1(h.h, a.data_ptr(), b.data_ptr(), B, T, loss.data_ptr(), scratch.data_ptr(),
                                     ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)))
    return float(loss)
