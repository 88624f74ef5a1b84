"""ORACLE -- TEST INFRASTRUCTURE ONLY.  numpy (float64) restatement of the reference's log-mel
(reference data/audio_processing.py:99-127: torchaudio MelSpectrogram(power=2, center=True, reflect padding,
periodic Hann, slaney mel scale, slaney norm) then log10(. + 1e-10); parameters configs/config.yaml:4-14) and of
VocoderLoss.mel_reconstruction_loss (reference models/losses.py:708-797).  Pinned to the value the live
reference returned (tests/golden/manifest.json "logmel_l1_pin") by tests/test_log_mel.py."""
import numpy as np

AUDIO = dict(sample_rate=22050, n_fft=1024, hop_length=256, win_length=1024, n_mels=80, fmin=0.0, fmax=8000.0)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def filterbank(cfg=AUDIO):
    """[n_freqs, n_mels] -- torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney')."""
    n_freqs = cfg["n_fft"] // 2 + 1
    all_freqs = np.linspace(0, cfg["sample_rate"] // 2, n_freqs)
    m_pts = np.linspace(_hz_to_mel(cfg["fmin"]), _hz_to_mel(cfg["fmax"]), cfg["n_mels"] + 2)
    f_pts = _mel_to_hz(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    return fb * (2.0 / (f_pts[2:cfg["n_mels"] + 2] - f_pts[:cfg["n_mels"]]))[None, :]


def log_mel(wav, cfg=AUDIO):
    """wav [B, T] -> [B, n_mels, T // hop + 1]."""
    wav = np.asarray(wav, dtype=np.float64)
    N, hop = cfg["n_fft"], cfg["hop_length"]
    x = np.pad(wav, ((0, 0), (N // 2, N // 2)), mode="reflect")
    frames = wav.shape[1] // hop + 1
    idx = np.arange(N)[None, :] + hop * np.arange(frames)[:, None]
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N) / N)
    spec = np.abs(np.fft.rfft(x[:, idx] * win, axis=-1)) ** 2           # [B, frames, n_freqs]
    mel = spec @ filterbank(cfg)                                         # [B, frames, n_mels]
    return np.log10(mel + 1e-10).transpose(0, 2, 1)


def log_mel_l1(wav_ref, wav_new, cfg=AUDIO):
    return float(np.mean(np.abs(log_mel(wav_new, cfg) - log_mel(wav_ref, cfg))))
