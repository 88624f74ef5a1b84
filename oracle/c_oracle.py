"""ctypes driver for oracle/hifigan_oracle.c (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhfg_oracle.so")
_MAX = 8


class _Cfg(ctypes.Structure):
    _fields_ = [("n_mels", ctypes.c_int32), ("n_up", ctypes.c_int32), ("c0", ctypes.c_int32),
                ("n_rk", ctypes.c_int32), ("up_rates", ctypes.c_int32 * _MAX),
                ("up_ks", ctypes.c_int32 * _MAX), ("rk", ctypes.c_int32 * _MAX),
                ("n_dil", ctypes.c_int32 * _MAX), ("dil", (ctypes.c_int32 * _MAX) * _MAX)]


def build_c_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "hifigan_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "_build/libhfg_oracle.so"])
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_c_oracle())
        lib.hfgo_out_len.restype = ctypes.c_int64
        lib.hfgo_out_len.argtypes = [ctypes.POINTER(_Cfg), ctypes.c_int]
        lib.hfgo_num_weights.restype = ctypes.c_int
        lib.hfgo_num_weights.argtypes = [ctypes.POINTER(_Cfg)]
        lib.hfgo_forward.restype = ctypes.c_int
        _lib = lib
    return _lib


def _mk_cfg(cfg: dict) -> _Cfg:
    c = _Cfg()
    c.n_mels = cfg["n_mels"]
    c.n_up = len(cfg["upsample_rates"])
    c.c0 = cfg["upsample_initial_channel"]
    c.n_rk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        c.up_rates[i], c.up_ks[i] = u, k
    for j, (rk, dils) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
        c.rk[j], c.n_dil[j] = rk, len(dils)
        for l, d in enumerate(dils):
            c.dil[j][l] = d
    return c


def forward_c(cfg: dict, weights: Dict[str, np.ndarray], names: List[str], mel: np.ndarray,
              stages: Optional[list] = None) -> np.ndarray:
    """Run the C oracle.  `weights` is a plain-schema state_dict (numpy fp32),
    `names` its key order (synth.weight_shapes order).  If `stages` is a list it
    receives the 2*n_up+1 stage-boundary activations [B,C,T]."""
    lib = _load()
    c = _mk_cfg(cfg)
    mel = np.ascontiguousarray(mel, dtype=np.float32)
    B, _, T = mel.shape
    arrs = [np.ascontiguousarray(weights[n], dtype=np.float32) for n in names]
    assert len(arrs) == lib.hfgo_num_weights(ctypes.byref(c))
    FP = ctypes.POINTER(ctypes.c_float)
    wptr = (FP * len(arrs))(*[a.ctypes.data_as(FP) for a in arrs])
    Tout = lib.hfgo_out_len(ctypes.byref(c), T)
    wav = np.empty((B, 1, Tout), dtype=np.float32)
    sptr = None
    bufs = []
    if stages is not None:
        C, Tc = cfg["upsample_initial_channel"], T
        bufs.append(np.empty((B, C, Tc), dtype=np.float32))
        for u, k in zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"]):
            C //= 2
            Tc = (Tc - 1) * u - 2 * ((k - u) // 2) + k
            bufs.append(np.empty((B, C, Tc), dtype=np.float32))
            bufs.append(np.empty((B, C, Tc), dtype=np.float32))
        sptr = (FP * len(bufs))(*[b.ctypes.data_as(FP) for b in bufs])
    rc = lib.hfgo_forward(ctypes.byref(c), wptr, mel.ctypes.data_as(FP), int(B), int(T),
                          wav.ctypes.data_as(FP), sptr)
    if rc != 0:
        raise MemoryError("hfgo_forward failed")
    if stages is not None:
        stages.extend(bufs)
    return wav
