"""Host-side mirror of the reference LengthRegulator and the duration rounding in front of it
 -- the integer frame indexing that produces the generator's input.

Reference: models/variance_adaptor.py:120-269 (`LengthRegulator`), :746-748
(`dur = clamp(round(exp(log_dur)).long(), min=1)`).  CUDA tensors only; the arithmetic is integer
prefix sums, a binary search and copies in libhfg_b200.so, so the result is bit-identical to
`torch.repeat_interleave` + zero padding."""
from __future__ import annotations

import ctypes
import os

import torch
import torch.nn as nn

from . import _capi


def _check(rc: int, what: str):
    if rc != 0:
        raise _capi.HfgError(rc, what)


def durations_from_log(log_dur: torch.Tensor) -> torch.Tensor:
    """log-durations [B, Tph] float32 (CUDA) -> int64 frame counts, >= 1 (reference :746-748)."""
    if not log_dur.is_cuda or log_dur.dtype != torch.float32:
        raise RuntimeError("durations_from_log expects a float32 CUDA tensor (no CPU path)")
    x = log_dur.contiguous()
    out = torch.empty(x.shape, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _check(_capi.load().hfg_durations_from_log(x.data_ptr(), x.numel(), out.data_ptr(),
                                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "hfg_durations_from_log")
    return out


class LengthRegulator(nn.Module):
    """Drop-in for the reference `LengthRegulator` (no parameters): forward(Henc [B,Tph,D], dur [B,Tph])
    -> Hlr [B, max_b sum(dur[b]), D], zero padded."""

    def __init__(self):
        super().__init__()
        self.debug_shapes = os.getenv("DEBUG_SHAPES", "0") == "1"

    def forward(self, Henc: torch.Tensor, dur: torch.Tensor) -> torch.Tensor:
        if Henc.dim() != 3 or dur.dim() != 2 or Henc.shape[:2] != dur.shape:
            raise RuntimeError(f"expected Henc [B,Tph,D] and dur [B,Tph], got {list(Henc.shape)} / {list(dur.shape)}")
        if not Henc.is_cuda:
            raise RuntimeError("LengthRegulator (B200) needs CUDA tensors: there is no CPU fallback")
        if self.debug_shapes:
            print(f"[LengthRegulator] Input Henc shape: {Henc.shape}")
            print(f"[LengthRegulator] Input dur shape: {dur.shape}")
        lib = _capi.load()
        h = Henc.contiguous().float()
        d = dur.to(device=Henc.device).long().contiguous()          # reference: dur.long() (:212)
        B, Tph, D = h.shape
        with torch.cuda.device(h.device):
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            frames = ctypes.c_int64()
            _check(lib.hfg_length_regulate_frames(d.data_ptr(), B, Tph, ctypes.byref(frames), st),
                   "hfg_length_regulate_frames")
            out = torch.empty((B, frames.value, D), dtype=torch.float32, device=h.device)
            if frames.value > 0:
                _check(lib.hfg_length_regulate(h.data_ptr(), d.data_ptr(), B, Tph, D, frames.value,
                                               out.data_ptr(), st), "hfg_length_regulate")
        if self.debug_shapes:
            print(f"[LengthRegulator] Output Hlr shape: {out.shape}")
        return out
