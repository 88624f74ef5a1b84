"""B200-native HiFi-GAN generator inference path: a drop-in for the reference's
models/hifigan.py::HiFiGANGenerator (mel [B,80,Tfrm] -> wav [B,1,Tfrm*256])
whose arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in
include/hfg.h."""
from . import synth  # noqa: F401
from . import _capi  # noqa: F401
from .generator import HiFiGANGenerator  # noqa: F401
from .length_regulator import LengthRegulator, durations_from_log  # noqa: F401
from .ar_decoder import PNCAARDecoder  # noqa: F401

__all__ = ["HiFiGANGenerator", "LengthRegulator", "PNCAARDecoder", "durations_from_log", "synth"]
