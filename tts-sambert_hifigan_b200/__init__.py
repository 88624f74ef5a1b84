"""B200-native HiFi-GAN generator (drop-in for the reference's
models/hifigan.py::HiFiGANGenerator inference path)."""
from . import synth  # noqa: F401
