"""Portable synthetic data for the HiFi-GAN generator path.

There is no network, no checkpoint and no dataset: every weight and every mel
the tests and the bench use is generated here from an integer seed.  The
generator is a counter-based splitmix64 hash evaluated with numpy uint64
arithmetic, so the same seed gives the same bits on every box (torch.manual_seed
streams are not guaranteed to be stable across builds, which would make the
committed golden vectors fragile).

* weights follow the scale of PyTorch's default Conv1d / ConvTranspose1d init
  (uniform in +-1/sqrt(fan_in), fan_in = weight.shape[1] * k) that the reference
  constructor applies (reference models/hifigan.py:177-222 builds plain
  nn.Conv1d / nn.ConvTranspose1d and never re-initialises them);
* mels are an Irwin-Hall(12) approximation of N(0,1), which is exact in float64
  and therefore bit-portable (reference tests use torch.randn mels:
  tests/test_hifigan_generator.py:56-61).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    """One splitmix64 finaliser round on a uint64 array (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = x + _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n float64 values in [0,1) with 24 random bits each."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([seed], dtype=np.uint64) * _GOLD
                           + np.array([stream], dtype=np.uint64) * _M1)
        idx = np.arange(n, dtype=np.uint64) + base
    z = _splitmix64(idx)
    return (z >> np.uint64(40)).astype(np.float64) * (1.0 / 16777216.0)


def uniform(seed: int, shape: Sequence[int], bound: float, stream: int = 0) -> np.ndarray:
    n = int(np.prod(shape)) if len(shape) else 1
    u = uniform01(seed, n, stream)
    return ((2.0 * u - 1.0) * bound).astype(np.float32).reshape(tuple(shape))


def normal(seed: int, shape: Sequence[int], stream: int = 0) -> np.ndarray:
    """Irwin-Hall(12) - 6: mean 0, variance 1, support [-6,6]; exact in float64."""
    n = int(np.prod(shape))
    acc = np.zeros(n, dtype=np.float64)
    for k in range(12):
        acc += uniform01(seed, n, stream * 16 + k + 1)
    return (acc - 6.0).astype(np.float32).reshape(tuple(shape))


# ----------------------------------------------------------------------------
# Generator geometry (shared by host module, tests and bench)
# ----------------------------------------------------------------------------

DEFAULT_CONFIG = dict(
    n_mels=80,
    upsample_rates=[8, 8, 2, 2],
    upsample_kernel_sizes=[16, 16, 4, 4],
    upsample_initial_channel=512,
    resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
)


def weight_shapes(cfg: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) for every tensor of the plain (156-key for the default
    config) state_dict schema, in the reference's registration order
    (reference models/hifigan.py:177-222, 47-70)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    c0 = cfg["upsample_initial_channel"]
    out.append(("conv_pre.weight", (c0, cfg["n_mels"], 7)))
    out.append(("conv_pre.bias", (c0,)))
    ups, mrfs = [], []
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        cin, cout = c0 // (2 ** i), c0 // (2 ** (i + 1))
        ups.append((f"ups.{i}.weight", (cin, cout, k)))
        ups.append((f"ups.{i}.bias", (cout,)))
        for j, (rk, dils) in enumerate(zip(cfg["resblock_kernel_sizes"],
                                           cfg["resblock_dilation_sizes"])):
            for which in ("convs1", "convs2"):
                for l in range(len(dils)):
                    p = f"mrfs.{i}.resblocks.{j}.{which}.{l}"
                    mrfs.append((p + ".weight", (cout, cout, rk)))
                    mrfs.append((p + ".bias", (cout,)))
    out += ups + mrfs
    cl = c0 // (2 ** len(cfg["upsample_rates"]))
    out.append(("conv_post.weight", (1, cl, 7)))
    out.append(("conv_post.bias", (1,)))
    return out


def make_weights(cfg: dict, seed: int, gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded weights in the reference state_dict schema (numpy float32).

    gain > 1 scales every weight (not bias) to push activations up; used by the
    tanh-saturation parity case."""
    sd: Dict[str, np.ndarray] = {}
    bound = 1.0
    for idx, (name, shape) in enumerate(weight_shapes(cfg)):
        if name.endswith(".weight"):
            fan_in = shape[1] * shape[2]
            bound = 1.0 / math.sqrt(fan_in)
            sd[name] = uniform(seed, shape, bound * gain, stream=idx)
        else:
            sd[name] = uniform(seed, shape, bound, stream=idx)
    return sd


def make_mel(seed: int, batch: int, n_mels: int, frames: int) -> np.ndarray:
    return normal(seed, (batch, n_mels, frames), stream=7)


def out_length(cfg: dict, frames: int) -> int:
    """Waveform length for `frames` mel frames: every ConvTranspose1d maps
    T -> (T-1)*u - 2*((k-u)//2) + k (reference models/hifigan.py:196-202)."""
    t = frames
    for u, k in zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"]):
        t = (t - 1) * u - 2 * ((k - u) // 2) + k
    return t


def flops_per_frame(cfg: dict) -> float:
    """Algorithmic FLOPs (2*MAC) per mel frame, asymptotic in T (SURVEY.md §8d:
    614.105 MFLOP for the default config)."""
    c0 = cfg["upsample_initial_channel"]
    mac = cfg["n_mels"] * c0 * 7
    scale = 1.0
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        cin, cout = c0 // (2 ** i), c0 // (2 ** (i + 1))
        mac += scale * cin * cout * k          # per input step: cin*cout*k MACs
        scale *= u
        for rk, dils in zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]):
            mac += scale * 2 * len(dils) * cout * cout * rk
    cl = c0 // (2 ** len(cfg["upsample_rates"]))
    mac += scale * cl * 7
    return 2.0 * mac


def make_weightnorm_weights(cfg: dict, seed: int) -> Dict[str, np.ndarray]:
    """The 232-key weight-normed schema (reference models/hifigan.py:274-283:
    `weight` -> `weight_g` [C0,1,1] + `weight_v` on ups / convs1 / convs2 only),
    with g perturbed by a seeded factor in [0.75, 1.25] so that folding
    g * v / ||v|| is not the identity.  Mirrors tests/golden/make_golden.py."""
    plain = make_weights(cfg, seed)
    out: Dict[str, np.ndarray] = {}
    g_keys = []
    for name, w in plain.items():
        if name.endswith(".weight") and not name.startswith(("conv_pre", "conv_post")):
            base = name[: -len("weight")]
            norm = np.sqrt((w.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
            out[base + "weight_g"] = norm.astype(np.float32)
            out[base + "weight_v"] = w
            g_keys.append(base + "weight_g")
        else:
            out[name] = w
    for i, k in enumerate(sorted(g_keys)):
        scale = uniform(seed, out[k].shape, 0.25, stream=1000 + i) + np.float32(1.0)
        out[k] = (out[k] * scale).astype(np.float32)
    return out


# ---------------------------------------------------------------------------
# autoregressive decoder (reference models/ar_decoder.py): portable weights
# ---------------------------------------------------------------------------
ARD_DEFAULT = dict(d_model=256, n_mels=80, n_layers=6, n_heads=8, d_ff=2048)


def ard_weight_shapes(cfg: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """Parameter keys / shapes of the reference PNCAARDecoder's state_dict, in its order, without the
    `pos_encoding.pe` buffer (which the module computes itself)."""
    d, m, ff = cfg["d_model"], cfg["n_mels"], cfg["d_ff"]
    out = [("prenet.0.weight", (d, m)), ("prenet.0.bias", (d,)), ("prenet.3.weight", (d, d)), ("prenet.3.bias", (d,))]
    for i in range(cfg["n_layers"]):
        L = f"decoder.layers.{i}."
        for a in ("self_attn.", "multihead_attn."):
            out += [(L + a + "in_proj_weight", (3 * d, d)), (L + a + "in_proj_bias", (3 * d,)),
                    (L + a + "out_proj.weight", (d, d)), (L + a + "out_proj.bias", (d,))]
        out += [(L + "linear1.weight", (ff, d)), (L + "linear1.bias", (ff,)),
                (L + "linear2.weight", (d, ff)), (L + "linear2.bias", (d,))]
        for n in ("norm1.", "norm2.", "norm3."):
            out += [(L + n + "weight", (d,)), (L + n + "bias", (d,))]
    out += [("mel_proj.weight", (m, d)), ("mel_proj.bias", (m,))]
    return out


def make_ard_weights(cfg: dict, seed: int) -> Dict[str, np.ndarray]:
    """Xavier-uniform matrices (the reference's own init, models/ar_decoder.py:91-95), small non-zero biases and
    LayerNorm gains around 1 -- from the portable generator, so both boxes build the same decoder."""
    sd = {}
    for idx, (name, shape) in enumerate(ard_weight_shapes(cfg)):
        if len(shape) == 2:
            bound = float(np.sqrt(6.0 / (shape[0] + shape[1])))
            sd[name] = uniform(seed, shape, bound, stream=idx + 1)
        elif ".norm" in name and name.endswith("weight"):
            sd[name] = (1.0 + uniform(seed, shape, 0.1, stream=idx + 1)).astype(np.float32)
        else:
            sd[name] = uniform(seed, shape, 0.05, stream=idx + 1)
    return sd
