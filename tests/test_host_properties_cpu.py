"""Property tests (hypothesis) of the host-side logic around the hot path, on random geometries:

* `hfg_receptive_radius` (host-only C entry point) against a closed-form interval propagation written
  independently here and, on small cases, against the radius observed on the oracle;
* `sharding.shard_bounds` / `plan_chunks`: cover, order, balance, halo clipping; time chunking with
  radius + 1 frames of halo reproduces the unchunked oracle exactly in fp64 (reference
  models/hifigan.py:224-261 has only local ops) and a halo below the radius is refused;
* the length-regulator oracle against `torch.repeat_interleave` (reference
  models/variance_adaptor.py:171-269) on random durations incl. zeros and negatives.

Sized to run in seconds; no GPU.
"""
import math

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle
from oracle import length_regulator as lr_oracle
from tts_sambert_hifigan_b200 import _capi, sharding, synth

# derandomize: the same examples on every run (the CPU suite must not depend on a random draw); wider sweeps: raise max_examples
COMMON = dict(deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow])
FAST = settings(max_examples=60, **COMMON)


@st.composite
def geometries(draw, max_stages=3, even_only=True):
    """Generator geometries the reference constructor accepts: k >= u; (k - u) even unless asked otherwise."""
    n_up = draw(st.integers(1, max_stages))
    rates, kernels = [], []
    for _ in range(n_up):
        u = draw(st.sampled_from([2, 3, 4, 8]))
        extra = draw(st.sampled_from([0, 2, 4] if even_only else [0, 1, 2, 3]))
        rates.append(u)
        kernels.append(u + extra if draw(st.booleans()) else 2 * u)
        if even_only and (kernels[-1] - u) % 2:
            kernels[-1] += 1
    n_rb = draw(st.integers(1, 3))
    rks = [draw(st.sampled_from([3, 5, 7, 11])) for _ in range(n_rb)]
    dils = [[draw(st.integers(1, 6)) for _ in range(draw(st.integers(1, 3)))] for _ in range(n_rb)]
    return dict(n_mels=8, upsample_rates=rates, upsample_kernel_sizes=kernels, upsample_initial_channel=16 * (2 ** n_up) // 2,
                resblock_kernel_sizes=rks, resblock_dilation_sizes=dils)


def closed_form_radius(cfg):
    """Propagate the sample interval one frame's output depends on, backwards through the network
    (conv_post k=7; per MRF the widest resblock, each pair reaching (d+1)(k-1)/2; ConvTranspose1d
    y[t] <- x[(t+p-j)/u], j in [0,k); conv_pre k=7) and express it in mel frames either side of the frame."""
    rates, kernels = cfg["upsample_rates"], cfg["upsample_kernel_sizes"]
    hop = int(np.prod(rates))
    lo, hi = 0, hop - 1                     # output samples of frame 0 (T_out == T*hop geometries)
    lo, hi = lo - 3, hi + 3                 # conv_post
    for i in reversed(range(len(rates))):
        reach = max(sum((d + 1) * (k - 1) // 2 for d in dils)
                    for k, dils in zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]))
        lo, hi = lo - reach, hi + reach     # MRF i
        u, k = rates[i], kernels[i]
        p = (k - u) // 2
        lo, hi = math.ceil((lo + p - (k - 1)) / u), math.floor((hi + p) / u)
    lo, hi = lo - 3, hi + 3                 # conv_pre
    return max(-lo, hi)


def observed_radius(cfg, frames, probe):
    sd = {k: torch.from_numpy(v).double() for k, v in synth.make_weights(cfg, 3).items()}
    sd = {k: (torch.zeros_like(v) if k.endswith(".bias") else v) for k, v in sd.items()}
    mel = torch.zeros(1, cfg["n_mels"], frames, dtype=torch.float64)
    mel[:, :, probe] = 1.0
    wav = oracle.forward_torch(cfg, sd, mel)
    hop = wav.shape[-1] // frames
    changed = (wav.abs() > 0).reshape(-1).nonzero().reshape(-1)
    return max(probe - int(changed.min()) // hop, int(changed.max()) // hop - probe)


@FAST
@given(geometries())
def test_receptive_radius_equals_closed_form(cfg):
    got = _capi.receptive_radius(_capi.make_config(**cfg))
    assert got == closed_form_radius(cfg), cfg


@settings(max_examples=10, **COMMON)
@given(geometries(max_stages=2))
def test_receptive_radius_equals_the_oracles_impulse_response(cfg):
    r = _capi.receptive_radius(_capi.make_config(**cfg))
    frames = 2 * r + 5
    assert observed_radius(cfg, frames, r + 2) == r, cfg


@FAST
@given(st.integers(0, 5000), st.integers(1, 64))
def test_shard_bounds_partition(n, world):
    spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        sharding.shard_bounds(n, world, world)


@FAST
@given(st.integers(1, 6000), st.integers(1, 16), st.integers(0, 40))
def test_plan_chunks_cover_and_clip(frames, n_chunks, halo):
    chunks = sharding.plan_chunks(frames, n_chunks, halo)
    assert len(chunks) == min(frames, n_chunks)
    assert chunks[0].start == 0 and chunks[-1].stop == frames
    for a, b in zip(chunks, chunks[1:]):
        assert a.stop == b.start
    for c in chunks:
        assert c.lo == max(0, c.start - halo) and c.hi == min(frames, c.stop + halo)
        assert 0 <= c.crop_front <= halo and c.frames > 0


@settings(max_examples=8, **COMMON)
@given(geometries(max_stages=2), st.integers(2, 5), st.integers(0, 2 ** 31 - 1))
def test_time_chunking_with_radius_plus_one_is_exact_in_fp64(cfg, n_chunks, seed):
    r = _capi.receptive_radius(_capi.make_config(**cfg))
    frames = n_chunks * (r + 3) + seed % 7
    sd = {k: torch.from_numpy(v).double() for k, v in synth.make_weights(cfg, seed % 1000).items()}
    mel = torch.from_numpy(synth.make_mel(seed % 1000 + 1, 1, cfg["n_mels"], frames)).double()
    hop = int(np.prod(cfg["upsample_rates"]))

    class Gen:                                   # a generator that knows its geometry, like the CUDA module
        receptive_radius = r

        def __call__(self, m):
            return oracle.forward_torch(cfg, sd, m)

    gen = Gen()
    full = gen(mel)
    got = sharding.generate_chunked(gen, mel, n_chunks, hop=hop)          # halo defaults to radius + 1
    assert got.shape == full.shape and float((got - full).abs().max()) == 0.0
    if r > 0:
        with pytest.raises(ValueError):
            sharding.generate_chunked(gen, mel, n_chunks, hop=hop, halo=r - 1)


@FAST
@given(st.integers(1, 5), st.integers(1, 12), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_length_regulator_oracle_equals_repeat_interleave(batch, n_ph, d_model, seed):
    rng = np.random.default_rng(seed)
    henc = rng.standard_normal((batch, n_ph, d_model)).astype(np.float32)
    dur = rng.integers(-2, 6, size=(batch, n_ph)).astype(np.int64)
    got = lr_oracle.length_regulate(henc, dur)
    rows = [torch.repeat_interleave(torch.from_numpy(henc[b]), torch.from_numpy(dur[b]).clamp(min=0), dim=0) for b in range(batch)]
    t_max = max(int(r.shape[0]) for r in rows)
    assert got.shape == (batch, t_max, d_model) or (t_max == 0 and got.shape[1] in (0, 1))
    for b, r in enumerate(rows):
        assert np.array_equal(got[b, :r.shape[0]], r.numpy())
        assert not got[b, r.shape[0]:].any()
