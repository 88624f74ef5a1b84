"""Generate the committed golden vectors from the LIVE reference module.

Run here (the container with /root/reference); the vectors travel, the
reference does not:

    python tests/golden/make_golden.py

Each case builds models.hifigan.HiFiGANGenerator (reference
models/hifigan.py:134-283) the way tests/test_hifigan_generator.py:26-33 does
(kwargs from configs/*.yaml for the default cases), loads the seeded synthetic
weights from tts-sambert_hifigan_b200/synth.py through load_state_dict, runs
forward(mel) on CPU in fp32 under no_grad and stores the waveform (and, for the
'stages' case, time-subsampled stage-boundary activations taken with forward
hooks).  Inputs are NOT stored: they are regenerated from the seeds.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import tts_sambert_hifigan_b200 as pkg            # noqa: E402
from tts_sambert_hifigan_b200 import synth        # noqa: E402
from models.hifigan import HiFiGANGenerator, HiFiGAN  # noqa: E402  (the reference)

STAGE_STRIDE = 5


def yaml_cfg():
    with open(os.path.join(REF, "configs/config.yaml")) as f:
        audio = yaml.safe_load(f)
    with open(os.path.join(REF, "configs/model_config.yaml")) as f:
        model = yaml.safe_load(f)
    g = model["vocoder"]["generator"]
    return dict(n_mels=audio["audio"]["n_mels"], upsample_rates=g["upsample_rates"],
                upsample_kernel_sizes=g["upsample_kernel_sizes"],
                upsample_initial_channel=g["upsample_initial_channel"],
                resblock_kernel_sizes=g["resblock_kernel_sizes"],
                resblock_dilation_sizes=g["resblock_dilation_sizes"])


CASES = {
    # name: (cfg or None for YAML default, weight seed, mel seed, B, T, extras)
    "default_b2_t24": (None, 11, 12, 2, 24, {}),
    "default_stages_b1_t9": (None, 21, 22, 1, 9, {"stages": True}),
    "default_config1_b1_t256": (None, 0, 1, 1, 256, {}),
    "default_weightnorm_b1_t16": (None, 31, 32, 1, 16, {"weight_norm": True}),
    "default_saturated_b1_t16": (None, 41, 42, 1, 16, {"gain": 2.25}),
    "default_ragged_b3_t7": (None, 51, 52, 3, 7, {}),
    "odd_upsample_b1_t20": (dict(n_mels=80, upsample_rates=[5, 5, 4, 2],
                                 upsample_kernel_sizes=[10, 10, 8, 4],
                                 upsample_initial_channel=512,
                                 resblock_kernel_sizes=[3, 7, 11],
                                 resblock_dilation_sizes=[[1, 3, 5]] * 3), 61, 62, 1, 20, {}),
    "small_custom_b3_t33": (dict(n_mels=16, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4],
                                 upsample_initial_channel=64, resblock_kernel_sizes=[3, 5],
                                 resblock_dilation_sizes=[[1, 2], [1, 3]]), 71, 72, 3, 33, {}),
}


def run_case(name, spec):
    out, meta = compute_case(name, spec)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, meta["wav_shape"], "absmax %.4f" % meta["wav_absmax"],
          "fp32-vs-fp64 %.2e" % meta["fp64_maxdiff"])
    return meta


def compute_case(name, spec):
    """Run one case on the live reference; returns (arrays of the .npz, manifest entry).  Also used by
    tests/test_goldens_are_live_cpu.py to re-derive committed vectors."""
    cfg, wseed, mseed, B, T, extra = spec
    cfg = cfg or yaml_cfg()
    gen = HiFiGANGenerator(**cfg).eval()
    sd = synth.make_weights(cfg, wseed, gain=extra.get("gain", 1.0))
    missing = gen.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    out = {}
    if extra.get("weight_norm"):
        # reference models/hifigan.py:274-283; then perturb g so folding is non-trivial
        gen.apply_weight_norm()
        wn_sd = gen.state_dict()
        i = 0
        for k in sorted(wn_sd):
            if k.endswith("weight_g"):
                scale = synth.uniform(wseed, wn_sd[k].shape, 0.25, stream=1000 + i) + 1.0
                wn_sd[k] = wn_sd[k] * torch.from_numpy(scale)
                i += 1
        gen.load_state_dict(wn_sd)
        out["n_state_keys"] = np.array(len(wn_sd))
    mel = torch.from_numpy(synth.make_mel(mseed, B, cfg["n_mels"], T))
    stages = []
    hooks = []
    if extra.get("stages"):
        mods = [gen.conv_pre]
        for u, m in zip(gen.ups, gen.mrfs):
            mods += [u, m]
        for m in mods:
            hooks.append(m.register_forward_hook(lambda _m, _i, o: stages.append(o.detach().clone())))
    with torch.no_grad():
        wav = gen(mel)
        wav64 = gen.double()(mel.double())
    for h in hooks:
        h.remove()
    out["wav"] = wav.numpy()
    out["wav_fp64_maxdiff"] = np.array(float((wav.double() - wav64).abs().max()))
    if extra.get("stages"):
        stages = stages[: len(stages) // 2]          # second half came from the fp64 run
        for i, s in enumerate(stages):
            out[f"stage{i}"] = s.numpy()[:, :, ::STAGE_STRIDE].copy()
            out[f"stage{i}_shape"] = np.array(s.shape)
    meta = dict(cfg=cfg, weight_seed=wseed, mel_seed=mseed, B=B, T=T, extra=extra,
                wav_shape=list(wav.shape), wav_absmax=float(wav.abs().max()),
                fp64_maxdiff=float(out["wav_fp64_maxdiff"]))
    return out, meta


def wrapper_case():
    """The boundary caller: reference HiFiGAN wrapper prints (models/hifigan.py:716-722)."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m = HiFiGAN(n_mels=80, upsample_rates=[8, 8, 2, 2], debug_shapes=True).eval()
        with torch.no_grad():
            m(torch.zeros(1, 80, 10))
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("[HiFiGAN")]
    return lines


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    manifest = {"torch": torch.__version__, "stage_stride": STAGE_STRIDE, "cases": {}}
    for name, spec in CASES.items():
        manifest["cases"][name] = run_case(name, spec)
    manifest["debug_print_lines"] = wrapper_case()
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


# The log-mel-L1 metric pin (manifest["logmel_l1_pin"]) is produced by running the reference's
# VocoderLoss(mel_config=metrics.AUDIO, loss_mode="mel_only").mel_reconstruction_loss(a, b)
# (reference models/losses.py:708-797) on a = 0.05*normal(101), b = a + 0.002*normal(102),
# shape [2,1,8192]; see tests/test_oracle.py::test_logmel_metric_matches_reference.
