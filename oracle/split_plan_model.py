"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Arithmetic model of the tf32 mode's *split plan* (csrc/tc_path.cuh tc_tf32_mixed):
the same network as torch_port.forward_torch (reference models/hifigan.py:224-261), with every rounding the
CUDA path performs placed where it performs it --

  * every MMA operand is fp16 (round to nearest): the packed weights, the mel, conv_pre's output, the
    on-chip intermediate H = lrelu(conv1 + b1) of a fused pair, an MRF output that feeds an upsampler;
  * products are accumulated in fp32 (modelled as an exact sum rounded once: float64 convolution -> float32);
  * the residual stream is stored as the fp16 pair hi = fp16(s), lo = fp16(s - hi) of s = lrelu(x): conv1 reads hi,
    the residual is lrelu_inv(hi + lo); finished resblock outputs (MRF sum) and the last MRF output (conv_post)
    are stored in fp32;
  * acc2 is pre-loaded with x + b2 (+ the other resblocks' outputs), conv2 accumulates on top, the closing pair
    scales by 1 / n_resblocks, everything is stored as leaky_relu(.) and inverted by min(a, 10 a).

It is the *specification* of the mode and it explains the mode's error against the reference: operand rounding,
nothing else (model vs reference 1.5e-5 ... 5.5e-5 on the goldens; CUDA vs reference 1.6e-5 ... 5.9e-5).  Two
realisations of one rounding scheme do not agree bit for bit -- fp32 summation order flips fp16 roundings, and
those flips ARE the error -- so the tests compare error LEVELS: tests/test_oracle.py bounds the model's error
against the reference goldens on CPU, tests/test_parity_gpu.py requires the CUDA path's max-abs / rms error
against the reference to stay within rounding statistics of the model's (a path that dropped the lo halves, the
fp16 mode, shows 2.6x).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .torch_port import fold_weight_norm

_SLOPE = torch.tensor(0.1, dtype=torch.float32)
_INV = torch.tensor(1.0, dtype=torch.float32) / _SLOPE            # 10.0f, as the kernels compute it


def _h(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> fp16 (round to nearest even, saturating like cvt.rn.satfinite) -> fp32."""
    return x.clamp(-65504.0, 65504.0).half().float()


def _lrelu(v):
    return torch.maximum(v, v * _SLOPE)


def _lrelu_inv(a):
    return torch.minimum(a, a * _INV)


def _acc32(fn, x, w, **kw):
    """Convolution of fp16-representable operands with fp32 accumulation, modelled as the exact sum rounded once."""
    return fn(x.double(), w.double(), None, **kw).float()


def _bf(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 (round to nearest even) -> fp32."""
    return x.bfloat16().float()


def forward_split_plan(cfg: dict, sd: Dict[str, torch.Tensor], mel: torch.Tensor) -> torch.Tensor:
    """mel [B, n_mels, T] fp32 -> wav [B, 1, T_out] fp32, rounding as HFG_MODE_TF32 does on its split plan."""
    return forward_mode_model(cfg, sd, mel, "tf32")


def forward_mode_model(cfg: dict, sd: Dict[str, torch.Tensor], mel: torch.Tensor, mode: str) -> torch.Tensor:
    """The same model for every tensor-core mode.  "tf32": the split plan described above.  "fp16" / "bf16": every plane
    -- the residual stream, the finished resblock outputs, the MRF outputs -- is stored in the operand dtype (no lo
    half, no fp32 planes); accumulation, pre-load and scaling as above; conv_post reads the 2-byte plane with fp32 FMAs."""
    assert mode in ("tf32", "fp16", "bf16")
    split = mode == "tf32"
    _h = _bf if mode == "bf16" else globals()["_h"]
    if any(k.endswith("weight_g") for k in sd):
        sd = fold_weight_norm(sd)
    sd = {k: v.float() for k, v in sd.items()}
    w16 = {k: _h(v) for k, v in sd.items() if k.endswith(".weight")}
    n_up = len(cfg["upsample_rates"])
    n_rb = len(cfg["resblock_kernel_sizes"])
    inv_div = torch.tensor(1.0, dtype=torch.float32) / torch.tensor(float(n_rb), dtype=torch.float32)

    # conv_pre: fp16 mel x fp16 weights; its output only feeds ups[0] -> stored as fp16(lrelu(.))
    v = _acc32(F.conv1d, _h(mel.float()), w16["conv_pre.weight"], padding=3) + sd["conv_pre.bias"].view(1, -1, 1)
    a16 = _h(_lrelu(v))
    y32 = None
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        v = _acc32(F.conv_transpose1d, a16, w16[f"ups.{i}.weight"], stride=u, padding=(k - u) // 2) \
            + sd[f"ups.{i}.bias"].view(1, -1, 1)
        s = _lrelu(v)
        x_hi = _h(s)
        x_lo = _h(s - x_hi) if split else torch.zeros_like(s)      # X: hi + lo (2-byte modes: hi only)
        finals = []
        closing = None
        for j, (rk, dils) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            hi, lo = x_hi, x_lo
            p = f"mrfs.{i}.resblocks.{j}."
            for l, d in enumerate(dils):
                last = l + 1 == len(dils)
                closes = last and j == n_rb - 1
                c1 = _acc32(F.conv1d, hi, w16[p + f"convs1.{l}.weight"], dilation=d, padding=(rk * d - d) // 2) \
                    + sd[p + f"convs1.{l}.bias"].view(1, -1, 1)
                hmid = _h(_lrelu(c1))                              # on-chip intermediate, fp16
                acc = _lrelu_inv(hi + lo) + sd[p + f"convs2.{l}.bias"].view(1, -1, 1)     # pre2: x + b2
                if closes:
                    for f32 in finals:                             # + the other resblocks' outputs (fp32 planes)
                        acc = acc + _lrelu_inv(f32)
                c2 = F.conv1d(hmid.double(), w16[p + f"convs2.{l}.weight"].double(), None, dilation=1,
                              padding=(rk - 1) // 2)
                xn = (acc.double() + c2).float()                   # conv2 accumulates on top of the pre-load
                if closes and n_rb > 1:
                    xn = xn * inv_div
                s = _lrelu(xn)
                if closes:
                    closing = s
                elif last:
                    finals.append(s if split else _h(s))           # fp32 plane (2-byte modes: operand dtype)
                else:
                    hi = _h(s)
                    lo = _h(s - hi) if split else torch.zeros_like(s)
        if i + 1 < n_up:
            a16 = _h(closing)                                      # MRF output feeds an upsampler: operand dtype only
        else:
            y32 = closing if split else _h(closing)                # last MRF output: fp32 (2-byte modes: operand dtype), read by conv_post
    y = F.conv1d(y32, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)     # fp32 FMAs on the device
    return torch.tanh(y)
