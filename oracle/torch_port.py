"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Functional restatement of the reference generator forward pass on the very ATen
operators the reference dispatches to (F.conv1d / F.conv_transpose1d /
F.leaky_relu / torch.tanh; reference models/hifigan.py:52-69,81-85,196-202,
238-256), driven by a flat state_dict instead of an nn.Module tree.  Because it
runs the same CPU kernels as the reference, it is also what bench.py times as
the CPU baseline on the GPU box (where /root/reference does not exist).

Pinned against the live reference module by tests/golden/make_golden.py +
tests/test_oracle.py (the reference's own tests pin no output value for this
path -- SURVEY.md section 8c).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """weight = g * v / ||v||, norm over every dim but 0 -- what
    nn.utils.weight_norm(dim=0) computes (reference models/hifigan.py:274-283
    applies it to ups / convs1 / convs2 only).  For ConvTranspose1d dim 0 is C_in."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            base = k[: -len("_g")]
            vv = sd[base + "_v"]
            norm = vv.reshape(vv.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (vv.dim() - 1)))
            out[base] = v * vv / norm
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


def forward_torch(cfg: dict, sd: Dict[str, torch.Tensor], mel: torch.Tensor,
                  stages: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """mel [B, n_mels, T] -> wav [B, 1, T_out]; dtype/device follow the inputs."""
    if any(k.endswith("weight_g") for k in sd):
        sd = fold_weight_norm(sd)
    slope = 0.1
    x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)      # :238
    if stages is not None:
        stages.append(x)
    n_rb = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.conv_transpose1d(F.leaky_relu(x, slope), sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"],
                               stride=u, padding=(k - u) // 2)                   # :244-245
        if stages is not None:
            stages.append(x)
        total = None
        for j, (rk, dils) in enumerate(zip(cfg["resblock_kernel_sizes"],
                                           cfg["resblock_dilation_sizes"])):
            r = x
            p = f"mrfs.{i}.resblocks.{j}."
            for l, d in enumerate(dils):                                          # :80-85
                h = F.conv1d(F.leaky_relu(r, slope), sd[p + f"convs1.{l}.weight"],
                             sd[p + f"convs1.{l}.bias"], dilation=d, padding=(rk * d - d) // 2)
                h = F.conv1d(F.leaky_relu(h, slope), sd[p + f"convs2.{l}.weight"],
                             sd[p + f"convs2.{l}.bias"], dilation=1, padding=(rk - 1) // 2)
                r = r + h
            total = r if total is None else total + r                            # :126-129
        x = total / n_rb                                                          # :131
        if stages is not None:
            stages.append(x)
    x = F.conv1d(F.leaky_relu(x, slope), sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(x)                                                          # :254-256
