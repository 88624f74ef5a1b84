"""Host-side mirror of the reference generator module.

`HiFiGANGenerator` here has the constructor, attributes, `forward(mel)`
signature, debug prints and state_dict schema of the reference class
(reference models/hifigan.py:134-283) but owns no arithmetic: `forward` hands
the mel to libhfg_b200.so (include/hfg.h) on the caller's CUDA stream.  PyTorch
is used for what it is good at here -- parameter bookkeeping, device memory and
streams.

    hifigan = HiFiGAN(...)                        # the reference wrapper
    new = HiFiGANGenerator(...).cuda()
    new.load_state_dict(hifigan.generator.state_dict())
    hifigan.generator = new                       # reference :681-689 / :719

Inference only: there is no autograd through the CUDA path (the reference's
tests/test_hifigan_generator.py:146-169 gradient test is out of contract).
"""
from __future__ import annotations

import math
import os
import warnings
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _capi

_SLOPE = 0.1  # reference models/hifigan.py:81,83,244,254


class _ConvParams(nn.Module):
    """Parameter holder with the reference's key names.  `transposed` only
    changes which dim is the fan-in for the default init."""

    def __init__(self, w_shape: Sequence[int], n_bias: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*w_shape))
        self.bias = nn.Parameter(torch.empty(n_bias))
        # same distribution as nn.Conv1d / nn.ConvTranspose1d default init
        bound = 1.0 / math.sqrt(w_shape[1] * w_shape[2])
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            self.bias.uniform_(-bound, bound)

    # weight-norm reparametrisation kept as plain parameters (reference :274-283)
    def split_weight_norm(self):
        if "weight" not in self._parameters:
            return
        w = self._parameters.pop("weight")
        g = w.detach().reshape(w.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
        self.weight_g = nn.Parameter(g)
        self.weight_v = nn.Parameter(w.detach().clone())

    def fold_weight_norm(self):
        if "weight_g" not in self._parameters:
            return
        g = self._parameters.pop("weight_g")
        v = self._parameters.pop("weight_v")
        norm = v.detach().reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
        bias = self._parameters.pop("bias")
        self.weight = nn.Parameter(v.detach() * (g.detach() / norm))
        self.bias = bias            # keep the plain schema's key order: weight, bias


class _ResBlockParams(nn.Module):
    def __init__(self, channels: int, kernel_size: int, n_dil: int):
        super().__init__()
        self.convs1 = nn.ModuleList(_ConvParams((channels, channels, kernel_size), channels) for _ in range(n_dil))
        self.convs2 = nn.ModuleList(_ConvParams((channels, channels, kernel_size), channels) for _ in range(n_dil))


class _MRFParams(nn.Module):
    def __init__(self, channels: int, kernel_sizes: Sequence[int], dilation_sizes: Sequence[Sequence[int]]):
        super().__init__()
        self.resblocks = nn.ModuleList(
            _ResBlockParams(channels, k, len(d)) for k, d in zip(kernel_sizes, dilation_sizes))


class HiFiGANGenerator(nn.Module):
    """B200-native drop-in for the reference HiFiGANGenerator.

    Shape contract (reference models/hifigan.py:144-147):
        mel [B, n_mels, Tfrm] float32  ->  wav [B, 1, T_wav] float32, T_wav = Tfrm * prod(rates)

    Extra keyword (not in the reference): `mode` in {"fp32", "tf32", "bf16", "fp16"} --
    arithmetic of the CUDA path (include/hfg.h hfg_mode); default from
    $HFG_MODE, else "tf32" (tensor cores, 10-bit-mantissa operands, fp32 accumulate, residual
    stream of >= 22 mantissa bits, parity <= 1e-3).
    "fp16" keeps tf32's 10-bit mantissa (parity <= 1e-3) at bf16's speed; its
    conversions saturate at +-65504.
    """

    def __init__(
        self,
        n_mels: int = 80,
        upsample_rates: List[int] = [8, 8, 2, 2],
        upsample_kernel_sizes: List[int] = [16, 16, 4, 4],
        upsample_initial_channel: int = 512,
        resblock_kernel_sizes: List[int] = [3, 7, 11],
        resblock_dilation_sizes: List[List[int]] = [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
        debug_shapes: bool = False,
        *,
        mode: Optional[str] = None,
    ):
        super().__init__()
        self.n_mels = n_mels
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)
        self.debug_shapes = debug_shapes or os.getenv("DEBUG_SHAPES", "0") == "1"   # reference :174
        self.mode = mode or os.getenv("HFG_MODE", "tf32")
        if self.mode not in _capi.MODES:
            raise ValueError(f"mode must be one of {sorted(_capi.MODES)}, got {self.mode!r}")
        self._geometry = dict(
            n_mels=n_mels, upsample_rates=list(upsample_rates),
            upsample_kernel_sizes=list(upsample_kernel_sizes),
            upsample_initial_channel=upsample_initial_channel,
            resblock_kernel_sizes=list(resblock_kernel_sizes),
            resblock_dilation_sizes=[list(d) for d in resblock_dilation_sizes])
        # validates list lengths / limits the way the C side will see them
        self._cfg = _capi.make_config(**self._geometry)

        c0 = upsample_initial_channel
        self.conv_pre = _ConvParams((c0, n_mels, 7), c0)
        self.ups = nn.ModuleList()
        self.mrfs = nn.ModuleList()
        for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
            cin, cout = c0 // (2 ** i), c0 // (2 ** (i + 1))
            self.ups.append(_ConvParams((cin, cout, k), cout))        # ConvTranspose1d: [C_in, C_out, k]
            self.mrfs.append(_MRFParams(cout, resblock_kernel_sizes, resblock_dilation_sizes))
        self.conv_post = _ConvParams((1, c0 // (2 ** self.num_upsamples), 7), 1)

        self._handles: Dict[int, _capi.Handle] = {}
        self._synced: Dict[int, tuple] = {}
        self._workspaces: Dict[int, torch.Tensor] = {}
        self.last_launch_count = 0

    # ------------------------------------------------------------------ weights
    def _normed_layers(self):
        for layer in self.ups:
            yield layer
        for mrf in self.mrfs:
            for rb in mrf.resblocks:
                yield from rb.convs1
                yield from rb.convs2

    def apply_weight_norm(self):
        """Reference models/hifigan.py:274-283: ups / convs1 / convs2 switch to the
        weight_g + weight_v schema (232 keys for the default config)."""
        for layer in self._normed_layers():
            layer.split_weight_norm()

    def remove_weight_norm(self):
        """Reference models/hifigan.py:263-272."""
        for layer in self._normed_layers():
            layer.fold_weight_norm()

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts both reference schemas: if the incoming dict is weight-normed
        and this module is not (or vice versa) the module is switched first."""
        has_g = any(k.endswith(".weight_g") for k in state_dict)
        mine_g = any(k.endswith(".weight_g") for k in self.state_dict())
        if has_g and not mine_g:
            self.apply_weight_norm()
        elif mine_g and not has_g:
            self.remove_weight_norm()
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    @classmethod
    def from_reference(cls, ref_module: nn.Module, **kw) -> "HiFiGANGenerator":
        """Build from a live reference HiFiGANGenerator (same geometry, same weights)."""
        ups = list(ref_module.ups)
        rbs = list(ref_module.mrfs[0].resblocks)
        new = cls(
            n_mels=ref_module.n_mels,
            upsample_rates=[m.stride[0] for m in ups],
            upsample_kernel_sizes=[m.kernel_size[0] for m in ups],
            upsample_initial_channel=ref_module.conv_pre.out_channels,
            resblock_kernel_sizes=[rb.convs1[0].kernel_size[0] for rb in rbs],
            resblock_dilation_sizes=[[c.dilation[0] for c in rb.convs1] for rb in rbs],
            debug_shapes=ref_module.debug_shapes, **kw)
        new.load_state_dict(ref_module.state_dict())
        return new

    def _weights_signature(self) -> tuple:
        return tuple((k, p.data_ptr(), p._version) for k, p in self.named_parameters())

    def _handle_for(self, device: torch.device) -> _capi.Handle:
        """One C handle per CUDA device; weights are (re)committed when any
        parameter changed since the last commit."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            with torch.cuda.device(idx):
                h = _capi.Handle(self._cfg)
            self._handles[idx] = h
        sig = self._weights_signature()
        if self._synced.get(idx) != sig:
            keep = []
            for name, p in self.named_parameters():
                t = p.detach().to(device="cpu", dtype=torch.float32).contiguous()
                keep.append(t)
                h.set_weight(name, t.data_ptr(), list(t.shape))
            with torch.cuda.device(idx):
                h.commit()
            self._synced[idx] = sig
        return h

    # ------------------------------------------------------------------ forward
    def _stage_shapes(self, batch: int, frames: int):
        c, t = self._geometry["upsample_initial_channel"], frames
        shapes = [(batch, c, t)]
        for u, k in zip(self._geometry["upsample_rates"], self._geometry["upsample_kernel_sizes"]):
            c //= 2
            t = (t - 1) * u - 2 * ((k - u) // 2) + k
            shapes.append((batch, c, t))
        return shapes

    def _check_input(self, mel: torch.Tensor):
        if mel.dim() != 3 or mel.shape[1] != self.n_mels:
            raise RuntimeError(
                f"expected mel of shape [B, {self.n_mels}, Tfrm], got {list(mel.shape)}")
        if mel.shape[0] == 0 or mel.shape[2] == 0:
            raise RuntimeError(f"empty mel {list(mel.shape)}: batch and frame count must be positive")
        if mel.dtype != torch.float32:
            raise RuntimeError(f"expected float32 mel, got {mel.dtype}")
        if mel.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError(
                "the B200 generator is inference-only: call it under torch.no_grad() "
                "(no autograd through the CUDA path)")

    def forward_frames_last(self, mel_pred: torch.Tensor) -> torch.Tensor:
        """mel_pred [B, Tfrm, n_mels] -- the layout SAMBERTAcousticModel.forward returns (reference
        models/acoustic_model.py:181-265) -- to wav [B, 1, T_wav].  Equivalent to
        forward(mel_pred.transpose(1, 2)) (reference design.md:905-906) without the transpose copy:
        the first kernel reads the frames-last layout directly."""
        if mel_pred.dim() != 3 or mel_pred.shape[2] != self.n_mels:
            raise RuntimeError(f"expected mel_pred of shape [B, Tfrm, {self.n_mels}], got {list(mel_pred.shape)}")
        # shape checks / dispatch work on the logical [B, n_mels, T] view; the buffer stays [B, T, C]
        return self.forward(mel_pred.contiguous().transpose(1, 2), _frames_last=True)

    @property
    def receptive_radius(self) -> int:
        """Mel frames on either side that one output frame depends on (13 for the default configuration),
        derived from the kernel sizes / dilations / rates by the library (hfg_receptive_radius)."""
        return _capi.receptive_radius(self._cfg)

    @torch.no_grad()
    def forward_ragged(self, mel: torch.Tensor, lengths, halo: Optional[int] = None) -> torch.Tensor:
        """Variable-length batch: mel [B, n_mels, Tmax] padded with whatever the producer left there,
        `lengths` the valid frame counts.  The reference has no masks and synthesises every utterance out to
        the batch maximum (reference models/variance_adaptor.py:240-264); here the
        kernels' tile schedulers skip every tile beyond (length + halo) frames (hfg_forward_lengths), in one
        launch sequence for the whole batch.  Samples [0, len*hop) of every utterance are bit-identical to
        forward(mel) of the same padded batch -- `halo` defaults to receptive_radius + 1 and smaller values are
        rejected -- and everything beyond is returned as zeros."""
        self._check_input(mel)
        if not mel.is_cuda:
            raise RuntimeError("forward_ragged needs a CUDA mel (the length table is built on the device)")
        B, _, T = mel.shape
        lens = torch.as_tensor(lengths)
        if lens.numel() != B:
            raise RuntimeError("lengths must hold one frame count per utterance")
        if int(lens.min()) < 1 or int(lens.max()) > T:
            raise RuntimeError("lengths must hold one frame count in [1, Tmax] per utterance")
        radius = self.receptive_radius
        halo = radius + 1 if halo is None else int(halo)
        if halo < radius:
            raise ValueError(f"halo={halo} is smaller than the receptive radius ({radius} frames) of this "
                             "configuration: the valid region would not match the full-length run")
        dev = mel.device
        lens_dev = lens.to(device=dev, dtype=torch.int32).contiguous()
        shapes = self._stage_shapes(B, T)
        mode = _capi.MODES[self.mode]
        buf = mel.contiguous()
        with torch.cuda.device(dev):
            h = self._handle_for(dev)
            h.set_mel_layout(False)
            ws = self._workspace(h, dev, B, T, mode)
            wav = torch.empty((B, 1, shapes[-1][2]), dtype=torch.float32, device=dev)
            h.forward_lengths(buf.data_ptr(), lens_dev.data_ptr(), halo, B, T, wav.data_ptr(), ws.data_ptr(),
                              ws.numel(), mode, torch.cuda.current_stream(dev).cuda_stream)
            self.last_launch_count = h.last_launch_count()
        return wav

    def _workspace(self, h, dev, B, T, mode):
        need = h.workspace_bytes(B, T, mode)
        ws = self._workspaces.get(dev.index)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            self._workspaces[dev.index] = ws
        return ws

    @torch.no_grad()
    def generate_stream(self, mels):
        """Batch after batch from host memory at the device rate: `mels` is an iterable of CPU float32 tensors
        [B, n_mels, Tfrm] (page-locked ones are used in place, others are staged into pinned memory first); yields
        one page-locked CPU waveform [B, 1, T_wav] per input, in order.  Two submissions are kept in flight
        (hfg_forward_host_submit / _wait), so the H2D copy of batch i+1 and the D2H copy of batch i-1 run under
        the kernels of batch i.  Results are identical to forward(mel_cpu)."""
        if not torch.cuda.is_available():
            raise RuntimeError("HiFiGANGenerator (B200) needs a CUDA device: there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        h = self._handle_for(dev)
        h.set_mel_layout(False)
        mode = _capi.MODES[self.mode]
        inflight = [None, None]                              # per slot: (pinned mel kept alive, pinned wav)
        order = []

        def finish(slot):
            h.forward_host_wait(slot)
            _, wav = inflight[slot]
            inflight[slot] = None
            return wav

        try:
            for i, mel in enumerate(mels):
                self._check_input(mel)
                if mel.is_cuda:
                    raise RuntimeError("generate_stream takes host tensors (use forward for CUDA tensors)")
                slot = i & 1
                if inflight[slot] is not None:
                    order.pop(0)
                    yield finish(slot)
                src = mel.contiguous()
                if not src.is_pinned():
                    src = src.pin_memory()
                B, _, T = src.shape
                wav = torch.empty((B, 1, self._stage_shapes(B, T)[-1][2]), dtype=torch.float32, pin_memory=True)
                h.forward_host_submit(slot, src.data_ptr(), B, T, wav.data_ptr(), mode)
                inflight[slot] = (src, wav)
                order.append(slot)
            while order:
                yield finish(order.pop(0))
        finally:
            for slot in (0, 1):                              # a consumer that stops early must not leave copies in flight
                if inflight[slot] is not None:
                    h.forward_host_wait(slot)
                    inflight[slot] = None
        self.last_launch_count = h.last_launch_count()

    def forward(self, mel: torch.Tensor, _stages: Optional[list] = None, _frames_last: bool = False) -> torch.Tensor:
        """Generate waveform from mel-spectrogram (reference models/hifigan.py:224-261).

        CUDA mel: runs asynchronously on the current stream, returns a CUDA tensor.
        CPU mel: staged through pinned memory to the current CUDA device and back
        (hfg_forward_host); returns a CPU tensor.  Without a CUDA device this
        raises -- there is no CPU implementation."""
        self._check_input(mel)
        B, _, T = mel.shape
        shapes = self._stage_shapes(B, T)
        if self.debug_shapes:
            print(f"[HiFiGANGenerator] Input mel shape: {mel.shape}")
            print(f"[HiFiGANGenerator] After conv_pre: {torch.Size(shapes[0])}")
            for i in range(self.num_upsamples):
                print(f"[HiFiGANGenerator] After upsample {i}: {torch.Size(shapes[i + 1])}")
                print(f"[HiFiGANGenerator] After MRF {i}: {torch.Size(shapes[i + 1])}")
        try:
            wav = self._dispatch(mel, shapes, _capi.MODES[self.mode], _stages, _frames_last)
        except _capi.HfgError as e:
            # geometry the UMMA shapes do not cover (channel counts not multiples of 16): the fp32
            # CUDA kernels handle any geometry.  Still the GPU -- there is no CPU path.
            if e.code != _capi.ERR_UNSUPPORTED or self.mode == "fp32":
                raise
            warnings.warn(f"HiFiGANGenerator: mode {self.mode!r} unsupported for this geometry "
                          f"({e}); using the fp32 CUDA kernels")
            self.mode = "fp32"
            if _stages is not None:
                _stages.clear()
            wav = self._dispatch(mel, shapes, _capi.MODES["fp32"], _stages, _frames_last)
        if self.debug_shapes:
            print(f"[HiFiGANGenerator] Output wav shape: {wav.shape}")
        return wav

    def _dispatch(self, mel, shapes, mode, stages, frames_last):
        if not torch.cuda.is_available():
            raise RuntimeError("HiFiGANGenerator (B200) needs a CUDA device: there is no CPU fallback")
        if frames_last:
            buf = mel.transpose(1, 2)                   # back to the caller's contiguous [B, T, C] buffer
            assert buf.is_contiguous()
        else:
            buf = mel.contiguous()
        if mel.is_cuda:
            return self._forward_cuda(buf, shapes, mode, stages, mel.shape, frames_last)
        return self._forward_host(buf, shapes, mode, mel.shape, frames_last)

    def _forward_cuda(self, mel, shapes, mode, stages, logical_shape, frames_last):
        dev = mel.device
        B, _, T = logical_shape
        with torch.cuda.device(dev):
            h = self._handle_for(dev)
            h.set_mel_layout(frames_last)
            ws = self._workspace(h, dev, B, T, mode)
            wav = torch.empty((B, 1, shapes[-1][2]), dtype=torch.float32, device=dev)
            stage_ptrs = None
            if stages is not None:
                for i, s in enumerate(shapes):
                    for _ in range(1 if i == 0 else 2):
                        stages.append(torch.empty(s, dtype=torch.float32, device=dev))
                stage_ptrs = [s.data_ptr() for s in stages]
            h.forward(mel.data_ptr(), B, T, wav.data_ptr(), ws.data_ptr(), ws.numel(), mode,
                      torch.cuda.current_stream(dev).cuda_stream, stage_ptrs)
            self.last_launch_count = h.last_launch_count()
        return wav

    def _forward_host(self, mel, shapes, mode, logical_shape, frames_last):
        dev = torch.device("cuda", torch.cuda.current_device())
        B, _, T = logical_shape
        h = self._handle_for(dev)
        h.set_mel_layout(frames_last)
        # the result is written by DMA straight into a page-locked tensor (torch caches these)
        wav = torch.empty((B, 1, shapes[-1][2]), dtype=torch.float32, pin_memory=True)
        h.forward_host(mel.data_ptr(), B, T, wav.data_ptr(), mode,
                       mel_pinned=mel.is_pinned(), wav_pinned=True)
        self.last_launch_count = h.last_launch_count()
        return wav
