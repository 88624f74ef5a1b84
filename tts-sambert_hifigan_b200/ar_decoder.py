"""Host-side mirror of the reference PNCAARDecoder for inference.

Same constructor, attribute names and state_dict schema as the reference class
(reference models/ar_decoder.py:14-95: `prenet.{0,3}`, `pos_encoding.pe`, `decoder.layers.{i}.*` of
torch.nn.TransformerDecoder, `mel_proj`), so `new.load_state_dict(old.state_dict())` is the migration.
`forward(Hvar, mel_gt=None, max_len=None)` in eval mode runs the KV-cached decode of libhfg_b200.so
(include/hfg_ard.h): the same frames as the reference's O(T^2) loop (models/ar_decoder.py:167-238) to fp32
round-off, one decoder position per step.  The torch submodules here only hold parameters; no torch op runs on
the data path.  Training (teacher forcing, :120-165) is out of contract."""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _capi


class _PositionalEncoding(nn.Module):
    """Buffer holder with the reference's key (`pos_encoding.pe`, reference models/ar_decoder.py:280-312)."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class PNCAARDecoder(nn.Module):
    """B200-native drop-in for the reference PNCAARDecoder in inference mode.

    Shape contract (reference models/ar_decoder.py:28-35): Hvar [B, Tfrm, d_model] -> mel_pred [B, Tfrm, n_mels].
    `verbose=True` reproduces the reference's unconditional progress prints (:187-236)."""

    def __init__(self, d_model=256, n_mels=80, n_layers=6, n_heads=8, d_ff=2048, dropout=0.1, chunk_size=1,
                 *, verbose: bool = True):
        super().__init__()
        self.d_model, self.n_mels, self.n_layers, self.n_heads, self.chunk_size = d_model, n_mels, n_layers, n_heads, chunk_size
        self.d_ff = d_ff
        self.verbose = verbose
        self.prenet = nn.Sequential(nn.Linear(n_mels, d_model), nn.ReLU(), nn.Dropout(dropout), nn.Linear(d_model, d_model))
        self.pos_encoding = _PositionalEncoding(d_model, dropout, max_len=5000)
        layer = nn.TransformerDecoderLayer(d_model=d_model, nhead=n_heads, dim_feedforward=d_ff, dropout=dropout,
                                           activation="relu", batch_first=True)
        self.decoder = nn.TransformerDecoder(layer, num_layers=n_layers)
        self.mel_proj = nn.Linear(d_model, n_mels)
        for p in self.parameters():                         # reference :91-95
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self._handles: Dict[int, "_Handle"] = {}
        self._synced: Dict[int, tuple] = {}
        self._workspaces: Dict[int, torch.Tensor] = {}
        self.last_launch_count = 0

    @classmethod
    def from_reference(cls, ref: nn.Module, **kw) -> "PNCAARDecoder":
        l0 = ref.decoder.layers[0]
        new = cls(d_model=ref.d_model, n_mels=ref.n_mels, n_layers=ref.n_layers, n_heads=ref.n_heads,
                  d_ff=l0.linear1.out_features, chunk_size=ref.chunk_size, **kw)
        new.load_state_dict(ref.state_dict())
        return new.eval()

    def _handle_for(self, dev: torch.device) -> "_Handle":
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            with torch.cuda.device(idx):
                h = _Handle(self.d_model, self.n_mels, self.n_layers, self.n_heads, self.d_ff, self.pos_encoding.pe.shape[1])
            self._handles[idx] = h
        sig = tuple((k, v.data_ptr(), v._version) for k, v in self.state_dict(keep_vars=True).items())
        if self._synced.get(idx) != sig:
            keep = []
            for name, t in self.state_dict().items():
                c = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
                keep.append(c)
                h.set_weight(name, c.data_ptr(), list(c.shape))
            with torch.cuda.device(idx):
                h.commit()
            self._synced[idx] = sig
        return h

    def forward(self, Hvar: torch.Tensor, mel_gt: Optional[torch.Tensor] = None, max_len: Optional[int] = None) -> torch.Tensor:
        if self.training and mel_gt is not None:
            raise NotImplementedError("the B200 PNCAARDecoder is inference-only (teacher forcing is out of contract)")
        if Hvar.dim() != 3 or Hvar.shape[2] != self.d_model:
            raise RuntimeError(f"expected Hvar of shape [B, Tfrm, {self.d_model}], got {list(Hvar.shape)}")
        if not Hvar.is_cuda:
            raise RuntimeError("PNCAARDecoder (B200) needs a CUDA tensor: there is no CPU fallback")
        B, Tfrm, _ = Hvar.shape
        if max_len is None:
            max_len = Tfrm
        if self.verbose:                                    # reference :187-192
            print(f"[PNCAARDecoder] Inference mode - Input Hvar shape: {Hvar.shape}")
            print(f"[PNCAARDecoder] Generating {max_len} frames autoregressively with chunk_size={self.chunk_size}")
            print(f"[PNCAARDecoder] Initial mel_pred shape: {torch.Size((B, 1, self.n_mels))}")
        dev = Hvar.device
        x = Hvar.contiguous().float()
        with torch.cuda.device(dev):
            h = self._handle_for(dev)
            need = h.workspace_bytes(B, Tfrm, max_len)
            ws = self._workspaces.get(dev.index)
            if ws is None or ws.numel() < need:
                ws = torch.empty(need, dtype=torch.uint8, device=dev)
                self._workspaces[dev.index] = ws
            mel = torch.empty((B, max_len, self.n_mels), dtype=torch.float32, device=dev)
            h.decode(x.data_ptr(), B, Tfrm, max_len, mel.data_ptr(), ws.data_ptr(), ws.numel(),
                     torch.cuda.current_stream(dev).cuda_stream)
            self.last_launch_count = h.last_launch_count()
        if self.verbose:                                    # reference :221-236
            done, chunk = 0, 0
            while done < max_len:
                n = min(self.chunk_size, max_len - done)
                done += n
                print(f"[PNCAARDecoder] Chunk {chunk}: Generated {n} frames, current shape: {torch.Size((B, done + 1, self.n_mels))}")
                chunk += 1
            print(f"[PNCAARDecoder] Final output mel_pred shape: {mel.shape}")
            print(f"[PNCAARDecoder] Total chunks generated: {chunk}")
        return mel


class _ArdConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("d_model", "n_mels", "n_layers", "n_heads", "d_ff", "max_pos")]


class _Handle:
    """Owns one hfg_ard_handle."""

    def __init__(self, d_model, n_mels, n_layers, n_heads, d_ff, max_pos):
        self._lib = _capi.load()
        self._h = ctypes.c_void_p()
        cfg = _ArdConfig(d_model, n_mels, n_layers, n_heads, d_ff, max_pos)
        rc = self._lib.hfg_ard_create(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != _capi.OK:
            raise _capi.HfgError(rc, {_capi.ERR_CUDA: "no usable CUDA device (this path has no CPU fallback)",
                                      _capi.ERR_UNSUPPORTED: "head_dim must be 16, 32, 64 or 128"}.get(rc, "invalid decoder configuration"))

    def _check(self, rc):
        if rc != _capi.OK:
            raise _capi.HfgError(rc, self._lib.hfg_ard_last_error(self._h).decode())

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.hfg_ard_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    def set_weight(self, name, host_ptr, shape):
        shp = (ctypes.c_int64 * len(shape))(*shape)
        self._check(self._lib.hfg_ard_set_weight(self._h, name.encode(), ctypes.c_void_p(host_ptr), shp, len(shape)))

    def commit(self):
        self._check(self._lib.hfg_ard_commit_weights(self._h))

    def workspace_bytes(self, batch, frames, max_len):
        out = ctypes.c_size_t()
        self._check(self._lib.hfg_ard_workspace_bytes(self._h, batch, frames, max_len, ctypes.byref(out)))
        return out.value

    def decode(self, hvar_ptr, batch, frames, max_len, mel_ptr, ws_ptr, ws_bytes, stream):
        self._check(self._lib.hfg_ard_decode(self._h, hvar_ptr, batch, frames, max_len, mel_ptr, ws_ptr, ws_bytes,
                                             ctypes.c_void_p(stream)))

    def last_launch_count(self):
        out = ctypes.c_int64()
        self._check(self._lib.hfg_ard_last_launch_count(self._h, ctypes.byref(out)))
        return out.value
