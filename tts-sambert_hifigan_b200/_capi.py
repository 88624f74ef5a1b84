"""ctypes binding of include/hfg.h (the C-ABI boundary).  No torch here."""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence

HFG_MAX_STAGES = 8
MODE_FP32, MODE_TF32, MODE_BF16 = 0, 1, 2
MODE_FP16 = 3
MODES = {"fp32": MODE_FP32, "tf32": MODE_TF32, "bf16": MODE_BF16, "fp16": MODE_FP16}

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_WORKSPACE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5

# every symbol include/hfg.h declares (tests check the .so exports them all)
SYMBOLS = [
    "hfg_abi_version", "hfg_create", "hfg_destroy", "hfg_last_error", "hfg_set_weight",
    "hfg_commit_weights", "hfg_out_len", "hfg_workspace_bytes", "hfg_forward",
    "hfg_forward_stages", "hfg_forward_host", "hfg_forward_host_ex", "hfg_last_launch_count",
    "hfg_set_profiling", "hfg_get_profile", "hfg_bench_layer", "hfg_set_mel_layout",
    "hfg_durations_from_log", "hfg_length_regulate_frames", "hfg_length_regulate",
    "hfg_forward_lengths", "hfg_receptive_radius", "hfg_forward_host_submit", "hfg_forward_host_wait",
    "hfg_tf32_plan",
]
# include/hfg_mel.h (on-device log-mel / log-mel L1)
MEL_SYMBOLS = ["hfg_mel_create", "hfg_mel_destroy", "hfg_mel_last_error", "hfg_mel_frames", "hfg_log_mel", "hfg_log_mel_l1"]
# include/hfg_ard.h (KV-cached autoregressive decoder)
ARD_SYMBOLS = [
    "hfg_ard_create", "hfg_ard_destroy", "hfg_ard_last_error", "hfg_ard_set_weight", "hfg_ard_commit_weights",
    "hfg_ard_workspace_bytes", "hfg_ard_decode", "hfg_ard_last_launch_count",
]


class HfgConfig(ctypes.Structure):
    _fields_ = [
        ("n_mels", ctypes.c_int32),
        ("num_upsamples", ctypes.c_int32),
        ("upsample_initial_channel", ctypes.c_int32),
        ("num_resblocks", ctypes.c_int32),
        ("upsample_rates", ctypes.c_int32 * HFG_MAX_STAGES),
        ("upsample_kernel_sizes", ctypes.c_int32 * HFG_MAX_STAGES),
        ("resblock_kernel_sizes", ctypes.c_int32 * HFG_MAX_STAGES),
        ("num_dilations", ctypes.c_int32 * HFG_MAX_STAGES),
        ("resblock_dilations", (ctypes.c_int32 * HFG_MAX_STAGES) * HFG_MAX_STAGES),
    ]


class HfgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"hfg error {code}: {msg}")
        self.code = code


# HFG_LIB_PATH: A/B-test another build of the same ABI (kernel tuning only)
_LIB_PATH = os.environ.get("HFG_LIB_PATH") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), "lib", "libhfg_b200.so")
_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load libhfg_b200.so.  Fails loudly if it has not been built: there is no
    fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(_LIB_PATH)
    vp, i32, i64p = ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.hfg_abi_version.restype = ctypes.c_int
    lib.hfg_create.restype = ctypes.c_int
    lib.hfg_create.argtypes = [ctypes.POINTER(HfgConfig), ctypes.POINTER(vp)]
    lib.hfg_destroy.restype = None
    lib.hfg_destroy.argtypes = [vp]
    lib.hfg_last_error.restype = ctypes.c_char_p
    lib.hfg_last_error.argtypes = [vp]
    lib.hfg_set_weight.restype = ctypes.c_int
    lib.hfg_set_weight.argtypes = [vp, ctypes.c_char_p, vp, i64p, i32]
    lib.hfg_commit_weights.restype = ctypes.c_int
    lib.hfg_commit_weights.argtypes = [vp]
    lib.hfg_out_len.restype = ctypes.c_int
    lib.hfg_out_len.argtypes = [vp, i32, i64p]
    lib.hfg_workspace_bytes.restype = ctypes.c_int
    lib.hfg_workspace_bytes.argtypes = [vp, i32, i32, i32, ctypes.POINTER(ctypes.c_size_t)]
    lib.hfg_forward.restype = ctypes.c_int
    lib.hfg_forward.argtypes = [vp, vp, i32, i32, vp, vp, ctypes.c_size_t, i32, vp]
    lib.hfg_forward_stages.restype = ctypes.c_int
    lib.hfg_forward_stages.argtypes = [vp, vp, i32, i32, vp, vp, ctypes.c_size_t, i32, vp,
                                       ctypes.POINTER(vp)]
    lib.hfg_forward_host.restype = ctypes.c_int
    lib.hfg_forward_host.argtypes = [vp, vp, i32, i32, vp, i32]
    lib.hfg_forward_host_ex.restype = ctypes.c_int
    lib.hfg_forward_host_ex.argtypes = [vp, vp, i32, i32, vp, i32, ctypes.c_uint32]
    lib.hfg_last_launch_count.restype = ctypes.c_int
    lib.hfg_last_launch_count.argtypes = [vp, i64p]
    lib.hfg_tf32_plan.restype = ctypes.c_int
    lib.hfg_tf32_plan.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
    lib.hfg_set_profiling.restype = ctypes.c_int
    lib.hfg_set_profiling.argtypes = [vp, i32]
    lib.hfg_get_profile.restype = ctypes.c_int
    lib.hfg_get_profile.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    lib.hfg_set_mel_layout.restype = ctypes.c_int
    lib.hfg_set_mel_layout.argtypes = [vp, i32]
    lib.hfg_durations_from_log.restype = ctypes.c_int
    lib.hfg_durations_from_log.argtypes = [vp, ctypes.c_int64, vp, vp]
    lib.hfg_length_regulate_frames.restype = ctypes.c_int
    lib.hfg_length_regulate_frames.argtypes = [vp, i32, i32, i64p, vp]
    lib.hfg_length_regulate.restype = ctypes.c_int
    lib.hfg_length_regulate.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    lib.hfg_forward_host_submit.restype = ctypes.c_int
    lib.hfg_forward_host_submit.argtypes = [vp, i32, vp, i32, i32, vp, i32]
    lib.hfg_forward_host_wait.restype = ctypes.c_int
    lib.hfg_forward_host_wait.argtypes = [vp, i32]
    lib.hfg_forward_lengths.restype = ctypes.c_int
    lib.hfg_forward_lengths.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, ctypes.c_size_t, i32, vp]
    lib.hfg_receptive_radius.restype = ctypes.c_int
    lib.hfg_receptive_radius.argtypes = [ctypes.POINTER(HfgConfig), ctypes.POINTER(ctypes.c_int32)]
    lib.hfg_mel_create.restype = ctypes.c_int
    lib.hfg_mel_create.argtypes = [vp, ctypes.POINTER(vp)]
    lib.hfg_mel_destroy.restype = None
    lib.hfg_mel_destroy.argtypes = [vp]
    lib.hfg_mel_last_error.restype = ctypes.c_char_p
    lib.hfg_mel_last_error.argtypes = [vp]
    lib.hfg_mel_frames.restype = ctypes.c_int
    lib.hfg_mel_frames.argtypes = [vp, ctypes.c_int64, i64p]
    lib.hfg_log_mel.restype = ctypes.c_int
    lib.hfg_log_mel.argtypes = [vp, vp, i32, ctypes.c_int64, vp, vp]
    lib.hfg_log_mel_l1.restype = ctypes.c_int
    lib.hfg_log_mel_l1.argtypes = [vp, vp, vp, i32, ctypes.c_int64, vp, vp, vp]
    lib.hfg_ard_create.restype = ctypes.c_int
    lib.hfg_ard_create.argtypes = [vp, ctypes.POINTER(vp)]
    lib.hfg_ard_destroy.restype = None
    lib.hfg_ard_destroy.argtypes = [vp]
    lib.hfg_ard_last_error.restype = ctypes.c_char_p
    lib.hfg_ard_last_error.argtypes = [vp]
    lib.hfg_ard_set_weight.restype = ctypes.c_int
    lib.hfg_ard_set_weight.argtypes = [vp, ctypes.c_char_p, vp, i64p, i32]
    lib.hfg_ard_commit_weights.restype = ctypes.c_int
    lib.hfg_ard_commit_weights.argtypes = [vp]
    lib.hfg_ard_workspace_bytes.restype = ctypes.c_int
    lib.hfg_ard_workspace_bytes.argtypes = [vp, i32, i32, i32, ctypes.POINTER(ctypes.c_size_t)]
    lib.hfg_ard_decode.restype = ctypes.c_int
    lib.hfg_ard_decode.argtypes = [vp, vp, i32, i32, i32, vp, vp, ctypes.c_size_t, vp]
    lib.hfg_ard_last_launch_count.restype = ctypes.c_int
    lib.hfg_ard_last_launch_count.argtypes = [vp, i64p]
    lib.hfg_bench_layer.restype = ctypes.c_int
    lib.hfg_bench_layer.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, fp]
    del fp
    _lib = lib
    return lib


def make_config(n_mels: int, upsample_rates: Sequence[int], upsample_kernel_sizes: Sequence[int],
                upsample_initial_channel: int, resblock_kernel_sizes: Sequence[int],
                resblock_dilation_sizes: Sequence[Sequence[int]]) -> HfgConfig:
    if len(upsample_rates) != len(upsample_kernel_sizes):
        raise ValueError("upsample_rates and upsample_kernel_sizes differ in length")
    if len(upsample_rates) > HFG_MAX_STAGES or len(resblock_kernel_sizes) > HFG_MAX_STAGES:
        raise ValueError(f"at most {HFG_MAX_STAGES} upsample stages / resblocks are supported")
    c = HfgConfig()
    c.n_mels = int(n_mels)
    c.num_upsamples = len(upsample_rates)
    c.upsample_initial_channel = int(upsample_initial_channel)
    # the reference zips kernel sizes with dilation lists (models/hifigan.py:111)
    pairs = list(zip(resblock_kernel_sizes, resblock_dilation_sizes))
    c.num_resblocks = len(pairs)
    for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
        c.upsample_rates[i], c.upsample_kernel_sizes[i] = int(u), int(k)
    for j, (k, dils) in enumerate(pairs):
        if len(dils) > HFG_MAX_STAGES:
            raise ValueError(f"at most {HFG_MAX_STAGES} dilations per resblock are supported")
        c.resblock_kernel_sizes[j] = int(k)
        c.num_dilations[j] = len(dils)
        for l, d in enumerate(dils):
            c.resblock_dilations[j][l] = int(d)
    return c


def receptive_radius(cfg: HfgConfig) -> int:
    """Receptive radius of one output frame in mel frames (host-only, no device needed)."""
    out = ctypes.c_int32()
    rc = load().hfg_receptive_radius(ctypes.byref(cfg), ctypes.byref(out))
    if rc != OK:
        raise HfgError(rc, "invalid generator configuration")
    return out.value


class Handle:
    """Owns one hfg_handle (one per CUDA device)."""

    def __init__(self, cfg: HfgConfig):
        self._lib = load()
        self._h = ctypes.c_void_p()
        rc = self._lib.hfg_create(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != OK:
            why = {ERR_CUDA: "no usable CUDA device (this path has no CPU fallback)",
                   ERR_INVALID: "invalid generator configuration"}.get(rc, "hfg_create failed")
            raise HfgError(rc, why)

    def _check(self, rc: int):
        if rc != OK:
            raise HfgError(rc, self._lib.hfg_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.hfg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weight(self, name: str, host_ptr: int, shape: Sequence[int]):
        shp = (ctypes.c_int64 * len(shape))(*shape)
        self._check(self._lib.hfg_set_weight(self._h, name.encode(), ctypes.c_void_p(host_ptr), shp, len(shape)))

    def commit(self):
        self._check(self._lib.hfg_commit_weights(self._h))

    def out_len(self, frames: int) -> int:
        out = ctypes.c_int64()
        self._check(self._lib.hfg_out_len(self._h, frames, ctypes.byref(out)))
        return out.value

    def workspace_bytes(self, batch: int, frames: int, mode: int) -> int:
        out = ctypes.c_size_t()
        self._check(self._lib.hfg_workspace_bytes(self._h, batch, frames, mode, ctypes.byref(out)))
        return out.value

    def forward(self, mel_ptr: int, batch: int, frames: int, wav_ptr: int, ws_ptr: int, ws_bytes: int,
                mode: int, stream: int, stage_ptrs: Optional[Sequence[int]] = None):
        if stage_ptrs is None:
            rc = self._lib.hfg_forward(self._h, mel_ptr, batch, frames, wav_ptr, ws_ptr, ws_bytes, mode,
                                       ctypes.c_void_p(stream))
        else:
            arr = (ctypes.c_void_p * len(stage_ptrs))(*stage_ptrs)
            rc = self._lib.hfg_forward_stages(self._h, mel_ptr, batch, frames, wav_ptr, ws_ptr, ws_bytes,
                                              mode, ctypes.c_void_p(stream), arr)
        self._check(rc)

    def forward_lengths(self, mel_ptr: int, lengths_ptr: int, halo: int, batch: int, frames: int, wav_ptr: int,
                        ws_ptr: int, ws_bytes: int, mode: int, stream: int):
        self._check(self._lib.hfg_forward_lengths(self._h, mel_ptr, lengths_ptr, halo, batch, frames, wav_ptr,
                                                  ws_ptr, ws_bytes, mode, ctypes.c_void_p(stream)))

    def forward_host(self, mel_ptr: int, batch: int, frames: int, wav_ptr: int, mode: int,
                     mel_pinned: bool = False, wav_pinned: bool = False):
        flags = (1 if mel_pinned else 0) | (2 if wav_pinned else 0)
        self._check(self._lib.hfg_forward_host_ex(self._h, mel_ptr, batch, frames, wav_ptr, mode, flags))

    def forward_host_submit(self, slot: int, mel_ptr: int, batch: int, frames: int, wav_ptr: int, mode: int):
        self._check(self._lib.hfg_forward_host_submit(self._h, slot, mel_ptr, batch, frames, wav_ptr, mode))

    def forward_host_wait(self, slot: int):
        self._check(self._lib.hfg_forward_host_wait(self._h, slot))

    def set_mel_layout(self, frames_last: bool):
        self._check(self._lib.hfg_set_mel_layout(self._h, 1 if frames_last else 0))

    def set_profiling(self, on):
        """False/0 off, True/1 per launch (serialised), 2 per stage (concurrent resblocks)."""
        self._check(self._lib.hfg_set_profiling(self._h, int(on)))

    def get_profile(self):
        import json
        need = ctypes.c_size_t()
        self._check(self._lib.hfg_get_profile(self._h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        self._check(self._lib.hfg_get_profile(self._h, buf, need.value, ctypes.byref(need)))
        return json.loads(buf.value.decode())

    def bench_layer(self, stage, resblock, pair, which, batch, rows, mode, iters=20) -> float:
        ms = ctypes.c_float()
        self._check(self._lib.hfg_bench_layer(self._h, stage, resblock, pair, which, batch, rows, mode, iters,
                                              ctypes.byref(ms)))
        return ms.value

    def last_launch_count(self) -> int:
        out = ctypes.c_int64()
        self._check(self._lib.hfg_last_launch_count(self._h, ctypes.byref(out)))
        return out.value

    def tf32_plan_is_split(self) -> bool:
        """True when HFG_MODE_TF32 runs on fp16 operand planes with an fp16 hi + lo residual stream (hfg_tf32_plan)."""
        out = ctypes.c_int32()
        self._check(self._lib.hfg_tf32_plan(self._h, ctypes.byref(out)))
        return bool(out.value)
