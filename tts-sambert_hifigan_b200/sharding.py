"""Spreading the generator over the GPUs of one box: one process per GPU,
utterance sharding or time chunking with a receptive-field halo, and an optional
final gather.  There is no collective on the data path: every rank can be handed
the (small) mel up front, computes its own slice of the output, and only the
finished waveform (256 samples per frame) is exchanged.

Why time chunking is exact: the reference generator has no normalisation,
attention or recurrence -- every op is a local convolution
(reference models/hifigan.py:72-86,116-131,224-261) -- and one output frame
depends on mel frames [f-13, f+13] only.  Generating frames
[a, b) from mel[a-h : b+h] with h >= 14 and cropping reproduces the unchunked
run; at true utterance edges the slice simply ends, so each layer applies its
own zero padding exactly as in the full run.

`generate` arguments are any callable mel[B, n_mels, T] -> wav[B, 1, T*hop]
(the B200 HiFiGANGenerator in production, the CPU oracle in the gloo tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

RECEPTIVE_HALO_FRAMES = 14      # default configuration: exact receptive radius 13 frames, + 1


def resolve_halo(generate, halo: Optional[int]) -> int:
    """Halo in frames for `generate`.  A generator that knows its own geometry (HiFiGANGenerator.receptive_radius,
    derived from kernel sizes / dilations / rates) supplies radius + 1 as the default and rejects anything
    smaller than the radius -- a too-small halo silently changes the output near chunk borders.  A plain callable
    (the CPU oracle in the gloo tests) gets the default configuration's 14."""
    owner = getattr(generate, "__self__", generate)
    radius = getattr(owner, "receptive_radius", None)
    if halo is None:
        return RECEPTIVE_HALO_FRAMES if radius is None else int(radius) + 1
    if radius is not None and halo < int(radius):
        raise ValueError(f"halo={halo} frames is smaller than the receptive radius {int(radius)} of this generator")
    return int(halo)


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of `rank` among `world` (sizes differ by <= 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class Chunk:
    start: int      # first output frame of this chunk
    stop: int       # one past the last output frame
    lo: int         # mel slice [lo, hi) fed to the generator
    hi: int

    @property
    def crop_front(self) -> int:       # frames to drop from the front of the chunk's output
        return self.start - self.lo

    @property
    def frames(self) -> int:
        return self.stop - self.start


def plan_chunks(frames: int, n_chunks: int, halo: int = RECEPTIVE_HALO_FRAMES) -> List[Chunk]:
    """Split [0, frames) into n_chunks contiguous ranges, each read with a
    `halo`-frame margin clipped at the true edges.  Empty ranges (more chunks
    than frames) are dropped."""
    if frames <= 0 or n_chunks <= 0 or halo < 0:
        raise ValueError("frames and n_chunks must be positive, halo non-negative")
    out = []
    for r in range(n_chunks):
        a, b = shard_bounds(frames, n_chunks, r)
        if b > a:
            out.append(Chunk(a, b, max(0, a - halo), min(frames, b + halo)))
    return out


def run_chunk(generate: Callable[[torch.Tensor], torch.Tensor], mel: torch.Tensor, c: Chunk, hop: int) -> torch.Tensor:
    """wav[B, 1, c.frames*hop] for one chunk (halo cropped)."""
    wav = generate(mel[:, :, c.lo:c.hi].contiguous())
    if wav.shape[-1] != (c.hi - c.lo) * hop:
        raise RuntimeError("chunked generation needs T_out == T*hop (upsample kernels with even k-u)")
    return wav[:, :, c.crop_front * hop:(c.crop_front + c.frames) * hop]


def generate_chunked(generate, mel: torch.Tensor, n_chunks: int, hop: int = 256,
                     halo: Optional[int] = None) -> torch.Tensor:
    """Single-process time chunking (bounded memory for long-form input)."""
    halo = resolve_halo(generate, halo)
    parts = [run_chunk(generate, mel, c, hop) for c in plan_chunks(mel.shape[-1], n_chunks, halo)]
    return torch.cat(parts, dim=-1)


# ----------------------------------------------------------------------------
# multi-process (one rank per GPU)
# ----------------------------------------------------------------------------

def _gather_var(local: torch.Tensor, sizes: Sequence[int], dim: int, group=None,
                dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """all_gather (or gather to dst) of tensors whose size differs along `dim`.
    The only collective of the path: the final waveform gather."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mx = max(sizes)
    pad_shape = list(local.shape)
    pad_shape[dim] = mx
    buf = local.new_zeros(pad_shape)
    buf.narrow(dim, 0, local.shape[dim]).copy_(local)
    if dst is None:
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf, group=group)
    else:
        outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, outs, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([o.narrow(dim, 0, s) for o, s in zip(outs, sizes) if s > 0], dim=dim)


def generate_utterance_sharded(generate, mel: torch.Tensor, group=None, gather: bool = True,
                               dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Every rank holds the full mel batch [B, n_mels, T], generates utterances
    shard_bounds(B, world, rank) and (optionally) gathers the waveforms.
    With gather=False the local shard is returned (no collective at all)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = mel.shape[0]
    a, b = shard_bounds(B, world, rank)
    if b > a:
        local = generate(mel[a:b].contiguous())
    else:
        probe = generate(mel[:1].contiguous())           # rank without work: shape only
        local = probe[:0]
    if not gather:
        return local
    sizes = [shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world)]
    if b == a:   # need a correctly shaped empty shard for the padded gather
        local = local.new_zeros((0,) + tuple(local.shape[1:]))
    return _gather_var(local, sizes, 0, group, dst)


def generate_time_sharded(generate, mel: torch.Tensor, hop: int = 256, halo: Optional[int] = None,
                          group=None, gather: bool = True, dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Long-form input: rank r generates frame range shard_bounds(T, world, r)
    from its mel slice with halo, then the cropped pieces are gathered along time."""
    halo = resolve_halo(generate, halo)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    T = mel.shape[-1]
    a, b = shard_bounds(T, world, rank)
    sizes = [(shard_bounds(T, world, r)[1] - shard_bounds(T, world, r)[0]) * hop for r in range(world)]
    if b > a:
        c = Chunk(a, b, max(0, a - halo), min(T, b + halo))
        local = run_chunk(generate, mel, c, hop)
    else:
        local = mel.new_zeros((mel.shape[0], 1, 0))
    if not gather:
        return local
    return _gather_var(local, sizes, 2, group, dst)
