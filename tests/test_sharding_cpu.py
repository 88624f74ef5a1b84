"""Host-side multi-GPU logic on CPU: chunk planning, exactness of halo chunking
(against the oracle), and the world_size-2 gloo path of both sharding modes."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from tts_sambert_hifigan_b200 import sharding, synth

SMALL = dict(n_mels=16, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4],
             upsample_initial_channel=32, resblock_kernel_sizes=[3, 7],
             resblock_dilation_sizes=[[1, 3], [1, 5]])
HOP = 8


def _gen(cfg, seed=3):
    sd = {k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, seed).items()}
    return lambda mel: oracle.forward_torch(cfg, sd, mel)


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 16, 256):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def test_plan_chunks_config4_geometry():
    # BASELINE.json configs[3]: 5168 frames over 8 GPUs, 14-frame halo
    chunks = sharding.plan_chunks(5168, 8, 14)
    assert [c.frames for c in chunks] == [646] * 8
    assert chunks[0].lo == 0 and chunks[0].hi == 646 + 14
    assert chunks[3].lo == 3 * 646 - 14 and chunks[3].hi == 4 * 646 + 14
    assert chunks[-1].hi == 5168
    # redundant compute: 14 halo frames per interior side
    total = sum(c.hi - c.lo for c in chunks)
    assert total == 5168 + 14 * 14
    assert [c.frames for c in sharding.plan_chunks(3, 8, 14)] == [1, 1, 1]     # more chunks than frames


def test_halo_chunking_is_exact_default_config():
    """fp64 so that 'exact' means exact: chunked == unchunked with the 14-frame halo,
    and a too-small halo is detectably wrong (receptive radius is 13 frames)."""
    cfg = synth.DEFAULT_CONFIG
    sd = {k: torch.from_numpy(v).double() for k, v in synth.make_weights(cfg, 1).items()}
    gen = lambda m: oracle.forward_torch(cfg, sd, m)
    mel = torch.from_numpy(synth.make_mel(2, 1, 80, 90)).double()
    full = gen(mel)
    got = sharding.generate_chunked(gen, mel, 3, hop=256, halo=14)
    assert got.shape == full.shape
    assert float((got - full).abs().max()) <= 1e-15
    bad = sharding.generate_chunked(gen, mel, 3, hop=256, halo=6)
    assert float((bad - full).abs().max()) > 1e-6


def test_chunking_rejects_odd_upsample_geometry():
    cfg = dict(synth.DEFAULT_CONFIG, upsample_rates=[5, 5, 4, 2], upsample_kernel_sizes=[10, 10, 8, 4])
    with pytest.raises(RuntimeError):
        sharding.generate_chunked(_gen(cfg), torch.zeros(1, 80, 40), 2, hop=200)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    gen = _gen(SMALL)
    mel = torch.from_numpy(synth.make_mel(5, 5, SMALL["n_mels"], 61))   # 5 utterances: ragged 3/2 split
    full = gen(mel)
    a = sharding.generate_utterance_sharded(gen, mel)
    b = sharding.generate_time_sharded(gen, mel, hop=HOP, halo=20)
    local = sharding.generate_utterance_sharded(gen, mel, gather=False)
    c = sharding.generate_utterance_sharded(gen, mel, dst=0)
    # more ranks than utterances: one rank has nothing to do
    d = sharding.generate_utterance_sharded(gen, mel[:1])
    res = dict(rank=rank,
               utt=float((a - full).abs().max()), time=float((b - full).abs().max()),
               local_shape=list(local.shape), dst_none=(c is None),
               dst_err=None if c is None else float((c - full).abs().max()),
               single=float((d - full[:1]).abs().max()))
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharding(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    for r in (r0, r1):
        assert r["utt"] <= 1e-6                      # ATen may block differently per batch size
        assert r["time"] <= 1e-6                     # halo chunking: equal up to fp32 re-association
        assert r["single"] <= 1e-6
    assert r0["local_shape"] == [3, 1, 61 * HOP] and r1["local_shape"] == [2, 1, 61 * HOP]
    assert r0["dst_none"] is False and r0["dst_err"] <= 1e-6 and r1["dst_none"] is True
