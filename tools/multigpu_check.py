"""Run under torchrun on N GPUs of one box: checks both sharding modes of the real CUDA generator
over NCCL against the single-GPU run (rank 0 recomputes everything locally).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import tts_sambert_hifigan_b200 as pkg
from tts_sambert_hifigan_b200 import sharding, synth

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = synth.DEFAULT_CONFIG
for mode in ("tf32", "fp16", "bf16"):
    gen = pkg.HiFiGANGenerator(**cfg, mode=mode).to(dev)
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_weights(cfg, 0).items()})
    mel = torch.from_numpy(synth.make_mel(3, 2 * world + 1, 80, 64)).to(dev)       # ragged utterance split
    long_mel = torch.from_numpy(synth.make_mel(4, 1, 80, 646 * world)).to(dev)     # config-4 style
    with torch.no_grad():
        a = sharding.generate_utterance_sharded(gen, mel)
        b = sharding.generate_time_sharded(gen, long_mel, hop=256, halo=14)
        full_a = gen(mel)
        full_b = gen(long_mel)
    torch.cuda.synchronize()
    ea = float((a - full_a).abs().max()); eb = float((b - full_b).abs().max())
    print(f"rank {rank}/{world} mode {mode}: utterance-sharded err {ea:.3e}  time-sharded err {eb:.3e}", flush=True)
    assert ea == 0.0 and eb == 0.0, (ea, eb)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("multigpu_check OK")
