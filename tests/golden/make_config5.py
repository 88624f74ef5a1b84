"""BASELINE.json configs[4] fixture: the UNMODIFIED reference SAM-BERT acoustic model
(reference models/acoustic_model.py:181-265, stdout silenced) on synthetic phoneme sequences ->
mel_pred [B, Tfrm, 80] and the integer frame durations (reference models/variance_adaptor.py:746-748).
The acoustic model cannot travel to the GPU box; its output can.

    python tests/golden/make_config5.py        # writes tests/golden/config5_acoustic_b8.npz

B = 8 sequences of 24 phonemes are stored (the 64-sequence batch of the config is 8 such groups);
the GPU test feeds them through the B200 generator and compares with the oracle."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from tts_sambert_hifigan_b200 import synth   # noqa: E402
from models.acoustic_model import SAMBERTAcousticModel   # noqa: E402  (the reference)

if __name__ == "__main__":
    torch.manual_seed(1234)
    B, Tph = 8, 24
    u = synth.uniform01(55, 3 * B * Tph).reshape(3, B, Tph)
    ph = torch.from_numpy((u[0] * 300).astype(np.int64))
    tone = torch.from_numpy((u[1] * 10).astype(np.int64))
    bnd = torch.from_numpy((u[2] * 5).astype(np.int64))
    with contextlib.redirect_stdout(io.StringIO()):
        model = SAMBERTAcousticModel().eval()
        mel_pred, pred = model.inference(ph, tone, bnd)
    dur = pred["dur"] if "dur" in pred else pred.get("duration")
    print("mel_pred", tuple(mel_pred.shape), "dur sum per utterance", dur.sum(dim=1).tolist())
    np.savez_compressed(os.path.join(HERE, "config5_acoustic_b8.npz"),
                        mel_pred=mel_pred.numpy().astype(np.float32), dur=dur.numpy().astype(np.int64),
                        ph_ids=ph.numpy(), tone_ids=tone.numpy(), boundary_ids=bnd.numpy())
