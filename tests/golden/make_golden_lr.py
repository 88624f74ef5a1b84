"""Golden vectors for the length regulator / duration rounding from the LIVE reference
(models/variance_adaptor.py:120-269, :746-748).  python tests/golden/make_golden_lr.py"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")
from tts_sambert_hifigan_b200 import synth                      # noqa: E402
from models.variance_adaptor import LengthRegulator             # noqa: E402  (the reference)

if __name__ == "__main__":
    B, Tph, D = 4, 37, 24
    henc = synth.normal(201, (B, Tph, D))
    dur = (synth.uniform01(202, B * Tph).reshape(B, Tph) * 9).astype(np.int64) - 1   # -1 .. 7: zeros and negatives
    dur[3] = 0
    dur[3, 5] = 2                                                                     # nearly empty utterance
    with contextlib.redirect_stdout(io.StringIO()):
        hlr = LengthRegulator()(torch.from_numpy(henc), torch.from_numpy(dur)).numpy()
    log_dur = synth.uniform(203, (B, 400), 3.0)                                       # exp -> 0.05 .. 20 frames
    d = torch.clamp(torch.exp(torch.from_numpy(log_dur)).round().long(), min=1).numpy()   # reference :746-748
    np.savez_compressed(os.path.join(HERE, "length_regulator.npz"), hlr=hlr, dur=dur, dur_from_log=d)
    print("hlr", hlr.shape, "dur sums", np.clip(dur, 0, None).sum(1).tolist(), "dur_from_log range", d.min(), d.max())
