"""Golden vectors for the KV-cached autoregressive decoder from the LIVE, unmodified
reference code: models/ar_decoder.py::PNCAARDecoder (inference loop :167-238) and, for the BASELINE config-5
case at its stated batch of 64, the whole models/acoustic_model.py::SAMBERTAcousticModel.

    python tests/golden/make_ar_decoder.py

The decoder's 9.4 M parameters cannot be committed, so they come from the portable generator
(synth.make_ard_weights) and are loaded into the reference module with load_state_dict -- reference code, custom
weights.  Stored:
  ar_decoder_b8.npz          mel_pred [8, 96, 80] of PNCAARDecoder on Hvar = synth.normal(301, (8, 96, 256))
  config5_acoustic_b64.npz   SAMBERTAcousticModel.inference on 64 synthetic phoneme sequences (its ar_decoder
                             holding the portable weights): the decoder's input Hvar (forward pre-hook),
                             mel_pred, dur, and the phoneme-level encoder output + durations for the length regulator
"""
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
from tts_sambert_hifigan_b200 import synth                       # noqa: E402
from models.ar_decoder import PNCAARDecoder                      # noqa: E402  (the reference)
from models.acoustic_model import SAMBERTAcousticModel           # noqa: E402  (the reference)


def portable_decoder_state(ref_decoder, seed):
    sd = {k: torch.from_numpy(v) for k, v in synth.make_ard_weights(synth.ARD_DEFAULT, seed).items()}
    sd["pos_encoding.pe"] = ref_decoder.state_dict()["pos_encoding.pe"]
    assert set(sd) == set(ref_decoder.state_dict())
    for k, v in ref_decoder.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    return sd


if __name__ == "__main__":
    quiet = contextlib.redirect_stdout(io.StringIO())
    torch.set_num_threads(os.cpu_count() or 1)
    # ---- decoder alone, B = 8, 96 frames ----
    with quiet:
        dec = PNCAARDecoder(**synth.ARD_DEFAULT).eval()
    dec.load_state_dict(portable_decoder_state(dec, 300))
    hvar = torch.from_numpy(synth.normal(301, (8, 96, 256)))
    with torch.no_grad(), quiet:
        mel = dec(hvar)
    np.savez_compressed(os.path.join(HERE, "ar_decoder_b8.npz"), mel_pred=mel.numpy().astype(np.float32))
    print("ar_decoder_b8", tuple(mel.shape), "peak", float(mel.abs().max()))

    # ---- config 5 at B = 64 through the whole acoustic model ----
    torch.manual_seed(4321)
    B, Tph = 64, 8
    u = synth.uniform01(56, 3 * B * Tph).reshape(3, B, Tph)
    ph = torch.from_numpy((u[0] * 300).astype(np.int64))
    tone = torch.from_numpy((u[1] * 10).astype(np.int64))
    bnd = torch.from_numpy((u[2] * 5).astype(np.int64))
    with quiet:
        model = SAMBERTAcousticModel().eval()
    model.ar_decoder.load_state_dict(portable_decoder_state(model.ar_decoder, 300))
    grabbed = {}
    model.ar_decoder.register_forward_pre_hook(lambda m, args: grabbed.__setitem__("hvar", args[0].detach().clone()))
    model.variance_adaptor.length_regulator.register_forward_pre_hook(
        lambda m, args: grabbed.update(henc=args[0].detach().clone(), lr_dur=args[1].detach().clone()))
    model.variance_adaptor.length_regulator.register_forward_hook(
        lambda m, args, out: grabbed.__setitem__("hlr", out.detach().clone()))
    with torch.no_grad(), quiet:
        mel_pred, pred = model.inference(ph, tone, bnd)
    dur = pred["dur"] if "dur" in pred else pred.get("duration")
    print("config5 B=64: mel_pred", tuple(mel_pred.shape), "Hvar", tuple(grabbed["hvar"].shape),
          "frames per utterance", dur.sum(dim=1).tolist()[:8], "...")
    np.savez_compressed(os.path.join(HERE, "config5_acoustic_b64.npz"),
                        hvar=grabbed["hvar"].numpy().astype(np.float32), mel_pred=mel_pred.numpy().astype(np.float32),
                        dur=dur.numpy().astype(np.int64), henc=grabbed["henc"].numpy().astype(np.float32),
                        hlr=grabbed["hlr"].numpy().astype(np.float32))
    print("sizes:", {f: os.path.getsize(os.path.join(HERE, f)) for f in ("ar_decoder_b8.npz", "config5_acoustic_b64.npz")})
